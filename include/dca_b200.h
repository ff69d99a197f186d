/*
 * dca_b200.h -- C ABI of pydca_b200/_plmdcaBackend_b200.so, the B200-native
 * replacement for pydca's DCA hot path.  Plain C: pointers and sizes only, no
 * torch / C++ types.  Every function returns 0 on success (or a non-NULL
 * pointer); on failure it returns non-zero / NULL and dca_last_error() holds the
 * reason.  Nothing here throws across the ABI and nothing falls back to the CPU:
 * without a CUDA device every compute entry point fails.
 *
 * Reference interfaces replaced (paths relative to the pydca source tree):
 *   plmdcaBackend / freeFieldsAndCouplings   pydca/plmdca/plmdcaBackend.cpp:151-223,
 *                                            bound by ctypes at pydca/plmdca/plmdca.py:79-89
 *   lbfgs_evaluate_t (objective operator)    pydca/plmdca/lbfgs/include/lbfgs.h:380-419
 *   PlmDCA::computeSeqsWeight                pydca/plmdca/plmdca_numerics.cpp:611-671
 *   msa_numerics operator functions          pydca/meanfield_dca/msa_numerics.py:14-342,
 *                                            called at pydca/meanfield_dca/meanfield_dca.py:371-548
 *   FN / APC scoring loops                   pydca/plmdca/plmdca.py:437-524,
 *                                            pydca/meanfield_dca/meanfield_dca.py:902-988
 */
#ifndef DCA_B200_H
#define DCA_B200_H

#include <stdbool.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- diagnostics -------------------------------------------------------- */
const char *dca_last_error(void);
int dca_device_count(void);                 /* CUDA devices visible (0 => compute calls fail) */
int dca_set_device(int device);
unsigned long long dca_kernel_launches(void); /* kernels launched by this library so far */

/* ---- the reference's own ABI (drop-in) ------------------------------------
 * plmdcaBackend.cpp:151-156.  biomolecule: 1 = protein, anything else = RNA.
 * msa_file: FASTA path.  Returns P = L*q + L(L-1)/2*q*q floats, fields first
 * (index a + q*i), then J_ij (i<j) blocks in pair order, element a*q + b
 * (plmdca_numerics.cpp:338-342).  The buffer is host memory owned by the
 * library; release it with freeFieldsAndCouplings (plmdcaBackend.cpp:204).
 * num_threads is accepted for signature compatibility and ignored (the work
 * runs on the GPU).  Returns NULL on error (the reference aborts instead).
 * GPUs: the call drives DCA_B200_GPUS GPUs of this node (default: all visible; dca_set_gpus
 * overrides) -- one host thread per GPU inside this one call, owner-sharded sites, peer-memory
 * exchange, every GPU copying its slice of the result into the returned vector.  Falls back to
 * one GPU when peer access is unavailable. */
float *plmdcaBackend(unsigned short biomolecule, unsigned short num_site_states,
                     const char *msa_file, unsigned int seqs_len, float seqid,
                     float lambda_h, float lambda_J, unsigned int max_iteration,
                     unsigned int num_threads, bool verbose);
void freeFieldsAndCouplings(void *h_and_J);
/* process-wide GPU count for plmdcaBackend(): n >= 1, or 0 = DCA_B200_GPUS / all visible */
int dca_set_gpus(int n);

/* ---- FASTA ingest (host logic) ----------------------------------------------
 * plmdca_numerics.cpp:685-767 / fasta_reader.py:122-163: upper-case, encode to
 * 0..q-1 (gap and every unknown letter = q-1), keep the first occurrence of
 * each distinct sequence in file order.  *msa_out is a u8[N][seqs_len] buffer owned by
 * the library (the reader's own output, no copy) until it is released with dca_free_host
 * -- not with free().  Large files are parsed on host threads (DCA_B200_READER_THREADS).
 * Returns N (>0) or a negative error. */
long dca_read_fasta(const char *msa_file, int biomolecule, int seqs_len, uint8_t **msa_out);
void dca_free_host(void *p);

/* ---- sequence reweighting ----------------------------------------------------
 * counts[i] = #{ j : ident(i,j)/L > seqid }, self included (bit-exact integers).
 * compare_f64 = 0: the C++ float32 compare (plmdca_numerics.cpp:642-643);
 * compare_f64 = 1: the numba float64 compare (msa_numerics.py:46).
 * Host buffers in, host buffer out; rows [row_lo,row_hi) only (0,N for all). */
int dca_weights(const uint8_t *msa, long num_seqs, int seqs_len, double seqid,
                int compare_f64, long row_lo, long row_hi, uint32_t *counts_out);

/* Symmetric variant for deep alignments and for several GPUs: ident(i,j) = ident(j,i), so only
 * the tiles on and above the diagonal are compared and a hit counts for both rows (half the
 * compares of dca_weights on a row range).  128-row tiles are dealt to `nparts` parts cyclically;
 * partial_counts_out[N] receives part `part`'s contribution to EVERY row -- the element-wise
 * sum over the parts (one all-reduce of N integers across GPUs) is counts[].  nparts = 1 gives
 * the counts directly (dca_weights on the whole alignment takes this path).  device_ms_out
 * (may be NULL): CUDA-event time of pack + compare kernels, H2D / D2H excluded. */
int dca_weights_part(const uint8_t *msa, long num_seqs, int seqs_len, double seqid,
                     int compare_f64, int part, int nparts, uint32_t *partial_counts_out,
                     float *device_ms_out);

/* ---- plmDCA objective / optimiser ------------------------------------------------
 * A dca_plm handle owns the device-resident MSA, weights and work buffers of one
 * objective (PlmDCA object of plmdca.h:16).  site_lo/site_hi shard the data term
 * by sites for multi-GPU use (0, L for the whole objective). */
typedef struct dca_plm dca_plm;
dca_plm *dca_plm_create(const uint8_t *msa, const float *weights, long num_seqs,
                        int seqs_len, int num_site_states, float lambda_h,
                        float lambda_J, int site_lo, int site_hi);
void dca_plm_destroy(dca_plm *p);
long dca_plm_num_params(const dca_plm *p);
/* carry = 1 (default) reproduces the reference's loop-carried prob_ni
 * (plmdca_numerics.cpp:492,:499); carry = 0 is textbook plmDCA.  chunk / warm:
 * chain length and warm-up overlap of the chunked scan (<=0 keeps the default 32/32). */
int dca_plm_set_carry(dca_plm *p, int carry, int chunk, int warm);
/* Arithmetic of the optimiser's inner products and of the L2 term of fx.
 *   mode 1 ("reference"): the reference's own -- SEQUENTIAL float32 sums over all P
 *     parameters (vecdot, arithmetic_ansi.h:114-121; fx += lambda*v*v,
 *     plmdca_numerics.cpp:463-486; fx += fxi in site order, :572), reproduced bit for bit
 *     by an exact block-parallel kernel (csrc/seqsum.cu) together with the reference's own
 *     sequence of two-loop passes (lbfgs.cpp:568-601).  Those sums carry a deterministic
 *     rounding bias (1.8 % of |g|^2 at N=5000, L=150) that steers the reference's L-BFGS
 *     trajectory; with it the fit reproduces the reference's to ~1e-6 on FN scores.
 *   mode 0 ("accurate"): float64 fixed-order reductions and the compact two-pass
 *     recursion -- faster, but a different (better conditioned) trajectory: FN scores differ
 *     from the reference's by a few 1e-4, exactly as the reference's own would with
 *     float64 sums.
 * Single-GPU handles only.  plmdcaBackend() picks the mode from DCA_B200_ARITHMETIC
 * (accurate | reference; default accurate). */
int dca_plm_set_fidelity(dca_plm *p, int mode);
/* process-wide choice for plmdcaBackend(): 0 accurate, 1 reference, -1 back to the environment
 * variable.  (The reference's entry point has no room for an extra argument.) */
int dca_set_arithmetic(int mode);
/* test / measurement hook of that kernel on HOST buffers: *out = the sequential float32
 * dot product of x and y (y != NULL), or the sequential L2 accumulation of x with lambda0 on
 * the first n0 entries and lambda1 on the rest (y == NULL). */
int dca_seq_sum(const float *x, const float *y, long n, long n0, float lambda0, float lambda1, float *out);
/* the same with the CUDA-event time of one (warm) launch and the pipeline's counters: stats8 =
 * {chunks applied from producer records, chunks done by the consumer CTA's exact path, ready
 * time-outs, stale / unusable records, failed prefix-range checks, consumer kilocycles (total,
 * exact path), chunks walked by the chain warp from the producers' row summaries} */
int dca_seq_sum_stats(const float *x, const float *y, long n, long n0, float lambda0, float lambda1,
                      float *out, unsigned *stats8, float *kernel_ms);
/* initial point, PlmDCA::initFieldsAndCouplings (plmdca_numerics.cpp:207-249) */
int dca_plm_init(dca_plm *p, float *x_out_host);
/* objective + gradient with HOST buffers (x in, g out, *fx out): PlmDCA::gradient
 * (plmdca_numerics.cpp:436-607).  H2D of x and D2H of g are part of the call. */
int dca_plm_eval(dca_plm *p, const float *x_host, float *g_host, double *fx_out);
/* the lbfgs_evaluate_t contract on DEVICE pointers (no copies, no sync):
 * scalars_out_host may be NULL; otherwise receives {fx, g.dir, g.g, x.x}. */
int dca_plm_eval_device(dca_plm *p, const float *x_dev, float *g_dev,
                        const float *dir_dev, double *scalars_out_host);
/* measurement hooks: one evaluation with CUDA events between its kernels, on the
 * stream they run on: ms_out[6] = unpack, logits, softmax scan, scatter, export
 * (owner-sharded handles only, else 0), assemble+finalize.  out10 = {N, L, q, ld, P, nsplit, nchunk, col_tiles, Npad, nscan}. */
int dca_plm_eval_timed(dca_plm *p, const float *x_host, float *ms_out);
int dca_plm_describe(const dca_plm *p, long *out10);
/* out4 = {1 if the handle uses the similarity-sorted row layout, 1 if its scatter uses four-row group sums, the
 * fraction of (site, group of four rows) pairs whose rows carry one residue -- what the equal-residue fast paths
 * of the logits / scatter kernels save --, columns per logits CTA}.  Measurement only. */
int dca_plm_layout_stats(const dca_plm *p, double *out4);
/* measurement hook for the HBM-bound part of the path: CUDA-event ms of the two L-BFGS
 * passes of one iteration and of the trial-step kernel on n floats with `stored_pairs`
 * (1..5) history pairs: ms_out3 = {history update + inner products, direction, step}. */
int dca_lbfgs_vector_bench(long n, int stored_pairs, int reps, float *ms_out3);
/* the cudaStream_t every kernel of this handle is launched on */
void *dca_plm_cuda_stream(const dca_plm *p);
/* optional collective for site-sharded objectives: sum g (n floats) and the
 * n_scalars doubles at scalars_dev over all ranks; both are DEVICE pointers.  The
 * callback must ENQUEUE the reduction in stream order on dca_plm_cuda_stream(p) (it is
 * called right after the evaluation kernels were launched there, without a host
 * synchronisation) and may return before it has completed. */
typedef int (*dca_allreduce_fn)(void *ctx, float *g_dev, long n, double *scalars_dev, int n_scalars);
typedef struct {
    int status;          /* libLBFGS code: 0 converged, -997 max iterations, -1001 line-search rounding ... */
    int iterations;
    int evaluations;     /* objective/gradient evaluations = 1 + sum of line-search trials */
    float fx;
    double device_ms;    /* CUDA-event time of the whole optimisation */
    double timed_ms;     /* CUDA-event time after `warmup_iterations` iterations */
    int timed_evaluations;
} dca_fit_result;
/* lbfgs() as configured at plmdcaBackend.cpp:68-81, starting from x_inout_host
 * (host buffer, P floats; overwritten with the result). */
int dca_plm_fit(dca_plm *p, float *x_inout_host, int max_iterations, int verbose,
                int warmup_iterations, dca_allreduce_fn allreduce, void *allreduce_ctx,
                dca_fit_result *result);
/* the same from the reference's initial point (plmdcaBackend.cpp:77) built on the device; the result
 * stays in the handle's device buffer (dca_plm_x_device): no host copy of the parameters at all.
 * On a sharded handle: collective, local slice. */
int dca_plm_fit_from_init(dca_plm *p, int max_iterations, int verbose, dca_fit_result *result);
/* ---- owner-sharded plmDCA for 2..16 GPUs of one node (one process per GPU) -----------------
 * The reference parallelises PlmDCA::gradient over sites with OpenMP
 * (plmdca_numerics.cpp:488); here rank r of `world` OWNS the sites
 * [bounds[r], bounds[r+1]) (bounds[0] = 0, bounds[world] = L, every rank at least one site)
 * and the slice of the parameter vector made of those sites' fields and of the blocks J_ij
 * whose first site i it owns -- two contiguous runs of the reference's vector
 * (dca_plm_local_layout).  On a sharded handle dca_plm_init / dca_plm_eval / dca_plm_fit
 * take and return that LOCAL slice (dca_plm_local_params floats); the optimiser's vectors are
 * sharded the same way.  Per evaluation the ranks exchange only the cross blocks, by P2P
 * stores over NVLink into each other's arena (fused into the unpack / export kernels), and
 * a few float64 scalars, summed in rank order so that every rank takes the same decisions.
 * Every rank must make the same sequence of eval / fit calls (they are collective).
 * Wiring: exchange dca_plm_ipc_handle() bytes between the processes (e.g. torch.distributed
 * all_gather) and pass all of them to dca_plm_connect_ipc; ranks living in one process pass
 * raw dca_plm_arena() pointers to dca_plm_connect_ptrs instead.  Call dca_plm_disconnect on
 * every rank (then synchronise the ranks) before any rank destroys its handle. */
dca_plm *dca_plm_create_sharded(const uint8_t *msa, const float *weights, long num_seqs,
                                int seqs_len, int num_site_states, float lambda_h,
                                float lambda_J, int rank, int world, const int *bounds);
long dca_plm_local_params(const dca_plm *p);
/* out4 = {first field, number of fields, first coupling element, number of coupling elements}
 * of the local slice inside the reference's vector (plmdca_numerics.cpp:226, :338-342) */
int dca_plm_local_layout(const dca_plm *p, long *out4);
/* the same arithmetic without a handle or a GPU (host only) */
int dca_shard_layout(int seqs_len, int num_site_states, int site_lo, int site_hi, long *out4);
void *dca_plm_arena(const dca_plm *p);                   /* device pointer, NULL unless sharded */
void *dca_plm_x_device(dca_plm *p);                      /* device buffer holding the last init / fit result */
int dca_plm_ipc_handle(const dca_plm *p, unsigned char *out64);
int dca_plm_connect_ipc(dca_plm *p, const unsigned char *handles /* world x 64 bytes, by rank */, int world);
int dca_plm_connect_ptrs(dca_plm *p, void *const *arena_ptrs /* world entries, by rank */);
int dca_plm_disconnect(dca_plm *p);

/* ---- asymmetric (per-site) estimator -- SURVEY 8e last row, what north_star words as "plmDCA's
 * independent per-site subproblems": NOT the reference's estimator (the reference fits one symmetric
 * J with all site conditionals coupled, plmdca_numerics.cpp:436-607); an opt-in fast mode.
 * Site i keeps its own copy J^(i)_ij of its couplings, which only its own conditional sees:
 *   f = sum_i [ -sum_n w_n log p_n(i, s_ni) + lambda_h |h_i|^2 + (lambda_J / 2) sum_{j != i} |J^(i)_ij|^2 ]
 * (the relaxation of the joint objective: equal to it when J^(i)_ij = (J^(j)_ji)^T), the same
 * conditional as the reference's including its carry (:492, :499).  The L problems share nothing,
 * so site ranges go to different GPUs with no exchange; at the end J_ij = (J^(i)_ij + (J^(j)_ji)^T) / 2.
 * A handle from dca_plm_create_asym covers the sites [site_lo, site_hi); its parameter vector
 * (dca_plm_local_params floats; dca_plm_init / eval / fit / fit_from_init work on it) is
 *   x = [ h : ld floats, column c = (i - site_lo) * q + a | W : L*q rows x ld ],  W[j*q + b][c] = J^(i)_ij(a, b)
 * with ld = the column count padded to a multiple of 128 (pad columns and the blocks j == i stay 0). */
dca_plm *dca_plm_create_asym(const uint8_t *msa, const float *weights, long num_seqs, int seqs_len,
                             int num_site_states, float lambda_h, float lambda_J, int site_lo, int site_hi);
/* copy the handle's columns (of x_local_host, or of the device-resident result when NULL) into host
 * arrays that hold every site: W_full [L*q][ld_full] (column i*q + a), fields_full [L*q] */
int dca_plm_asym_export(dca_plm *p, const float *x_local_host, float *W_full_host, int ld_full,
                        float *fields_full_host);
/* symmetrise: the reference's parameter vector (P floats, plmdca_numerics.cpp:226, :338-342) */
int dca_plm_asym_pack(const float *W_full_host, int ld_full, const float *fields_full_host, int seqs_len,
                      int num_site_states, float *x_out_host);
/* all of it on `gpus` GPUs of this node (0 = DCA_B200_GPUS / all visible): one host thread per GPU, equal
 * site ranges, independent L-BFGS runs (same parameters as plmdcaBackend.cpp:68-75), one gather through
 * host memory, symmetrisation on the caller's GPU.  result: the slowest GPU's iterations /
 * evaluations / device time, the sum of the objectives.  plmdcaBackend() takes this path when
 * DCA_B200_ESTIMATOR=asymmetric or after dca_set_estimator(1). */
int dca_plm_fit_asymmetric(const uint8_t *msa, const float *weights, long num_seqs, int seqs_len,
                           int num_site_states, float lambda_h, float lambda_J, int max_iterations,
                           int gpus, int verbose, float *x_out_host, dca_fit_result *result);
/* process-wide estimator of plmdcaBackend(): 0 joint (the reference's), 1 asymmetric, -1 = environment */
int dca_set_estimator(int mode);

/* statistics of this thread's most recent plmdcaBackend() call (the reference's
 * entry point discards the L-BFGS return code, plmdcaBackend.cpp:197). */
int dca_last_fit(dca_fit_result *result);

/* ---- scoring --------------------------------------------------------------------
 * Frobenius norm of the double-centred non-gap block of every J_ij, then APC
 * (plmdca.py:437-524).  fn_out / fn_apc_out: double[L(L-1)/2] in pair order. */
int dca_plm_fn_apc(const float *x_host, int seqs_len, int num_site_states,
                   double *fn_out, double *fn_apc_out);

/* ---- mean-field DCA operators (msa_numerics.py) ------------------------------------
 * weights: float64[N] (1/counts).  fi_out: double[L][q] (gap column kept,
 * :54-89).  fij_out: double[L(L-1)/2][q-1][q-1] in pair order (:183-229);
 * may be NULL. */
int dca_mf_freqs(const uint8_t *msa, const double *weights, long num_seqs, int seqs_len,
                 int num_site_states, double *fi_out, double *fij_out);
/* fused: frequencies -> pseudocount (:93-125, :232-267) -> covariance matrix
 * C (n x n, n = L(q-1), :271-318) written row-major to corr_out. */
int dca_mf_corr_mat(const uint8_t *msa, const double *weights, long num_seqs, int seqs_len,
                    int num_site_states, double pseudocount, double *reg_fi_out,
                    double *corr_out);
/* couplings = -inv(C), C symmetric positive definite (:322-342). */
int dca_mf_couplings(const double *corr, int n, double *couplings_out);
/* FN + APC of the couplings matrix (meanfield_dca.py:902-988). */
int dca_mf_fn_apc(const double *couplings, int seqs_len, int num_site_states,
                  double *fn_out, double *fn_apc_out);
/* the whole pipeline of MeanFieldDCA.compute_sorted_FN / _DI (meanfield_dca.py:793-988)
 * with every intermediate resident in HBM: frequencies -> pseudocount -> covariance ->
 * -inverse -> FN (+APC) and optionally DI.  reg_fi_out [L][q], couplings_out [n][n],
 * fn_apc_out, di_out and stage_ms_out may be NULL.  stage_ms_out[5] = CUDA-event ms of
 * {pair frequencies (incl. H2D), covariance, inverse, FN, DI}. */
int dca_mf_scores(const uint8_t *msa, const double *weights, long num_seqs, int seqs_len,
                  int num_site_states, double pseudocount, double *reg_fi_out,
                  double *couplings_out, double *fn_out, double *fn_apc_out, double *di_out,
                  float *stage_ms_out);
/* measurement hook: best-of-reps TFLOP/s of a register-only mma.sync.m8n8k4.f64 stream on
 * every SM -- the measured FP64 tensor-pipe peak the inverse is reported against
 * ("the builder micro-benchmarks it before reporting the inverse"). */
int dca_fp64_mma_peak(int reps, double *tflops_out);
/* construct_corr_mat on caller-supplied regularised frequencies (:271-318):
 * reg_fi double[L][q], reg_fij double[L(L-1)/2][q-1][q-1] -> corr_out n x n. */
int dca_mf_corr_from_freqs(const double *reg_fi, const double *reg_fij, int seqs_len,
                           int num_site_states, double *corr_out);

/* ---- reference-sequence matching -----------------------------------------------------------
 * scores_out[n] = Smith-Waterman / Gotoh local alignment score of the reference sequence
 * (ref_codes: states 0..q-2, no gaps) against row n of the alignment with its gaps removed:
 * the search loop of SequenceBackmapper.find_matching_seqs_from_alignment
 * (sequence_backmapper.py:228-283; Biopython pairwise2.align.localds with BLOSUM62, -10/-1 for
 * proteins and +5/-4, -8/0 for RNA, :207-214), one GPU thread per sequence.  Integers: exact. */
int dca_sw_scores(const uint8_t *msa, long num_seqs, int seqs_len, int num_site_states,
                  const uint8_t *ref_codes, int ref_len, int *scores_out);

/* ---- direct information ------------------------------------------------------------
 * two-site-model fields (fixed point to 1e-4) + DI per pair:
 * msa_numerics.py:378-533 (mf: couplings = n x n matrix) and
 * pydca/plmdca/msa_numerics.py:156-311 (plm: parameter vector x, gap stripped).
 * di_out double[L(L-1)/2]; fields_out double[L(L-1)/2][2][q] or NULL. */
int dca_mf_direct_info(const double *couplings, const double *reg_fi, int seqs_len,
                       int num_site_states, double *di_out, double *fields_out);
int dca_plm_direct_info(const float *x_host, const double *reg_fi, int seqs_len,
                        int num_site_states, double *di_out, double *fields_out);

#ifdef __cplusplus
}
#endif
#endif /* DCA_B200_H */
