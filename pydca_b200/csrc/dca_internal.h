// dca_internal.h -- declarations shared by the translation units of
// _plmdcaBackend_b200.so (not part of the public C ABI; see include/dca_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <memory>
#include <new>
#include <utility>
#include <vector>

#include "xrank.cuh"

namespace dca {

// ---- alloc.cu ------------------------------------------------------------------
void *dev_alloc(size_t bytes);          // caching cudaMalloc; nullptr + error on failure
void dev_free(void *p);                 // only after the using stream was synchronised
void dev_cache_release();
void *host_pinned_alloc(size_t bytes);   // cached cudaMallocHost
bool host_pinned_free(void *p);          // false if p is not ours
template <typename T>
static inline bool dev_alloc_t(T **out, size_t count) { *out = static_cast<T *>(dev_alloc(sizeof(T) * count)); return *out != nullptr; }

// scope guard for a cached device block: released on every return path
template <typename T>
struct DevBuf {
    T *p = nullptr;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { dev_free(p); }
    bool alloc(size_t count) { p = static_cast<T *>(dev_alloc(sizeof(T) * count)); return p != nullptr; }
    operator T *() const { return p; }
};

// row-major 2-D tensor map: inner dimension dim0 (contiguous), dim1 rows of stride1_bytes;
// elem_bytes 4 (float32) or 1 (uint8); box = box0 x box1 elements.  0 on success.
int make_tensor_map_2d(void *out_CUtensorMap, int elem_bytes, const void *base, unsigned long long dim0,
                       unsigned long long dim1, unsigned long long stride1_bytes, unsigned box0, unsigned box1);

// index of the first residue code >= q, or -1.  The common case (none) is one branch-free max over the
// bytes, which the host compiler vectorises; a loop with an early exit per byte cost 1.5 ms of the
// 4 ms frequency stage at N = 20000, L = 300.
static inline long first_bad_residue(const uint8_t *msa, long count, int q)
{
    uint8_t mx = 0;
    for (long k = 0; k < count; ++k) mx = msa[k] > mx ? msa[k] : mx;
    if ((int)mx < q) return -1;
    for (long k = 0; k < count; ++k)
        if (msa[k] >= q) return k;
    return -1;
}

// ---- capi.cu ------------------------------------------------------------------
// GPUs one call may drive: dca_set_gpus() > DCA_B200_GPUS > all visible, capped by `limit`
int backend_gpu_count(int limit);

// ---- fasta.cpp -----------------------------------------------------------------
// FASTA -> u8 [N][L] (states 0..q-1, gap q-1), unique rows in first-occurrence order.
// bytes whose resize() does not zero-fill (the reader's threads write every byte themselves: a serial memset
// and the page faults of a 25 MB alignment were most of the de-duplication pass)
template <class T>
struct default_init_allocator : std::allocator<T> {
    template <class U> struct rebind { using other = default_init_allocator<U>; };
    template <class U> void construct(U *p) noexcept { ::new (static_cast<void *>(p)) U; }
    template <class U, class... A> void construct(U *p, A &&...a) { ::new (static_cast<void *>(p)) U(std::forward<A>(a)...); }
};
using ByteVec = std::vector<uint8_t, default_init_allocator<uint8_t>>;
long read_fasta_unique(const char *path, int biomolecule, int L, ByteVec &out);
int encode_residue(int biomolecule, int c);

// ---- reweight.cu -----------------------------------------------------------------
int reweight_min_ident(int L, double seqid, int compare_f64);
// part >= 0: symmetric tiles, part `part` of `nparts` (partial counts of all N rows);
// part < 0: rows [row_lo,row_hi) against all rows
int reweight_counts_device(const uint8_t *d_msa, long N, int L, double seqid, int compare_f64,
                           long row_lo, long row_hi, unsigned *d_counts, cudaStream_t stream,
                           int part = -1, int nparts = 1);

// ---- plm_eval.cu -------------------------------------------------------------------
// similarity order of the rows of a device alignment (u8 [N][L]): order[k] = file index of sorted row k
int similarity_order(const uint8_t *d_msa, long N, int L, cudaStream_t st, std::vector<int> &order, std::vector<int> &rowidx);
struct PlmProblem;   // device-resident state of one plmDCA objective

// scalars every evaluation leaves on the device (double): see plm_eval.cu
enum PlmScalar { PS_FX = 0, PS_DG = 1, PS_GG = 2, PS_XX = 3, PS_COUNT = 4 };
// slot PS_COUNT holds the peer-timeout flag; slot PS_DGINIT the g.d of the direction the optimiser built last
constexpr int PS_DGINIT = PS_COUNT + 1;

PlmProblem *plm_create(const uint8_t *h_msa, const float *h_w, long N, int L, int q,
                       float lambda_h, float lambda_J, int site_lo, int site_hi);
// owner-sharded objective for multi-GPU runs (xrank.cuh): rank r of `world` owns the sites
// [bounds[r], bounds[r+1]); x, g and the optimiser's vectors hold the local slice only
PlmProblem *plm_create_sharded(const uint8_t *h_msa, const float *h_w, long N, int L, int q,
                               float lambda_h, float lambda_J, int rank, int world, const int *bounds);
PlmProblem *plm_create_asym(const uint8_t *h_msa, const float *h_w, long N, int L, int q,
                            float lambda_h, float lambda_J, int site_lo, int site_hi);
bool plm_is_asym(const PlmProblem *p);
int plm_asym_export(PlmProblem *p, const float *d_x, float *h_W_full, int ld_full, float *h_fields_full);
int plm_asym_pack_host(const float *h_W_full, int ld_full, const float *h_fields_full, int L, int q, float *h_xref);
int plm_connect(PlmProblem *p, void *const *arenas, bool shared_process);    // device pointers of every rank's arena
void plm_disconnect(PlmProblem *p);
void *plm_arena(const PlmProblem *p);
size_t plm_arena_bytes(const PlmProblem *p);
bool plm_is_sharded(const PlmProblem *p);
long plm_local_params(const PlmProblem *p);             // == plm_num_params unless sharded
void plm_local_layout(const PlmProblem *p, long *out4);  // {field off, count, coupling off, count} in the full vector
struct XrCall { XrPeers P; XrLayout lay; unsigned epoch, seq; unsigned *err; };
XrCall plm_xr_next(PlmProblem *p);
int plm_xr_barrier(PlmProblem *p);                       // flag barrier on the stream (no-op unless sharded)                       // arguments of the next cross-rank reduction
void plm_destroy(PlmProblem *p);
long plm_num_params(const PlmProblem *p);
cudaStream_t plm_stream(const PlmProblem *p);
void plm_set_carry(PlmProblem *p, int carry, int chunk, int warm);
int plm_set_fidelity(PlmProblem *p, int on);             // 1 = reference arithmetic (sequential float32 sums), single GPU
int plm_fidelity(const PlmProblem *p);
void plm_fx_layout(const PlmProblem *p, const double **fxsite, long *nscan, int *L, int *q, float *lh, float *lJ);
// objective + gradient at device x -> device g ; scalars (fx, g.d, g.g, x.x) left in
// p's scalar buffer; d_dir may be null (then PS_DG = 0).  No host sync.
int plm_eval_device(PlmProblem *p, const float *d_x, float *d_g, const float *d_dir);
// copy the PS_COUNT scalars to the host (one small D2H + stream sync)
int plm_read_scalars(PlmProblem *p, double *out, double *dginit = nullptr);
// initial point (plmdca_numerics.cpp:207-249) written to device x
int plm_init_x_device(PlmProblem *p, float *d_x);
double *plm_scalar_buffer(PlmProblem *p);
constexpr int PLM_NUM_KERNEL_TIMES = 6;   // unpack, logits, softmax, scatter, export (sharded), assemble+finalize
int plm_eval_timed(PlmProblem *p, const float *d_x, float *d_g, float *ms);
void plm_describe(const PlmProblem *p, long *out);
void plm_layout_stats(const PlmProblem *p, double *out4);

// ---- lbfgs.cu --------------------------------------------------------------------------
struct LbfgsResult {
    int status;        // libLBFGS status code (lbfgs.h:78-150 numbering)
    int iterations;
    int evaluations;   // objective/gradient evaluations (1 + sum of line-search trials)
    float fx;
    double device_ms;  // CUDA-event time of the optimisation loop
};
typedef int (*allreduce_fn_t)(void *ctx, float *d_buf, long n, double *d_scalars, int n_scalars);
int lbfgs_minimize(PlmProblem *p, float *d_x, int max_iterations, int verbose,
                   allreduce_fn_t allreduce, void *allreduce_ctx, LbfgsResult *res,
                   int warmup_iterations, double *timed_ms, int *timed_evals);

int lbfgs_vector_bench(long n, int bound, int reps, float *ms_out);
int lbfgs_preload_kernels();   // force-load the optimiser's kernels (see plm_setup_t)

// ---- seqsum.cu -----------------------------------------------------------------------
// exact sequential float32 sums (the reference's vecdot / L2 accumulation), one CTA per job
enum { SEQ_DOT = 0, SEQ_REG = 1 };
constexpr int SEQ_MAX_JOBS = 4;
struct SeqJob {
    const float *x;
    const float *y;          // SEQ_DOT: t_k = fl(x_k * y_k)
    long n;
    int kind;
    long n0;                 // SEQ_REG: t_k = fl(fl(lam * x_k) * x_k), lam = lam0 for k < n0, else lam1
    float lam0, lam1;
    int out;                 // result slot in the float block handed to seqsum_launch
    const double *fxsite;    // SEQ_REG epilogue (may be null): s += (float) sum_k fxsite[k*L + i], i = 0..L-1
    long nscan;
    int L;
};
int seqsum_prepare();        // once per device before the first launch (dynamic shared memory attribute)
struct SeqPipe;              // device buffers of the producer / consumer pipeline (one per optimiser instance)
SeqPipe *seqsum_pipe_create(long nmax);                  // null: the single-CTA kernel is used
void seqsum_pipe_destroy(SeqPipe *p);
int seqsum_pipe_stats(SeqPipe *p, unsigned *out_jobs_x8, cudaStream_t st);
int seqsum_launch(const SeqJob *jobs, int njobs, float *d_out, cudaStream_t st, SeqPipe *pipe = nullptr);
int seqsum_host(const float *x, const float *y, long n, long n0, float lam0, float lam1, float *out,
                unsigned *stats8 = nullptr, float *kernel_ms = nullptr);

// ---- align.cu -----------------------------------------------------------------------------
int sw_scores_host(const uint8_t *msa, long N, int L, int q, const uint8_t *ref, int ref_len, int *scores_out);

// ---- scoring.cu ---------------------------------------------------------------------------
int plm_fn_apc_host(const float *h_x, int L, int q, double *fn, double *fn_apc);

}  // namespace dca
