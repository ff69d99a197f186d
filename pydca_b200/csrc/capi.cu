// capi.cu -- the C ABI declared in include/dca_b200.h.
//
// Drop-in boundary of the reference: plmdcaBackend / freeFieldsAndCouplings
// (pydca/plmdca/plmdcaBackend.cpp:151-223) plus array-based entry points for
// each operator of the hot path.  Nothing throws across this boundary and there
// is no CPU fallback: without a CUDA device every compute call fails.
#include "dca_common.cuh"
#include "dca_internal.h"
#include "../../include/dca_b200.h"

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <ctime>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>
#include <unordered_map>
#include <vector>

#ifndef DCA_DEFAULT_ARITHMETIC
#define DCA_DEFAULT_ARITHMETIC 0      /* 1 = reference, 0 = accurate */
#endif

namespace dca {

std::atomic<unsigned long long> g_launches{0};

static thread_local std::string t_error;
static std::string g_error;
static std::mutex g_error_mu;

void set_error(const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    t_error = buf;
    std::lock_guard<std::mutex> lk(g_error_mu);
    g_error = buf;
}

const char *last_error() { return t_error.empty() ? g_error.c_str() : t_error.c_str(); }

int require_device()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        set_error("no CUDA device available (%s): pydca_b200 has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return -1;
    }
    return 0;
}

int sm_count()
{
    // cached per device ordinal: one process may drive several GPUs (plmdcaBackend with DCA_B200_GPUS)
    static std::atomic<int> cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int v = cached[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cached[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

// mf.cu
int mf_freqs_host(const uint8_t *msa, const double *w, long N, int L, int q, double *fi, double *fij);
int mf_corr_mat_host(const uint8_t *msa, const double *w, long N, int L, int q, double pseudocount,
                     double *reg_fi, double *corr);
int mf_couplings_host(const double *corr, int n, double *out);
int mf_fn_apc_host(const double *h_J, int L, int q, double *fn, double *fn_apc);
int mf_corr_from_freqs_host(const double *reg_fi, const double *reg_fij, int L, int q, double *corr);
int mf_scores_host(const uint8_t *msa, const double *w, long N, int L, int q, double pseudocount,
                   double *reg_fi, double *couplings, double *fn, double *fn_apc, double *di, float *stage_ms);
int fp64_mma_peak(int reps, double *tflops_out);
// scoring.cu
int plm_di_host(const float *h_x, const double *h_reg_fi, int L, int q, double *di, double *fields);
int mf_di_host(const double *h_J, const double *h_reg_fi, int L, int q, double *di, double *fields);

}  // namespace dca

using namespace dca;

struct dca_plm {
    PlmProblem *prob;
    float *d_x, *d_g;                       // whole vector, or the local slice of a sharded objective
};

static thread_local dca_fit_result t_last_fit = {};

// peer arenas opened through CUDA IPC stay mapped for the life of the process, keyed by handle:
// the owners recycle their arenas (plm_eval.cu), and re-opening costs ~1.5 ms per peer
static std::mutex g_ipc_mu;
static std::vector<std::pair<std::string, void *>> g_ipc_open;

extern "C" void dca_plm_destroy(dca_plm *p);

// arithmetic of plmdcaBackend(): DCA_B200_ARITHMETIC = accurate (default: float64 reductions,
// fastest; FN a few 1e-4 from the reference's, like the reference's own with float64 sums) |
// reference (the reference's sequential float32 sums bit for bit: its trajectory, FN ~1e-6)
static std::atomic<int> g_arithmetic{-1};      // dca_set_arithmetic(): overrides the environment
static std::atomic<int> g_gpus{0};             // dca_set_gpus(): overrides DCA_B200_GPUS (0 = environment)
static std::atomic<int> g_estimator{-1};       // dca_set_estimator(): overrides DCA_B200_ESTIMATOR

// how many GPUs plmdcaBackend() drives: dca_set_gpus() > DCA_B200_GPUS > all visible
static int backend_gpus(int L);
int dca::backend_gpu_count(int limit) { return backend_gpus(limit); }

static int backend_gpus(int L)
{
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess) { cudaGetLastError(); visible = 1; }
    int want = g_gpus.load();
    if (want <= 0) { const char *e = getenv("DCA_B200_GPUS"); want = e ? atoi(e) : visible; }
    if (want <= 0) want = visible;
    return std::max(1, std::min(std::min(want, visible), std::min(L, (int)XR_MAX_RANKS)));
}

static int backend_arithmetic()
{
    const int forced = g_arithmetic.load();
    if (forced == 0 || forced == 1) return forced;
    const char *e = getenv("DCA_B200_ARITHMETIC");
    if (e && (strcmp(e, "accurate") == 0 || strcmp(e, "0") == 0)) return 0;
    if (e && (strcmp(e, "reference") == 0 || strcmp(e, "1") == 0)) return 1;
    return DCA_DEFAULT_ARITHMETIC;
}

// estimator of plmdcaBackend(): joint (0, the reference's: one symmetric J, default) or asymmetric (1:
// every site fits its own copy of its couplings, independent problems, symmetrised at the end)
static int backend_estimator()
{
    const int forced = g_estimator.load();
    if (forced == 0 || forced == 1) return forced;
    const char *e = getenv("DCA_B200_ESTIMATOR");
    if (e && (strcmp(e, "asymmetric") == 0 || strcmp(e, "1") == 0)) return 1;
    return 0;
}

static dca_plm *wrap_problem(PlmProblem *prob)
{
    dca_plm *h = new (std::nothrow) dca_plm{prob, nullptr, nullptr};
    if (!h) { plm_destroy(prob); set_error("out of memory"); return nullptr; }
    const size_t bytes = sizeof(float) * (size_t)plm_local_params(prob);
    h->d_x = static_cast<float *>(dev_alloc(bytes));
    h->d_g = static_cast<float *>(dev_alloc(bytes));
    if (!h->d_x || !h->d_g) { dca_plm_destroy(h); return nullptr; }
    return h;
}

extern "C" {

const char *dca_last_error(void) { return last_error(); }

int dca_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int dca_set_device(int device)
{
    DCA_CUDA(cudaSetDevice(device));
    return 0;
}

unsigned long long dca_kernel_launches(void) { return g_launches.load(); }

static std::mutex g_host_vec_mu;
static std::unordered_map<void *, ByteVec *> g_host_vecs;      // buffers handed out by dca_read_fasta

long dca_read_fasta(const char *msa_file, int biomolecule, int seqs_len, uint8_t **msa_out)
{
    if (!msa_file || !msa_out || seqs_len <= 0) { set_error("dca_read_fasta: bad arguments"); return -1; }
    // the reader's own buffer is handed out (no copy of the alignment); dca_free_host finds it again
    ByteVec *buf = nullptr;
    try {
        buf = new ByteVec();
        const long n = read_fasta_unique(msa_file, biomolecule, seqs_len, *buf);
        if (n <= 0) { delete buf; return n; }
        {
            std::lock_guard<std::mutex> lk(g_host_vec_mu);
            g_host_vecs[buf->data()] = buf;
        }
        *msa_out = buf->data();
        return n;
    } catch (const std::exception &e) {                    // no exception crosses the C ABI
        delete buf;
        set_error("dca_read_fasta: %s", e.what());
        return -1;
    }
}

void dca_free_host(void *p)
{
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(g_host_vec_mu);
        auto it = g_host_vecs.find(p);
        if (it != g_host_vecs.end()) { delete it->second; g_host_vecs.erase(it); return; }
    }
    free(p);
}

int dca_weights(const uint8_t *msa, long num_seqs, int seqs_len, double seqid, int compare_f64,
                long row_lo, long row_hi, uint32_t *counts_out)
{
    if (require_device() != 0) return -1;
    if (!msa || !counts_out || num_seqs <= 0 || seqs_len <= 0) { set_error("dca_weights: empty alignment"); return -1; }
    if (row_lo < 0 || row_hi > num_seqs || row_lo >= row_hi) { set_error("dca_weights: bad row range"); return -1; }
    if (!(seqid > 0.0) || seqid > 1.0) { set_error("dca_weights: seqid %g not in (0,1]", seqid); return -1; }
    DevBuf<uint8_t> d_msa;               // released on every return path
    DevBuf<unsigned> d_counts;
    if (!d_msa.alloc((size_t)num_seqs * seqs_len) || !d_counts.alloc((size_t)num_seqs)) return -1;
    DCA_CUDA(cudaMemcpy(d_msa, msa, (size_t)num_seqs * seqs_len, cudaMemcpyHostToDevice));
    // the whole alignment: symmetric tiles (half the compares); a row shard: those rows against all
    const bool whole = (row_lo == 0 && row_hi == num_seqs);
    if (reweight_counts_device(d_msa, num_seqs, seqs_len, seqid, compare_f64, row_lo, row_hi, d_counts, 0,
                               whole ? 0 : -1, 1) != 0) return -1;
    DCA_CUDA(cudaMemcpy(counts_out + row_lo, d_counts.p + row_lo, sizeof(unsigned) * (size_t)(row_hi - row_lo),
                        cudaMemcpyDeviceToHost));
    return 0;
}

int dca_weights_part(const uint8_t *msa, long num_seqs, int seqs_len, double seqid, int compare_f64,
                     int part, int nparts, uint32_t *partial_counts_out, float *device_ms_out)
{
    if (require_device() != 0) return -1;
    if (!msa || !partial_counts_out || num_seqs <= 0 || seqs_len <= 0) { set_error("dca_weights_part: empty alignment"); return -1; }
    if (nparts < 1 || part < 0 || part >= nparts) { set_error("dca_weights_part: bad part %d of %d", part, nparts); return -1; }
    if (!(seqid > 0.0) || seqid > 1.0) { set_error("dca_weights_part: seqid %g not in (0,1]", seqid); return -1; }
    DevBuf<uint8_t> d_msa;
    DevBuf<unsigned> d_counts;
    if (!d_msa.alloc((size_t)num_seqs * seqs_len) || !d_counts.alloc((size_t)num_seqs)) return -1;
    DCA_CUDA(cudaMemcpy(d_msa, msa, (size_t)num_seqs * seqs_len, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    DCA_CUDA(cudaEventCreate(&e0)); DCA_CUDA(cudaEventCreate(&e1));
    cudaEventRecord(e0, 0);
    const int rc = reweight_counts_device(d_msa, num_seqs, seqs_len, seqid, compare_f64, 0, num_seqs, d_counts, 0, part, nparts);
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (rc != 0) return -1;
    if (device_ms_out) *device_ms_out = ms;
    DCA_CUDA(cudaMemcpy(partial_counts_out, d_counts.p, sizeof(unsigned) * (size_t)num_seqs, cudaMemcpyDeviceToHost));
    return 0;
}

dca_plm *dca_plm_create(const uint8_t *msa, const float *weights, long num_seqs, int seqs_len,
                        int num_site_states, float lambda_h, float lambda_J, int site_lo, int site_hi)
{
    if (!msa || !weights) { set_error("dca_plm_create: null input"); return nullptr; }
    PlmProblem *prob = plm_create(msa, weights, num_seqs, seqs_len, num_site_states, lambda_h, lambda_J,
                                  site_lo, site_hi);
    if (!prob) return nullptr;
    return wrap_problem(prob);
}

dca_plm *dca_plm_create_sharded(const uint8_t *msa, const float *weights, long num_seqs, int seqs_len,
                                int num_site_states, float lambda_h, float lambda_J, int rank, int world,
                                const int *bounds)
{
    if (!msa || !weights || !bounds) { set_error("dca_plm_create_sharded: null input"); return nullptr; }
    PlmProblem *prob = plm_create_sharded(msa, weights, num_seqs, seqs_len, num_site_states, lambda_h, lambda_J,
                                          rank, world, bounds);
    if (!prob) return nullptr;
    return wrap_problem(prob);
}

dca_plm *dca_plm_create_asym(const uint8_t *msa, const float *weights, long num_seqs, int seqs_len,
                             int num_site_states, float lambda_h, float lambda_J, int site_lo, int site_hi)
{
    if (!msa || !weights) { set_error("dca_plm_create_asym: null input"); return nullptr; }
    PlmProblem *prob = plm_create_asym(msa, weights, num_seqs, seqs_len, num_site_states, lambda_h, lambda_J,
                                       site_lo, site_hi);
    if (!prob) return nullptr;
    return wrap_problem(prob);
}

int dca_plm_asym_export(dca_plm *p, const float *x_local_host, float *W_full_host, int ld_full, float *fields_full_host)
{
    if (!p || !W_full_host || !fields_full_host) { set_error("dca_plm_asym_export: null argument"); return -1; }
    if (x_local_host)
        DCA_CUDA(cudaMemcpyAsync(p->d_x, x_local_host, sizeof(float) * (size_t)plm_local_params(p->prob),
                                 cudaMemcpyHostToDevice, plm_stream(p->prob)));
    return plm_asym_export(p->prob, p->d_x, W_full_host, ld_full, fields_full_host);
}

int dca_plm_asym_pack(const float *W_full_host, int ld_full, const float *fields_full_host, int seqs_len,
                      int num_site_states, float *x_out_host)
{
    if (!W_full_host || !fields_full_host || !x_out_host) { set_error("dca_plm_asym_pack: null argument"); return -1; }
    if (ld_full < seqs_len * num_site_states) { set_error("dca_plm_asym_pack: ld_full %d < L*q", ld_full); return -1; }
    return plm_asym_pack_host(W_full_host, ld_full, fields_full_host, seqs_len, num_site_states, x_out_host);
}

long dca_plm_local_params(const dca_plm *p) { return p ? plm_local_params(p->prob) : -1; }

int dca_plm_local_layout(const dca_plm *p, long *out4)
{
    if (!p || !out4) { set_error("dca_plm_local_layout: null argument"); return -1; }
    plm_local_layout(p->prob, out4);
    return 0;
}

int dca_shard_layout(int seqs_len, int num_site_states, int site_lo, int site_hi, long *out4)
{
    const long L = seqs_len, q = num_site_states;
    if (!out4 || L < 2 || q < 2 || site_lo < 0 || site_hi > L || site_lo >= site_hi) {
        set_error("dca_shard_layout: bad arguments"); return -1;
    }
    out4[0] = site_lo * q;
    out4[1] = (site_hi - site_lo) * q;
    out4[2] = L * q + pair_index(L, site_lo, site_lo + 1) * q * q;
    out4[3] = (pair_index(L, site_hi, site_hi + 1) - pair_index(L, site_lo, site_lo + 1)) * q * q;
    return 0;
}

void *dca_plm_arena(const dca_plm *p) { return p ? plm_arena(p->prob) : nullptr; }
void *dca_plm_x_device(dca_plm *p) { return p ? (void *)p->d_x : nullptr; }

int dca_plm_ipc_handle(const dca_plm *p, unsigned char *out64)
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    if (!p || !out64 || !plm_arena(p->prob)) { set_error("dca_plm_ipc_handle: not a sharded objective"); return -1; }
    cudaIpcMemHandle_t h;
    DCA_CUDA(cudaIpcGetMemHandle(&h, plm_arena(p->prob)));
    memcpy(out64, &h, 64);
    return 0;
}

int dca_plm_connect_ptrs(dca_plm *p, void *const *arena_ptrs)
{
    if (!p || !arena_ptrs) { set_error("dca_plm_connect_ptrs: null argument"); return -1; }
    return plm_connect(p->prob, arena_ptrs, true);
}

int dca_plm_connect_ipc(dca_plm *p, const unsigned char *handles, int world)
{
    if (!p || !handles || world < 1 || world > XR_MAX_RANKS) { set_error("dca_plm_connect_ipc: bad arguments"); return -1; }
    void *ptrs[XR_MAX_RANKS] = {nullptr};
    void *own = plm_arena(p->prob);
    cudaIpcMemHandle_t mine;
    DCA_CUDA(cudaIpcGetMemHandle(&mine, own));
    for (int r = 0; r < world; ++r) {
        if (memcmp(handles + 64 * r, &mine, 64) == 0) { ptrs[r] = own; continue; }
        const std::string key(reinterpret_cast<const char *>(handles + 64 * r), 64);
        {
            std::lock_guard<std::mutex> lk(g_ipc_mu);
            for (auto &kv : g_ipc_open) if (kv.first == key) ptrs[r] = kv.second;
        }
        if (ptrs[r]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + 64 * r, 64);
        cudaError_t e = cudaIpcOpenMemHandle(&ptrs[r], h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            set_error("cudaIpcOpenMemHandle(rank %d) -> %s", r, cudaGetErrorString(e));
            cudaGetLastError();
            return -1;
        }
        std::lock_guard<std::mutex> lk(g_ipc_mu);
        g_ipc_open.emplace_back(key, ptrs[r]);
    }
    return plm_connect(p->prob, ptrs, false);
}

int dca_plm_disconnect(dca_plm *p)
{
    if (!p) { set_error("null handle"); return -1; }
    plm_disconnect(p->prob);          // forget the peers; their mappings stay cached (see g_ipc_open)
    return 0;
}

void dca_plm_destroy(dca_plm *p)
{
    if (!p) return;
    cudaStreamSynchronize(plm_stream(p->prob));
    dca_plm_disconnect(p);
    dev_free(p->d_x); dev_free(p->d_g);
    plm_destroy(p->prob);
    delete p;
}

long dca_plm_num_params(const dca_plm *p) { return p ? plm_num_params(p->prob) : -1; }

int dca_plm_set_carry(dca_plm *p, int carry, int chunk, int warm)
{
    if (!p) { set_error("null handle"); return -1; }
    plm_set_carry(p->prob, carry, chunk, warm);
    return 0;
}

int dca_set_arithmetic(int mode)
{
    if (mode < -1 || mode > 1) { set_error("dca_set_arithmetic: mode must be -1 (environment), 0 (accurate) or 1 (reference)"); return -1; }
    g_arithmetic.store(mode);
    return 0;
}

int dca_set_estimator(int mode)
{
    if (mode < -1 || mode > 1) { set_error("dca_set_estimator: mode %d (0 joint, 1 asymmetric, -1 environment)", mode); return -1; }
    g_estimator.store(mode);
    return 0;
}

int dca_set_gpus(int n)
{
    if (n < 0) { set_error("dca_set_gpus: n must be >= 0 (0 = DCA_B200_GPUS / all visible)"); return -1; }
    g_gpus.store(n);
    return 0;
}

int dca_plm_set_fidelity(dca_plm *p, int mode)
{
    if (!p) { set_error("null handle"); return -1; }
    return plm_set_fidelity(p->prob, mode);
}

int dca_seq_sum(const float *x, const float *y, long n, long n0, float lambda0, float lambda1, float *out)
{
    return seqsum_host(x, y, n, n0, lambda0, lambda1, out);
}

int dca_seq_sum_stats(const float *x, const float *y, long n, long n0, float lambda0, float lambda1, float *out,
                      unsigned *stats8, float *kernel_ms)
{
    return seqsum_host(x, y, n, n0, lambda0, lambda1, out, stats8, kernel_ms);
}

int dca_plm_init(dca_plm *p, float *x_out_host)
{
    if (!p || !x_out_host) { set_error("dca_plm_init: null argument"); return -1; }
    if (plm_init_x_device(p->prob, p->d_x) != 0) return -1;
    DCA_CUDA(cudaMemcpyAsync(x_out_host, p->d_x, sizeof(float) * (size_t)plm_local_params(p->prob),
                             cudaMemcpyDeviceToHost, plm_stream(p->prob)));
    DCA_CUDA(cudaStreamSynchronize(plm_stream(p->prob)));
    return 0;
}

int dca_plm_eval(dca_plm *p, const float *x_host, float *g_host, double *fx_out)
{
    if (!p || !x_host || !g_host) { set_error("dca_plm_eval: null argument"); return -1; }
    cudaStream_t st = plm_stream(p->prob);
    const size_t bytes = sizeof(float) * (size_t)plm_local_params(p->prob);
    DCA_CUDA(cudaMemcpyAsync(p->d_x, x_host, bytes, cudaMemcpyHostToDevice, st));
    if (plm_eval_device(p->prob, p->d_x, p->d_g, nullptr) != 0) return -1;
    // scalars first (pinned buffer + stream sync), then the gradient on an idle stream: a copy
    // to pageable memory waits inside the driver, and when several sharded ranks live in one
    // process that wait must not sit in front of a peer's kernel launches
    double h[PS_COUNT];
    if (plm_read_scalars(p->prob, h) != 0) return -1;
    DCA_CUDA(cudaMemcpyAsync(g_host, p->d_g, bytes, cudaMemcpyDeviceToHost, st));
    DCA_CUDA(cudaStreamSynchronize(st));
    if (fx_out) *fx_out = h[PS_FX];
    return 0;
}

int dca_plm_eval_device(dca_plm *p, const float *x_dev, float *g_dev, const float *dir_dev,
                        double *scalars_out_host)
{
    if (!p || !x_dev || !g_dev) { set_error("dca_plm_eval_device: null argument"); return -1; }
    if (plm_eval_device(p->prob, x_dev, g_dev, dir_dev) != 0) return -1;
    if (scalars_out_host) return plm_read_scalars(p->prob, scalars_out_host);
    return 0;
}

int dca_plm_eval_timed(dca_plm *p, const float *x_host, float *ms_out)
{
    if (!p || !x_host || !ms_out) { set_error("dca_plm_eval_timed: null argument"); return -1; }
    cudaStream_t st = plm_stream(p->prob);
    DCA_CUDA(cudaMemcpyAsync(p->d_x, x_host, sizeof(float) * (size_t)plm_local_params(p->prob),
                             cudaMemcpyHostToDevice, st));
    return plm_eval_timed(p->prob, p->d_x, p->d_g, ms_out);
}

int dca_lbfgs_vector_bench(long n, int stored_pairs, int reps, float *ms_out3)
{
    if (!ms_out3) { set_error("dca_lbfgs_vector_bench: null argument"); return -1; }
    return lbfgs_vector_bench(n, stored_pairs, reps, ms_out3);
}

void *dca_plm_cuda_stream(const dca_plm *p) { return p ? (void *)plm_stream(p->prob) : nullptr; }

int dca_plm_describe(const dca_plm *p, long *out10)
{
    if (!p || !out10) { set_error("dca_plm_describe: null argument"); return -1; }
    plm_describe(p->prob, out10);
    return 0;
}

int dca_plm_layout_stats(const dca_plm *p, double *out4)
{
    if (!p || !out4) { set_error("dca_plm_layout_stats: null argument"); return -1; }
    plm_layout_stats(p->prob, out4);
    return 0;
}

int dca_plm_fit(dca_plm *p, float *x_inout_host, int max_iterations, int verbose, int warmup_iterations,
                dca_allreduce_fn allreduce, void *allreduce_ctx, dca_fit_result *result)
{
    if (!p || !x_inout_host) { set_error("dca_plm_fit: null argument"); return -1; }
    cudaStream_t st = plm_stream(p->prob);
    const size_t bytes = sizeof(float) * (size_t)plm_local_params(p->prob);
    DCA_CUDA(cudaMemcpyAsync(p->d_x, x_inout_host, bytes, cudaMemcpyHostToDevice, st));
    LbfgsResult r{};
    double timed_ms = 0.0;
    int timed_evals = 0;
    if (lbfgs_minimize(p->prob, p->d_x, max_iterations, verbose, (dca::allreduce_fn_t)allreduce,
                       allreduce_ctx, &r, warmup_iterations, &timed_ms, &timed_evals) != 0)
        return -1;
    DCA_CUDA(cudaMemcpyAsync(x_inout_host, p->d_x, bytes, cudaMemcpyDeviceToHost, st));
    DCA_CUDA(cudaStreamSynchronize(st));
    if (result) {
        result->status = r.status; result->iterations = r.iterations; result->evaluations = r.evaluations;
        result->fx = r.fx; result->device_ms = r.device_ms; result->timed_ms = timed_ms;
        result->timed_evaluations = timed_evals;
    }
    return 0;
}

int dca_plm_fit_from_init(dca_plm *p, int max_iterations, int verbose, dca_fit_result *result)
{
    if (!p) { set_error("dca_plm_fit_from_init: null handle"); return -1; }
    if (plm_init_x_device(p->prob, p->d_x) != 0) return -1;
    LbfgsResult r{};
    double timed_ms = 0.0;
    int timed_evals = 0;
    if (lbfgs_minimize(p->prob, p->d_x, max_iterations, verbose, nullptr, nullptr, &r, 0, &timed_ms, &timed_evals) != 0)
        return -1;
    if (result) {
        result->status = r.status; result->iterations = r.iterations; result->evaluations = r.evaluations;
        result->fx = r.fx; result->device_ms = r.device_ms; result->timed_ms = timed_ms;
        result->timed_evaluations = timed_evals;
    }
    return 0;
}

int dca_plm_fn_apc(const float *x_host, int seqs_len, int num_site_states, double *fn_out, double *fn_apc_out)
{
    if (!x_host || !fn_out) { set_error("dca_plm_fn_apc: null argument"); return -1; }
    return plm_fn_apc_host(x_host, seqs_len, num_site_states, fn_out, fn_apc_out);
}

// ---- several GPUs behind the reference's entry point -------------------------------------
// plmdcaBackend() is what stock pydca calls (plmdcaBackend.cpp:151); it is one synchronous call
// in one process.  With G > 1 GPUs it starts one host thread per GPU: each creates its
// owner-sharded objective (sites [bounds[r], bounds[r+1])), enables peer access, hands its
// arena pointer to the others (same address space: no IPC), runs the same collective L-BFGS
// as the one-process-per-GPU path, and copies ITS slice of the result straight into the
// caller's host vector over its own PCIe link -- no gather collective at all.
namespace {

class ThreadBarrier {
    std::mutex mu;
    std::condition_variable cv;
    int n, waiting = 0;
    unsigned long gen = 0;
public:
    explicit ThreadBarrier(int count) : n(count) {}
    void wait()
    {
        std::unique_lock<std::mutex> lk(mu);
        const unsigned long g = gen;
        if (++waiting == n) { waiting = 0; ++gen; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
};

struct MgShared {
    int G = 1;
    const uint8_t *msa = nullptr;
    long N = 0;
    int L = 0, q = 0;
    float lh = 0.f, lJ = 0.f;
    double seqid = 0.8;
    int max_it = 0, verbose = 0;
    bool deep_reweight = false;
    std::vector<int> bounds;
    std::vector<std::vector<uint32_t>> partial;     // symmetric reweighting parts (deep alignments)
    std::vector<float> w;
    void *arenas[XR_MAX_RANKS] = {nullptr};
    float *x_host = nullptr;
    std::atomic<int> err{0};
    ThreadBarrier *bar = nullptr;
    LbfgsResult res0{};
    double timed_ms = 0.0;
    int timed_evals = 0;
    double stage_ms[4] = {0, 0, 0, 0};              // rank 0: reweight, create+wire, fit, d2h
};

static double wall_ms()
{
    timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec * 1e3 + t.tv_nsec * 1e-6;
}

static void mg_worker(MgShared *S, int r)
{
    PlmProblem *prob = nullptr;
    float *d_x = nullptr;
    auto fail = [&](const char *what) {
        S->err.store(1);
        fprintf(stderr, "plmdcaBackend: GPU %d: %s: %s\n", r, what, last_error());
    };
    double t0 = wall_ms();
    if (cudaSetDevice(r) != cudaSuccess) { set_error("cudaSetDevice(%d) failed", r); fail("device"); }
    for (int p = 0; p < S->G && !S->err.load(); ++p) {
        if (p == r) continue;
        int can = 0;
        cudaDeviceCanAccessPeer(&can, r, p);
        if (!can) { set_error("GPU %d cannot access GPU %d's memory (no P2P)", r, p); fail("peer access"); break; }
        const cudaError_t e = cudaDeviceEnablePeerAccess(p, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
            set_error("cudaDeviceEnablePeerAccess(%d -> %d): %s", r, p, cudaGetErrorString(e)); fail("peer access"); break;
        }
        cudaGetLastError();
    }
    // ---- reweighting: deep alignments split the symmetric tile pairs over the GPUs -------------
    if (S->deep_reweight && !S->err.load()) {
        DevBuf<uint8_t> d_msa;
        DevBuf<unsigned> d_counts;
        S->partial[r].assign((size_t)S->N, 0u);
        if (!d_msa.alloc((size_t)S->N * S->L) || !d_counts.alloc((size_t)S->N) ||
            cudaMemcpy(d_msa, S->msa, (size_t)S->N * S->L, cudaMemcpyHostToDevice) != cudaSuccess ||
            reweight_counts_device(d_msa, S->N, S->L, S->seqid, 0, 0, S->N, d_counts, 0, r, S->G) != 0 ||
            cudaMemcpy(S->partial[r].data(), d_counts.p, sizeof(unsigned) * (size_t)S->N, cudaMemcpyDeviceToHost) != cudaSuccess)
            fail("reweighting");
    }
    S->bar->wait();
    if (S->deep_reweight && r == 0 && !S->err.load()) {
        for (long n = 0; n < S->N; ++n) {
            unsigned c = 0;
            for (int p = 0; p < S->G; ++p) c += S->partial[p][(size_t)n];
            S->w[(size_t)n] = 1.f / (float)c;                                    // plmdca_numerics.cpp:669
        }
    }
    S->bar->wait();
    if (r == 0) { S->stage_ms[0] = wall_ms() - t0; t0 = wall_ms(); }
    // ---- owner-sharded objective + peer wiring -------------------------------------------------------
    if (!S->err.load()) {
        prob = plm_create_sharded(S->msa, S->w.data(), S->N, S->L, S->q, S->lh, S->lJ, r, S->G, S->bounds.data());
        if (!prob) fail("create");
        else S->arenas[r] = plm_arena(prob);
    }
    S->bar->wait();
    if (!S->err.load() && plm_connect(prob, S->arenas, true) != 0) fail("connect");
    if (!S->err.load()) {
        d_x = static_cast<float *>(dev_alloc(sizeof(float) * (size_t)plm_local_params(prob)));
        if (!d_x) fail("allocation");
    }
    S->bar->wait();                       // every arena is known everywhere before the first push
    if (r == 0) { S->stage_ms[1] = wall_ms() - t0; t0 = wall_ms(); }
    // ---- collective fit; on an error every rank still reaches the barriers below (a dead peer
    // shows up as the 4 s flag time-out inside the kernels, not as a hang)
    if (!S->err.load()) {
        LbfgsResult res{};
        double tms = 0.0;
        int tev = 0;
        if (plm_init_x_device(prob, d_x) != 0 ||
            lbfgs_minimize(prob, d_x, S->max_it, (r == 0) ? S->verbose : 0, nullptr, nullptr, &res, 0, &tms, &tev) != 0)
            fail("fit");
        else if (r == 0) { S->res0 = res; S->timed_ms = tms; S->timed_evals = tev; }
    }
    if (r == 0) { S->stage_ms[2] = wall_ms() - t0; t0 = wall_ms(); }
    if (!S->err.load()) {
        long lay[4];
        plm_local_layout(prob, lay);
        cudaStream_t st = plm_stream(prob);
        if (cudaMemcpyAsync(S->x_host + lay[0], d_x, sizeof(float) * (size_t)lay[1], cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaMemcpyAsync(S->x_host + lay[2], d_x + lay[1], sizeof(float) * (size_t)lay[3], cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) {
            set_error("D2H of the slice of GPU %d failed: %s", r, cudaGetErrorString(cudaGetLastError()));
            fail("d2h");
        }
    }
    S->bar->wait();                       // nobody releases an arena a peer may still push into
    if (r == 0) S->stage_ms[3] = wall_ms() - t0;
    if (prob) { cudaStreamSynchronize(plm_stream(prob)); plm_disconnect(prob); }
    dev_free(d_x);
    if (prob) plm_destroy(prob);
}

// contiguous, balanced site ranges (distributed.py: shard_range)
static std::vector<int> even_bounds(int L, int G)
{
    std::vector<int> b((size_t)G + 1, 0);
    const int base = L / G, extra = L % G;
    for (int r = 0; r < G; ++r) b[(size_t)r + 1] = b[(size_t)r] + base + (r < extra ? 1 : 0);
    return b;
}

// returns 0 and fills x_host / res on success; -1 when the multi-GPU path could not run
static int backend_multi_gpu(int G, const ByteVec &msa, std::vector<float> &w, bool have_w, long N, int L,
                             int q, double seqid, float lh, float lJ, int max_it, int verbose, float *x_host,
                             dca_fit_result *res, bool trace)
{
    int saved = 0;
    cudaGetDevice(&saved);
    MgShared S;
    S.G = G; S.msa = msa.data(); S.N = N; S.L = L; S.q = q; S.lh = lh; S.lJ = lJ; S.seqid = seqid;
    S.max_it = max_it; S.verbose = verbose;
    S.deep_reweight = !have_w;
    S.bounds = even_bounds(L, G);
    S.partial.resize((size_t)G);
    S.w.swap(w);
    if (!have_w) S.w.assign((size_t)N, 0.f);
    S.x_host = x_host;
    ThreadBarrier bar(G);
    S.bar = &bar;
    std::vector<std::thread> th;
    for (int r = 1; r < G; ++r) th.emplace_back(mg_worker, &S, r);
    mg_worker(&S, 0);
    for (auto &t : th) t.join();
    cudaSetDevice(saved);
    w.swap(S.w);
    if (trace)
        fprintf(stderr, "[dca trace] %d GPUs: reweight %.2f ms, create+wire %.2f ms, fit %.2f ms, d2h %.2f ms\n",
                G, S.stage_ms[0], S.stage_ms[1], S.stage_ms[2], S.stage_ms[3]);
    if (S.err.load()) return -1;
    res->status = S.res0.status; res->iterations = S.res0.iterations; res->evaluations = S.res0.evaluations;
    res->fx = S.res0.fx; res->device_ms = S.res0.device_ms; res->timed_ms = S.timed_ms;
    res->timed_evaluations = S.timed_evals;
    return 0;
}

// ---- asymmetric estimator over G GPUs: independent fits of contiguous site ranges (every site costs the
// same: L - 1 blocks), no exchange at all during the fits; at the end every GPU copies its columns of
// the directed couplings into one host matrix -- the single gather of north_star -- and the caller's
// GPU symmetrises it into the reference's vector.
struct AsymShared {
    int G = 1, dev0 = 0;
    const uint8_t *msa = nullptr;
    const float *w = nullptr;
    long N = 0;
    int L = 0, q = 0, max_it = 0, verbose = 0, ld_full = 0;
    float lh = 0.f, lJ = 0.f;
    std::vector<int> bounds;
    float *W_full = nullptr, *h_full = nullptr;
    std::vector<LbfgsResult> res;
    std::atomic<int> err{0};
};

static void asym_worker(AsymShared *S, int r)
{
    auto fail = [&](const char *what) {
        S->err.store(1);
        fprintf(stderr, "plmdca asymmetric: GPU %d: %s: %s\n", r, what, last_error());
    };
    if (S->G > 1 && cudaSetDevice(r) != cudaSuccess) { set_error("cudaSetDevice(%d) failed", r); fail("device"); return; }
    PlmProblem *prob = plm_create_asym(S->msa, S->w, S->N, S->L, S->q, S->lh, S->lJ, S->bounds[(size_t)r], S->bounds[(size_t)r + 1]);
    if (!prob) { fail("create"); return; }
    float *d_x = static_cast<float *>(dev_alloc(sizeof(float) * (size_t)plm_local_params(prob)));
    if (!d_x) fail("allocation");
    else if (plm_init_x_device(prob, d_x) != 0 ||
             lbfgs_minimize(prob, d_x, S->max_it, (r == 0) ? S->verbose : 0, nullptr, nullptr, &S->res[(size_t)r], 0, nullptr, nullptr) != 0)
        fail("fit");
    else if (plm_asym_export(prob, d_x, S->W_full, S->ld_full, S->h_full) != 0)
        fail("export");
    cudaStreamSynchronize(plm_stream(prob));
    dev_free(d_x);
    plm_destroy(prob);
}

static int fit_asymmetric(const uint8_t *msa, const float *w, long N, int L, int q, float lh, float lJ, int max_it,
                          int gpus, int verbose, float *x_out, dca_fit_result *result)
{
    if (require_device() != 0) return -1;
    int visible = 1;
    if (cudaGetDeviceCount(&visible) != cudaSuccess) { cudaGetLastError(); visible = 1; }
    AsymShared S;
    S.G = std::max(1, std::min(std::min(gpus <= 0 ? backend_gpus(L) : gpus, visible), L));
    cudaGetDevice(&S.dev0);
    S.msa = msa; S.w = w; S.N = N; S.L = L; S.q = q; S.lh = lh; S.lJ = lJ; S.max_it = max_it; S.verbose = verbose;
    S.bounds = even_bounds(L, S.G);
    S.res.assign((size_t)S.G, LbfgsResult{});
    const long Lq = (long)L * q;
    S.ld_full = (int)((Lq + 3) & ~3L);
    S.W_full = static_cast<float *>(host_pinned_alloc(sizeof(float) * (size_t)Lq * S.ld_full));
    S.h_full = static_cast<float *>(host_pinned_alloc(sizeof(float) * (size_t)Lq));
    int rc = (S.W_full && S.h_full) ? 0 : -1;
    if (rc == 0) {
        std::vector<std::thread> th;
        for (int r = 1; r < S.G; ++r) th.emplace_back(asym_worker, &S, r);
        asym_worker(&S, 0);
        for (auto &t : th) t.join();
        if (S.G > 1) cudaSetDevice(S.dev0);
        if (S.err.load()) rc = -1;
    }
    if (rc == 0) rc = plm_asym_pack_host(S.W_full, S.ld_full, S.h_full, L, q, x_out);
    if (S.W_full) host_pinned_free(S.W_full);
    if (S.h_full) host_pinned_free(S.h_full);
    if (rc == 0 && result) {
        // the fits are independent: report the slowest one's counts, the sum of the objectives, and
        // the first status that is neither success nor the float32 line-search floor (-1001)
        dca_fit_result out{};
        out.status = 0;
        for (const LbfgsResult &r : S.res) {
            out.iterations = std::max(out.iterations, r.iterations);
            out.evaluations = std::max(out.evaluations, r.evaluations);
            out.device_ms = std::max(out.device_ms, (double)r.device_ms);
            out.fx += r.fx;
            if (out.status == 0 || (out.status == -1001 && r.status != 0)) out.status = r.status;
        }
        out.timed_ms = out.device_ms; out.timed_evaluations = out.evaluations;
        *result = out;
    }
    return rc;
}

}  // namespace

// ---- the reference's own entry point -------------------------------------------
float *plmdcaBackend(unsigned short biomolecule, unsigned short num_site_states, const char *msa_file,
                     unsigned int seqs_len, float seqid, float lambda_h, float lambda_J,
                     unsigned int max_iteration, unsigned int num_threads, bool verbose)
{
    (void)num_threads;   // kept for signature compatibility: the work runs on the GPU
    const bool trace = getenv("DCA_B200_TRACE") != nullptr;
    auto now = []() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e3 + t.tv_nsec * 1e-6; };
    double t_prev = now();
    auto mark = [&](const char *what) {
        if (!trace) return;
        const double t = now();
        fprintf(stderr, "[dca trace] %-10s %8.2f ms\n", what, t - t_prev);
        t_prev = t;
    };
    if (require_device() != 0) { fprintf(stderr, "plmdcaBackend: %s\n", last_error()); return nullptr; }
    const int q = (int)num_site_states, L = (int)seqs_len;
    if ((biomolecule == 1 && q != 21) || (biomolecule != 1 && q != 5)) {
        set_error("plmdcaBackend: num_site_states %d does not match biomolecule %d", q, (int)biomolecule);
        fprintf(stderr, "plmdcaBackend: %s\n", last_error());
        return nullptr;
    }
    ByteVec msa;
    const long N = read_fasta_unique(msa_file, biomolecule == 1 ? 1 : 2, L, msa);       // plmdca_numerics.cpp:43
    if (N <= 0) { fprintf(stderr, "plmdcaBackend: %s\n", last_error()); return nullptr; }
    mark("read_fasta");
    // GPUs: DCA_B200_GPUS / dca_set_gpus(), default all visible; the reference-arithmetic mode is
    // single-GPU (its sequential sums span the whole vector)
    int G = backend_gpus(L);
    if (G > 1 && backend_arithmetic() == 1) {
        if (verbose) fprintf(stderr, "plmdcaBackend: arithmetic=reference runs on one GPU\n");
        G = 1;
    }
    const bool deep = (double)N * (double)N * (double)L >= 1.0e11;    // reweighting worth splitting over the GPUs
    std::vector<float> w((size_t)N);
    if (!(G > 1 && deep)) {
        std::vector<uint32_t> counts((size_t)N);
        if (dca_weights(msa.data(), N, L, (double)seqid, 0, 0, N, counts.data()) != 0) {    // :45
            fprintf(stderr, "plmdcaBackend: %s\n", last_error());
            return nullptr;
        }
        for (long n = 0; n < N; ++n) w[n] = 1.f / (float)counts[n];                          // :669
        mark("reweight");
    }
    if (backend_estimator() == 1) {        // independent per-site fits (not the reference's estimator): opt-in
        if (G > 1 && deep) {
            std::vector<uint32_t> counts((size_t)N);
            if (dca_weights(msa.data(), N, L, (double)seqid, 0, 0, N, counts.data()) != 0) {
                fprintf(stderr, "plmdcaBackend: %s\n", last_error());
                return nullptr;
            }
            for (long n = 0; n < N; ++n) w[n] = 1.f / (float)counts[n];
        }
        const size_t Pa = (size_t)L * q + (size_t)L * (L - 1) / 2 * q * q;
        float *xa = (float *)host_pinned_alloc(sizeof(float) * Pa);
        if (!xa) { fprintf(stderr, "ERROR: Failed to allocate a memory block for variables.\n"); return nullptr; }
        dca_fit_result resa{};
        if (fit_asymmetric(msa.data(), w.data(), N, L, q, lambda_h, lambda_J, (int)max_iteration, G, verbose ? 1 : 0, xa, &resa) != 0) {
            fprintf(stderr, "plmdcaBackend: %s\n", last_error());
            host_pinned_free(xa);
            return nullptr;
        }
        mark("asymmetric");
        t_last_fit = resa;
        return xa;
    }
    if (G > 1) {
        const size_t Pm = (size_t)L * q + (size_t)L * (L - 1) / 2 * q * q;
        float *xm = (float *)host_pinned_alloc(sizeof(float) * Pm);
        if (!xm) { fprintf(stderr, "ERROR: Failed to allocate a memory block for variables.\n"); return nullptr; }
        dca_fit_result resm{};
        if (backend_multi_gpu(G, msa, w, !deep, N, L, q, (double)seqid, lambda_h, lambda_J, (int)max_iteration,
                              verbose ? 1 : 0, xm, &resm, trace) == 0) {
            mark("multi-gpu");
            t_last_fit = resm;
            return xm;
        }
        host_pinned_free(xm);
        fprintf(stderr, "plmdcaBackend: the %d-GPU path failed (%s); using one GPU\n", G, last_error());
        if (deep) {     // weights were to be computed by the workers
            std::vector<uint32_t> counts((size_t)N);
            if (dca_weights(msa.data(), N, L, (double)seqid, 0, 0, N, counts.data()) != 0) {
                fprintf(stderr, "plmdcaBackend: %s\n", last_error());
                return nullptr;
            }
            for (long n = 0; n < N; ++n) w[n] = 1.f / (float)counts[n];
        }
    }
    dca_plm *h = dca_plm_create(msa.data(), w.data(), N, L, q, lambda_h, lambda_J, 0, L);
    if (!h) { fprintf(stderr, "plmdcaBackend: %s\n", last_error()); return nullptr; }
    mark("create");
    const size_t P = (size_t)dca_plm_num_params(h);
    // result buffer (plmdcaBackend.cpp:61 lbfgs_malloc): pinned so the final D2H runs at
    // full PCIe rate; released by freeFieldsAndCouplings
    float *x = (float *)host_pinned_alloc(sizeof(float) * P);
    if (!x) { fprintf(stderr, "ERROR: Failed to allocate a memory block for variables.\n"); dca_plm_destroy(h); return nullptr; }
    dca_fit_result res{};
    // initial point (:77) and L-BFGS (:81) without leaving the device
    cudaStream_t st = plm_stream(h->prob);
    int rc = plm_set_fidelity(h->prob, backend_arithmetic());
    if (rc == 0) rc = plm_init_x_device(h->prob, h->d_x);
    mark("init");
    LbfgsResult r{};
    double timed_ms = 0.0;
    int timed_evals = 0;
    if (rc == 0) rc = lbfgs_minimize(h->prob, h->d_x, (int)max_iteration, verbose ? 1 : 0, nullptr, nullptr, &r, 0,
                                     &timed_ms, &timed_evals);
    if (rc == 0 && cudaMemcpyAsync(x, h->d_x, sizeof(float) * P, cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = -1;
    if (rc == 0 && cudaStreamSynchronize(st) != cudaSuccess) rc = -1;
    res.status = r.status; res.iterations = r.iterations; res.evaluations = r.evaluations; res.fx = r.fx;
    res.device_ms = r.device_ms; res.timed_ms = timed_ms; res.timed_evaluations = timed_evals;
    mark("fit+d2h");
    dca_plm_destroy(h);
    mark("destroy");
    if (rc != 0) { fprintf(stderr, "plmdcaBackend: %s\n", last_error()); host_pinned_free(x); return nullptr; }
    t_last_fit = res;
    return x;
}

int dca_plm_fit_asymmetric(const uint8_t *msa, const float *weights, long num_seqs, int seqs_len, int num_site_states,
                           float lambda_h, float lambda_J, int max_iterations, int gpus, int verbose,
                           float *x_out_host, dca_fit_result *result)
{
    if (!msa || !weights || !x_out_host) { set_error("dca_plm_fit_asymmetric: null argument"); return -1; }
    return fit_asymmetric(msa, weights, num_seqs, seqs_len, num_site_states, lambda_h, lambda_J, max_iterations, gpus,
                          verbose, x_out_host, result);
}

int dca_last_fit(dca_fit_result *result)
{
    if (!result) { set_error("dca_last_fit: null argument"); return -1; }
    *result = t_last_fit;
    return 0;
}

void freeFieldsAndCouplings(void *h_and_J)
{
    if (h_and_J && !host_pinned_free(h_and_J)) free(h_and_J);
}

// ---- mean-field operators ---------------------------------------------------------
int dca_mf_freqs(const uint8_t *msa, const double *weights, long num_seqs, int seqs_len,
                 int num_site_states, double *fi_out, double *fij_out)
{
    if (!msa || !weights || !fi_out) { set_error("dca_mf_freqs: null argument"); return -1; }
    return mf_freqs_host(msa, weights, num_seqs, seqs_len, num_site_states, fi_out, fij_out);
}

int dca_mf_corr_mat(const uint8_t *msa, const double *weights, long num_seqs, int seqs_len,
                    int num_site_states, double pseudocount, double *reg_fi_out, double *corr_out)
{
    if (!msa || !weights || !corr_out) { set_error("dca_mf_corr_mat: null argument"); return -1; }
    return mf_corr_mat_host(msa, weights, num_seqs, seqs_len, num_site_states, pseudocount, reg_fi_out, corr_out);
}

int dca_mf_couplings(const double *corr, int n, double *couplings_out)
{
    if (!corr || !couplings_out || n <= 0) { set_error("dca_mf_couplings: bad argument"); return -1; }
    return mf_couplings_host(corr, n, couplings_out);
}

int dca_mf_fn_apc(const double *couplings, int seqs_len, int num_site_states, double *fn_out, double *fn_apc_out)
{
    if (!couplings || !fn_out) { set_error("dca_mf_fn_apc: null argument"); return -1; }
    return mf_fn_apc_host(couplings, seqs_len, num_site_states, fn_out, fn_apc_out);
}

int dca_mf_scores(const uint8_t *msa, const double *weights, long num_seqs, int seqs_len, int num_site_states,
                  double pseudocount, double *reg_fi_out, double *couplings_out, double *fn_out,
                  double *fn_apc_out, double *di_out, float *stage_ms_out)
{
    if (!msa || !weights || !fn_out) { set_error("dca_mf_scores: null argument"); return -1; }
    return mf_scores_host(msa, weights, num_seqs, seqs_len, num_site_states, pseudocount, reg_fi_out,
                          couplings_out, fn_out, fn_apc_out, di_out, stage_ms_out);
}

int dca_sw_scores(const uint8_t *msa, long num_seqs, int seqs_len, int num_site_states, const uint8_t *ref_codes,
                  int ref_len, int *scores_out)
{
    return sw_scores_host(msa, num_seqs, seqs_len, num_site_states, ref_codes, ref_len, scores_out);
}

int dca_fp64_mma_peak(int reps, double *tflops_out) { return fp64_mma_peak(reps, tflops_out); }

int dca_mf_corr_from_freqs(const double *reg_fi, const double *reg_fij, int seqs_len, int num_site_states,
                           double *corr_out)
{
    if (!reg_fi || !reg_fij || !corr_out) { set_error("dca_mf_corr_from_freqs: null argument"); return -1; }
    return mf_corr_from_freqs_host(reg_fi, reg_fij, seqs_len, num_site_states, corr_out);
}

int dca_mf_direct_info(const double *couplings, const double *reg_fi, int seqs_len, int num_site_states,
                       double *di_out, double *fields_out)
{
    if (!couplings || !reg_fi || !di_out) { set_error("dca_mf_direct_info: null argument"); return -1; }
    return mf_di_host(couplings, reg_fi, seqs_len, num_site_states, di_out, fields_out);
}

int dca_plm_direct_info(const float *x_host, const double *reg_fi, int seqs_len, int num_site_states,
                        double *di_out, double *fields_out)
{
    if (!x_host || !reg_fi || !di_out) { set_error("dca_plm_direct_info: null argument"); return -1; }
    return plm_di_host(x_host, reg_fi, seqs_len, num_site_states, di_out, fields_out);
}

}  // extern "C"
