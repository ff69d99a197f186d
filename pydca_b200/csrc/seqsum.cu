// seqsum.cu -- exact evaluation of SEQUENTIAL float32 sums on the GPU.
//
// Replaces, for the reference-arithmetic mode of the optimiser (lbfgs.cu):
//   vecdot / vec2norm   pydca/plmdca/lbfgs/include/arithmetic_ansi.h:114-133
//   the L2 accumulation of PlmDCA::gradient   pydca/plmdca/plmdca_numerics.cpp:463-486
//   fx += fxi (site order, num_threads = 1)   :572
// These run s = fl32(s + t_k) over all P parameters; the rounding bias of that chain (1.8 % of
// |g|^2 at config 2) is part of the reference's results, so it is reproduced bit for bit.
//
// One CTA per sum (the chain is serial; independent sums of one optimiser step run as
// separate CTAs of one launch).  A CTA walks the vector in groups of 4096 terms staged by a
// 4-deep cp.async ring.  Per group every thread rounds its 4 terms to integers in the
// binade of the running value (seqsum_core.h), a warp scan + one cross-warp scan give the
// group's total and prefix range, and if the running value provably stays inside its binade
// the whole group is applied in ONE integer addition.  Otherwise warp 0 redoes the group
// in blocks of 128, then 32 terms, and finally term by term -- always exact.
#include "dca_common.cuh"
#include "dca_internal.h"
#include "seqsum_core.h"

namespace dca {

namespace {

constexpr int SQ_THREADS = 1024;
constexpr int SQ_GROUP = 4 * SQ_THREADS;
constexpr int SQ_STAGES = 4;
constexpr int SQ_WARPS = SQ_THREADS / 32;
constexpr size_t SQ_SMEM = sizeof(float) * 2 * SQ_STAGES * SQ_GROUP;    // x and y rings: 128 KB

struct SeqLaunch { SeqJob job[SEQ_MAX_JOBS]; };

__device__ __forceinline__ SeqSum shfl_up_sum(const SeqSum &v, int o, bool ties)
{
    SeqSum r;
    r.total[0] = __shfl_up_sync(0xffffffffu, v.total[0], o);
    r.pmin[0] = __shfl_up_sync(0xffffffffu, v.pmin[0], o);
    r.pmax[0] = __shfl_up_sync(0xffffffffu, v.pmax[0], o);
    if (ties) {
        r.total[1] = __shfl_up_sync(0xffffffffu, v.total[1], o);
        r.pmin[1] = __shfl_up_sync(0xffffffffu, v.pmin[1], o);
        r.pmax[1] = __shfl_up_sync(0xffffffffu, v.pmax[1], o);
    } else {
        r.total[1] = r.total[0]; r.pmin[1] = r.pmin[0]; r.pmax[1] = r.pmax[0];
    }
    return r;
}

// inclusive scan over the lanes with seq_compose: lane 31 ends up with the warp's summary.
// `ties` (warp-uniform): some summary depends on the parity; otherwise both halves are equal
// and only one is moved.
__device__ __forceinline__ SeqSum warp_scan(SeqSum v, bool ties)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const SeqSum up = shfl_up_sum(v, o, ties);
        if (lane >= o) v = seq_compose(up, v);
    }
    return v;
}

// summary of NT consecutive terms held by one lane
template <int NT>
__device__ __forceinline__ SeqSum lane_summary(const SeqCtx &c, const float *t, bool *bad, bool *tie)
{
    int r0[NT], r1[NT];
    bool anytie = false;
#pragma unroll
    for (int k = 0; k < NT; ++k) {
        seq_round_term(c, t[k], &r0[k], &r1[k], bad);
        anytie |= (r0[k] != r1[k]);
    }
    *tie = anytie;
    SeqSum s;
    if (!anytie) {          // common case: plain prefix sums
        int run = r0[0], mn = run, mx = run;
#pragma unroll
        for (int k = 1; k < NT; ++k) { run += r0[k]; mn = min(mn, run); mx = max(mx, run); }
        s.total[0] = s.total[1] = run; s.pmin[0] = s.pmin[1] = mn; s.pmax[0] = s.pmax[1] = mx;
    } else {
        s = seq_single(r0[0], r1[0]);
#pragma unroll
        for (int k = 1; k < NT; ++k) s = seq_compose(s, seq_single(r0[k], r1[k]));
    }
    return s;
}

// Summary of the warp's 32*NT consecutive terms from the lanes' summaries; valid in every lane
// (tie-free path) or in lane 31 (general path) -- read it from lane 31.  Without ties the
// summaries do not depend on the parity: one integer scan of the totals and two hardware
// warp reductions (per-lane totals are below 2^24, the warp's below 2^29: no overflow).
__device__ __forceinline__ SeqSum warp_summary(SeqSum v, bool ties)
{
    if (ties) return warp_scan(v, true);
    const int lane = threadIdx.x & 31;
    int inc = v.total[0];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += up;
    }
    const int exc = inc - v.total[0];
    const int wmin = __reduce_min_sync(0xffffffffu, exc + v.pmin[0]);
    const int wmax = __reduce_max_sync(0xffffffffu, exc + v.pmax[0]);
    const int wtot = __shfl_sync(0xffffffffu, inc, 31);
    SeqSum o;
    o.total[0] = o.total[1] = max(-SEQ_SAT, min(SEQ_SAT, wtot));      // composed fields stay within +/-SEQ_SAT
    o.pmin[0] = o.pmin[1] = max(wmin, -SEQ_SAT);
    o.pmax[0] = o.pmax[1] = min(wmax, SEQ_SAT);
    return o;
}

// warp 0's exact fallback: try NT terms per lane (32*NT consecutive terms) in the integer
// picture from the current running value s (uniform over the warp); true = applied
template <int NT>
__device__ __forceinline__ bool try_warp(float &s, const float *t)
{
    SeqCtx c;
    if (!seq_ctx_make(s, &c)) return false;
    bool bad = false, tie = false;
    SeqSum v = lane_summary<NT>(c, t, &bad, &tie);
    if (__any_sync(0xffffffffu, bad)) return false;
    v = warp_summary(v, __any_sync(0xffffffffu, tie));
    const int p = c.M & 1;
    const int total = __shfl_sync(0xffffffffu, v.total[p], 31);
    const int pmin = __shfl_sync(0xffffffffu, v.pmin[p], 31);
    const int pmax = __shfl_sync(0xffffffffu, v.pmax[p], 31);
    if (!seq_block_ok(c, pmin, pmax)) return false;
    s = seq_apply(c, total);
    return true;
}

// the group [0, 4096) of tbuf, from running value s, exactly; executed by one full warp
__device__ float fallback_group(float s, const float *tbuf)
{
    const int lane = threadIdx.x & 31;
    for (int b = 0; b < SQ_GROUP / 128; ++b) {
        const float4 q = *reinterpret_cast<const float4 *>(tbuf + b * 128 + 4 * lane);
        const float t4[4] = {q.x, q.y, q.z, q.w};
        if (try_warp<4>(s, t4)) continue;
        for (int h = 0; h < 4; ++h) {
            const float t1 = tbuf[b * 128 + h * 32 + lane];
            if (try_warp<1>(s, &t1)) continue;
            // the plain chain: every lane runs it on broadcast terms, s stays warp-uniform
#pragma unroll 8
            for (int k = 0; k < 32; ++k) s = __fadd_rn(s, __shfl_sync(0xffffffffu, t1, k));
        }
    }
    return s;
}

__global__ void __launch_bounds__(SQ_THREADS, 1)
seqsum_kernel(SeqLaunch jobs, float *__restrict__ out)
{
    extern __shared__ __align__(16) float ring[];          // [SQ_STAGES][2][SQ_GROUP]
    __shared__ __align__(16) float tbuf[SQ_GROUP];
    __shared__ SeqSum wsum[SQ_WARPS];
    __shared__ int wflag[SQ_WARPS];                        // bit 0: bad term, bit 1: tie
    __shared__ float s_sh;

    const SeqJob J = jobs.job[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long n = J.n;
    const long ngroups = (n + SQ_GROUP - 1) / SQ_GROUP;
    const bool two = (J.kind == SEQ_DOT);

    auto issue = [&](long g) {
        if (g < ngroups) {
            const long base = g * SQ_GROUP + 4 * tid;
            if (base + 4 <= n) {
                float *dst = ring + (size_t)(g % SQ_STAGES) * 2 * SQ_GROUP + 4 * tid;
                cp_async16(dst, J.x + base);
                if (two) cp_async16(dst + SQ_GROUP, J.y + base);
            }
        }
        cp_async_commit();
    };

    if (tid == 0) s_sh = 0.f;
    for (int g = 0; g < SQ_STAGES - 1; ++g) issue(g);
    __syncthreads();

    for (long g = 0; g < ngroups; ++g) {
        issue(g + SQ_STAGES - 1);
        cp_async_wait<SQ_STAGES - 1>();                    // my own copies of group g have landed
        const long base = g * SQ_GROUP + 4 * tid;
        float xv[4], yv[4], t[4];
        if (base + 4 <= n) {
            const float *src = ring + (size_t)(g % SQ_STAGES) * 2 * SQ_GROUP + 4 * tid;
            const float4 a = *reinterpret_cast<const float4 *>(src);
            xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w;
            if (two) {
                const float4 b = *reinterpret_cast<const float4 *>(src + SQ_GROUP);
                yv[0] = b.x; yv[1] = b.y; yv[2] = b.z; yv[3] = b.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                xv[k] = (base + k < n) ? J.x[base + k] : 0.f;
                yv[k] = (two && base + k < n) ? J.y[base + k] : 0.f;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (two) t[k] = __fmul_rn(xv[k], yv[k]);                       // arithmetic_ansi.h:119
            else {                                                         // plmdca_numerics.cpp:470, :482
                const float lam = (base + k < J.n0) ? J.lam0 : J.lam1;
                t[k] = __fmul_rn(__fmul_rn(lam, xv[k]), xv[k]);
            }
            if (base + k >= n) t[k] = 0.f;
        }
        *reinterpret_cast<float4 *>(tbuf + 4 * tid) = make_float4(t[0], t[1], t[2], t[3]);

        const float s = s_sh;
        SeqCtx c;
        const bool ctx_ok = seq_ctx_make(s, &c);           // uniform over the CTA
        if (ctx_ok) {
            bool bad = false, tie = false;
            SeqSum v = lane_summary<4>(c, t, &bad, &tie);
            const bool wbad = __any_sync(0xffffffffu, bad), wtie = __any_sync(0xffffffffu, tie);
            v = warp_summary(v, wtie);
            if (lane == 31) { wsum[warp] = v; wflag[warp] = (wbad ? 1 : 0) | (wtie ? 2 : 0); }
        }
        __syncthreads();                                   // tbuf and the warp summaries are visible
        if (warp == 0) {
            float s_new = s;
            bool pass = false;
            if (ctx_ok) {
                SeqSum v = wsum[lane];
                const int fl = wflag[lane];
                const bool anybad = __any_sync(0xffffffffu, (fl & 1) != 0);
                v = warp_scan(v, __any_sync(0xffffffffu, (fl & 2) != 0));
                const int p = c.M & 1;
                const int total = __shfl_sync(0xffffffffu, v.total[p], 31);
                const int pmin = __shfl_sync(0xffffffffu, v.pmin[p], 31);
                const int pmax = __shfl_sync(0xffffffffu, v.pmax[p], 31);
                pass = !anybad && seq_block_ok(c, pmin, pmax);
                if (pass) s_new = seq_apply(c, total);
            }
            if (!pass) s_new = fallback_group(s, tbuf);
            if (lane == 0) s_sh = s_new;
        }
        __syncthreads();
    }

    float s = s_sh;
    if (J.fxsite) {
        // fx = reg; fx += fxi for i = 0..L-1 (plmdca_numerics.cpp:572 in site order, i.e. the
        // reference at num_threads = 1).  fxi: float64 sum of the per-(scan chunk, site)
        // partials, rounded to the reference's float once.
        for (int i0 = 0; i0 < J.L; i0 += SQ_GROUP) {
            __syncthreads();
            for (int i = i0 + tid; i < min(J.L, i0 + SQ_GROUP); i += SQ_THREADS) {
                double a = 0.0;
                for (long k = 0; k < J.nscan; ++k) a += J.fxsite[k * J.L + i];
                tbuf[i - i0] = (float)a;
            }
            __syncthreads();
            if (tid == 0) {
                for (int i = i0; i < min(J.L, i0 + SQ_GROUP); ++i) s = __fadd_rn(s, tbuf[i - i0]);
                s_sh = s;
            }
            __syncthreads();
            s = s_sh;
        }
    }
    if (tid == 0) out[J.out] = s;
}

}  // namespace

int seqsum_prepare()
{
    DCA_CUDA(cudaFuncSetAttribute(seqsum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SQ_SMEM));
    return 0;
}

int seqsum_launch(const SeqJob *jobs, int njobs, float *d_out, cudaStream_t st)
{
    if (njobs < 1 || njobs > SEQ_MAX_JOBS) { set_error("seqsum_launch: %d jobs (1..%d)", njobs, SEQ_MAX_JOBS); return -1; }
    SeqLaunch l{};
    for (int k = 0; k < njobs; ++k) l.job[k] = jobs[k];
    DCA_LAUNCH(seqsum_kernel, njobs, SQ_THREADS, SQ_SMEM, st, l, d_out);
    return 0;
}

// host-buffer entry for the parity tests: out = sequential float32 dot(x, y) (y != null) or the
// L2 accumulation of x with lam0 on the first n0 entries and lam1 on the rest (y == null)
int seqsum_host(const float *x, const float *y, long n, long n0, float lam0, float lam1, float *out)
{
    if (require_device() != 0) return -1;
    if (!x || !out || n <= 0) { set_error("seqsum_host: bad arguments"); return -1; }
    if (seqsum_prepare() != 0) return -1;
    float *dx = nullptr, *dy = nullptr, *dout = nullptr;
    int rc = -1;
    cudaStream_t st = nullptr;
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { set_error("seqsum_host: stream"); return -1; }
    do {
        if (!dev_alloc_t(&dx, (size_t)n) || !dev_alloc_t(&dout, (size_t)4)) break;
        if (y && !dev_alloc_t(&dy, (size_t)n)) break;
        if (cudaMemcpyAsync(dx, x, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, st) != cudaSuccess) break;
        if (y && cudaMemcpyAsync(dy, y, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, st) != cudaSuccess) break;
        SeqJob j{};
        j.x = dx; j.y = dy; j.n = n; j.kind = y ? SEQ_DOT : SEQ_REG; j.n0 = n0; j.lam0 = lam0; j.lam1 = lam1; j.out = 0;
        if (seqsum_launch(&j, 1, dout, st) != 0) break;
        if (cudaMemcpyAsync(out, dout, sizeof(float), cudaMemcpyDeviceToHost, st) != cudaSuccess) break;
        if (cudaStreamSynchronize(st) != cudaSuccess) break;
        rc = 0;
    } while (0);
    if (rc != 0 && cudaPeekAtLastError() != cudaSuccess) set_error("seqsum_host: %s", cudaGetErrorString(cudaGetLastError()));
    cudaStreamSynchronize(st);
    dev_free(dx); dev_free(dy); dev_free(dout);
    cudaStreamDestroy(st);
    return rc;
}

}  // namespace dca
