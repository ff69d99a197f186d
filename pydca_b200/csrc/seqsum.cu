// seqsum.cu -- exact evaluation of SEQUENTIAL float32 sums on the GPU.
//
// Replaces, for the reference-arithmetic mode of the optimiser (lbfgs.cu):
//   vecdot / vec2norm   pydca/plmdca/lbfgs/include/arithmetic_ansi.h:114-133
//   the L2 accumulation of PlmDCA::gradient   pydca/plmdca/plmdca_numerics.cpp:463-486
//   fx += fxi (site order, num_threads = 1)   :572
// These run s = fl32(s + t_k) over all P parameters; the rounding bias of that chain (1.8 % of
// |g|^2 at config 2) is part of the reference's results, so it is reproduced bit for bit.
//
// Two kernels.  seqsum_pipe_kernel (the one the optimiser uses) is a producer / consumer
// pipeline per sum: P producer CTAs take the 4096-term chunks round-robin and reduce each to a
// 32-byte record (total and prefix range, for both parities) for the THREE binades around the
// exponent predicted for the running value there (pre-pass: float64 prefix of the chunk sums,
// corrected by the drift the consumer publishes), and keep the 32 row summaries behind it.  In
// the consumer CTA four fetch warps copy the records into a shared-memory ring as they become
// ready; one chain warp applies 32 of them per step -- candidate matching the running value's
// exponent, records composed by a warp scan, range check against M, one integer addition.  A
// record that does not fit (binade crossed inside the chunk) is walked by the chain warp from
// the row summaries, the one row that does not fit by the plain chain; what that cannot do (no
// candidate for the exponent, running value 0: the first chunk of every sum) goes through the
// exact in-CTA path below.  The result is exact in every case; only the speed depends on the
// speculation.
// seqsum_kernel is the single-CTA version (no speculation), kept as the fallback.
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

#include <algorithm>

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

// The same scan for summaries that do not depend on the parity (no exact tie inside: 99.8 % of the rows of the
// config-2 fit) and are small (|fields| < 2^24: 32 of them add up without overflow): an integer scan of the totals
// and prefix minima / maxima of the absolute prefix ranges -- 15 shuffles and ~30 integer operations in three short
// dependent chains instead of five rounds of seq_compose on both parities (the chain warp runs alone on its
// scheduler: ~1.2 k of the ~1.7 k cycles of one of its steps were this scan).  `valid` = false: neutral element.
// Equal to warp_scan wherever no field saturates; a saturated field only ever turns "fits" into "does not fit".
__device__ __forceinline__ SeqSum warp_scan_plain(const SeqSum &v, bool valid)
{
    const int lane = threadIdx.x & 31;
    const int tot = valid ? v.total[0] : 0;
    int inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += up;
    }
    const int exc = inc - tot;
    int a = valid ? exc + v.pmin[0] : (1 << 30), b = valid ? exc + v.pmax[0] : -(1 << 30);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int ua = __shfl_up_sync(0xffffffffu, a, o), ub = __shfl_up_sync(0xffffffffu, b, o);
        if (lane >= o) { a = min(a, ua); b = max(b, ub); }
    }
    SeqSum r;
    r.total[0] = r.total[1] = inc;
    r.pmin[0] = r.pmin[1] = a;
    r.pmax[0] = r.pmax[1] = b;
    return r;
}

// warp-uniform: can warp_scan_plain take these summaries?  (lanes that are not valid do not count)
__device__ __forceinline__ bool scan_is_plain(const SeqSum &v, bool valid)
{
    const int lim = 1 << 24;
    const bool ok = !valid || (v.total[0] == v.total[1] && v.pmin[0] == v.pmin[1] && v.pmax[0] == v.pmax[1] &&
                               v.total[0] > -lim && v.total[0] < lim && v.pmin[0] > -lim && v.pmin[0] < lim &&
                               v.pmax[0] > -lim && v.pmax[0] < lim);
    return __all_sync(0xffffffffu, ok);
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

// the plain chain over 32 terms in shared memory: every lane runs it on the same (broadcast)
// loads, so s stays warp-uniform; the loads do not depend on s and pipeline ahead of the adds
__device__ __forceinline__ float serial32(float s, const float *t)
{
#pragma unroll
    for (int k = 0; k < 32; k += 4) {
        const float4 q = *reinterpret_cast<const float4 *>(t + k);
        s = __fadd_rn(s, q.x); s = __fadd_rn(s, q.y); s = __fadd_rn(s, q.z); s = __fadd_rn(s, q.w);
    }
    return s;
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
            s = serial32(s, tbuf + b * 128 + h * 32);
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

// ---- producer / consumer pipeline -------------------------------------------------------------
constexpr int SEQ_MAX_ROUNDS = 6;            // exponent changes inside one chunk after which the row walk hands it to the CTA's exact path
constexpr int PC_RING = 128;                 // chunks in the consumer's shared-memory ring of records
constexpr int PC_FETCH = 4;                  // fetch warps of the consumer CTA (PC_RING / 32 batches in flight)
constexpr int PC_CHUNK = SQ_GROUP;             // 4096 terms: one row of 128 per warp
struct SeqRec { int total[2], pmin[2], pmax[2]; int flags; int e_base; };      // 32 bytes
static_assert(sizeof(SeqRec) == 32, "record layout");

struct SeqPipeDev {
    unsigned long long *state;                 // [jobs][2]: (launch << 32) | progress, (launch << 32) | exponent field
    unsigned *ready;                           // [jobs][maxchunks] = (launch id << 8) | middle candidate's exponent, once the records are written
    SeqRec *rec;                               // [jobs][maxchunks][3]
    int4 *rowrec;                              // [jobs][maxchunks][3][32 rows][2]: the row (128-term) summaries behind each record + flags
    double *csum;                              // [jobs][maxchunks] float64 sums of the chunks' terms (pre-pass)
    int *epred;                                // [jobs][maxchunks] exponent field predicted for the running value entering chunk c (0: none)
    unsigned *ticket;                          // [jobs] arrival counter of the pre-pass
    unsigned *stats;                           // [jobs][8]: chunks applied from records, chunks done by the CTA's exact path, ready time-outs, stale records, failed range checks, consumer kcycles (total, exact path), chunks done by the row walk
    long maxchunks;
};

__device__ __forceinline__ void st_release_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.gpu.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int lds_acquire_cta(const volatile int *p)
{
    int v;
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];\n" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(const_cast<const int *>(p))) : "memory");
    return v;
}
__device__ __forceinline__ void sts_release_cta(volatile int *p, int v)
{
    asm volatile("st.release.cta.shared.s32 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(const_cast<int *>(p))), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release_u32(unsigned *p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// the four terms of this thread at `base` (arithmetic_ansi.h:119 / plmdca_numerics.cpp:470, :482)
__device__ __forceinline__ void load_terms(const SeqJob &J, long base, float *t)
{
    const long n = J.n;
    const bool two = (J.kind == SEQ_DOT);
    float xv[4], yv[4];
    if (base + 4 <= n) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(J.x + base));
        xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w;
        if (two) {
            const float4 b = __ldg(reinterpret_cast<const float4 *>(J.y + base));
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
        if (two) t[k] = __fmul_rn(xv[k], yv[k]);
        else {
            const float lam = (base + k < J.n0) ? J.lam0 : J.lam1;
            t[k] = __fmul_rn(__fmul_rn(lam, xv[k]), xv[k]);
        }
        if (base + k >= n) t[k] = 0.f;
    }
}

// The same in two steps, for a prefetch that does not wait: issue the loads of the four terms at `base`
// (raw operands in registers), form the products when they are needed.  load_terms multiplies at once --
// a warp that calls it stalls until the loads are back.
__device__ __forceinline__ void load_raw(const SeqJob &J, long base, float4 &xr, float4 &yr)
{
    xr = make_float4(0.f, 0.f, 0.f, 0.f);
    yr = xr;
    if (base + 4 <= J.n) {
        xr = __ldg(reinterpret_cast<const float4 *>(J.x + base));
        if (J.kind == SEQ_DOT) yr = __ldg(reinterpret_cast<const float4 *>(J.y + base));
    } else {
        float xv[4], yv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            xv[k] = (base + k < J.n) ? J.x[base + k] : 0.f;
            yv[k] = (J.kind == SEQ_DOT && base + k < J.n) ? J.y[base + k] : 0.f;
        }
        xr = make_float4(xv[0], xv[1], xv[2], xv[3]);
        yr = make_float4(yv[0], yv[1], yv[2], yv[3]);
    }
}
__device__ __forceinline__ void terms_from_raw(const SeqJob &J, long base, const float4 &xr, const float4 &yr, float *t)
{
    const float xv[4] = {xr.x, xr.y, xr.z, xr.w}, yv[4] = {yr.x, yr.y, yr.z, yr.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (J.kind == SEQ_DOT) t[k] = __fmul_rn(xv[k], yv[k]);
        else {
            const float lam = (base + k < J.n0) ? J.lam0 : J.lam1;
            t[k] = __fmul_rn(__fmul_rn(lam, xv[k]), xv[k]);
        }
        if (base + k >= J.n) t[k] = 0.f;
    }
}

__device__ __forceinline__ int exp_field(float s)           // 0 when the integer picture does not apply
{
    const int ef = (int)((seq_bits(s) >> 23) & 0xffu);
    return (ef < 40 || ef > 214) ? 0 : ef;
}

// warp 0 of the consumer CTA: the chunk in tbuf from running value s, exactly.  Every row (128
// terms) was summarised for the three exponents around the one the running value is PREDICTED to
// have on entering it (float64 prefix of the row sums).  Lane l takes row l; the rows from w on
// that have a candidate for the current exponent are composed by a warp scan (both parities) and
// the range check of the composed prefixes says how many leading rows can be applied in ONE
// integer addition -- as long as it holds the exponent cannot change.  The first row that does not
// fit goes through the plain chain (128 dependent adds: ~600 cycles, cheaper than any finer
// speculation), then the same again from the next row with the exponent reached: one pass through
// the chunk however often the running value changes its binade (a sum starts from 0 and crosses
// half a dozen in its first chunk).
__device__ __noinline__ void walk_rows(float &s, const SeqSum (*ws)[SQ_WARPS], const int (*wf)[SQ_WARPS], const int *erow,
                          const float *tbuf)
{
    const int lane = threadIdx.x & 31;
    int w = 0;
    while (w < SQ_WARPS) {
        SeqCtx cc;
        if (seq_ctx_make(s, &cc)) {
            const int ef = exp_field(s);
            SeqSum v;
            int bad = 0;
            if (lane < w) {                                // rows already done: neutral element of seq_compose
                v.total[0] = v.total[1] = 0; v.pmin[0] = v.pmin[1] = SEQ_SAT; v.pmax[0] = v.pmax[1] = -SEQ_SAT;
            } else {
                const int em = erow[lane];
                const int idx = ef - (em - 1);
                if (em != 0 && idx >= 0 && idx <= 2) { v = ws[idx][lane]; bad = wf[idx][lane] & 1; }
                else { v.total[0] = v.total[1] = v.pmin[0] = v.pmin[1] = v.pmax[0] = v.pmax[1] = 0; bad = 1; }
            }
            const unsigned badmask = __ballot_sync(0xffffffffu, bad != 0);
            const int nusable = (badmask == 0u) ? 32 : (__ffs(badmask) - 1);
            {
                const bool valid = lane >= w && bad == 0;      // (bad rows hold zeros: harmless in either scan)
                v = scan_is_plain(v, valid) ? warp_scan_plain(v, valid) : warp_scan(v, true);
            }
            const int p = cc.M & 1;
            const bool okl = lane < w || (lane < nusable && seq_block_ok(cc, v.pmin[p], v.pmax[p]));
            const unsigned okmask = __ballot_sync(0xffffffffu, okl);
            const int nok = (okmask == 0xffffffffu) ? 32 : (__ffs(~okmask) - 1);
            if (nok > w) {
                const int total = __shfl_sync(0xffffffffu, v.total[p], nok - 1);
                s = seq_apply(cc, total);
                w = nok;
                if (w >= SQ_WARPS) break;
            }
        }
        // row w: the plain chain; every lane runs it on the same (broadcast) loads, s stays warp-uniform
#pragma unroll
        for (int h = 0; h < 4; ++h) s = serial32(s, tbuf + w * 128 + h * 32);
        ++w;
    }
}

// summaries of this warp's row (4 terms per lane) for the exponent fields e - 1, e, e + 1 (e = 0: none)
__device__ __forceinline__ void summarise_row3(int e, const float *t, SeqSum (*ws)[SQ_WARPS], int (*wf)[SQ_WARPS])
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int ef = e - 1 + k;
        SeqCtx cx;
        cx.M = 0; cx.u = 0.f;
        const bool valid = (e != 0 && ef >= 40 && ef <= 214);
        cx.scale = seq_from_bits((uint32_t)(277 - (valid ? ef : 127)) << 23);      // 2^(23 - (ef - 127))
        bool bad = !valid, tie = false;
        SeqSum v = lane_summary<4>(cx, t, &bad, &tie);
        const bool wbad = __any_sync(0xffffffffu, bad), wtie = __any_sync(0xffffffffu, tie);
        v = warp_summary(v, wtie);
        if (lane == 31) { ws[k][warp] = v; wf[k][warp] = (wbad ? 1 : 0) | (wtie ? 2 : 0); }
    }
}

// The chain warp's own way through a chunk whose record does not fit (typically: the running value
// crosses a binade inside it), from the producers' ROW summaries: lane l holds row l for the candidate
// of the current exponent; the rows from w on are composed by a warp scan and the leading ones that
// fit are applied in one integer addition; the first row that does not fit goes through the plain
// chain (its 128 terms are loaded here); then the same again from the next row with the candidate
// of the exponent reached.  No barrier, no re-summarising.  false (s untouched): a candidate is
// missing, the integer picture does not apply, or the exponent keeps changing -- the CTA's exact
// path does the chunk.
__device__ __noinline__ bool row_walk(float &s_io, const SeqJob &J, long c, int e_mid, const int4 *rows, float *scratch)
{
    const int lane = threadIdx.x & 31;
    if (e_mid == 0) return false;
    int4 A = make_int4(0, 0, 0, 0), B = make_int4(0, 0, 1, 0);
    int have = -1;                                         // candidate held in A, B
    float s = s_io;
    int w = 0, changes = 0;
    while (w < SQ_WARPS) {
        SeqCtx cc;
        if (!seq_ctx_make(s, &cc)) return false;
        const int ef = exp_field(s);
        const int idx = ef - (e_mid - 1);
        if (idx < 0 || idx > 2) return false;
        if (idx != have) {                                 // on entry and after a change of exponent
            A = __ldcg(rows + (idx * 32 + lane) * 2);
            B = __ldcg(rows + (idx * 32 + lane) * 2 + 1);
            have = idx;
        }
        SeqSum v;
        int bad = 0;
        if (lane < w) {                                    // rows already done: neutral element of seq_compose
            v.total[0] = v.total[1] = 0; v.pmin[0] = v.pmin[1] = SEQ_SAT; v.pmax[0] = v.pmax[1] = -SEQ_SAT;
        } else {
            v.total[0] = A.x; v.total[1] = A.y; v.pmin[0] = A.z; v.pmin[1] = A.w; v.pmax[0] = B.x; v.pmax[1] = B.y;
            bad = B.z & 1;
        }
        const unsigned badmask = __ballot_sync(0xffffffffu, bad != 0);
        const int nusable = (badmask == 0u) ? 32 : (__ffs(badmask) - 1);
        {
            const bool valid = lane >= w;
            v = scan_is_plain(v, valid) ? warp_scan_plain(v, valid) : warp_scan(v, true);
        }
        const int p = cc.M & 1;
        const bool okl = lane < w || (lane < nusable && seq_block_ok(cc, v.pmin[p], v.pmax[p]));
        const unsigned okmask = __ballot_sync(0xffffffffu, okl);
        const int nok = (okmask == 0xffffffffu) ? 32 : (__ffs(~okmask) - 1);
        if (nok > w) {
            const int total = __shfl_sync(0xffffffffu, v.total[p], nok - 1);
            s = seq_apply(cc, total);
            w = nok;
            if (w >= SQ_WARPS) break;
        }
        float t[4];
        load_terms(J, c * PC_CHUNK + (long)w * 128 + 4 * lane, t);
        __syncwarp();
        *reinterpret_cast<float4 *>(scratch + 4 * lane) = make_float4(t[0], t[1], t[2], t[3]);
        __syncwarp();
#pragma unroll
        for (int h = 0; h < 4; ++h) s = serial32(s, scratch + h * 32);
        ++w;
        if (exp_field(s) != ef && ++changes > SEQ_MAX_ROUNDS) return false;
    }
    s_io = s;
    return true;
}

// Pre-pass: float64 sum of every chunk's terms; the last CTA of a job turns them into the
// exponent the running value is EXPECTED to have when it enters each chunk (exclusive prefix,
// rounded to float).  The sequential float32 value drifts away from this (stagnation), which
// the producers correct with the drift the consumer has observed so far.
__global__ void __launch_bounds__(256)
seqsum_prepass_kernel(SeqLaunch jobs, SeqPipeDev pd)
{
    __shared__ double red[8];
    __shared__ bool last;
    const int job = blockIdx.y;
    const SeqJob J = jobs.job[job];
    const long nchunks = (J.n + PC_CHUNK - 1) / PC_CHUNK;
    const long c = blockIdx.x;
    if (c >= nchunks) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double acc = 0.0;
#pragma unroll
    for (int r = 0; r < PC_CHUNK / (4 * 256); ++r) {
        float t[4];
        load_terms(J, c * PC_CHUNK + (long)r * 1024 + 4 * tid, t);
        acc += ((double)t[0] + (double)t[1]) + ((double)t[2] + (double)t[3]);
    }
    acc = warp_sum(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    double *csum = pd.csum + (size_t)job * pd.maxchunks;
    int *epred = pd.epred + (size_t)job * pd.maxchunks;
    if (tid == 0) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += red[w];
        csum[c] = v;
        __threadfence();
        last = (atomicInc(pd.ticket + job, (unsigned)nchunks - 1u) == (unsigned)nchunks - 1u);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    // exclusive prefix over the chunk sums: each thread a contiguous run, then a scan of the runs
    __shared__ double runs[256];
    const long per = (nchunks + 255) / 256;
    const long lo = min(nchunks, (long)tid * per), hi = min(nchunks, lo + per);
    double run = 0.0;
    for (long k = lo; k < hi; ++k) run += __ldcg(csum + k);
    runs[tid] = run;
    __syncthreads();
    if (tid == 0) {
        double a = 0.0;
        for (int k = 0; k < 256; ++k) { const double v = runs[k]; runs[k] = a; a += v; }
    }
    __syncthreads();
    double pre = runs[tid];
    for (long k = lo; k < hi; ++k) {
        epred[k] = exp_field((float)pre);
        pre += __ldcg(csum + k);
    }
}

__global__ void __launch_bounds__(SQ_THREADS, 1)
seqsum_pipe_kernel(SeqLaunch jobs, float *__restrict__ out, SeqPipeDev pd, int P, unsigned launch_id)
{
    __shared__ __align__(16) float tbuf[PC_CHUNK];
    __shared__ SeqSum wsum[2][3][SQ_WARPS];
    __shared__ int wflag[2][3][SQ_WARPS];
    __shared__ float s_sh;
    __shared__ long c_sh;
    __shared__ double rowsum_sh[SQ_WARPS];                 // consumer's exact path: float64 row sums and
    __shared__ int erow_sh[SQ_WARPS];                      // the exponent field predicted on entering each row
    // consumer only: ring of fetched records (three candidates per chunk), filled by the fetch warps
    __shared__ int4 ring[PC_RING][6];
    __shared__ volatile int ring_tag[PC_RING];             // chunk index + 1 once the slot holds that chunk
    __shared__ volatile int ring_e[PC_RING];               // exponent field of the chunk's middle candidate
    __shared__ volatile unsigned long long posexp_sh;      // the chain warp's position (low word) and the exponent field of its running value (high word)
    __shared__ volatile int stop_sh;                       // the chain warp has left the fast-forward loop

    const int job = blockIdx.x / (P + 1), role = blockIdx.x % (P + 1);
    const SeqJob J = jobs.job[job];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long nchunks = (J.n + PC_CHUNK - 1) / PC_CHUNK;
    unsigned long long *st_prog = pd.state + 2 * job, *st_exp = pd.state + 2 * job + 1;
    unsigned *ready = pd.ready + (size_t)job * pd.maxchunks;
    SeqRec *rec = pd.rec + (size_t)job * pd.maxchunks * 3;
    int4 *rowrec = pd.rowrec + (size_t)job * pd.maxchunks * 3 * 64;
    const unsigned long long tag = (unsigned long long)launch_id << 32;

    if (role > 0) {
        // ================= producer: chunks role-1, role-1+P, ... =================
        // No flow control: the exponent of the running value at chunk c is predicted by the
        // pre-pass (epred) and corrected by the drift the consumer has measured so far
        // (published exponent minus prediction at the consumer's position).
        // ONE barrier per chunk.  The exponent of the next chunk is worked out by thread 0 (from hints read two
        // chunks ahead with relaxed loads: they are hints) while the warps summarise the current one; a chunk is
        // published one barrier later, after the next chunk's summaries -- by then its records have been on their
        // way for a whole chunk and the release does not wait; the operands of the next chunk are prefetched RAW
        // (a product formed at the prefetch would stall the warp until the loads are back); and the chunk's record
        // is a reduction of the 32 row summaries, not a scan.  ncu before (profiles/r2q_seqsum_pipe_*): 21 % of the
        // samples at the barrier behind the fence, 10 % at the first product of the prefetched terms, 6 % at the
        // barrier behind thread 0's hints.
        const int *epred = pd.epred + (size_t)job * pd.maxchunks;
        __shared__ int e_next_sh[2], skip_next_sh[2];
        int buf = 0;
        long c = role - 1;
        float4 xn, yn;
        if (c < nchunks) load_raw(J, c * PC_CHUNK + warp * 128 + 4 * lane, xn, yn);
        unsigned long long a_n = 0, b_n = 0;               // thread 0: hint words and prediction for the NEXT chunk
        int ep_n = 0;
        auto decode_prog = [&](unsigned long long a) -> long { return ((a >> 32) == launch_id) ? (long)(a & 0xffffffffu) : 0; };
        auto predicted = [&](int ep, unsigned long long b, long prog, int et) -> int {
            const int e_seen = ((b >> 32) == launch_id) ? (int)(b & 0xffffffffu) : 0;
            int e = ep;
            if (e != 0 && e_seen != 0 && prog > 0 && prog < nchunks && et != 0) e += e_seen - et;
            return e;
        };
        if (tid == 0 && c < nchunks) {
            const unsigned long long a = ld_relaxed_u64(st_prog), b = ld_relaxed_u64(st_exp);
            const long prog = decode_prog(a);
            const int et = (prog > 0 && prog < nchunks) ? __ldcg(epred + prog) : 0;
            e_next_sh[0] = predicted(__ldcg(epred + c), b, prog, et);
            skip_next_sh[0] = (prog > c) ? 1 : 0;
            if (c + P < nchunks) { a_n = ld_relaxed_u64(st_prog); b_n = ld_relaxed_u64(st_exp); ep_n = __ldcg(epred + c + P); }
        }
        __syncthreads();
        long c_prev = -1;
        int e_prev = 0;
        for (; c < nchunks; c += P, buf ^= 1) {
            float t[4];
            terms_from_raw(J, c * PC_CHUNK + warp * 128 + 4 * lane, xn, yn, t);
            if (c + P < nchunks) load_raw(J, (c + P) * PC_CHUNK + warp * 128 + 4 * lane, xn, yn);      // next chunk in flight
            const int e = e_next_sh[buf];
            const bool skip = skip_next_sh[buf] != 0;       // the consumer did this chunk itself
            int et_n = 0;
            long prog_n = 0;
            if (tid == 0 && c + P < nchunks) {              // the hint words have arrived: prediction at the consumer's position
                prog_n = decode_prog(a_n);
                et_n = (prog_n > 0 && prog_n < nchunks) ? __ldcg(epred + prog_n) : 0;
            }
            if (!skip) summarise_row3(e, t, wsum[buf], wflag[buf]);
            if (tid == 0 && c + P < nchunks) {
                e_next_sh[buf ^ 1] = predicted(ep_n, b_n, prog_n, et_n);
                skip_next_sh[buf ^ 1] = (prog_n > c + P) ? 1 : 0;
            }
            __syncthreads();
            if (tid == 0) {
                // the previous chunk's records were written before this barrier by the three record warps
                if (c_prev >= 0) st_release_u32(ready + c_prev, (launch_id << 8) | (unsigned)(e_prev & 0xff));
                if (c + 2 * P < nchunks) { a_n = ld_relaxed_u64(st_prog); b_n = ld_relaxed_u64(st_exp); ep_n = __ldcg(epred + c + 2 * P); }
            }
            if (!skip && warp < 3) {
                SeqSum v = wsum[buf][warp][lane];
                const int fl = wflag[buf][warp][lane];
                {   // the row summaries themselves, for the consumer's walk through a chunk whose record does not fit
                    int4 *dst = rowrec + ((c * 3 + warp) * 32 + lane) * 2;
                    dst[0] = make_int4(v.total[0], v.total[1], v.pmin[0], v.pmin[1]);
                    dst[1] = make_int4(v.pmax[0], v.pmax[1], fl, 0);
                }
                // A row that moves the running integer by 2^24 or more cannot be part of a record that fits, and
                // the plain reduction below needs |row totals| < 2^24 (32 of them: no 32-bit overflow): such a
                // chunk is marked unusable like one with a term too large.
                const int lim = 1 << 24;
                const bool big = v.total[0] >= lim || v.total[0] <= -lim || v.total[1] >= lim || v.total[1] <= -lim ||
                                 v.pmax[0] >= lim || v.pmax[1] >= lim || v.pmin[0] <= -lim || v.pmin[1] <= -lim;
                const bool anybad = __any_sync(0xffffffffu, (fl & 1) != 0 || big);
                v = warp_summary(v, __any_sync(0xffffffffu, (fl & 2) != 0));        // valid in lane 31
                if (lane == 31) {
                    int4 *dst = reinterpret_cast<int4 *>(rec + c * 3 + warp);
                    dst[0] = make_int4(v.total[0], v.total[1], v.pmin[0], v.pmin[1]);
                    dst[1] = make_int4(v.pmax[0], v.pmax[1], anybad ? 1 : 0, e - 1);
                }
            }
            c_prev = skip ? -1 : c;
            e_prev = e;
        }
        __syncthreads();                                   // the last chunk: published right away
        if (tid == 0 && c_prev >= 0) st_release_u32(ready + c_prev, (launch_id << 8) | (unsigned)(e_prev & 0xff));
        return;
    }

    // ================= consumer =================
    if (tid == 0) { s_sh = 0.f; c_sh = 0; posexp_sh = 0ull; stop_sh = 0; st_release_u64(st_prog, tag); st_release_u64(st_exp, tag); }
    for (int i = tid; i < PC_RING; i += SQ_THREADS) ring_tag[i] = 0;
    __syncthreads();
    const long long clk_begin = clock64();
    long long clk_exact = 0;
    for (;;) {
        if (warp == 0) {
            // ---- fast-forward over the producers' records, 32 chunks per step ----
            // The chain warp reads nothing but shared memory: the fetch warps below copy the records
            // into a ring as the producers publish them (decoupled access / execute -- a global round
            // trip inside the chain, and two more after every exact chunk, was most of its time).
            // The records of a batch are composed by a warp scan (both parities); the range check of
            // the composed prefix tells how many leading chunks can be applied at once, and as long
            // as it holds the exponent -- hence the choice of candidate -- cannot change.
            float s = s_sh;
            long c = c_sh;
            bool stop = false;
            unsigned n_fast = 0, n_timeout = 0, n_stale = 0, n_range = 0, n_rows = 0;
            while (c < nchunks && !stop) {
                SeqCtx cc;
                if (!seq_ctx_make(s, &cc)) break;
                const int ef = exp_field(s);
                const long k = c + lane;
                const int slot = (int)(k % PC_RING);
                const bool rdy = k < nchunks && lds_acquire_cta(&ring_tag[slot]) == (int)(k + 1);      // tag before data
                const unsigned rmask = __ballot_sync(0xffffffffu, rdy);
                const int nready = (rmask == 0xffffffffu) ? 32 : (__ffs(~rmask) - 1);
                if (nready == 0) {
                    bool ok = false;
                    const long long t0 = clock64();
                    while (clock64() - t0 < 40000) {
                        if (lds_acquire_cta(&ring_tag[(int)(c % PC_RING)]) == (int)(c + 1)) { ok = true; break; }
                    }
                    if (!ok) { ++n_timeout; stop = true; }
                    continue;
                }
                SeqSum v;
                v.total[0] = v.total[1] = v.pmin[0] = v.pmin[1] = v.pmax[0] = v.pmax[1] = 0;
                int bad = 1;
                if (lane < nready) {
                    const int e_mid = ring_e[slot];        // exponent field of the middle candidate (0: none)
                    const int idx = ef - (e_mid - 1);
                    if (e_mid != 0 && idx >= 0 && idx <= 2) {
                        const int4 A = ring[slot][2 * idx], B = ring[slot][2 * idx + 1];
                        v.total[0] = A.x; v.total[1] = A.y; v.pmin[0] = A.z; v.pmin[1] = A.w;
                        v.pmax[0] = B.x; v.pmax[1] = B.y;
                        bad = B.z & 1;
                    }
                }
                const unsigned badmask = __ballot_sync(0xffffffffu, bad != 0);
                const int nusable = (badmask == 0u) ? 32 : (__ffs(badmask) - 1);      // leading usable records
                v = scan_is_plain(v, true) ? warp_scan_plain(v, true) : warp_scan(v, true);
                const int p = cc.M & 1;
                const bool okl = lane < nusable && seq_block_ok(cc, v.pmin[p], v.pmax[p]);
                const unsigned okmask = __ballot_sync(0xffffffffu, okl);
                const int nok = (okmask == 0xffffffffu) ? 32 : (__ffs(~okmask) - 1);
                if (nok > 0) {
                    const int total = __shfl_sync(0xffffffffu, v.total[p], nok - 1);
                    s = seq_apply(cc, total);
                    c += nok;
                    n_fast += (unsigned)nok;
                    // frees the ring slots behind c; fetch warp 0 forwards it to the producers as their hint
                    if (lane == 0) posexp_sh = ((unsigned long long)(unsigned)exp_field(s) << 32) | (unsigned long long)c;
                }
                if (nok < nready) {                        // the record of chunk c does not fit
                    if (nok >= nusable) ++n_stale; else ++n_range;
                    const int e_mid = ring_e[(int)(c % PC_RING)];
                    if (row_walk(s, J, c, e_mid, rowrec + c * 3 * 64, tbuf)) {
                        ++c; ++n_rows;
                        if (lane == 0) posexp_sh = ((unsigned long long)(unsigned)exp_field(s) << 32) | (unsigned long long)c;
                    } else stop = true;                    // the CTA's exact path
                }
            }
            if (lane == 0) {
                stop_sh = 1;
                s_sh = s; c_sh = c;
                unsigned *stt = pd.stats + 8 * job;
                stt[0] += n_fast; stt[2] += n_timeout; stt[3] += n_stale; stt[4] += n_range;
                stt[7] += n_rows;
            }
        } else if (warp <= PC_FETCH) {
            // ---- fetch warp: batches b = warp - 1 (mod PC_FETCH) of 32 chunks; every lane polls the ready
            // word of its chunk and copies the three candidate records into the ring slot.  Stateless:
            // what is already in the ring is told by the tags, what is no longer wanted by posexp_sh. ----
            const int fj = warp - 1;
            const unsigned want = launch_id << 8;
            long b = -1;
            unsigned long long pe_sent = ~0ull;
            while (stop_sh == 0) {
                const unsigned long long pe = posexp_sh;
                const long cp = (long)(pe & 0xffffffffull);
                if (fj == 0 && lane == 0 && pe != pe_sent) {       // hints for the producers: no ordering needed
                    st_relaxed_u64(st_prog, tag | (pe & 0xffffffffull));
                    st_relaxed_u64(st_exp, tag | (pe >> 32));
                    pe_sent = pe;
                }
                const long bmin = cp / 32;
                if (b < bmin) b = bmin + (((long)fj - bmin) % PC_FETCH + PC_FETCH) % PC_FETCH;
                if (b * 32 >= nchunks || b * 32 + 32 > cp + PC_RING) { __nanosleep(64); continue; }
                const long k = b * 32 + lane;
                const int slot = (int)(k % PC_RING);
                bool have = k >= nchunks || k < cp || ring_tag[slot] == (int)(k + 1);
                if (!have) {
                    const unsigned fl = ld_acquire_u32(ready + k);
                    if ((fl & 0xffffff00u) == want) {
                        const int4 *src = reinterpret_cast<const int4 *>(rec + k * 3);
                        int4 r[6];
#pragma unroll
                        for (int i = 0; i < 6; ++i) r[i] = __ldcg(src + i);
#pragma unroll
                        for (int i = 0; i < 6; ++i) ring[slot][i] = r[i];
                        ring_e[slot] = (int)(fl & 0xffu);
                        sts_release_cta(&ring_tag[slot], (int)(k + 1));      // data before tag
                        have = true;
                    }
                }
                if (__all_sync(0xffffffffu, have)) b += PC_FETCH;
            }
        }
        __syncthreads();
        const long c = c_sh;
        const float s = s_sh;
        if (c >= nchunks) break;
        const long long clk_e0 = clock64();
        // ---- exact path for chunk c: the whole CTA summarises its rows for the exponents the running
        // value is predicted to have on entering each of them, warp 0 walks the rows ----
        float t[4];
        load_terms(J, c * PC_CHUNK + 4 * tid, t);
        *reinterpret_cast<float4 *>(tbuf + 4 * tid) = make_float4(t[0], t[1], t[2], t[3]);
        {
            const double rs = warp_sum(((double)t[0] + (double)t[1]) + ((double)t[2] + (double)t[3]));
            if (lane == 0) rowsum_sh[warp] = rs;
        }
        __syncthreads();
        {
            const double pre = warp_sum((lane < warp) ? rowsum_sh[lane] : 0.0);
            const int e_row = exp_field((float)((double)s + pre));
            summarise_row3(e_row, t, wsum[0], wflag[0]);
            if (lane == 0) erow_sh[warp] = e_row;
        }
        __syncthreads();
        if (warp == 0) {
            float s_new = s;
            walk_rows(s_new, wsum[0], wflag[0], erow_sh, tbuf);
            if (lane == 0) s_sh = s_new;
        }
        __syncthreads();
        const float sc = s_sh;
        if (tid == 0) {
            c_sh = c + 1;
            posexp_sh = ((unsigned long long)(unsigned)exp_field(sc) << 32) | (unsigned long long)(c + 1);
            stop_sh = 0;                                   // the fetch warps resume behind the next barrier
            clk_exact += clock64() - clk_e0;
            pd.stats[8 * job + 1] += 1u;
            st_release_u64(st_prog, tag | (unsigned long long)(c + 1));
            st_release_u64(st_exp, tag | (unsigned long long)(unsigned)exp_field(sc));
        }
        __syncthreads();
    }

    if (tid == 0) {
        pd.stats[8 * job + 5] += (unsigned)((clock64() - clk_begin) >> 10);
        pd.stats[8 * job + 6] += (unsigned)(clk_exact >> 10);
    }
    float s = s_sh;
    if (J.fxsite) {                                        // as in seqsum_kernel
        for (int i0 = 0; i0 < J.L; i0 += PC_CHUNK) {
            __syncthreads();
            for (int i = i0 + tid; i < min(J.L, i0 + PC_CHUNK); i += SQ_THREADS) {
                double a = 0.0;
                for (long k = 0; k < J.nscan; ++k) a += J.fxsite[k * J.L + i];
                tbuf[i - i0] = (float)a;
            }
            __syncthreads();
            if (tid == 0) {
                for (int i = i0; i < min(J.L, i0 + PC_CHUNK); ++i) s = __fadd_rn(s, tbuf[i - i0]);
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

struct SeqPipe {
    SeqPipeDev dev{};
    unsigned launch_id = 0;
};

SeqPipe *seqsum_pipe_create(long nmax)
{
    const char *e = getenv("DCA_B200_SEQSUM");
    if (e && strcmp(e, "serial") == 0) return nullptr;       // single-CTA kernel only (measurements)
    SeqPipe *p = new SeqPipe();
    p->dev.maxchunks = (nmax + PC_CHUNK - 1) / PC_CHUNK;
    const size_t nrec = (size_t)SEQ_MAX_JOBS * p->dev.maxchunks;
    p->dev.state = static_cast<unsigned long long *>(dev_alloc(sizeof(unsigned long long) * 2 * SEQ_MAX_JOBS));
    p->dev.ready = static_cast<unsigned *>(dev_alloc(sizeof(unsigned) * nrec));
    p->dev.rec = static_cast<SeqRec *>(dev_alloc(sizeof(SeqRec) * 3 * nrec));
    p->dev.rowrec = static_cast<int4 *>(dev_alloc(sizeof(int4) * 3 * 64 * nrec));
    p->dev.stats = static_cast<unsigned *>(dev_alloc(sizeof(unsigned) * 8 * SEQ_MAX_JOBS));
    p->dev.csum = static_cast<double *>(dev_alloc(sizeof(double) * nrec));
    p->dev.epred = static_cast<int *>(dev_alloc(sizeof(int) * nrec));
    p->dev.ticket = static_cast<unsigned *>(dev_alloc(sizeof(unsigned) * SEQ_MAX_JOBS));
    if (!p->dev.state || !p->dev.ready || !p->dev.rec || !p->dev.rowrec || !p->dev.stats || !p->dev.csum || !p->dev.epred || !p->dev.ticket ||
        cudaMemset(p->dev.ticket, 0, sizeof(unsigned) * SEQ_MAX_JOBS) != cudaSuccess ||
        cudaMemset(p->dev.stats, 0, sizeof(unsigned) * 8 * SEQ_MAX_JOBS) != cudaSuccess ||
        cudaMemset(p->dev.state, 0, sizeof(unsigned long long) * 2 * SEQ_MAX_JOBS) != cudaSuccess ||
        cudaMemset(p->dev.ready, 0, sizeof(unsigned) * nrec) != cudaSuccess) {
        seqsum_pipe_destroy(p);
        return nullptr;
    }
    return p;
}

// accumulated counters of the pipeline since creation: out[jobs][8] (see SeqPipeDev::stats)
int seqsum_pipe_stats(SeqPipe *p, unsigned *out, cudaStream_t st)
{
    if (!p) return -1;
    DCA_CUDA(cudaMemcpyAsync(out, p->dev.stats, sizeof(unsigned) * 8 * SEQ_MAX_JOBS, cudaMemcpyDeviceToHost, st));
    DCA_CUDA(cudaStreamSynchronize(st));
    return 0;
}

void seqsum_pipe_destroy(SeqPipe *p)
{
    if (!p) return;
    dev_free(p->dev.state); dev_free(p->dev.ready); dev_free(p->dev.rec); dev_free(p->dev.rowrec); dev_free(p->dev.stats);
    dev_free(p->dev.csum); dev_free(p->dev.epred); dev_free(p->dev.ticket);
    delete p;
}

static int seqsum_launch_impl(const SeqJob *jobs, int njobs, float *d_out, cudaStream_t st, SeqPipe *pipe);

// DCA_B200_SEQSUM_TIMING=1 (measurements only): every launch is bracketed by CUDA events and a
// synchronisation; per (first output slot, job count) the accumulated time goes to stderr.
int seqsum_launch(const SeqJob *jobs, int njobs, float *d_out, cudaStream_t st, SeqPipe *pipe)
{
    static const bool timing = getenv("DCA_B200_SEQSUM_TIMING") != nullptr;
    if (!timing) return seqsum_launch_impl(jobs, njobs, d_out, st, pipe);
    static double acc_ms[32][SEQ_MAX_JOBS + 1];
    static long acc_n[32][SEQ_MAX_JOBS + 1];
    static long calls = 0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    const int rc = seqsum_launch_impl(jobs, njobs, d_out, st, pipe);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    const int slot = jobs[0].out & 31;
    acc_ms[slot][njobs] += ms; acc_n[slot][njobs] += 1;
    if (++calls % 100 == 0) {
        fprintf(stderr, "[seqsum timing] after %ld launches:", calls);
        for (int a = 0; a < 32; ++a)
            for (int b = 1; b <= SEQ_MAX_JOBS; ++b)
                if (acc_n[a][b]) fprintf(stderr, " out%d x%d: %ld x %.3f ms;", a, b, acc_n[a][b], acc_ms[a][b] / acc_n[a][b]);
        fprintf(stderr, "\n");
    }
    return rc;
}

static int seqsum_launch_impl(const SeqJob *jobs, int njobs, float *d_out, cudaStream_t st, SeqPipe *pipe)
{
    if (njobs < 1 || njobs > SEQ_MAX_JOBS) { set_error("seqsum_launch: %d jobs (1..%d)", njobs, SEQ_MAX_JOBS); return -1; }
    SeqLaunch l{};
    long nmax = 0;
    for (int k = 0; k < njobs; ++k) { l.job[k] = jobs[k]; nmax = std::max(nmax, jobs[k].n); }
    const long chunks = (nmax + PC_CHUNK - 1) / PC_CHUNK;
    if (!pipe || chunks > pipe->dev.maxchunks || chunks < 64) {
        DCA_LAUNCH(seqsum_kernel, njobs, SQ_THREADS, SQ_SMEM, st, l, d_out);
        return 0;
    }
    // every CTA of the grid must be resident (producers and consumer talk through memory):
    // 1024-thread CTAs take one SM each, so (P + 1) * njobs <= SM count
    int P = sm_count() / njobs - 1;
    P = (int)std::min<long>(P, std::max<long>(1, chunks / 4));
    if (P < 2) { DCA_LAUNCH(seqsum_kernel, njobs, SQ_THREADS, SQ_SMEM, st, l, d_out); return 0; }
    ++pipe->launch_id;
    if ((pipe->launch_id & 0xffffffu) == 0u) {
        // the ready words carry 24 bits of the launch id: start a new era with clean flags
        ++pipe->launch_id;
        DCA_CUDA(cudaMemsetAsync(pipe->dev.ready, 0, sizeof(unsigned) * (size_t)SEQ_MAX_JOBS * pipe->dev.maxchunks, st));
    }
    // the pre-pass needs equal chunk counts per job only for its grid; shorter jobs return early
    DCA_LAUNCH(seqsum_prepass_kernel, dim3((unsigned)chunks, (unsigned)njobs), 256, 0, st, l, pipe->dev);
    DCA_LAUNCH(seqsum_pipe_kernel, njobs * (P + 1), SQ_THREADS, 0, st, l, d_out, pipe->dev, P, pipe->launch_id);
    return 0;
}

// host-buffer entry for the parity tests: out = sequential float32 dot(x, y) (y != null) or the
// L2 accumulation of x with lam0 on the first n0 entries and lam1 on the rest (y == null)
int seqsum_host(const float *x, const float *y, long n, long n0, float lam0, float lam1, float *out,
                unsigned *stats8, float *kernel_ms)
{
    if (require_device() != 0) return -1;
    if (!x || !out || n <= 0) { set_error("seqsum_host: bad arguments"); return -1; }
    if (seqsum_prepare() != 0) return -1;
    float *dx = nullptr, *dy = nullptr, *dout = nullptr;
    SeqPipe *pipe = nullptr;
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
        pipe = seqsum_pipe_create(n);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        if (seqsum_launch(&j, 1, dout, st, pipe) != 0) break;          // warm (module load)
        if (pipe) cudaMemsetAsync(pipe->dev.stats, 0, sizeof(unsigned) * 8 * SEQ_MAX_JOBS, st);
        cudaEventRecord(e0, st);
        if (seqsum_launch(&j, 1, dout, st, pipe) != 0) break;
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        if (kernel_ms) cudaEventElapsedTime(kernel_ms, e0, e1);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        if (stats8) {
            memset(stats8, 0, sizeof(unsigned) * 8);
            if (pipe) cudaMemcpyAsync(stats8, pipe->dev.stats, sizeof(unsigned) * 8, cudaMemcpyDeviceToHost, st);
        }
        if (cudaMemcpyAsync(out, dout, sizeof(float), cudaMemcpyDeviceToHost, st) != cudaSuccess) break;
        if (cudaStreamSynchronize(st) != cudaSuccess) break;
        rc = 0;
    } while (0);
    if (rc != 0 && cudaPeekAtLastError() != cudaSuccess) set_error("seqsum_host: %s", cudaGetErrorString(cudaGetLastError()));
    cudaStreamSynchronize(st);
    dev_free(dx); dev_free(dy); dev_free(dout);
    seqsum_pipe_destroy(pipe);
    cudaStreamDestroy(st);
    return rc;
}

}  // namespace dca
