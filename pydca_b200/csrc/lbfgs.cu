// lbfgs.cu -- L-BFGS driver with device-resident vectors for the plmDCA objective.
//
// Replaces  lbfgs()                     pydca/plmdca/lbfgs/lib/lbfgs.cpp:248-644
//           line_search_morethuente()   :815-1004
//           update_trial_interval()     :1128-1295
//           vector ops                  pydca/plmdca/lbfgs/include/arithmetic_ansi.h:51-133
//           ObjectiveFunction::run      pydca/plmdca/plmdcaBackend.cpp:47-95 (parameters)
//
// Same algorithm and parameters as the reference's call site: m = 5,
// epsilon = 1e-3, ftol = 1e-4, gtol = 0.9, xtol = 1e-16, max_linesearch = 5,
// min_step = 1e-20, max_step = 1e20, More-Thuente line search, first step
// 1/|g|, H0 = (y.s / y.y) I, revert to the previous point when the line search
// fails.  Line-search scalars are float like the reference's lbfgsfloatval_t
// (lbfgs.h:50-56 forces float).  What differs by construction:
//   * the 16 vectors of P floats live in HBM; x/xp and g/gp are pointer-swapped
//     instead of copied;
//   * the two-loop recursion is evaluated in its compact form: two HBM passes per
//     iteration (history update + the new inner products; one linear combination)
//     instead of the reference's ~78 unfused vector passes -- see
//     vec_update_gram_kernel;
//   * inner products are reduced in float64 in a fixed order (the reference's are
//     sequential float32 sums, so its trajectory is not reproducible even against
//     itself, SURVEY finding 3).
//
// Host <-> device traffic per objective evaluation: one 48-byte D2H read of
// {fx, g.d, g.g, x.x}, the peer-timeout flag and the g.d of the direction built last
// (read together with the first trial's scalars: no round trip of its own).
#include "dca_common.cuh"
#include "dca_internal.h"

#include <algorithm>
#include <cmath>

namespace dca {

namespace {

constexpr int VEC_THREADS = 256;
constexpr int LB_M = 5;

// small device scalar block kept for the single-vector kernels below
enum { SC_ALPHA = 0, SC_YS = 5, SC_COEF = 10, SC_SCALE = 11, SC_FLOATS = 16 };

// Every kernel below streams P floats as float4 (the vectors come from cudaMalloc,
// 256-byte aligned; the <= 3 trailing elements are handled by one CTA), and the
// kernels that produce a dot product finish it themselves: each CTA writes its
// float64 partial, takes a ticket, and the LAST CTA reduces the partials in a
// fixed order and applies the scalar step of the recursion (no separate
// reduction launch, no host round trip, deterministic).
//
// scalar steps (op):
//   0: ys_slot = y.s ; scale = ys / yy                  (lbfgs.cpp:556-558, :591)
//   1: alpha_slot = (s.d) / ys_slot ; coef = -alpha      (:585-588)
//   2: beta = (y.d) / ys_slot ; coef = alpha_slot - beta (:596-599)
//   3: dscal[PS_DG] = g.d                                (:845, line search)
//   4: dscal[PS_DG] = g.d, dscal[PS_GG] = g.g            (multi-GPU, after the all-reduce)
//  -1: none
__device__ __forceinline__ void finish_dot(double r0, double r1, double *__restrict__ partial,
                                           int op, int slot, float *__restrict__ sc,
                                           double *__restrict__ dscal, unsigned *__restrict__ ticket,
                                           double *scratch)
{
    __shared__ bool last;
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x] = r0;
        partial[2 * blockIdx.x + 1] = r1;
        __threadfence();
        last = (atomicInc(ticket, gridDim.x - 1u) == gridDim.x - 1u);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    double s0 = 0.0, s1 = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
        s0 += partial[2 * b];
        s1 += partial[2 * b + 1];
    }
    s0 = block_sum(s0, scratch);
    s1 = block_sum(s1, scratch);
    if (threadIdx.x != 0) return;
    if (op == 0) {
        const float ys = (float)s0, yy = (float)s1;
        sc[SC_YS + slot] = ys;
        sc[SC_SCALE] = ys / yy;
    } else if (op == 1) {
        const float alpha = (float)s0 / sc[SC_YS + slot];
        sc[SC_ALPHA + slot] = alpha;
        sc[SC_COEF] = -alpha;
    } else if (op == 2) {
        const float beta = (float)s0 / sc[SC_YS + slot];
        sc[SC_COEF] = sc[SC_ALPHA + slot] - beta;
    } else if (op == 3) {
        dscal[PS_DG] = s0;
    } else if (op == 4) {
        dscal[PS_DG] = s0;
        dscal[PS_GG] = s1;
    }
}

#define VEC_LOOP4(n)                                                                              \
    const long n4__ = (n) >> 2;                                                                   \
    for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < n4__; v += (long)gridDim.x * blockDim.x)
#define VEC_TAIL(n, k)                                                                            \
    const long k = ((n) & ~3L) + threadIdx.x;                                                     \
    if (blockIdx.x == gridDim.x - 1 && k < (n))

// x = xp + stp * d with separate multiply and add (arithmetic_ansi.h:71-88:
// veccpy then vecadd, no fused multiply-add on the reference's x86 target)
__global__ void __launch_bounds__(VEC_THREADS)
vec_step_kernel(float *__restrict__ x, const float *__restrict__ xp, const float *__restrict__ d,
                float stp, long n)
{
    VEC_LOOP4(n) {
        const float4 a = reinterpret_cast<const float4 *>(xp)[v], b = reinterpret_cast<const float4 *>(d)[v];
        float4 o;
        o.x = __fadd_rn(a.x, __fmul_rn(stp, b.x)); o.y = __fadd_rn(a.y, __fmul_rn(stp, b.y));
        o.z = __fadd_rn(a.z, __fmul_rn(stp, b.z)); o.w = __fadd_rn(a.w, __fmul_rn(stp, b.w));
        reinterpret_cast<float4 *>(x)[v] = o;
    }
    VEC_TAIL(n, k) x[k] = __fadd_rn(xp[k], __fmul_rn(stp, d[k]));
}

// one fused pass of the two-loop recursion (lbfgs.cpp:568-601):
//   mode 0: d = -u                       (vecncpy :575)
//   mode 1: d += coef * u                (vecadd  :588 / :599), coef from device
//   mode 2: d *= scale                   (vecscale :591), scale from device
// followed, when v != null, by the dot v.d of the UPDATED d -- what the next
// step of the recursion needs (:585 / :596), or g.d for the line search (:845)
// -- finished by the last CTA with scalar step `op`.
__global__ void __launch_bounds__(VEC_THREADS)
vec_twoloop_kernel(float *__restrict__ d, const float *__restrict__ u, const float *__restrict__ vv,
                   float *__restrict__ sc, int mode, long n, double *__restrict__ partial,
                   int op, int slot, double *__restrict__ dscal, unsigned *__restrict__ ticket)
{
    __shared__ double scratch[32];
    const float c = (mode == 1) ? sc[SC_COEF] : ((mode == 2) ? sc[SC_SCALE] : 0.f);
    double dot = 0.0;
    VEC_LOOP4(n) {
        float4 dv;
        if (mode == 0) {
            const float4 a = reinterpret_cast<const float4 *>(u)[v];
            dv = make_float4(-a.x, -a.y, -a.z, -a.w);
        } else if (mode == 1) {
            const float4 a = reinterpret_cast<const float4 *>(u)[v], o = reinterpret_cast<const float4 *>(d)[v];
            dv.x = __fadd_rn(o.x, __fmul_rn(c, a.x)); dv.y = __fadd_rn(o.y, __fmul_rn(c, a.y));
            dv.z = __fadd_rn(o.z, __fmul_rn(c, a.z)); dv.w = __fadd_rn(o.w, __fmul_rn(c, a.w));
        } else {
            const float4 o = reinterpret_cast<const float4 *>(d)[v];
            dv = make_float4(__fmul_rn(o.x, c), __fmul_rn(o.y, c), __fmul_rn(o.z, c), __fmul_rn(o.w, c));
        }
        reinterpret_cast<float4 *>(d)[v] = dv;
        if (vv) {
            const float4 w = reinterpret_cast<const float4 *>(vv)[v];
            dot += (double)w.x * dv.x + (double)w.y * dv.y + (double)w.z * dv.z + (double)w.w * dv.w;
        }
    }
    VEC_TAIL(n, k) {
        float dv;
        if (mode == 0) dv = -u[k];
        else if (mode == 1) dv = __fadd_rn(d[k], __fmul_rn(c, u[k]));
        else dv = __fmul_rn(d[k], c);
        d[k] = dv;
        if (vv) dot += (double)vv[k] * dv;
    }
    if (vv) {
        const double r0 = block_sum(dot, scratch);
        finish_dot(r0, 0.0, partial, op, slot, sc, dscal, ticket, scratch);
    }
}

// {a.b, a.a} -> op 4 (multi-GPU path: g.d and g.g after the all-reduce)
__global__ void __launch_bounds__(VEC_THREADS)
vec_dots_kernel(const float *__restrict__ a, const float *__restrict__ b, long n,
                double *__restrict__ partial, float *__restrict__ sc, double *__restrict__ dscal,
                unsigned *__restrict__ ticket)
{
    __shared__ double scratch[32];
    double ab = 0.0, aa = 0.0;
    VEC_LOOP4(n) {
        const float4 p = reinterpret_cast<const float4 *>(a)[v];
        aa += (double)p.x * p.x + (double)p.y * p.y + (double)p.z * p.z + (double)p.w * p.w;
        if (b) {
            const float4 q = reinterpret_cast<const float4 *>(b)[v];
            ab += (double)p.x * q.x + (double)p.y * q.y + (double)p.z * q.z + (double)p.w * q.w;
        }
    }
    VEC_TAIL(n, k) {
        aa += (double)a[k] * a[k];
        if (b) ab += (double)a[k] * b[k];
    }
    const double r0 = block_sum(ab, scratch), r1 = block_sum(aa, scratch);
    finish_dot(r0, r1, partial, 4, 0, sc, dscal, ticket, scratch);
}

// ---- compact two-loop recursion ---------------------------------------------------
// The recursion of lbfgs.cpp:568-601 only ever needs inner products between the
// vectors {s_k, y_k, g}.  With the Gram entries SY[i][k] = s_i.y_k, YY[i][k] = y_i.y_k,
// sg[k] = s_k.g, yg[k] = y_k.g, gg = g.g every alpha_j, beta_j follows from scalar
// recurrences (exactly the quantities the reference computes on the updated vector):
//   s_j.q_t = -sg[j] - sum_{u<t} alpha_u SY[j][u]
//   y_j.r_t = gamma (-yg[j] - sum_u alpha_u YY[j][u]) + sum_{v older than t} c_v SY[v][j]
// and the direction is ONE linear combination
//   d = -gamma g - gamma sum_u alpha_u y_u + sum_v (alpha_v - beta_v) s_v,
// g.d (needed by the line search, :845) included.  Per iteration that is two passes
// over HBM -- (A) s,y update + the 5 new inner products per stored pair, (B) the
// combination -- instead of 2m+3 = 13 dependent passes: 28 vs 50 vector reads/writes of
// P floats at m = 5.  The scalar recursion runs in the last CTA of pass A (float64).
struct HistPtrs {
    const float *s[LB_M];
    const float *y[LB_M];
    int slot[LB_M];         // ring slot of entry t, t = 0 newest (the pair being written) ... bound-1 oldest
    int bound;
};

constexpr int GR_SY = 0;                       // [LB_M][LB_M]
constexpr int GR_YY = LB_M * LB_M;             // [LB_M][LB_M]
constexpr int GR_SG = 2 * LB_M * LB_M;         // [LB_M]
constexpr int GR_YG = GR_SG + LB_M;            // [LB_M]
constexpr int GR_GG = GR_YG + LB_M;
constexpr int GR_DOUBLES = GR_GG + 1;
constexpr int GR_NPART = 5 * LB_M;             // partial sums per CTA
// float coefficients for pass B: [0] g, [1..m] y by entry t, [1+m..2m] s by entry t
constexpr int CF_FLOATS = 1 + 2 * LB_M;

// Gram update + scalar two-loop recursion from the 25 freshly reduced inner products t[]
// (t may live in shared memory).  Whole CTA participates; thread 0 runs the recursion on a
// shared-memory copy of the 56-double Gram block (the recursion is ~100 dependent loads:
// from global memory it was a 30 us serial tail).
__device__ __forceinline__ void gram_recursion(const double *t, const HistPtrs &H, double *__restrict__ gram,
                                               float *__restrict__ coef, double *__restrict__ dscal)
{
    const int nh = H.bound - 1;
    __shared__ double gs[GR_DOUBLES];
    for (int i = threadIdx.x; i < GR_DOUBLES; i += blockDim.x) gs[i] = gram[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        const int e = H.slot[0];
        gs[GR_SY + e * LB_M + e] = t[0];
        gs[GR_YY + e * LB_M + e] = t[1];
        gs[GR_SG + e] = t[2];
        gs[GR_YG + e] = t[3];
        gs[GR_GG] = t[4];
        for (int h = 0; h < nh; ++h) {
            const int kk = H.slot[h + 1];
            gs[GR_SY + e * LB_M + kk] = t[5 + 5 * h + 0];
            gs[GR_SY + kk * LB_M + e] = t[5 + 5 * h + 1];
            gs[GR_YY + e * LB_M + kk] = t[5 + 5 * h + 2];
            gs[GR_YY + kk * LB_M + e] = t[5 + 5 * h + 2];
            gs[GR_SG + kk] = t[5 + 5 * h + 3];
            gs[GR_YG + kk] = t[5 + 5 * h + 4];
        }
        const int bound = H.bound;
        const double gamma = (double)((float)t[0] / (float)t[1]);          // ys / yy in float (:591)
        double alpha[LB_M], c[LB_M];
        for (int tt = 0; tt < bound; ++tt) {                                // newest -> oldest (:581-589)
            const int j = H.slot[tt];
            double num = -gs[GR_SG + j];
            for (int u = 0; u < tt; ++u) num -= alpha[u] * gs[GR_SY + j * LB_M + H.slot[u]];
            alpha[tt] = num / gs[GR_SY + j * LB_M + j];
        }
        for (int tt = bound - 1; tt >= 0; --tt) {                           // oldest -> newest (:593-601)
            const int j = H.slot[tt];
            double num = -gs[GR_YG + j];
            for (int u = 0; u < bound; ++u) num -= alpha[u] * gs[GR_YY + j * LB_M + H.slot[u]];
            num *= gamma;
            for (int v = bound - 1; v > tt; --v) num += c[v] * gs[GR_SY + H.slot[v] * LB_M + j];
            const double beta = num / gs[GR_SY + j * LB_M + j];
            c[tt] = alpha[tt] - beta;
        }
        double dg = -gamma * gs[GR_GG];
        coef[0] = (float)(-gamma);
        for (int tt = bound; tt < LB_M; ++tt) { coef[1 + tt] = 0.f; coef[1 + LB_M + tt] = 0.f; }
        for (int tt = 0; tt < bound; ++tt) {
            const int j = H.slot[tt];
            coef[1 + tt] = (float)(-gamma * alpha[tt]);
            coef[1 + LB_M + tt] = (float)c[tt];
            dg += -gamma * alpha[tt] * gs[GR_YG + j] + c[tt] * gs[GR_SG + j];
        }
        dscal[PS_DG] = dg;
        dscal[PS_DGINIT] = dg;      // survives the next evaluation, which overwrites PS_DG
    }
    __syncthreads();
    for (int i = threadIdx.x; i < GR_DOUBLES; i += blockDim.x) gram[i] = gs[i];
}

__global__ void __launch_bounds__(VEC_THREADS)
vec_update_gram_kernel(float *__restrict__ s_new, float *__restrict__ y_new, const float *__restrict__ x,
                       const float *__restrict__ xp, const float *__restrict__ g, const float *__restrict__ gp,
                       HistPtrs H, long n, double *__restrict__ partial, double *__restrict__ gram,
                       float *__restrict__ coef, double *__restrict__ dscal, unsigned *__restrict__ ticket,
                       double *__restrict__ gram_loc, XrCall xc)
{   // gram_loc != null (owner-sharded multi-GPU): the last CTA first sums this rank's 25 partial
    // inner products over all ranks through peer memory (rank order: same bits everywhere),
    // then every rank runs the same recursion
    __shared__ double red[VEC_THREADS / 32][GR_NPART];
    __shared__ bool last;
    double acc[GR_NPART];
#pragma unroll
    for (int k = 0; k < GR_NPART; ++k) acc[k] = 0.0;
    const int nh = H.bound - 1;                // stored pairs other than the new one

    // one element: the five products with g / the new pair, then five per stored pair
    auto elem = [&](float sv, float yv, float gv) {
        acc[0] += (double)yv * sv; acc[1] += (double)yv * yv; acc[2] += (double)sv * gv;
        acc[3] += (double)yv * gv; acc[4] += (double)gv * gv;
    };
    auto elem_h = [&](int h, float sv, float yv, float gv, float hs, float hy) {
        acc[5 + 5 * h + 0] += (double)sv * hy;      // s_new . y_h
        acc[5 + 5 * h + 1] += (double)hs * yv;      // s_h . y_new
        acc[5 + 5 * h + 2] += (double)yv * hy;      // y_new . y_h
        acc[5 + 5 * h + 3] += (double)hs * gv;      // s_h . g
        acc[5 + 5 * h + 4] += (double)hy * gv;      // y_h . g
    };
    VEC_LOOP4(n) {
        // issue every load of this iteration first (predicated, one basic block): with the
        // loads behind per-pair branches only ~6 of the 12 streams were in flight per thread
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 X = reinterpret_cast<const float4 *>(x)[v], XP = reinterpret_cast<const float4 *>(xp)[v];
        const float4 G = reinterpret_cast<const float4 *>(g)[v], GP = reinterpret_cast<const float4 *>(gp)[v];
        float4 HS[LB_M - 1], HY[LB_M - 1];
#pragma unroll
        for (int h = 0; h < LB_M - 1; ++h) {
            HS[h] = (h < nh) ? reinterpret_cast<const float4 *>(H.s[h + 1])[v] : z4;
            HY[h] = (h < nh) ? reinterpret_cast<const float4 *>(H.y[h + 1])[v] : z4;
        }
        const float4 S = make_float4(X.x - XP.x, X.y - XP.y, X.z - XP.z, X.w - XP.w);
        const float4 Y = make_float4(G.x - GP.x, G.y - GP.y, G.z - GP.z, G.w - GP.w);
        reinterpret_cast<float4 *>(s_new)[v] = S;
        reinterpret_cast<float4 *>(y_new)[v] = Y;
        // 4-element partial products in float (FMA), accumulated in float64: a quarter of the
        // FP64 work, and still far tighter than the reference's sequential float32 sums
        auto dot4 = [](const float4 &a, const float4 &b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; };
        acc[0] += (double)dot4(Y, S); acc[1] += (double)dot4(Y, Y); acc[2] += (double)dot4(S, G);
        acc[3] += (double)dot4(Y, G); acc[4] += (double)dot4(G, G);
#pragma unroll
        for (int h = 0; h < LB_M - 1; ++h) {       // pairs beyond nh contribute exact zeros
            acc[5 + 5 * h + 0] += (double)dot4(S, HY[h]);
            acc[5 + 5 * h + 1] += (double)dot4(HS[h], Y);
            acc[5 + 5 * h + 2] += (double)dot4(Y, HY[h]);
            acc[5 + 5 * h + 3] += (double)dot4(HS[h], G);
            acc[5 + 5 * h + 4] += (double)dot4(HY[h], G);
        }
    }
    VEC_TAIL(n, k) {
        const float sv = x[k] - xp[k], yv = g[k] - gp[k], gv = g[k];
        s_new[k] = sv; y_new[k] = yv;
        elem(sv, yv, gv);
#pragma unroll
        for (int h = 0; h < LB_M - 1; ++h)
            if (h < nh) elem_h(h, sv, yv, gv, H.s[h + 1][k], H.y[h + 1][k]);
    }

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < GR_NPART; ++k) {
        const double v = warp_sum(acc[k]);
        if (lane == 0) red[wid][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < GR_NPART) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < VEC_THREADS / 32; ++w) v += red[w][threadIdx.x];
        partial[(size_t)blockIdx.x * GR_NPART + threadIdx.x] = v;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicInc(ticket, gridDim.x - 1u) == gridDim.x - 1u);
    __syncthreads();
    if (!last) return;
    __threadfence();
    // last CTA: fixed-order reduction of the partials, 10 row groups x 25 sums in parallel ...
    {
        __shared__ double red2[10][GR_NPART];
        const int col = threadIdx.x % GR_NPART, grp = threadIdx.x / GR_NPART;
        if (grp < 10) {
            double v = 0.0;
            for (unsigned b = grp; b < gridDim.x; b += 10) v += partial[(size_t)b * GR_NPART + col];
            red2[grp][col] = v;
        }
        __syncthreads();
        if (threadIdx.x < GR_NPART) {
            double v = 0.0;
#pragma unroll
            for (int r = 0; r < 10; ++r) v += red2[r][threadIdx.x];
            red[0][threadIdx.x] = v;
        }
    }
    __syncthreads();
    if (gram_loc) {
        if (threadIdx.x < GR_NPART) gram_loc[threadIdx.x] = red[0][threadIdx.x];
        __syncthreads();
        xr_reduce_body(xc.P, xc.lay, xc.epoch, xc.seq, gram_loc, GR_NPART, red[0], xc.err);
    }
    gram_recursion(red[0], H, gram, coef, dscal);
}

// pass B: d = coef_g g + sum_t (coef_y[t] y_t + coef_s[t] s_t)
__global__ void __launch_bounds__(VEC_THREADS)
vec_direction_kernel(float *__restrict__ d, const float *__restrict__ g, HistPtrs H,
                     const float *__restrict__ coef, long n)
{
    float cy[LB_M], cs[LB_M];
    const float cg = coef[0];
#pragma unroll
    for (int t = 0; t < LB_M; ++t) { cy[t] = coef[1 + t]; cs[t] = coef[1 + LB_M + t]; }
    const int bound = H.bound;
    VEC_LOOP4(n) {
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 gv = reinterpret_cast<const float4 *>(g)[v];
        float4 yv[LB_M], sv[LB_M];
#pragma unroll
        for (int t = 0; t < LB_M; ++t) {           // all loads first, predicated
            yv[t] = (t < bound) ? reinterpret_cast<const float4 *>(H.y[t])[v] : z4;
            sv[t] = (t < bound) ? reinterpret_cast<const float4 *>(H.s[t])[v] : z4;
        }
        float4 o = make_float4(cg * gv.x, cg * gv.y, cg * gv.z, cg * gv.w);
#pragma unroll
        for (int t = 0; t < LB_M; ++t) {
            o.x += cy[t] * yv[t].x; o.y += cy[t] * yv[t].y; o.z += cy[t] * yv[t].z; o.w += cy[t] * yv[t].w;
        }
#pragma unroll
        for (int t = LB_M - 1; t >= 0; --t) {
            o.x += cs[t] * sv[t].x; o.y += cs[t] * sv[t].y; o.z += cs[t] * sv[t].z; o.w += cs[t] * sv[t].w;
        }
        reinterpret_cast<float4 *>(d)[v] = o;
    }
    VEC_TAIL(n, k) {
        float o = cg * g[k];
        for (int t = 0; t < bound; ++t) o += cy[t] * H.y[t][k];
        for (int t = bound - 1; t >= 0; --t) o += cs[t] * H.s[t][k];
        d[k] = o;
    }
}

// ---- reference-arithmetic mode ("fidelity"): the reference's own sequence of vector passes ----
// Inner products come from seqsum.cu (bit-exact sequential float32 sums) as floats in the
// device block fd[]; the element-wise passes below read the scalars they need from it, so
// the two-loop recursion (lbfgs.cpp:568-601) runs without a host round trip.
enum { FD_FX = 0, FD_DG = 1, FD_GG = 2, FD_XX = 3, FD_DOT = 4, FD_YS = 5, FD_ALPHA = 10, FD_YY = 15,
       FD_DGINIT = 16, FD_FLOATS = 20 };

// s = x - xp, y = g - gp (vecdiff, :548-549) and d = -g (vecncpy, :575)
__global__ void __launch_bounds__(VEC_THREADS)
fid_diff_kernel(float *__restrict__ s_new, float *__restrict__ y_new, float *__restrict__ d,
                const float *__restrict__ x, const float *__restrict__ xp, const float *__restrict__ g,
                const float *__restrict__ gp, long n)
{
    VEC_LOOP4(n) {
        const float4 X = reinterpret_cast<const float4 *>(x)[v], XP = reinterpret_cast<const float4 *>(xp)[v];
        const float4 G = reinterpret_cast<const float4 *>(g)[v], GP = reinterpret_cast<const float4 *>(gp)[v];
        reinterpret_cast<float4 *>(s_new)[v] = make_float4(X.x - XP.x, X.y - XP.y, X.z - XP.z, X.w - XP.w);
        reinterpret_cast<float4 *>(y_new)[v] = make_float4(G.x - GP.x, G.y - GP.y, G.z - GP.z, G.w - GP.w);
        reinterpret_cast<float4 *>(d)[v] = make_float4(-G.x, -G.y, -G.z, -G.w);
    }
    VEC_TAIL(n, k) { s_new[k] = x[k] - xp[k]; y_new[k] = g[k] - gp[k]; d[k] = -g[k]; }
}

// one vecadd / vecscale of the recursion with its scalar taken from fd[]:
//   mode 0: alpha_j = fd[DOT] / ys_j ; d += (-alpha_j) * u      (u = y_j, :585-588)
//   mode 1: d *= ys_new / yy                                      (:591)
//   mode 2: beta = fd[DOT] / ys_j ; d += (alpha_j - beta) * u    (u = s_j, :596-599)
__global__ void __launch_bounds__(VEC_THREADS)
fid_axpy_kernel(float *__restrict__ d, const float *__restrict__ u, float *__restrict__ fd, int mode, int j, long n)
{
    float c;
    if (mode == 0) {
        const float alpha = __fdiv_rn(fd[FD_DOT], fd[FD_YS + j]);
        // stored for mode 2; FD_ALPHA is not read by this launch and FD_DOT not written by it
        if (blockIdx.x == 0 && threadIdx.x == 0) fd[FD_ALPHA + j] = alpha;
        c = -alpha;
    } else if (mode == 1) {
        c = __fdiv_rn(fd[FD_YS + j], fd[FD_YY]);
    } else {
        c = __fsub_rn(fd[FD_ALPHA + j], __fdiv_rn(fd[FD_DOT], fd[FD_YS + j]));
    }
    VEC_LOOP4(n) {
        float4 o = reinterpret_cast<const float4 *>(d)[v];
        if (mode == 1) {
            o = make_float4(__fmul_rn(o.x, c), __fmul_rn(o.y, c), __fmul_rn(o.z, c), __fmul_rn(o.w, c));
        } else {
            const float4 a = reinterpret_cast<const float4 *>(u)[v];
            o.x = __fadd_rn(o.x, __fmul_rn(c, a.x)); o.y = __fadd_rn(o.y, __fmul_rn(c, a.y));
            o.z = __fadd_rn(o.z, __fmul_rn(c, a.z)); o.w = __fadd_rn(o.w, __fmul_rn(c, a.w));
        }
        reinterpret_cast<float4 *>(d)[v] = o;
    }
    VEC_TAIL(n, k) d[k] = (mode == 1) ? __fmul_rn(d[k], c) : __fadd_rn(d[k], __fmul_rn(c, u[k]));
}

// libLBFGS status codes (lbfgs.h:78-150; the enum counts up from -1024)
enum {
    ST_SUCCESS = 0, ST_STOP = 1, ST_ALREADY_MINIMIZED = 2,
    ST_UNKNOWN = -1024, ST_LOGIC = -1023, ST_OUTOFMEMORY = -1022, ST_CANCELED = -1021,
    ST_OUTOFINTERVAL = -1003, ST_INCORRECT_TMINMAX = -1002, ST_ROUNDING = -1001,
    ST_MINSTEP = -1000, ST_MAXSTEP = -999, ST_MAXLINESEARCH = -998, ST_MAXITER = -997,
    ST_WIDTHTOOSMALL = -996, ST_INVALIDPARAMS = -995, ST_INCREASEGRADIENT = -994
};

struct Params {
    float epsilon = 1e-3f;          // plmdcaBackend.cpp:70
    int max_linesearch = 5;         // :72
    float ftol = 1e-4f;             // :73
    float gtol = 0.9f;              // lbfgs.cpp:119 default
    float xtol = 1.0e-16f;
    float min_step = 1e-20f;
    float max_step = 1e20f;
};

inline float max3f(float a, float b, float c) { return std::max(std::max(a, b), c); }

// minimiser of the cubic through (u,fu,du),(v,fv,dv)            (lbfgs.cpp:1024-1038)
inline float cubic_min(float u, float fu, float du, float v, float fv, float dv)
{
    const float d = v - u;
    const float theta = (fu - fv) * 3 / d + du + dv;
    const float s = max3f(std::fabs(theta), std::fabs(du), std::fabs(dv));
    const float a = theta / s;
    float gamma = s * std::sqrt(a * a - (du / s) * (dv / s));
    if (v < u) gamma = -gamma;
    const float p = gamma - du + theta;
    const float q = gamma - du + gamma + dv;
    return u + (p / q) * d;
}

// variant used when the derivative keeps its sign and shrinks    (:1052-1072)
inline float cubic_min_bounded(float u, float fu, float du, float v, float fv, float dv,
                               float xmin, float xmax)
{
    const float d = v - u;
    const float theta = (fu - fv) * 3 / d + du + dv;
    const float s = max3f(std::fabs(theta), std::fabs(du), std::fabs(dv));
    const float a = theta / s;
    float gamma = s * std::sqrt(std::max(0.f, a * a - (du / s) * (dv / s)));
    if (u < v) gamma = -gamma;
    const float p = gamma - dv + theta;
    const float q = gamma - dv + gamma + du;
    const float r = p / q;
    if (r < 0.f && gamma != 0.f) return v - r * d;
    return (a < 0) ? xmax : xmin;
}

inline float quad_min(float u, float fu, float du, float v, float fv)
{   // :1083-1085
    const float a = v - u;
    return u + du / ((fu - fv) / a + du) / 2 * a;
}

inline float secant_min(float u, float du, float v, float dv)
{   // :1095-1097
    const float a = u - v;
    return v + dv / (dv - du) * a;
}

// More-Thuente safeguarded step + interval update               (:1128-1295)
int update_interval(float &x, float &fx, float &dx, float &y, float &fy, float &dy,
                    float &t, float &ft, float &dt, float tmin, float tmax, int &brackt)
{
    const bool dsign = (dt * (dx / std::fabs(dx))) < 0.;     // fsigndiff, non-IEEE branch (:64)
    if (brackt) {
        if (t <= std::min(x, y) || std::max(x, y) <= t) return ST_OUTOFINTERVAL;
        if (0. <= dx * (t - x)) return ST_INCREASEGRADIENT;
        if (tmax < tmin) return ST_INCORRECT_TMINMAX;
    }
    bool bound;
    float newt;
    if (fx < ft) {                       // higher value: minimum bracketed
        brackt = 1; bound = true;
        const float mc = cubic_min(x, fx, dx, t, ft, dt);
        const float mq = quad_min(x, fx, dx, t, ft);
        newt = (std::fabs(mc - x) < std::fabs(mq - x)) ? mc : mc + 0.5 * (mq - mc);
    } else if (dsign) {                  // lower value, derivative changed sign
        brackt = 1; bound = false;
        const float mc = cubic_min(x, fx, dx, t, ft, dt);
        const float mq = secant_min(x, dx, t, dt);
        newt = (std::fabs(mc - t) > std::fabs(mq - t)) ? mc : mq;
    } else if (std::fabs(dt) < std::fabs(dx)) {   // lower value, same sign, derivative shrinks
        bound = true;
        const float mc = cubic_min_bounded(x, fx, dx, t, ft, dt, tmin, tmax);
        const float mq = secant_min(x, dx, t, dt);
        if (brackt) newt = (std::fabs(t - mc) < std::fabs(t - mq)) ? mc : mq;
        else        newt = (std::fabs(t - mc) > std::fabs(t - mq)) ? mc : mq;
    } else {                             // lower value, same sign, derivative does not shrink
        bound = false;
        if (brackt) newt = cubic_min(t, ft, dt, y, fy, dy);
        else newt = (x < t) ? tmax : tmin;
    }
    if (fx < ft) { y = t; fy = ft; dy = dt; }
    else {
        if (dsign) { y = x; fy = fx; dy = dx; }
        x = t; fx = ft; dx = dt;
    }
    if (tmax < newt) newt = tmax;
    if (newt < tmin) newt = tmin;
    if (brackt && bound) {
        const float mq = x + 0.66 * (y - x);
        if (x < y) { if (mq < newt) newt = mq; }
        else       { if (newt < mq) newt = mq; }
    }
    t = newt;
    return 0;
}

struct Driver {
    PlmProblem *p;
    long n;
    cudaStream_t st;
    int blocks;
    float *x = nullptr, *xp = nullptr, *g = nullptr, *gp = nullptr, *d = nullptr;
    float *s[LB_M] = {nullptr}, *y[LB_M] = {nullptr};
    float *sc = nullptr, *coef = nullptr;
    double *partial = nullptr, *gram = nullptr, *gram_loc = nullptr;
    unsigned *ticket = nullptr;
    bool sharded = false;          // owner-sharded multi-GPU: n = local slice, scalars summed over peer memory
    allreduce_fn_t allreduce = nullptr;
    void *ctx = nullptr;
    int evals = 0;
    bool fidelity = false;         // reference arithmetic: sequential float32 sums (seqsum.cu)
    float *fd = nullptr, *h_fd = nullptr;
    SeqPipe *pipe = nullptr;

    int alloc()
    {
        if (fidelity) {
            if (!dev_alloc_t(&fd, (size_t)FD_FLOATS)) return -1;
            DCA_CUDA(cudaMemsetAsync(fd, 0, sizeof(float) * FD_FLOATS, st));
            h_fd = static_cast<float *>(host_pinned_alloc(sizeof(float) * FD_FLOATS));
            if (!h_fd) return -1;
            if (seqsum_prepare() != 0) return -1;
            pipe = seqsum_pipe_create(n);
        }
        float **v[] = {&xp, &g, &gp, &d};
        for (auto pp : v) if (!dev_alloc_t(pp, (size_t)n)) return -1;
        for (int i = 0; i < LB_M; ++i) {
            if (!dev_alloc_t(&s[i], (size_t)n) || !dev_alloc_t(&y[i], (size_t)n)) return -1;
        }
        if (!dev_alloc_t(&sc, (size_t)SC_FLOATS)) return -1;
        DCA_CUDA(cudaMemsetAsync(sc, 0, sizeof(float) * SC_FLOATS, st));
        if (!dev_alloc_t(&partial, (size_t)GR_NPART * blocks) || !dev_alloc_t(&ticket, (size_t)1) ||
            !dev_alloc_t(&gram, (size_t)GR_DOUBLES) || !dev_alloc_t(&coef, (size_t)CF_FLOATS + 3) ||
            !dev_alloc_t(&gram_loc, (size_t)XR_SLOT)) return -1;
        DCA_CUDA(cudaMemsetAsync(gram, 0, sizeof(double) * GR_DOUBLES, st));
        DCA_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned), st));
        return 0;
    }
    void release(float *keep_x)
    {
        float *v[] = {x, xp, g, gp, d};
        for (float *q : v) if (q && q != keep_x) dev_free(q);
        for (int i = 0; i < LB_M; ++i) { dev_free(s[i]); dev_free(y[i]); }
        dev_free(sc); dev_free(partial); dev_free(ticket); dev_free(gram); dev_free(coef); dev_free(gram_loc);
        dev_free(fd);
        if (h_fd) host_pinned_free(h_fd);
        seqsum_pipe_destroy(pipe);
    }

    // objective at x -> g ; returns {fx, g.d, g.g, x.x} on the host
    SeqJob dot_job(const float *a, const float *b, int out) const
    {
        SeqJob j{};
        j.x = a; j.y = b; j.n = n; j.kind = SEQ_DOT; j.out = out;
        return j;
    }

    // reference arithmetic: fx = sequential L2 sum + per-site terms in site order
    // (plmdca_numerics.cpp:463-486, :572), g.d / g.g / x.x as the reference's vecdot
    int evaluate_fidelity(const float *dir, double *h, double *dginit)
    {
        if (plm_eval_device(p, x, g, nullptr) != 0) return -1;
        SeqJob jobs[4];
        int nj = 0;
        {
            SeqJob r{};
            const double *fxsite; long nscan; int L, q; float lh, lJ;
            plm_fx_layout(p, &fxsite, &nscan, &L, &q, &lh, &lJ);
            r.x = x; r.y = nullptr; r.n = n; r.kind = SEQ_REG; r.n0 = (long)L * q; r.lam0 = lh; r.lam1 = lJ;
            r.out = FD_FX; r.fxsite = fxsite; r.nscan = nscan; r.L = L;
            jobs[nj++] = r;
        }
        if (dir) jobs[nj++] = dot_job(g, dir, FD_DG);
        jobs[nj++] = dot_job(g, g, FD_GG);
        jobs[nj++] = dot_job(x, x, FD_XX);
        if (seqsum_launch(jobs, nj, fd, st, pipe) != 0) return -1;
        DCA_CUDA(cudaMemcpyAsync(h_fd, fd, sizeof(float) * FD_FLOATS, cudaMemcpyDeviceToHost, st));
        DCA_CUDA(cudaStreamSynchronize(st));
        h[PS_FX] = h_fd[FD_FX]; h[PS_DG] = dir ? h_fd[FD_DG] : 0.0; h[PS_GG] = h_fd[FD_GG]; h[PS_XX] = h_fd[FD_XX];
        if (dginit) *dginit = h_fd[FD_DGINIT];
        return 0;
    }

    // history update + two-loop recursion exactly as lbfgs.cpp:546-601, nothing read back
    int direction_fidelity(int end, int bound)
    {
        DCA_LAUNCH(fid_diff_kernel, blocks, VEC_THREADS, 0, st, s[end], y[end], d, x, xp, g, gp, n);
        SeqJob two[2] = {dot_job(y[end], s[end], FD_YS + end), dot_job(y[end], y[end], FD_YY)};
        if (seqsum_launch(two, 2, fd, st, pipe) != 0) return -1;
        int j = (end + 1) % LB_M;
        for (int i = 0; i < bound; ++i) {
            j = (j + LB_M - 1) % LB_M;
            SeqJob a = dot_job(s[j], d, FD_DOT);
            if (seqsum_launch(&a, 1, fd, st, pipe) != 0) return -1;
            DCA_LAUNCH(fid_axpy_kernel, blocks, VEC_THREADS, 0, st, d, y[j], fd, 0, j, n);
        }
        DCA_LAUNCH(fid_axpy_kernel, blocks, VEC_THREADS, 0, st, d, (const float *)nullptr, fd, 1, end, n);
        for (int i = 0; i < bound; ++i) {
            SeqJob b = dot_job(y[j], d, FD_DOT);
            if (seqsum_launch(&b, 1, fd, st, pipe) != 0) return -1;
            DCA_LAUNCH(fid_axpy_kernel, blocks, VEC_THREADS, 0, st, d, s[j], fd, 2, j, n);
            j = (j + 1) % LB_M;
        }
        SeqJob dg = dot_job(g, d, FD_DGINIT);             // vecdot(&dginit, g, s, n), lbfgs.cpp:830
        return seqsum_launch(&dg, 1, fd, st, pipe);
    }

    int evaluate(const float *dir, double *h, double *dginit = nullptr)
    {
        ++evals;
        if (fidelity) return evaluate_fidelity(dir, h, dginit);
        if (!allreduce) {
            if (plm_eval_device(p, x, g, dir) != 0) return -1;
            return plm_read_scalars(p, h, dginit);
        }
        if (plm_eval_device(p, x, g, nullptr) != 0) return -1;
        // the callback enqueues the collective in stream order behind the evaluation on the
        // library's stream (dca_plm_cuda_stream): no host synchronisation here
        if (allreduce(ctx, g, n, plm_scalar_buffer(p), PS_COUNT) != 0) {
            set_error("lbfgs: all-reduce callback failed");
            return -1;
        }
        DCA_LAUNCH(vec_dots_kernel, blocks, VEC_THREADS, 0, st, g, dir, n, partial, sc, plm_scalar_buffer(p), ticket);
        return plm_read_scalars(p, h, dginit);
    }
};

}  // namespace

int lbfgs_preload_kernels()
{
    cudaFuncAttributes fa;
    DCA_CUDA(cudaFuncGetAttributes(&fa, vec_step_kernel));
    DCA_CUDA(cudaFuncGetAttributes(&fa, vec_twoloop_kernel));
    DCA_CUDA(cudaFuncGetAttributes(&fa, vec_dots_kernel));
    DCA_CUDA(cudaFuncGetAttributes(&fa, vec_update_gram_kernel));
    DCA_CUDA(cudaFuncGetAttributes(&fa, vec_direction_kernel));
    DCA_CUDA(cudaFuncGetAttributes(&fa, fid_diff_kernel));
    DCA_CUDA(cudaFuncGetAttributes(&fa, fid_axpy_kernel));
    return 0;
}

// Measurement hook: the two HBM-bound passes of one L-BFGS iteration (and the trial-step
// kernel) on n floats with `bound` stored pairs, CUDA events on the launching stream.
// ms_out[3] = {vec_update_gram, vec_direction, vec_step} averaged over reps.
int lbfgs_vector_bench(long n, int bound, int reps, float *ms_out)
{
    if (require_device() != 0) return -1;
    if (bound < 1 || bound > LB_M || n <= 0 || reps <= 0) { set_error("lbfgs_vector_bench: bad arguments"); return -1; }
    cudaStream_t st;
    DCA_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    float *v[5 + 2 * LB_M] = {nullptr};
    float *coef = nullptr;
    double *partial = nullptr, *gram = nullptr, *dscal = nullptr;
    unsigned *ticket = nullptr;
    const int blocks = (int)std::min<long>(ceil_div<long>(n, VEC_THREADS * 4), 8L * sm_count());
    bool ok = true;
    for (auto &q : v) ok = ok && dev_alloc_t(&q, (size_t)n);
    ok = ok && dev_alloc_t(&coef, (size_t)CF_FLOATS + 3) && dev_alloc_t(&partial, (size_t)GR_NPART * blocks) &&
         dev_alloc_t(&gram, (size_t)GR_DOUBLES) && dev_alloc_t(&dscal, (size_t)16) && dev_alloc_t(&ticket, (size_t)1);
    int rc = ok ? 0 : -1;
    if (ok) {
        for (auto q : v) cudaMemsetAsync(q, 0, sizeof(float) * (size_t)n, st);
        cudaMemsetAsync(coef, 0, sizeof(float) * (CF_FLOATS + 3), st);
        cudaMemsetAsync(gram, 0, sizeof(double) * GR_DOUBLES, st);
        cudaMemsetAsync(ticket, 0, sizeof(unsigned), st);
        HistPtrs H;
        H.bound = bound;
        for (int t = 0; t < LB_M; ++t) { H.slot[t] = t; H.s[t] = v[5 + t]; H.y[t] = v[5 + LB_M + t]; }
        cudaEvent_t e[4];
        for (auto &x : e) cudaEventCreate(&x);
        float acc[3] = {0.f, 0.f, 0.f};
        for (int r = -2; r < reps; ++r) {          // two warm-up rounds
            cudaEventRecord(e[0], st);
            DCA_LAUNCH(vec_update_gram_kernel, blocks, VEC_THREADS, 0, st, v[5], v[5 + LB_M], v[0], v[1], v[2], v[3], H, n,
                       partial, gram, coef, dscal, ticket, (double *)nullptr, XrCall{});
            cudaEventRecord(e[1], st);
            DCA_LAUNCH(vec_direction_kernel, blocks, VEC_THREADS, 0, st, v[4], v[2], H, coef, n);
            cudaEventRecord(e[2], st);
            DCA_LAUNCH(vec_step_kernel, blocks, VEC_THREADS, 0, st, v[0], v[1], v[4], 1.0f, n);
            cudaEventRecord(e[3], st);
            cudaStreamSynchronize(st);
            if (r >= 0)
                for (int k = 0; k < 3; ++k) { float ms = 0.f; cudaEventElapsedTime(&ms, e[k], e[k + 1]); acc[k] += ms; }
        }
        for (int k = 0; k < 3; ++k) ms_out[k] = acc[k] / reps;
        for (auto &x : e) cudaEventDestroy(x);
        if (cudaGetLastError() != cudaSuccess) { set_error("lbfgs_vector_bench: kernel failure"); rc = -1; }
    }
    cudaStreamSynchronize(st);
    for (auto q : v) dev_free(q);
    dev_free(coef); dev_free(partial); dev_free(gram); dev_free(dscal); dev_free(ticket);
    cudaStreamDestroy(st);
    return rc;
}

int lbfgs_minimize(PlmProblem *p, float *d_x, int max_iterations, int verbose,
                   allreduce_fn_t allreduce, void *allreduce_ctx, LbfgsResult *res,
                   int warmup_iterations, double *timed_ms, int *timed_evals)
{
    const Params prm;
    Driver D;
    D.p = p; D.n = plm_local_params(p); D.st = plm_stream(p);
    D.sharded = plm_is_sharded(p);
    D.fidelity = plm_fidelity(p) != 0;
    if (D.fidelity && (D.sharded || allreduce)) { set_error("lbfgs: the reference-arithmetic mode runs on one GPU"); return -1; }
    if (D.sharded && allreduce) { set_error("lbfgs: an owner-sharded objective exchanges over peer memory; no all-reduce callback"); return -1; }
    D.blocks = std::min<long>(ceil_div<long>(D.n, VEC_THREADS * 4), 8L * sm_count());
    D.x = d_x; D.allreduce = allreduce; D.ctx = allreduce_ctx;
    if (D.alloc() != 0) { D.release(d_x); return -1; }
    cudaStream_t st = D.st;
    const long n = D.n;
    double *dsc = plm_scalar_buffer(p);

    // sharded: start the clock on every rank at the same moment (the ranks reach this point
    // with a skew of their unequal H2D copies; without it the fastest rank's time includes it)
    if (plm_xr_barrier(p) != 0) { D.release(d_x); return -1; }
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evw = nullptr;
    cudaEventCreate(&ev0); cudaEventCreate(&ev1); cudaEventCreate(&evw);
    cudaEventRecord(ev0, st);
    bool warm_marked = (warmup_iterations <= 0);
    int evals_at_warm = 0;
    if (warm_marked) cudaEventRecord(evw, st);

    int ret = ST_UNKNOWN;
    int k = 1, end = 0, ls = 0;
    double h[PS_COUNT];
    float fx = 0.f, xnorm = 0.f, gnorm = 0.f, step = 0.f, dginit = 0.f;
    bool dginit_deferred = false;
    int rc = 0;

#define LB_CHECK(expr) do { if ((expr) != 0) { rc = -1; goto done; } } while (0)

    LB_CHECK(D.evaluate(nullptr, h));                                    // lbfgs.cpp:415
    fx = (float)h[PS_FX];
    DCA_LAUNCH(vec_twoloop_kernel, D.blocks, VEC_THREADS, 0, st, D.d, D.g, (const float *)nullptr, D.sc, 0, n,
               D.partial, -1, 0, dsc, D.ticket);
    xnorm = (float)std::sqrt(h[PS_XX]);
    gnorm = (float)std::sqrt(h[PS_GG]);
    if (xnorm < 1.0f) xnorm = 1.0f;
    if (gnorm / xnorm <= prm.epsilon) { ret = ST_ALREADY_MINIMIZED; goto done; }   // :450-454
    step = (float)(1.0 / (double)gnorm);                                 // :459 vec2norminv (d = -g): 1 / the float norm
    dginit = -(float)h[PS_GG];

    for (;;) {
        if (!warm_marked && k > warmup_iterations) {
            cudaEventRecord(evw, st); warm_marked = true; evals_at_warm = D.evals;
        }
        std::swap(D.x, D.xp);                                            // veccpy(xp,x) :465
        std::swap(D.g, D.gp);                                            // veccpy(gp,g) :466

        // ---- More-Thuente line search (:815-1004) on phi(t) = f(xp + t d) ----
        // dginit (g.d of the direction just built on the device) is not read back before the first
        // trial: nothing needs it until that trial's scalars are in, so it rides along in the same
        // 48-byte D2H -- one host round trip per iteration less.  Should it turn out positive
        // (:833-836 would have returned before evaluating) the trial is discarded and not counted.
        {
            int count = 0, brackt = 0, stage1 = 1, uinfo = 0;
            float dg = 0.f, f = fx;
            float stx = 0.f, fxb = fx, dgx = dginit;
            float sty = 0.f, fy = fx, dgy = dginit;
            const float finit = fx;
            float stmin, stmax;
            bool accepted = false;
            if (step <= 0.f) { ls = ST_INVALIDPARAMS; goto ls_done; }
            if (!dginit_deferred && 0 < dginit) { ls = ST_INCREASEGRADIENT; goto ls_done; }
            {
                float dgtest = prm.ftol * dginit;
                float width = prm.max_step - prm.min_step;
                float prev_width = 2.0 * width;
                for (;;) {
                    if (brackt) { stmin = std::min(stx, sty); stmax = std::max(stx, sty); }
                    else { stmin = stx; stmax = step + 4.0 * (step - stx); }
                    if (step < prm.min_step) step = prm.min_step;
                    if (prm.max_step < step) step = prm.max_step;
                    if ((brackt && ((step <= stmin || stmax <= step) || prm.max_linesearch <= count + 1 || uinfo != 0)) ||
                        (brackt && (stmax - stmin <= prm.xtol * stmax)))
                        step = stx;

                    DCA_LAUNCH(vec_step_kernel, D.blocks, VEC_THREADS, 0, st, D.x, D.xp, D.d, step, n);
                    if (dginit_deferred) {
                        double dgi = 0.0;
                        LB_CHECK(D.evaluate(D.d, h, &dgi));
                        dginit_deferred = false;
                        dginit = (float)dgi;
                        dgx = dginit; dgy = dginit;
                        dgtest = prm.ftol * dginit;
                        if (0 < dginit) { --D.evals; ls = ST_INCREASEGRADIENT; break; }
                    } else {
                        LB_CHECK(D.evaluate(D.d, h));
                    }
                    f = (float)h[PS_FX];
                    dg = (float)h[PS_DG];
                    const float ftest1 = finit + step * dgtest;
                    ++count;

                    if (brackt && ((step <= stmin || stmax <= step) || uinfo != 0)) { ls = ST_ROUNDING; accepted = true; break; }
                    if (step == prm.max_step && f <= ftest1 && dg <= dgtest) { ls = ST_MAXSTEP; accepted = true; break; }
                    if (step == prm.min_step && (ftest1 < f || dgtest <= dg)) { ls = ST_MINSTEP; accepted = true; break; }
                    if (brackt && (stmax - stmin) <= prm.xtol * stmax) { ls = ST_WIDTHTOOSMALL; accepted = true; break; }
                    if (prm.max_linesearch <= count) { ls = ST_MAXLINESEARCH; accepted = true; break; }
                    if (f <= ftest1 && std::fabs(dg) <= prm.gtol * (-dginit)) { ls = count; accepted = true; break; }

                    if (stage1 && f <= ftest1 && std::min(prm.ftol, prm.gtol) * dginit <= dg) stage1 = 0;
                    if (stage1 && ftest1 < f && f <= fxb) {
                        float fm = f - step * dgtest, fxm = fxb - stx * dgtest, fym = fy - sty * dgtest;
                        float dgm = dg - dgtest, dgxm = dgx - dgtest, dgym = dgy - dgtest;
                        uinfo = update_interval(stx, fxm, dgxm, sty, fym, dgym, step, fm, dgm, stmin, stmax, brackt);
                        fxb = fxm + stx * dgtest; fy = fym + sty * dgtest;
                        dgx = dgxm + dgtest; dgy = dgym + dgtest;
                    } else {
                        uinfo = update_interval(stx, fxb, dgx, sty, fy, dgy, step, f, dg, stmin, stmax, brackt);
                    }
                    if (brackt) {
                        if (0.66 * prev_width <= std::fabs(sty - stx)) step = stx + 0.5 * (sty - stx);
                        prev_width = width;
                        width = std::fabs(sty - stx);
                    }
                }
            }
            if (accepted) fx = f;        // every exit after an evaluation reports that trial's value (:622)
        ls_done:;
        }
        if (ls < 0) {                                                    // :478-484
            std::swap(D.x, D.xp);
            std::swap(D.g, D.gp);
            ret = ls;
            // fx reported by the reference here is the failed trial's value (:622)
            break;
        }
        xnorm = (float)std::sqrt(h[PS_XX]);                              // :487-492
        gnorm = (float)std::sqrt(h[PS_GG]);
        if (verbose) {                                                   // plmdcaBackend.cpp:137-146
            fprintf(stderr, "Iteration %d:\n", k);
            fprintf(stderr, "fx = %f, xnorm = %f, gnorm = %f, step = %f\n", fx, xnorm, gnorm, step);
            fprintf(stderr, "\n");
        }
        if (xnorm < 1.0f) xnorm = 1.0f;
        if (gnorm / xnorm <= prm.epsilon) { ret = ST_SUCCESS; break; }   // :506-511
        if (max_iterations != 0 && max_iterations < k + 1) { ret = ST_MAXITER; break; }   // :535-539

        // ---- history update + two-loop recursion, all on the device (:546-601) ----
        {
            const int bound = (LB_M <= k) ? LB_M : k;
            HistPtrs H;
            H.bound = bound;
            for (int t = 0; t < LB_M; ++t) {
                const int slot = (end + LB_M - t) % LB_M;      // t = 0: the pair written now
                H.slot[t] = slot; H.s[t] = D.s[slot]; H.y[t] = D.y[slot];
            }
            if (D.fidelity) {
                LB_CHECK(D.direction_fidelity(end, bound));
            } else {
            // pass A: s = x - xp, y = g - gp, the new inner products, Gram update and the scalar
            // recursion (alpha, beta, gamma -> coefficients, g.d) in the last CTA
            DCA_LAUNCH(vec_update_gram_kernel, D.blocks, VEC_THREADS, 0, st,
                       D.s[end], D.y[end], D.x, D.xp, D.g, D.gp, H, n, D.partial, D.gram, D.coef, dsc, D.ticket,
                       D.sharded ? D.gram_loc : (double *)nullptr, D.sharded ? plm_xr_next(p) : XrCall{});
            // pass B: the direction as one linear combination
            DCA_LAUNCH(vec_direction_kernel, D.blocks, VEC_THREADS, 0, st, D.d, D.g, H, D.coef, n);
            }
            ++k;
            end = (end + 1) % LB_M;
        }
        dginit_deferred = true;      // g.d of the new direction (:845) is read with the first trial's scalars
        dginit = 0.f;
        step = 1.0f;                                                     // :617
    }

done:
#undef LB_CHECK
    cudaEventRecord(ev1, st);
    cudaStreamSynchronize(st);
    if (D.fidelity && D.pipe && getenv("DCA_B200_TRACE")) {
        unsigned stt[8 * SEQ_MAX_JOBS] = {0};
        if (seqsum_pipe_stats(D.pipe, stt, st) == 0)
            for (int j = 0; j < SEQ_MAX_JOBS; ++j)
                fprintf(stderr, "[dca trace] seqsum job slot %d: %u chunks from records, %u exact, %u time-outs, %u stale, %u range\n",
                        j, stt[8 * j], stt[8 * j + 1], stt[8 * j + 2], stt[8 * j + 3], stt[8 * j + 4]);
    }
    if (!warm_marked) { cudaEventRecord(evw, st); cudaStreamSynchronize(st); evals_at_warm = D.evals; }
    float ms_total = 0.f, ms_timed = 0.f;
    cudaEventElapsedTime(&ms_total, ev0, ev1);
    cudaEventElapsedTime(&ms_timed, evw, ev1);
    cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaEventDestroy(evw);
    if (rc == 0 && D.x != d_x) {
        // the caller's buffer must hold the result: copy, then swap ownership back
        cudaMemcpyAsync(d_x, D.x, sizeof(float) * (size_t)n, cudaMemcpyDeviceToDevice, st);
        cudaStreamSynchronize(st);
        if (D.xp == d_x) D.xp = D.x;
        D.x = d_x;
    } else if (D.x != d_x) {
        if (D.xp == d_x) D.xp = D.x;
        D.x = d_x;
    }
    D.release(d_x);
    if (res) {
        res->status = ret; res->iterations = k; res->evaluations = D.evals;
        res->fx = fx; res->device_ms = ms_total;
    }
    if (timed_ms) *timed_ms = ms_timed;
    if (timed_evals) *timed_evals = D.evals - evals_at_warm;
    if (rc == 0 && verbose) {                                            // plmdcaBackend.cpp:82-91
        if (ret == ST_ROUNDING) fprintf(stderr, "L-BFGS optimization completed\n");
        else {
            fprintf(stderr, "L-BFGS optimization terminated with status code = %d\n", ret);
            fprintf(stderr, "fx = %f\n", fx);
        }
    }
    return rc;
}

}  // namespace dca
