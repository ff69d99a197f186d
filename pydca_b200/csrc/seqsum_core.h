// seqsum_core.h -- the arithmetic behind the EXACT block-parallel evaluation of a
// sequential float32 sum
//        s = 0;  for k in 0..n-1:  s = fl32(s + t_k)
// which is what the reference's optimiser and objective do over all P parameters:
//   vecdot     pydca/plmdca/lbfgs/include/arithmetic_ansi.h:114-121   (t_k = fl32(x_k * y_k))
//   L2 terms   pydca/plmdca/plmdca_numerics.cpp:463-486               (t_k = fl32(fl32(lambda * x_k) * x_k))
// At P = 4.9e6 (config 2) such a sum is far from the true sum: once the running value is
// 2^24 times larger than a term, the term is rounded away ("stagnation"); the reference's
// |g|^2 is 1.8 % low at the first iterate.  That bias is deterministic and steers the
// whole L-BFGS trajectory, so reproducing the reference's fit needs
// the same sums, bit for bit.
//
// The trick.  While the running value stays inside one binade [2^e, 2^(e+1)), every
// float32 there is an integer multiple of u = 2^(e-23): s = M*u with 2^23 <= |M| < 2^24.
// Adding t gives the float nearest to (M + t/u)*u, i.e. (M + rint(t/u))*u, as long as that
// result is still inside the closed binade and with exact ties (t/u = x.5) resolved by M's parity.  Inside a
// binade the recursion is therefore INTEGER addition of r_k = rint(t_k/u) -- associative,
// so a block of terms can be rounded and prefix-summed in parallel, and applied in one
// step if every prefix state M + r_0 + ... + r_k stays in [2^23 + 2, 2^24 - 2].  Blocks
// that fail the test (binade crossing, a term too large, s = 0 / subnormal) are
// re-done at a finer granularity and finally by the plain serial chain, so the result is
// exact in all cases.
//
// This header is plain C++ usable from host and device code; seqsum.cu builds the kernel
// on it and oracle/seqsum_model.cpp (test infrastructure) the CPU model that pins the
// claim above against the naive loop.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define SEQ_HD __host__ __device__ __forceinline__
#else
#define SEQ_HD static inline
#endif

struct SeqCtx {
    float scale;   // 2^(23-e): t * scale = t / u, exact
    float u;       // 2^(e-23)
    int M;         // s / u, signed, 2^23 <= |M| < 2^24
};

constexpr int SEQ_M_LO = (1 << 23) + 2;     // prefix states must stay inside [LO, HI] (in |M|)
constexpr int SEQ_M_HI = (1 << 24) - 2;
constexpr int SEQ_R_MAX = 1 << 22;          // |t/u| beyond this: the block is handed down
constexpr int SEQ_SAT = 1 << 26;            // saturation of composed block summaries

SEQ_HD uint32_t seq_bits(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t b; memcpy(&b, &f, 4); return b;
#endif
}
SEQ_HD float seq_from_bits(uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(b);
#else
    float f; memcpy(&f, &b, 4); return f;
#endif
}

// binade context of the running value; false when the integer picture does not apply
// (zero, subnormal, inf/nan, or an exponent so extreme that 2^(23-e) is not a normal float)
SEQ_HD bool seq_ctx_make(float s, SeqCtx *c)
{
    const uint32_t b = seq_bits(s);
    const int ef = (int)((b >> 23) & 0xffu);          // biased exponent
    if (ef < 40 || ef > 214) return false;            // |s| in [2^-87, 2^88): scale and u are normal floats
    const int e = ef - 127;
    c->scale = seq_from_bits((uint32_t)(127 + 23 - e) << 23);
    c->u = seq_from_bits((uint32_t)(127 + e - 23) << 23);
    const int mant = (int)((b & 0x7fffffu) | 0x800000u);
    c->M = (b >> 31) ? -mant : mant;
    return true;
}

// Rounded term for both parities of the running integer M.  r0 = rint(t/u) (half to even) is
// the increment when M is even.  On an exact tie (t/u = m + 1/2) the adder rounds M + t/u to
// the even neighbour, so an ODD M takes the other neighbour of t/u: r1 = 2*floor(t/u) + 1 - r0.
// Ties are common when the terms are much larger than u (1 in ~2^9 products), so they are
// part of the integer picture, not a fallback.  bad = too large for it, or nan.
SEQ_HD void seq_round_term(const SeqCtx &c, float t, int *r0, int *r1, bool *bad)
{
    const float z = t * c.scale;                      // exact (power of two), may overflow to inf
    if (!(fabsf(z) < (float)SEQ_R_MAX)) { *bad = true; *r0 = 0; *r1 = 0; return; }
    const float r = rintf(z);                         // round half to even, like the adder
    const float d = z - r;                            // exact
    const int ri = (int)r;
    *r0 = ri;
    // tie: z = m + 1/2; rintf took the even neighbour, the other one lies on d's side
    *r1 = (fabsf(d) == 0.5f) ? (d > 0.f ? ri + 1 : ri - 1) : ri;
}

// can a block whose inclusive prefix sums span [pmin, pmax] be applied in one step?
SEQ_HD bool seq_block_ok(const SeqCtx &c, long long pmin, long long pmax)
{
    if (c.M > 0) return c.M + pmin >= SEQ_M_LO && c.M + pmax <= SEQ_M_HI;
    return c.M + pmax <= -SEQ_M_LO && c.M + pmin >= -SEQ_M_HI;
}

SEQ_HD float seq_apply(const SeqCtx &c, long long total)
{
    return (float)(c.M + (int)total) * c.u;           // |M + total| < 2^24: both steps exact
}

SEQ_HD int seq_sat(long long v)
{
    return v > SEQ_SAT ? SEQ_SAT : (v < -SEQ_SAT ? -SEQ_SAT : (int)v);
}

// a + b clamped to +/-SEQ_SAT, for |a|, |b| <= SEQ_SAT (no 32-bit overflow: 2^27 < 2^31)
SEQ_HD int seq_add_sat(int a, int b)
{
    const int v = a + b;
    return v > SEQ_SAT ? SEQ_SAT : (v < -SEQ_SAT ? -SEQ_SAT : v);
}

// Summary of a block of terms for a fixed context, as a function of the parity p of the
// integer state entering the block: total increment and the range of inclusive prefixes.
struct SeqSum { int total[2], pmin[2], pmax[2]; };

SEQ_HD SeqSum seq_single(int r0, int r1)
{
    SeqSum o;
    o.total[0] = o.pmin[0] = o.pmax[0] = r0;
    o.total[1] = o.pmin[1] = o.pmax[1] = r1;
    return o;
}

// summary of (a followed by b); every field stays within +/-SEQ_SAT (single terms are below
// SEQ_R_MAX), saturating, which can only turn "ok" into "not ok"
// (a saturated total keeps |pmin| or |pmax| >= SEQ_SAT > 2^24; its parity no longer matters)
SEQ_HD SeqSum seq_compose(const SeqSum &a, const SeqSum &b)
{
    SeqSum o;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int p = 0; p < 2; ++p) {
        const int pa = (p + a.total[p]) & 1;          // parity of the state entering b
        o.total[p] = seq_add_sat(a.total[p], b.total[pa]);
        const int bmin = seq_add_sat(a.total[p], b.pmin[pa]);
        const int bmax = seq_add_sat(a.total[p], b.pmax[pa]);
        o.pmin[p] = a.pmin[p] < bmin ? a.pmin[p] : bmin;
        o.pmax[p] = a.pmax[p] > bmax ? a.pmax[p] : bmax;
    }
    return o;
}
