// seqsum_model.cpp -- TEST INFRASTRUCTURE ONLY.
//
// CPU model of the block-parallel sequential-float32-sum kernel
// (pydca_b200/csrc/seqsum.cu) built on the SAME arithmetic header
// (pydca_b200/csrc/seqsum_core.h), next to the naive loop it must reproduce bit for bit:
//   vecdot    /root/reference/pydca/plmdca/lbfgs/include/arithmetic_ansi.h:114-121
//   L2 terms  /root/reference/pydca/plmdca/plmdca_numerics.cpp:463-486
// tests/test_seqsum_model.py compares the two on adversarial inputs (cancelling sums,
// binade crossings, exact ties, zeros, huge / tiny terms).  Nothing under pydca_b200/
// links or loads this file.
//
// Build: g++ -O2 -ffp-contract=off -fPIC -shared (oracle/Makefile, target `seqmodel`).
#include <cstdint>
#include <cstring>
#include <vector>

#include "../pydca_b200/csrc/seqsum_core.h"

namespace {

constexpr int GROUP = 4096;   // terms per CTA step  (SQ_GROUP in seqsum.cu)
constexpr int SUB = 128;      // terms per warp step (32 lanes x 4)

// one block of `len` terms tried in the integer picture; true = applied
bool try_block(float &s, const float *t, int len)
{
    SeqCtx c;
    if (!seq_ctx_make(s, &c)) return false;
    bool bad = false;
    SeqSum acc = seq_single(0, 0);
    // the kernel composes per-thread (4 terms) -> warp -> CTA summaries; composition is
    // associative, so one left fold over single-term summaries models all of them
    for (int k = 0; k < len; ++k) {
        int r0, r1;
        seq_round_term(c, t[k], &r0, &r1, &bad);
        const SeqSum one = seq_single(r0, r1);
        acc = (k == 0) ? one : seq_compose(acc, one);
    }
    const int p = c.M & 1;
    if (bad || !seq_block_ok(c, acc.pmin[p], acc.pmax[p])) return false;
    s = seq_apply(c, acc.total[p]);
    return true;
}

float model_sum(const float *t, long n, long *stats)
{
    float s = 0.f;
    long g_ok = 0, w_ok = 0, serial = 0;
    for (long g0 = 0; g0 < n; g0 += GROUP) {
        const int glen = (int)((n - g0 < GROUP) ? (n - g0) : GROUP);
        if (try_block(s, t + g0, glen)) { ++g_ok; continue; }
        for (int b0 = 0; b0 < glen; b0 += SUB) {
            const int blen = (glen - b0 < SUB) ? (glen - b0) : SUB;
            if (try_block(s, t + g0 + b0, blen)) { ++w_ok; continue; }
            ++serial;
            for (int k = 0; k < blen; ++k) s = s + t[g0 + b0 + k];
        }
    }
    if (stats) { stats[0] = g_ok; stats[1] = w_ok; stats[2] = serial; }
    return s;
}

}  // namespace

extern "C" {

// the reference's vecdot, verbatim semantics (float product, float running sum)
float seq_naive_dot(const float *x, const float *y, long n)
{
    volatile float s = 0.f;
    for (long i = 0; i < n; ++i) { const float p = x[i] * y[i]; s = s + p; }
    return s;
}

// the reference's L2 accumulation: fx += lambda * v * v, lambda0 for the first n0 entries
float seq_naive_reg(const float *x, long n, long n0, float lam0, float lam1)
{
    volatile float s = 0.f;
    for (long i = 0; i < n; ++i) {
        const float lam = (i < n0) ? lam0 : lam1;
        const float a = lam * x[i];
        const float p = a * x[i];
        s = s + p;
    }
    return s;
}

float seq_model_dot(const float *x, const float *y, long n, long *stats)
{
    std::vector<float> t((size_t)n);
    for (long i = 0; i < n; ++i) t[(size_t)i] = x[i] * y[i];
    return model_sum(t.data(), n, stats);
}

float seq_model_reg(const float *x, long n, long n0, float lam0, float lam1, long *stats)
{
    std::vector<float> t((size_t)n);
    for (long i = 0; i < n; ++i) { const float a = ((i < n0) ? lam0 : lam1) * x[i]; t[(size_t)i] = a * x[i]; }
    return model_sum(t.data(), n, stats);
}

// plain sums of given terms (tests craft the terms directly: ties, crossings)
float seq_naive_sum(const float *t, long n)
{
    volatile float s = 0.f;
    for (long i = 0; i < n; ++i) s = s + t[i];
    return s;
}
float seq_model_sum(const float *t, long n, long *stats) { return model_sum(t, n, stats); }

}  // extern "C"
