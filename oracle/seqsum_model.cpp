// seqsum_model.cpp -- TEST INFRASTRUCTURE ONLY.
//
// CPU model of the block-parallel sequential-float32-sum kernel
// (pydca_b200/csrc/seqsum.cu) built on the SAME arithmetic header
// (pydca_b200/csrc/seqsum_core.h), next to the naive loop it must reproduce bit for bit:
//   vecdot    /root/reference/pydca/plmdca/lbfgs/include/arithmetic_ansi.h:114-121
//   L2 terms  /root/reference/pydca/plmdca/plmdca_numerics.cpp:463-486
// tests/test_seqsum_model.py compares the two on adversarial inputs (cancelling sums,
// binade crossings, exact ties, zeros, huge / tiny terms).  seq_pipe_model_* restates the
// decisions of the producer / consumer pipeline (seqsum_pipe_kernel) the same way.  Nothing under pydca_b200/
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

// ---- model of the producer / consumer pipeline (seqsum_pipe_kernel) ---------------------------------
// Same decisions as the kernel, taken sequentially: chunk records for the three exponents around a
// (deliberately stale) prediction, the chain warp's steps over up to 32 records with the plain integer
// scan where the summaries allow it, the row walk from the producers' row summaries, and the consumer
// CTA's exact path with per-row predicted exponents.  What a GPU adds -- timing -- cannot change the
// result: every step either applies a block whose range check holds or runs the plain chain.
constexpr int ROW = 128, ROWS = GROUP / ROW;

int exp_field_of(float s)
{
    const int ef = (int)((seq_bits(s) >> 23) & 0xffu);
    return (ef < 40 || ef > 214) ? 0 : ef;
}

struct Summ { SeqSum v; bool bad; };

Summ summarise(const float *t, int len, int ef)          // one block for the exponent field ef (0: none)
{
    Summ r;
    r.v = seq_single(0, 0);
    r.bad = !(ef >= 40 && ef <= 214);
    if (r.bad) return r;
    SeqCtx cx;
    cx.M = 0; cx.u = 0.f;
    cx.scale = seq_from_bits((uint32_t)(277 - ef) << 23);
    for (int k = 0; k < len; ++k) {
        int r0, r1;
        seq_round_term(cx, t[k], &r0, &r1, &r.bad);
        const SeqSum one = seq_single(r0, r1);
        r.v = (k == 0) ? one : seq_compose(r.v, one);
    }
    return r;
}

// how many leading summaries of v[from..) can be applied from state cc, and their total: the kernel's
// warp_scan_plain where every summary is parity-independent and small, seq_compose otherwise
int leading_ok(const SeqCtx &cc, const std::vector<Summ> &v, int from, int *total, long *plain_used)
{
    const int lim = 1 << 24, p = cc.M & 1, n = (int)v.size();
    bool plain = true;
    for (int k = from; k < n; ++k) {
        if (v[k].bad) continue;                          // (the kernel feeds zeros for unusable lanes)
        const SeqSum &q = v[k].v;
        plain = plain && q.total[0] == q.total[1] && q.pmin[0] == q.pmin[1] && q.pmax[0] == q.pmax[1] &&
                q.total[0] > -lim && q.total[0] < lim && q.pmin[0] > -lim && q.pmin[0] < lim && q.pmax[0] > -lim && q.pmax[0] < lim;
    }
    int nok = 0;
    *total = 0;
    if (plain) {
        if (plain_used) ++*plain_used;
        long long inc = 0, mn = (1 << 30), mx = -(1 << 30);
        for (int k = from; k < n; ++k) {
            if (v[k].bad) break;
            const long long exc = inc;
            inc += v[k].v.total[0];
            if (exc + v[k].v.pmin[0] < mn) mn = exc + v[k].v.pmin[0];
            if (exc + v[k].v.pmax[0] > mx) mx = exc + v[k].v.pmax[0];
            if (!seq_block_ok(cc, mn, mx)) break;
            ++nok; *total = (int)inc;
        }
    } else {
        SeqSum acc = seq_single(0, 0);
        for (int k = from; k < n; ++k) {
            if (v[k].bad) break;
            acc = (k == from) ? v[k].v : seq_compose(acc, v[k].v);
            if (!seq_block_ok(cc, acc.pmin[p], acc.pmax[p])) break;
            ++nok; *total = acc.total[p];
        }
    }
    return nok;
}

void serial(float &s, const float *t, int len) { for (int k = 0; k < len; ++k) s = s + t[k]; }

// rows of one chunk from running value s with row summaries made for the exponents erow[r] - 1 .. + 1
// (rows[r][k]); a row without a candidate for the current exponent, or that does not fit, takes the chain
void walk_rows_model(float &s, const float *t, int len, const std::vector<int> &erow,
                     const std::vector<std::vector<Summ>> &rows, long *plain_used)
{
    const int nrows = (len + ROW - 1) / ROW;
    int w = 0;
    while (w < nrows) {
        SeqCtx cc;
        if (seq_ctx_make(s, &cc)) {
            const int ef = exp_field_of(s);
            std::vector<Summ> v((size_t)nrows);
            for (int r = 0; r < nrows; ++r) {
                const int idx = ef - (erow[r] - 1);
                if (erow[r] != 0 && idx >= 0 && idx <= 2) v[r] = rows[r][idx];
                else { v[r].v = seq_single(0, 0); v[r].bad = true; }
            }
            int total;
            const int nok = leading_ok(cc, v, w, &total, plain_used);
            if (nok > 0) { s = seq_apply(cc, total); w += nok; if (w >= nrows) break; }
        }
        const int b0 = w * ROW, blen = (len - b0 < ROW) ? (len - b0) : ROW;
        serial(s, t + b0, blen);
        ++w;
    }
}

float pipe_model_sum(const float *t, long n, long *stats)
{
    const long nchunks = (n + GROUP - 1) / GROUP;
    long n_rec = 0, n_rowwalk = 0, n_exact = 0, n_plain = 0;
    // pre-pass: exponent predicted for the running value entering each chunk (float64 prefix)
    std::vector<int> epred((size_t)nchunks);
    {
        double pre = 0.0;
        for (long c = 0; c < nchunks; ++c) {
            epred[(size_t)c] = exp_field_of((float)pre);
            const long e = (c + 1) * GROUP < n ? (c + 1) * GROUP : n;
            for (long k = c * GROUP; k < e; ++k) pre += (double)t[k];
        }
    }
    auto chunk_len = [&](long c) { return (int)((n - c * GROUP < GROUP) ? (n - c * GROUP) : GROUP); };
    // producers: the prediction corrected by the drift seen LAG chunks earlier (hints are stale by design)
    const long LAG = 24;
    std::vector<int> e_mid((size_t)nchunks, 0);
    std::vector<int> seen_exp((size_t)nchunks + 1, 0);    // exponent field of the running value entering chunk c
    float s = 0.f;
    long c = 0;
    auto e_for = [&](long k) {
        int e = epred[(size_t)k];
        const long at = k - LAG;
        if (e != 0 && at > 0 && at <= c && seen_exp[(size_t)at] != 0 && epred[(size_t)at] != 0) e += seen_exp[(size_t)at] - epred[(size_t)at];
        return e;
    };
    auto exact_chunk = [&](long k) {                       // the consumer CTA's path
        const float *tc = t + k * GROUP;
        const int len = chunk_len(k), nrows = (len + ROW - 1) / ROW;
        std::vector<int> erow((size_t)nrows);
        std::vector<std::vector<Summ>> rows((size_t)nrows, std::vector<Summ>(3));
        double pre = 0.0;
        for (int r = 0; r < nrows; ++r) {
            const int blen = (len - r * ROW < ROW) ? (len - r * ROW) : ROW;
            erow[(size_t)r] = exp_field_of((float)((double)s + pre));
            for (int q = 0; q < 3; ++q) rows[(size_t)r][(size_t)q] = summarise(tc + r * ROW, blen, erow[(size_t)r] ? erow[(size_t)r] - 1 + q : 0);
            for (int k2 = 0; k2 < blen; ++k2) pre += (double)tc[r * ROW + k2];
        }
        walk_rows_model(s, tc, len, erow, rows, &n_plain);
        ++n_exact;
    };
    while (c < nchunks) {
        seen_exp[(size_t)c] = exp_field_of(s);
        SeqCtx cc;
        if (!seq_ctx_make(s, &cc)) { exact_chunk(c); ++c; continue; }
        const int ef = exp_field_of(s);
        const int nb = (int)((nchunks - c < 32) ? (nchunks - c) : 32);
        std::vector<Summ> v((size_t)nb);
        for (int b = 0; b < nb; ++b) {
            const int e = e_for(c + b);
            e_mid[(size_t)(c + b)] = e;
            const int idx = ef - (e - 1);
            if (e != 0 && idx >= 0 && idx <= 2) {
                v[(size_t)b] = summarise(t + (c + b) * GROUP, chunk_len(c + b), ef);
                // the record is marked unusable when a row moves the integer by 2^24 or more
                const int len = chunk_len(c + b);
                for (int r0 = 0; r0 < len && !v[(size_t)b].bad; r0 += ROW) {
                    const Summ row = summarise(t + (c + b) * GROUP + r0, (len - r0 < ROW) ? (len - r0) : ROW, ef);
                    const int lim = 1 << 24;
                    if (row.bad || row.v.total[0] >= lim || row.v.total[0] <= -lim || row.v.total[1] >= lim || row.v.total[1] <= -lim ||
                        row.v.pmax[0] >= lim || row.v.pmax[1] >= lim || row.v.pmin[0] <= -lim || row.v.pmin[1] <= -lim) v[(size_t)b].bad = true;
                }
            } else { v[(size_t)b].v = seq_single(0, 0); v[(size_t)b].bad = true; }
        }
        int total;
        const int nok = leading_ok(cc, v, 0, &total, &n_plain);
        if (nok > 0) {
            s = seq_apply(cc, total);
            for (int b = 1; b <= nok; ++b) seen_exp[(size_t)(c + b)] = 0;      // (not observed at these boundaries)
            c += nok; n_rec += nok;
        }
        if (nok == nb) continue;
        // chunk c: row walk from the producers' row summaries (candidates e_mid - 1 .. + 1 for every row)
        seen_exp[(size_t)c] = exp_field_of(s);
        const int e = e_mid[(size_t)c];
        const int idx0 = exp_field_of(s) - (e - 1);
        if (e != 0 && idx0 >= 0 && idx0 <= 2 && seq_ctx_make(s, &cc)) {
            const float *tc = t + c * GROUP;
            const int len = chunk_len(c), nrows = (len + ROW - 1) / ROW;
            std::vector<int> erow((size_t)nrows, e);
            std::vector<std::vector<Summ>> rows((size_t)nrows, std::vector<Summ>(3));
            for (int r = 0; r < nrows; ++r)
                for (int q = 0; q < 3; ++q)
                    rows[(size_t)r][(size_t)q] = summarise(tc + r * ROW, (len - r * ROW < ROW) ? (len - r * ROW) : ROW, e - 1 + q);
            walk_rows_model(s, tc, len, erow, rows, &n_plain);
            ++n_rowwalk;
        } else exact_chunk(c);
        ++c;
    }
    if (stats) { stats[0] = n_rec; stats[1] = n_rowwalk; stats[2] = n_exact; stats[3] = n_plain; }
    return s;
}

}  // namespace

extern "C" {

// the pipeline's decisions, sequentially (stats: chunks from records, row walks, exact chunks, plain scans)
float seq_pipe_model_sum(const float *t, long n, long *stats) { return pipe_model_sum(t, n, stats); }
float seq_pipe_model_dot(const float *x, const float *y, long n, long *stats)
{
    std::vector<float> t((size_t)n);
    for (long i = 0; i < n; ++i) t[(size_t)i] = x[i] * y[i];
    return pipe_model_sum(t.data(), n, stats);
}


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
