// mf.cu -- mean-field DCA operators on sm_100a.
//
// Replaces (pydca/meanfield_dca/msa_numerics.py):
//   compute_single_site_freqs :54-89     get_reg_single_site_freqs :93-125
//   compute_pair_site_freqs   :183-229   get_reg_pair_site_freqs   :232-267
//   construct_corr_mat        :271-318   compute_couplings         :322-342
//
// 1. pair frequencies.  F[(j,b)][(i,a)] = sum_n w_n [s_nj==b][s_ni==a] in
//    float64 as weighted histograms over per-site residue buckets (see
//    mf_pairhist_kernel): N*L*L/32 warp steps instead of the N*L*L*q adds of a
//    dense one-hot product; no atomics, sequential summation order.  The
//    diagonal of F holds the single-site counts (F[(i,a)][(i,a)] = sum_n w_n [s_ni==a]).
// 2. pseudocount + covariance assembly fused in one pass over F.
// 3. couplings = -C^{-1}: blocked Gauss-Jordan inversion without pivoting (C is
//    symmetric positive definite by construction when pseudocount > 0), block
//    size 128; all O(n^3) work is rank-128 updates done by a float64
//    tensor-core GEMM kernel (mma.sync.m8n8k4.f64, "DMMA"; tcgen05 has no
//    FP64).  The reference uses LAPACK getrf/getri through numba; results
//    agree to O(cond * eps).
#include "dca_common.cuh"
#include "dca_internal.h"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

namespace dca {

constexpr int MF_TC = 64;          // F row padding (columns)
constexpr int MF_HW = 8;           // warps per CTA of the histogram kernel

// Per site: the row indices of the alignment grouped by residue, ascending n inside a group
// (stable counting sort), plus the q+1 group offsets.  One warp per site.
template <int Q>
__global__ void __launch_bounds__(128)
mf_site_buckets_kernel(const uint8_t *__restrict__ msa, long N, int L, unsigned *__restrict__ rows,
                       unsigned *__restrict__ off, int site_lo, int site_hi)
{   // sites [site_lo, site_hi) only (multi-GPU: every GPU buckets the sites it histograms)
    __shared__ unsigned cnt[4][Q + 1];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int i = site_lo + blockIdx.x * 4 + wib;
    if (i >= site_hi) return;
    if (lane <= Q) cnt[wib][lane] = 0;
    __syncwarp();
    // one warp walks N rows with a stride of L bytes: the loop is bound by the latency of its loads (two warps
    // per SM), so eight of them are kept in flight per lane
    for (long n0 = lane; n0 < N; n0 += 32 * 8) {
        unsigned r8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const long n = n0 + 32 * u; r8[u] = (n < N) ? msa[n * L + i] : (unsigned)Q; }
#pragma unroll
        for (int u = 0; u < 8; ++u) if (r8[u] < (unsigned)Q) atomicAdd(&cnt[wib][r8[u]], 1u);
    }
    __syncwarp();
    if (lane == 0) {
        unsigned run = 0;
        for (int b = 0; b < Q; ++b) { const unsigned c = cnt[wib][b]; cnt[wib][b] = run; off[(size_t)i * (Q + 1) + b] = run; run += c; }
        off[(size_t)i * (Q + 1) + Q] = run;
    }
    __syncwarp();
    const unsigned lt = (1u << lane) - 1u;
    for (long nb = 0; nb < N; nb += 32 * 8) {
        unsigned r8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const long n = nb + 32 * u + lane; r8[u] = (n < N) ? msa[n * L + i] : (unsigned)Q; }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const long n = nb + 32 * u + lane;
            if (nb + 32 * u >= N) break;                                     // warp-uniform
            const bool valid = n < N;
            const unsigned r = r8[u];
            const unsigned peers = __match_any_sync(0xffffffffu, r);
            const unsigned base = cnt[wib][r];
            __syncwarp();
            if (valid) {
                rows[(size_t)i * N + base + __popc(peers & lt)] = (unsigned)n;
                if ((peers & lt) == 0u) cnt[wib][r] = base + __popc(peers);     // lowest lane of the group
            }
            __syncwarp();
        }
    }
}

// Pair counts as weighted histograms.  F[(j,b)][(i,a)] = sum_n w_n [s_ni = a][s_nj = b] has only
// N non-zero terms per (i,j) -- one per sequence -- so instead of multiplying dense one-hot
// rows (N*L*L*q adds: 3.8e10 at N = 20000, L = 300; the first version, 18 ms) a warp takes
// (site i, residue a, 32 consecutive sites j), walks the rows of bucket (i,a) in ascending n and
// adds w_n into bins[s_nj][lane]: N*L*L/32 warp steps (6e7), every step one coalesced 32-byte
// read of the MSA row.  The bins live in shared memory ([q][32] doubles per warp, lane-interleaved:
// conflict-free for any residues); each bin is owned by one lane: no atomics, and the order of the
// additions is the sequential order of the reference's loop (msa_numerics.py:54-89, :183-229).
// Only tiles holding some j >= i are computed (the consumers read the canonical i <= j entries).
template <int Q>
__global__ void __launch_bounds__(MF_HW * 32)
mf_pairhist_kernel(const uint8_t *__restrict__ msa, const double *__restrict__ w,
                   const unsigned *__restrict__ rows, const unsigned *__restrict__ off,
                   long N, int L, int ld, double *__restrict__ F, int site_lo)
{   // F may live on ANOTHER GPU (multi-GPU mfDCA: every GPU stores the columns of its sites
    // straight into the solver GPU's matrix over NVLink)
    __shared__ double bins[MF_HW][Q][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int i = site_lo + blockIdx.y;
    const int njt = (L + 31) / 32;
    const int item = blockIdx.x * MF_HW + wib;
    if (item >= Q * njt) return;
    const int a = item / njt, jt = item - a * njt;
    const int j0 = jt * 32;
    if (j0 + 31 < i) return;
    const int j = j0 + lane;
    const bool jok = j < L;
#pragma unroll
    for (int b = 0; b < Q; ++b) bins[wib][b][lane] = 0.0;
    const unsigned beg = off[(size_t)i * (Q + 1) + a], end = off[(size_t)i * (Q + 1) + a + 1];
    const unsigned *ri = rows + (size_t)i * N;
    for (unsigned k = beg; k < end; k += 32) {
        const unsigned cntk = min(32u, end - k);
        unsigned nk = 0;
        double wk = 0.0;
        if (lane < cntk) { nk = ri[k + lane]; wk = w[nk]; }
        unsigned t = 0;
        for (; t + 4 <= cntk; t += 4) {            // four MSA reads in flight, then the four updates
            unsigned n4[4], b4[4];
            double w4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                n4[u] = __shfl_sync(0xffffffffu, nk, t + u);
                w4[u] = __shfl_sync(0xffffffffu, wk, t + u);
                b4[u] = jok ? msa[(size_t)n4[u] * L + j] : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) bins[wib][b4[u]][lane] += w4[u];
        }
        for (; t < cntk; ++t) {
            const unsigned n = __shfl_sync(0xffffffffu, nk, t);
            const double wv = __shfl_sync(0xffffffffu, wk, t);
            const unsigned b = jok ? msa[(size_t)n * L + j] : 0u;
            bins[wib][b][lane] += wv;
        }
    }
    if (jok) {
        double *f = F + ((size_t)j * Q) * ld + (size_t)i * Q + a;
#pragma unroll
        for (int b = 0; b < Q; ++b) f[(size_t)b * ld] = bins[wib][b][lane];
    }
}

// fi[i][a] = F[(i,a)][(i,a)] / Meff  (all q states, gap included), regularised copy
__global__ void mf_single_kernel(const double *__restrict__ F, size_t f_split_stride, int nsplit,
                                 int L, int q, int ld, double meff, double pseudocount,
                                 double *__restrict__ fi, double *__restrict__ reg_fi)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= L * q) return;
    double s = 0.0;
    for (int z = 0; z < nsplit; ++z) s += F[(size_t)z * f_split_stride + (size_t)idx * ld + idx];
    const double f = s / meff;
    if (fi) fi[idx] = f;
    if (reg_fi) reg_fi[idx] = pseudocount / (double)q + (1.0 - pseudocount) * f;     // msa_numerics.py:120-124
}

// fij[pair][a][b] for non-gap a,b (msa_numerics.py:183-229 layout)
__global__ void mf_pack_fij_kernel(const double *__restrict__ F, size_t f_split_stride, int nsplit,
                                   int L, int q, int ld, double meff, double *__restrict__ fij)
{
    const int qm1 = q - 1;
    const long total = (long)L * (L - 1) / 2 * qm1 * qm1;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const long pair = idx / (qm1 * qm1);
    const int ab = (int)(idx - pair * qm1 * qm1);
    const int a = ab / qm1, b = ab - a * qm1;
    int i = 0; long rem = pair;
    while (rem >= L - 1 - i) { rem -= L - 1 - i; ++i; }
    const int j = i + 1 + (int)rem;
    double s = 0.0;
    for (int z = 0; z < nsplit; ++z)
        s += F[(size_t)z * f_split_stride + ((size_t)j * q + b) * ld + (size_t)i * q + a];
    fij[idx] = s / meff;
}

// covariance matrix (msa_numerics.py:271-318) with the pair pseudocount
// (:232-267) fused; C is npad x npad, identity on the padding diagonal.
__global__ void mf_corr_kernel(const double *__restrict__ F, size_t f_split_stride, int nsplit,
                               const double *__restrict__ reg_fi, int L, int q, int ld, double meff,
                               double pseudocount, int n, int npad, double *__restrict__ C)
{
    const int qm1 = q - 1;
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = blockIdx.y;
    if (col >= npad) return;
    double out;
    if (row >= n || col >= n) out = (row == col) ? 1.0 : 0.0;
    else {
        int i = row / qm1, a = row - i * qm1, j = col / qm1, b = col - j * qm1;
        if (i == j) {
            const double fia = reg_fi[i * q + a], fib = reg_fi[i * q + b];
            out = (a == b) ? fia * (1.0 - fia) : -1.0 * fia * fib;
        } else {
            if (i > j) { int t = i; i = j; j = t; t = a; a = b; b = t; }   // canonical i<j source: exact symmetry
            double s = 0.0;
            for (int z = 0; z < nsplit; ++z)
                s += F[(size_t)z * f_split_stride + ((size_t)j * q + b) * ld + (size_t)i * q + a];
            const double fij = s / meff;
            const double rfij = pseudocount / (double)(q * q) + (1.0 - pseudocount) * fij;
            out = rfij - reg_fi[i * q + a] * reg_fi[j * q + b];
        }
    }
    C[(size_t)row * npad + col] = out;
}

// ------------------------------------------------------------------------------
// FP64 tensor-core GEMM: C[BM x 128 tile] = beta*C + alpha * A[BM x 128] * B[128 x 128], K = 128.
// BM = 128 (8 warps) or 64 (4 warps, up to three CTAs per SM: with one 128 x 128 CTA per SM the
// epilogue's read-modify-write of C -- 256 KB per tile at ~44 GB/s per SM, 20 % of the CTA's life --
// overlapped with nothing).  Which tiles a launch touches is a `mode`:
//   GM_ALL    grid (nb, rows/BM): every tile
//   GM_BLK    step k of the symmetric Gauss-Jordan, the tiles of "block m" = row block m (tiles (m, j <= m)) and
//             column block m (tiles (i > m, m)); grid (nb, 128/BM), blockIdx.x enumerates them.  m = k + 1 is what
//             the pivot panel of step k + 1 is built from, m = k + 2 what the step after needs first (the tile
//             (k + 2, k + 1) belongs to block k + 1 and is skipped when m = k + 2)
//   GM_REST   the other tiles (i, j <= i) of every row block i != k (not in blocks k + 1, k + 2); row block k
//             itself is the pivot row of the step: its tiles are copied from B (= Rk).  grid
//             (nb (nb + 1) / 2, 128/BM), blockIdx.x enumerates the lower tiles
constexpr int GJ_B = 128;
constexpr int GM_KC = 16;
constexpr int GM_AS = 20;      // A smem row stride (doubles): 4 mod 16 -> conflict-free LDS.64
constexpr int GM_BS = 132;     // B smem row stride (doubles)
enum { GM_ALL = 0, GM_BLK = 1, GM_REST = 2 };

__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int BM>
__global__ void __launch_bounds__(BM * 2, (BM == 64) ? 3 : 1)
dgemm128_kernel(double *__restrict__ C, int ldc, const double *__restrict__ A, int lda,
                const double *__restrict__ B, int ldb, double alpha, double beta, int mode, int k, int m)
{
    constexpr int NT = BM * 2;                 // threads: one warp per 32 x 64 patch
    constexpr int SUB = GJ_B / BM;             // row sub-tiles per 128-row block
    int tr, tc, sub;                           // 128-blocked tile coordinates + row sub-tile
    if (mode == GM_ALL) {
        tc = blockIdx.x; tr = blockIdx.y / SUB; sub = blockIdx.y % SUB;
    } else if (mode == GM_BLK) {
        const int t = blockIdx.x;
        sub = blockIdx.y;
        if (t <= m) { tr = m; tc = t; } else { tr = t; tc = m; }
        if (m == k + 2 && tr == k + 2 && tc == k + 1) return;       // belongs to block k + 1
    } else {
        // blockIdx.x enumerates the tiles (tr, tc <= tr) row by row
        const int t = blockIdx.x;
        sub = blockIdx.y;
        tr = (int)((sqrtf(8.f * (float)t + 1.f) - 1.f) * 0.5f);
        while ((tr + 1) * (tr + 2) / 2 <= t) ++tr;
        while (tr * (tr + 1) / 2 > t) --tr;
        tc = t - tr * (tr + 1) / 2;
        if (tr == k + 1 || (tc == k + 1 && tr > k + 1)) return;          // block k + 1
        if (tr == k + 2 || (tc == k + 2 && tr > k + 2)) return;          // block k + 2
        if (tr == k) {
            // row block k of the result is Rk itself (B): copy this CTA's BM x 128 piece
            const double *src = B + (size_t)sub * BM * ldb + (size_t)tc * GJ_B;
            double *dst = C + ((size_t)k * GJ_B + (size_t)sub * BM) * ldc + (size_t)tc * GJ_B;
            for (int e = threadIdx.x; e < BM * (GJ_B / 2); e += NT) {
                const int r = e / (GJ_B / 2), c2 = e - r * (GJ_B / 2);
                reinterpret_cast<double2 *>(dst + (size_t)r * ldc)[c2] = reinterpret_cast<const double2 *>(src + (size_t)r * ldb)[c2];
            }
            return;
        }
    }
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *As = reinterpret_cast<double *>(smem_raw);                 // [2][BM][GM_AS]
    double *Bs = As + 2 * BM * GM_AS;                                  // [2][GM_KC][GM_BS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;                           // (BM/32) x 2 warps: 32 x 64 per warp
    const int g = lane >> 2, t4 = lane & 3;
    const size_t row0 = (size_t)tr * GJ_B + (size_t)sub * BM;
    const double *Ab = A + row0 * lda;
    const double *Bb = B + (size_t)tc * GJ_B;

    double acc[4][8][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }

    auto stage = [&](int buf, int kc) {
        // A chunk: BM rows x 16 doubles (8 x 16B per row)
        for (int t = tid; t < BM * 8; t += NT) {
            const int row = t >> 3, seg = t & 7;
            cp_async16(As + ((size_t)buf * BM + row) * GM_AS + seg * 2,
                       Ab + (size_t)row * lda + kc * GM_KC + seg * 2);
        }
        // B chunk: 16 rows x 128 doubles (64 x 16B per row)
        for (int t = tid; t < GM_KC * 64; t += NT) {
            const int row = t >> 6, seg = t & 63;
            cp_async16(Bs + ((size_t)buf * GM_KC + row) * GM_BS + seg * 2,
                       Bb + (size_t)(kc * GM_KC + row) * ldb + seg * 2);
        }
        cp_async_commit();
    };

    constexpr int NK = GJ_B / GM_KC;
    stage(0, 0);
    for (int kc = 0; kc < NK; ++kc) {
        const int buf = kc & 1;
        if (kc + 1 < NK) { stage(buf ^ 1, kc + 1); cp_async_wait<1>(); }
        else             { cp_async_wait<0>(); }
        __syncthreads();
        const double *Ap = As + (size_t)buf * BM * GM_AS + (wm * 32 + g) * GM_AS + t4;
        const double *Bp = Bs + (size_t)buf * GM_KC * GM_BS + t4 * GM_BS + wn * 64 + g;
#pragma unroll
        for (int k4 = 0; k4 < GM_KC / 4; ++k4) {
            double a[4], b[8];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) a[mi] = Ap[mi * 8 * GM_AS + k4 * 4];
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) b[ni] = Bp[k4 * 4 * GM_BS + ni * 8];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 8; ++ni) dmma_m8n8k4(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
        __syncthreads();
    }
    double *Cb = C + row0 * ldc + (size_t)tc * GJ_B;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) {
            double2 *p = reinterpret_cast<double2 *>(Cb + (size_t)(wm * 32 + mi * 8 + g) * ldc + wn * 64 + ni * 8 + 2 * t4);
            double2 v = make_double2(alpha * acc[mi][ni][0], alpha * acc[mi][ni][1]);
            if (beta != 0.0) { const double2 c = *p; v.x += beta * c.x; v.y += beta * c.y; }
            *p = v;
        }
}

// inverse of one 128x128 diagonal block (unblocked Gauss-Jordan, no pivoting), one CTA.
// The block lives in REGISTERS: thread (tr, tc) = (warp, lane) owns the 8 x 4 patch rows tr*8.., cols
// tc*4..; per pivot the owners publish the old pivot row and column through two small shared arrays
// (double-buffered: one barrier per pivot) and every thread updates its patch from 4 + 8 values.
// The update of a pivot step is  a'[i][j] = a[i][j] - col[i] * row[j]  everywhere except in the pivot
// row (row[j], pivot element 1/d) and the pivot column (-col[i]/d).  Selecting those cases per element
// cost three 64-bit selects per FMA and made this single-CTA kernel issue-bound (0.185 ms per block,
// the critical path of the look-ahead).  Instead the owners patch their registers BEFORE the update --
// pivot column := 0 with row[p] := 1/d, pivot row := 0 with col[p] := -1 -- after which the one
// formula gives exactly the special values (0 - c * (1/d), 0 - (-1) * r, 0 - (-1) * (1/d)): the
// update is 32 plain FMAs per thread, the patches are switch statements only the owners run.
__global__ void __launch_bounds__(512, 1)
gj_diag_inv_kernel(const double *__restrict__ A, int ld, int k, double *__restrict__ Dinv, int *__restrict__ info,
                   double pivot_tol)
{
    __shared__ double prow[2][GJ_B], pcol[2][GJ_B];
    const int tid = threadIdx.x, tr = tid >> 5, tc = tid & 31;
    const double *Ak = A + (size_t)k * GJ_B * ld + (size_t)k * GJ_B;
    double a[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const double2 lo = *reinterpret_cast<const double2 *>(Ak + (size_t)(tr * 8 + i) * ld + tc * 4);
        const double2 hi = *reinterpret_cast<const double2 *>(Ak + (size_t)(tr * 8 + i) * ld + tc * 4 + 2);
        a[i][0] = lo.x; a[i][1] = lo.y; a[i][2] = hi.x; a[i][3] = hi.y;
    }
#define GJ_ROW_CASE(I) case I: prow[buf][tc * 4 + 0] = a[I][0]; prow[buf][tc * 4 + 1] = a[I][1]; \
                               prow[buf][tc * 4 + 2] = a[I][2]; prow[buf][tc * 4 + 3] = a[I][3]; \
                               a[I][0] = 0.0; a[I][1] = 0.0; a[I][2] = 0.0; a[I][3] = 0.0; break;
#define GJ_COL_CASE(J) case J: _Pragma("unroll") for (int i = 0; i < 8; ++i) { pcol[buf][tr * 8 + i] = a[i][J]; a[i][J] = 0.0; } break;
    for (int p = 0; p < GJ_B; ++p) {
        const int buf = p & 1;
        const bool own_row = (tr == (p >> 3)), own_col = (tc == (p >> 2));
        // the element (p, p) belongs to both owners: publish the column first (it still holds d),
        // then the row -- whose copy of d is read back below as prow[p] -- and zero both
        if (own_col) {                                 // my lane owns pivot column p (register column p & 3)
            switch (p & 3) { GJ_COL_CASE(0) GJ_COL_CASE(1) GJ_COL_CASE(2) GJ_COL_CASE(3) }
        }
        if (own_row) {                                 // my warp owns pivot row p (register row p & 7)
            switch (p & 7) { GJ_ROW_CASE(0) GJ_ROW_CASE(1) GJ_ROW_CASE(2) GJ_ROW_CASE(3)
                             GJ_ROW_CASE(4) GJ_ROW_CASE(5) GJ_ROW_CASE(6) GJ_ROW_CASE(7) }
        }
        __syncthreads();
        // prow[p] was zeroed by the column owner before the row was published: d comes from pcol
        const double d = pcol[buf][p];
        // a symmetric positive definite matrix has positive pivots (Schur complements); anything at or
        // below pivot_tol = 1e-14 * max diagonal is singular / indefinite to working precision (NaN included)
        if (tid == 0 && !(d > pivot_tol) && info && *info == 0) *info = k * GJ_B + p + 1;
        const double piv = 1.0 / d;
        double rowv[4], colv[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double r = prow[buf][tc * 4 + j] * piv;                       // scaled pivot row
            rowv[j] = (tc * 4 + j == p) ? piv : r;                              // (prow[p] is 0)
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const double c = pcol[buf][tr * 8 + i];
            colv[i] = (tr * 8 + i == p) ? -1.0 : c;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) a[i][j] = a[i][j] - colv[i] * rowv[j];
    }
#undef GJ_ROW_CASE
#undef GJ_COL_CASE
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        double *o = Dinv + (size_t)(tr * 8 + i) * GJ_B + tc * 4;
        *reinterpret_cast<double2 *>(o) = make_double2(a[i][0], a[i][1]);
        *reinterpret_cast<double2 *>(o + 2) = make_double2(a[i][2], a[i][3]);
    }
}

// Symmetric Gauss-Jordan keeps only the tiles (i, j <= i).  For an SPD matrix the intermediates
// obey, with K the pivoted and U the unpivoted block indices,  B_KK and B_UU symmetric and
// B_KU = -B_UK^T, so every tile above the diagonal is (+/-) the transpose of one below it.
// This kernel assembles what step k needs from the lower tiles:
//   Ck   [npad][128]  = column block k:   rows i >= k direct, rows i < k = -tile(k,i)^T
//   Arow [128][npad]  = row block k:      cols j < k direct, j = k the identity, j > k = tile(j,k)^T
// and then resets the lower part of column block k to the unit columns, as the unsymmetric
// algorithm does.  Every element it overwrites is read by the overwriting thread only.
__global__ void gj_gather_kernel(double *__restrict__ A, int ld, int npad, int k, double *__restrict__ Ck,
                                 double *__restrict__ Arow)
{
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)npad * GJ_B) return;
    const int r = (int)(idx >> 7), c = (int)(idx & 127);
    const int kb = k * GJ_B;
    if (r >= kb) {
        double *p = A + (size_t)r * ld + kb + c;
        const double v = *p;
        Ck[idx] = v;
        const bool diag_tile = r < kb + GJ_B;
        Arow[(size_t)c * npad + r] = diag_tile ? ((r == kb + c) ? 1.0 : 0.0) : v;
        *p = (r == kb + c) ? 1.0 : 0.0;
    } else {
        const double v = A[(size_t)(kb + c) * ld + r];
        Ck[idx] = -v;
        Arow[(size_t)c * npad + r] = v;
    }
}

// after the last step every index is pivoted and the matrix is symmetric: fill the upper tiles
__global__ void gj_mirror_kernel(double *__restrict__ A, int ld, int npad)
{
    __shared__ double t[32][33];
    const int bx = blockIdx.x, by = blockIdx.y;              // 32 x 32 element tiles
    if (bx <= by) {
        if (bx < by) return;
        // diagonal 32-tile: mirror inside the tile
    }
    const int x = threadIdx.x, y = threadIdx.y;              // 32 x 8 threads
    // source tile (row block bx, col block by) lies below the diagonal (bx > by), or is the diagonal one
    for (int yy = y; yy < 32; yy += 8) t[yy][x] = A[(size_t)(bx * 32 + yy) * ld + by * 32 + x];
    __syncthreads();
    for (int yy = y; yy < 32; yy += 8) {
        const int row = by * 32 + yy, col = bx * 32 + x;     // transposed destination
        if (col > row) A[(size_t)row * ld + col] = t[x][yy];
    }
}

__global__ void negate_copy_kernel(const double *__restrict__ A, int ld, int n, double *__restrict__ out)
{
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)n * n) return;
    const int r = (int)(idx / n), c = (int)(idx - (long)r * n);
    out[idx] = -1.0 * A[(size_t)r * ld + c];
}

// d_A: npad x npad (row-major, ld = npad), SYMMETRIC positive definite, overwritten by its inverse.
// Only the tiles on and below the diagonal are read and updated (see gj_gather_kernel).
// Look-ahead schedule.  Step k needs the pivot panel (Ck, Rk = Dinv_k * row block k), whose producers
// -- the single-CTA diagonal inverse, the gather and a 128-row GEMM -- are latency-bound kernels
// that leave the GPU empty (0.06 + 0.19 ms per step when they ran in line).  Everything they read
// lives in "block k+1" (row block k+1 and column block k+1), so the rank-128 update of step k is cut in
// three launches -- U1(k): block k+1, U2(k): block k+2, REST(k): everything else plus the store of row
// block k -- and the small ones run on two high-priority side streams, up to two steps ahead of the
// main stream, which only ever runs the big update:
//   main:    wait panel(k)                  | REST(k)                          | record rest(k)
//   side A:  wait u2(k)                     | U1(k+1) | panel(k+2) = Dinv, gather, Rk | record panel(k+2)
//   side B:  wait rest(k), panel(k+1)       | U2(k+1)                          | record u2(k+1)
// U1(k+1) needs panel(k+1) (side A order) and step k on block k+2, i.e. U2(k) (+ the tile (k+2,k+1), which
// U1(k) updated: side A order).  U2(k+1) needs panel(k+1) and step k on block k+3: REST(k), U1(k) for the
// tile (k+3,k+1) (before panel(k+1) on side A) and U2(k) for (k+3,k+2) (side B order).  REST(k) touches neither
// block k+1 nor k+2, so everything on the side streams overlaps it and fills its last partial wave;
// panel(k+2) resets column block k+2 of A, which REST(k), REST(k+1) skip.  Ck / Arow / Rk are
// triple-buffered: panel(k+2) overwrites the buffers of panel(k-1), whose last readers REST(k-1) (main) and
// U2(k-1) (side B) precede U2(k), which side A waits for.
static int gj_invert_device(double *d_A, int npad, cudaStream_t st)
{
    const int nb = npad / GJ_B;
    double pivot_tol = 0.0;
    {   // scale of the problem: the largest diagonal entry (n doubles, strided copy)
        std::vector<double> diag((size_t)npad);
        DCA_CUDA(cudaMemcpy2DAsync(diag.data(), sizeof(double), d_A, sizeof(double) * (size_t)(npad + 1), sizeof(double),
                                   (size_t)npad, cudaMemcpyDeviceToHost, st));
        DCA_CUDA(cudaStreamSynchronize(st));
        double mx = 0.0;
        for (double v : diag) {
            if (!(v > 0.0)) { set_error("mf_couplings: non-positive diagonal entry %g: the matrix is singular or not positive definite", v); return -1; }
            mx = std::max(mx, v);
        }
        pivot_tol = 1e-14 * mx;
    }
    DevBuf<double> Dinv, Ck[3], Rk[3], Arow[3];
    DevBuf<int> info_buf;
    if (!Dinv.alloc((size_t)GJ_B * GJ_B) || !info_buf.alloc(1)) return -1;
    for (int b = 0; b < 3; ++b)
        if (!Ck[b].alloc((size_t)npad * GJ_B) || !Rk[b].alloc((size_t)npad * GJ_B) || !Arow[b].alloc((size_t)npad * GJ_B))
            return -1;
    int *d_info = info_buf.p;
    cudaStream_t side = nullptr;
    std::vector<cudaEvent_t> ev_panel((size_t)nb, nullptr), ev_u2((size_t)nb, nullptr), ev_rest((size_t)nb, nullptr);
    cudaEvent_t ev_start = nullptr;
    cudaStream_t sideB = nullptr;
    {   // highest priority: the small grids must get SMs while the big update is still queued
        int prio_lo = 0, prio_hi = 0;
        DCA_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        DCA_CUDA(cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, prio_hi));
        DCA_CUDA(cudaStreamCreateWithPriority(&sideB, cudaStreamNonBlocking, prio_hi));
    }
    for (int k = 0; k < nb; ++k) {
        DCA_CUDA(cudaEventCreateWithFlags(&ev_panel[(size_t)k], cudaEventDisableTiming));
        DCA_CUDA(cudaEventCreateWithFlags(&ev_u2[(size_t)k], cudaEventDisableTiming));
        DCA_CUDA(cudaEventCreateWithFlags(&ev_rest[(size_t)k], cudaEventDisableTiming));
    }
    DCA_CUDA(cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
    DCA_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int), st));
    constexpr int BM = 64, SUB = GJ_B / BM;
    const size_t smem_gemm = sizeof(double) * (2 * BM * GM_AS + 2 * GM_KC * GM_BS);
    DCA_CUDA(cudaFuncSetAttribute(dgemm128_kernel<BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_gemm));
    // panel of step kk on side stream A: Dinv_kk, column / row block kk rebuilt from the lower tiles (column
    // block kk reset to unit columns), Rk (128 x npad) = Dinv_kk * row block kk
    auto panel = [&](int kk) -> int {
        const int b = kk % 3;
        DCA_LAUNCH(gj_diag_inv_kernel, 1, 512, 0, side, d_A, npad, kk, Dinv.p, d_info, pivot_tol);
        DCA_LAUNCH(gj_gather_kernel, (unsigned)ceil_div<long>((long)npad * GJ_B, 256), 256, 0, side, d_A, npad, npad, kk,
                   Ck[b].p, Arow[b].p);
        DCA_LAUNCH(dgemm128_kernel<BM>, dim3(nb, SUB), BM * 2, smem_gemm, side,
                   Rk[b].p, npad, Dinv.p, GJ_B, Arow[b].p, npad, 1.0, 0.0, (int)GM_ALL, 0, 0);
        DCA_CUDA(cudaEventRecord(ev_panel[(size_t)kk], side));
        return 0;
    };
    // A[i,j] -= Ck[i] * Rk[j] of step kk on the tiles of block m
    auto update_block = [&](int kk, int m, cudaStream_t s) -> int {
        const int b = kk % 3;
        DCA_LAUNCH(dgemm128_kernel<BM>, dim3(nb, SUB), BM * 2, smem_gemm, s,
                   d_A, npad, Ck[b].p, GJ_B, Rk[b].p, npad, -1.0, 1.0, (int)GM_BLK, kk, m);
        return 0;
    };
    DCA_CUDA(cudaEventRecord(ev_start, st));               // the matrix (and the memset) precede this point on st
    DCA_CUDA(cudaStreamWaitEvent(side, ev_start, 0));
    int rc = panel(0);
    if (rc == 0 && nb > 1) { rc = update_block(0, 1, side); if (rc == 0) rc = panel(1); }
    if (rc == 0 && nb > 2) {                                // U2(0)
        DCA_CUDA(cudaStreamWaitEvent(sideB, ev_panel[0], 0));
        rc = update_block(0, 2, sideB);
        if (rc == 0) DCA_CUDA(cudaEventRecord(ev_u2[0], sideB));
    }
    for (int k = 0; k < nb && rc == 0; ++k) {
        const int b = k % 3;
        // the tiles j <= i of every row block i != k outside blocks k+1, k+2 (half the rank-128 updates of the
        // unsymmetric algorithm: n^3 instead of 2 n^3 flops) and the store of row block k, below / on the diagonal
        DCA_CUDA(cudaStreamWaitEvent(st, ev_panel[(size_t)k], 0));
        DCA_LAUNCH(dgemm128_kernel<BM>, dim3(nb * (nb + 1) / 2, SUB), BM * 2, smem_gemm, st,
                   d_A, npad, Ck[b].p, GJ_B, Rk[b].p, npad, -1.0, 1.0, (int)GM_REST, k, 0);
        DCA_CUDA(cudaEventRecord(ev_rest[(size_t)k], st));
        if (k + 2 < nb) {                                   // side A: U1(k+1), panel(k+2)
            DCA_CUDA(cudaStreamWaitEvent(side, ev_u2[(size_t)k], 0));
            rc = update_block(k + 1, k + 2, side);
            if (rc == 0) rc = panel(k + 2);
            if (rc != 0) break;
        }
        if (k + 3 < nb) {                                   // side B: U2(k+1)
            DCA_CUDA(cudaStreamWaitEvent(sideB, ev_rest[(size_t)k], 0));
            DCA_CUDA(cudaStreamWaitEvent(sideB, ev_panel[(size_t)k + 1], 0));
            rc = update_block(k + 1, k + 3, sideB);
            if (rc == 0) DCA_CUDA(cudaEventRecord(ev_u2[(size_t)k + 1], sideB));
        }
    }
    if (rc == 0) DCA_LAUNCH(gj_mirror_kernel, dim3(npad / 32, npad / 32), dim3(32, 8), 0, st, d_A, npad, npad);
    int info = 0;
    if (rc == 0) DCA_CUDA(cudaMemcpyAsync(&info, d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
    cudaStreamSynchronize(side);
    cudaStreamSynchronize(sideB);
    cudaStreamSynchronize(st);
    cudaStreamDestroy(side);
    cudaStreamDestroy(sideB);
    for (auto e : ev_panel) if (e) cudaEventDestroy(e);
    for (auto e : ev_u2) if (e) cudaEventDestroy(e);
    for (auto e : ev_rest) if (e) cudaEventDestroy(e);
    cudaEventDestroy(ev_start);
    if (rc != 0) return -1;
    if (info != 0) {
        set_error("mf_couplings: pivot at row %d is not positive: the correlation matrix is singular or indefinite to "
                  "working precision (increase the pseudocount)", info - 1);
        return -1;
    }
    return 0;
}

// ------------------------------------------------------------------------------
struct MfWork {
    uint8_t *d_msa = nullptr;
    unsigned *d_rows = nullptr, *d_off = nullptr;     // per site: rows grouped by residue, group offsets
    double *d_w = nullptr, *d_F = nullptr, *d_fi = nullptr, *d_rfi = nullptr;
    int ld = 0, nsplit = 1;                           // one copy of F (the consumers' nsplit argument)
    int gpus = 1;                                     // GPUs that computed the pair histograms
    size_t fstride = 0;
    double meff = 0.0;
    ~MfWork() { cudaDeviceSynchronize(); dev_free(d_msa); dev_free(d_rows); dev_free(d_off); dev_free(d_w); dev_free(d_F); dev_free(d_fi); dev_free(d_rfi); }
};

template <int Q>
static int mf_pairfreq_range_t(const uint8_t *d_msa, const double *d_w, unsigned *d_rows, unsigned *d_off, long N, int L,
                               int ld, double *d_F, int site_lo, int site_hi)
{
    if (site_hi <= site_lo) return 0;
    DCA_LAUNCH((mf_site_buckets_kernel<Q>), ceil_div(site_hi - site_lo, 4), 128, 0, 0, d_msa, N, L, d_rows, d_off,
               site_lo, site_hi);
    const int njt = ceil_div(L, 32);
    dim3 grid(ceil_div(Q * njt, MF_HW), site_hi - site_lo);
    DCA_LAUNCH((mf_pairhist_kernel<Q>), grid, MF_HW * 32, 0, 0, d_msa, d_w, d_rows, d_off, N, L, ld, d_F, site_lo);
    DCA_CUDA(cudaGetLastError());
    return 0;
}

static int mf_pairfreq_range(int q, const uint8_t *d_msa, const double *d_w, unsigned *d_rows, unsigned *d_off, long N,
                             int L, int ld, double *d_F, int site_lo, int site_hi)
{
    return (q == 21) ? mf_pairfreq_range_t<21>(d_msa, d_w, d_rows, d_off, N, L, ld, d_F, site_lo, site_hi)
                     : mf_pairfreq_range_t<5>(d_msa, d_w, d_rows, d_off, N, L, ld, d_F, site_lo, site_hi);
}

// Contiguous site ranges of equal histogram cost: site i costs the 32-site tiles holding some
// j >= i, i.e. njt - i/32 (the kernel skips the tiles left of the diagonal).
static std::vector<int> mf_site_bounds(int L, int G)
{
    const int njt = ceil_div(L, 32);
    std::vector<double> cum((size_t)L + 1, 0.0);
    for (int i = 0; i < L; ++i) cum[(size_t)i + 1] = cum[(size_t)i] + (double)(njt - i / 32);
    std::vector<int> b((size_t)G + 1, L);
    b[0] = 0;
    for (int r = 1; r < G; ++r) {
        const double target = cum[(size_t)L] * r / G;
        int k = (int)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
        b[(size_t)r] = std::min(L, std::max(k, b[(size_t)r - 1]));
    }
    return b;
}

// the pair histograms are independent per (site i, tile of j), so GPU r takes
// the sites [bounds[r], bounds[r+1]) -- MSA and weights replicated -- and stores its columns of F
// directly into the solver GPU's matrix (peer stores from the kernel; no gather step).  Ranks
// are host threads of this call; GPU `home` (the caller's) owns F and does the rest of the pipeline.
static void mf_pairfreq_worker(int dev, int home, const uint8_t *h_msa, const double *h_w, long N, int L, int q, int ld,
                               double *d_F_home, int site_lo, int site_hi, std::atomic<int> *err)
{
    if (cudaSetDevice(dev) != cudaSuccess) { err->store(1); return; }
    int can = 0;
    cudaDeviceCanAccessPeer(&can, dev, home);
    if (!can) { set_error("mf: GPU %d cannot access GPU %d's memory", dev, home); err->store(1); return; }
    const cudaError_t pe = cudaDeviceEnablePeerAccess(home, 0);
    if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) { set_error("mf: enabling peer access failed"); err->store(1); return; }
    cudaGetLastError();
    DevBuf<uint8_t> d_msa;
    DevBuf<double> d_w;
    DevBuf<unsigned> d_rows, d_off;
    if (!d_msa.alloc((size_t)N * L) || !d_w.alloc((size_t)N) || !d_rows.alloc((size_t)L * N) ||
        !d_off.alloc((size_t)L * (q + 1)) ||
        cudaMemcpy(d_msa, h_msa, (size_t)N * L, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(d_w, h_w, sizeof(double) * (size_t)N, cudaMemcpyHostToDevice) != cudaSuccess ||
        mf_pairfreq_range(q, d_msa, d_w, d_rows, d_off, N, L, ld, d_F_home, site_lo, site_hi) != 0 ||
        cudaDeviceSynchronize() != cudaSuccess) {
        if (cudaPeekAtLastError() != cudaSuccess) set_error("mf: GPU %d: %s", dev, cudaGetErrorString(cudaGetLastError()));
        err->store(1);
    }
}

static int mf_prepare(MfWork &W, const uint8_t *msa, const double *w, long N, int L, int q)
{
    if (require_device() != 0) return -1;
    if (q != 21 && q != 5) { set_error("mf: num_site_states must be 21 or 5 (got %d)", q); return -1; }
    if (N <= 0 || L < 2) { set_error("mf: empty alignment (N=%ld L=%d)", N, L); return -1; }
    {
        const long k = first_bad_residue(msa, N * (long)L, q);
        if (k >= 0) { set_error("mf: residue code %d >= q=%d at flat index %ld", msa[k], q, k); return -1; }
    }
    const int Lq = L * q;
    W.ld = ceil_div(Lq, MF_TC) * MF_TC;
    W.nsplit = 1;
    W.fstride = (size_t)Lq * W.ld;
    double m = 0.0;
    for (long n = 0; n < N; ++n) m += w[n];
    W.meff = m;
    if (!dev_alloc_t(&W.d_msa, (size_t)N * L) || !dev_alloc_t(&W.d_w, (size_t)N) ||
        !dev_alloc_t(&W.d_rows, (size_t)L * N) || !dev_alloc_t(&W.d_off, (size_t)L * (q + 1)) ||
        !dev_alloc_t(&W.d_F, W.fstride * W.nsplit) || !dev_alloc_t(&W.d_fi, (size_t)Lq) ||
        !dev_alloc_t(&W.d_rfi, (size_t)Lq))
        return -1;
    DCA_CUDA(cudaMemcpy(W.d_msa, msa, (size_t)N * L, cudaMemcpyHostToDevice));
    DCA_CUDA(cudaMemcpy(W.d_w, w, sizeof(double) * (size_t)N, cudaMemcpyHostToDevice));
    // GPUs: DCA_B200_GPUS / dca_set_gpus (default all visible), worth it only for real work
    int G = backend_gpu_count(L);
    if ((double)N * L * L < 2.0e8) G = 1;
    int home = 0;
    DCA_CUDA(cudaGetDevice(&home));
    if (G > 1 && home != 0) G = 1;                 // the helpers are devices 1..G-1
    W.gpus = 1;
    if (G > 1) {
        const std::vector<int> b = mf_site_bounds(L, G);
        std::atomic<int> err{0};
        std::vector<std::thread> th;
        for (int r = 1; r < G; ++r)
            th.emplace_back(mf_pairfreq_worker, r, home, msa, w, N, L, q, W.ld, W.d_F, b[(size_t)r], b[(size_t)r + 1], &err);
        const int rc0 = mf_pairfreq_range(q, W.d_msa, W.d_w, W.d_rows, W.d_off, N, L, W.ld, W.d_F, b[0], b[1]);
        for (auto &t : th) t.join();
        DCA_CUDA(cudaSetDevice(home));
        if (rc0 != 0) return -1;
        if (!err.load()) { W.gpus = G; return 0; }
        // a helper failed (no peer access, out of memory ...): redo everything here
    }
    return mf_pairfreq_range(q, W.d_msa, W.d_w, W.d_rows, W.d_off, N, L, W.ld, W.d_F, 0, L);
}

int mf_freqs_host(const uint8_t *msa, const double *w, long N, int L, int q, double *fi, double *fij)
{
    MfWork W;
    if (mf_prepare(W, msa, w, N, L, q) != 0) return -1;
    const int Lq = L * q;
    DCA_LAUNCH(mf_single_kernel, ceil_div(Lq, 128), 128, 0, 0, W.d_F, W.fstride, W.nsplit, L, q, W.ld, W.meff, 0.0,
               W.d_fi, (double *)nullptr);
    DCA_CUDA(cudaMemcpy(fi, W.d_fi, sizeof(double) * (size_t)Lq, cudaMemcpyDeviceToHost));
    if (fij) {
        const long total = (long)L * (L - 1) / 2 * (q - 1) * (q - 1);
        DevBuf<double> d_fij;
        if (!d_fij.alloc((size_t)total)) return -1;
        DCA_LAUNCH(mf_pack_fij_kernel, (unsigned)ceil_div<long>(total, 256), 256, 0, 0,
                   W.d_F, W.fstride, W.nsplit, L, q, W.ld, W.meff, d_fij.p);
        cudaError_t e = cudaMemcpy(fij, d_fij.p, sizeof(double) * (size_t)total, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { set_error("mf_freqs: D2H failed: %s", cudaGetErrorString(e)); return -1; }
    }
    return 0;
}

// builds C on the device; returns the device pointer (npad x npad) through d_C
static int mf_corr_device(MfWork &W, int L, int q, double pseudocount, double **d_C, int *npad_out)
{
    const int Lq = L * q, n = L * (q - 1);
    const int npad = ceil_div(n, GJ_B) * GJ_B;
    DCA_LAUNCH(mf_single_kernel, ceil_div(Lq, 128), 128, 0, 0, W.d_F, W.fstride, W.nsplit, L, q, W.ld, W.meff,
               pseudocount, W.d_fi, W.d_rfi);
    if (!dev_alloc_t(d_C, (size_t)npad * npad)) return -1;
    dim3 grid(ceil_div(npad, 256), npad);
    DCA_LAUNCH(mf_corr_kernel, grid, 256, 0, 0, W.d_F, W.fstride, W.nsplit, W.d_rfi, L, q, W.ld, W.meff,
               pseudocount, n, npad, *d_C);
    DCA_CUDA(cudaGetLastError());
    *npad_out = npad;
    return 0;
}

int mf_corr_mat_host(const uint8_t *msa, const double *w, long N, int L, int q, double pseudocount,
                     double *reg_fi, double *corr)
{
    MfWork W;
    if (mf_prepare(W, msa, w, N, L, q) != 0) return -1;
    double *d_C = nullptr;
    int npad = 0;
    if (mf_corr_device(W, L, q, pseudocount, &d_C, &npad) != 0) { dev_free(d_C); return -1; }
    const int n = L * (q - 1);
    cudaError_t e = cudaMemcpy2D(corr, sizeof(double) * (size_t)n, d_C, sizeof(double) * (size_t)npad,
                                 sizeof(double) * (size_t)n, (size_t)n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && reg_fi)
        e = cudaMemcpy(reg_fi, W.d_rfi, sizeof(double) * (size_t)L * q, cudaMemcpyDeviceToHost);
    dev_free(d_C);
    if (e != cudaSuccess) { set_error("mf_corr_mat: D2H failed: %s", cudaGetErrorString(e)); return -1; }
    return 0;
}

// construct_corr_mat as a standalone operator on caller-supplied regularised
// frequencies (msa_numerics.py:271-318): reg_fi [L][q], reg_fij [pairs][q-1][q-1].
__global__ void mf_corr_from_freqs_kernel(const double *__restrict__ reg_fi, const double *__restrict__ reg_fij,
                                          int L, int q, int n, double *__restrict__ C)
{
    const int qm1 = q - 1;
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = blockIdx.y;
    if (col >= n) return;
    int i = row / qm1, a = row - i * qm1, j = col / qm1, b = col - j * qm1;
    double out;
    if (i == j) {
        const double fia = reg_fi[i * q + a], fib = reg_fi[i * q + b];
        out = (a == b) ? fia * (1.0 - fia) : -1.0 * fia * fib;
    } else {
        if (i > j) { int t = i; i = j; j = t; t = a; a = b; b = t; }
        out = reg_fij[(pair_index(L, i, j) * qm1 + a) * qm1 + b] - reg_fi[i * q + a] * reg_fi[j * q + b];
    }
    C[(size_t)row * n + col] = out;
}

int mf_corr_from_freqs_host(const double *reg_fi, const double *reg_fij, int L, int q, double *corr)
{
    if (require_device() != 0) return -1;
    if (L < 2 || q < 2) { set_error("construct_corr_mat: bad shape L=%d q=%d", L, q); return -1; }
    const int qm1 = q - 1, n = L * qm1;
    const size_t nfij = (size_t)L * (L - 1) / 2 * qm1 * qm1;
    DevBuf<double> d_fi, d_fij, d_C;          // cached blocks, released on every return path
    if (!d_fi.alloc((size_t)L * q) || !d_fij.alloc(nfij) || !d_C.alloc((size_t)n * n)) return -1;
    DCA_CUDA(cudaMemcpy(d_fi.p, reg_fi, sizeof(double) * (size_t)L * q, cudaMemcpyHostToDevice));
    DCA_CUDA(cudaMemcpy(d_fij.p, reg_fij, sizeof(double) * nfij, cudaMemcpyHostToDevice));
    dim3 grid(ceil_div(n, 256), n);
    DCA_LAUNCH(mf_corr_from_freqs_kernel, grid, 256, 0, 0, d_fi.p, d_fij.p, L, q, n, d_C.p);
    DCA_CUDA(cudaGetLastError());
    DCA_CUDA(cudaMemcpy(corr, d_C.p, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToHost));
    return 0;
}

// scoring.cu
int fn_matrix_device(const double *d_J, long ldm, int L, int q, double *d_fn, cudaStream_t st);
int di_matrix_device(const double *d_J, long ldm, const double *d_rfi, int L, int q, double *d_di, cudaStream_t st);
void apc_host(const double *fn, int L, double *out);

__global__ void negate_inplace_kernel(double *__restrict__ A, long count)
{
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < count) A[idx] = -1.0 * A[idx];
}

// The whole mean-field pipeline with every intermediate resident in HBM
// (meanfield_dca.py:902-988 end to end): MSA + weights in, scores out.  Optional
// outputs: reg_fi [L][q], couplings [n][n], di [pairs]; stage_ms[5] = pair
// frequencies, covariance assembly, inverse, FN, DI (CUDA events).
int mf_scores_host(const uint8_t *msa, const double *w, long N, int L, int q, double pseudocount,
                   double *reg_fi, double *couplings, double *fn, double *fn_apc, double *di, float *stage_ms)
{
    cudaEvent_t ev[6];
    for (auto &e : ev) DCA_CUDA(cudaEventCreate(&e));
    MfWork W;
    DCA_CUDA(cudaEventRecord(ev[0], 0));
    if (mf_prepare(W, msa, w, N, L, q) != 0) return -1;
    DCA_CUDA(cudaEventRecord(ev[1], 0));
    double *d_C = nullptr, *d_fn = nullptr, *d_di = nullptr;
    int npad = 0;
    int rc = mf_corr_device(W, L, q, pseudocount, &d_C, &npad);
    DCA_CUDA(cudaEventRecord(ev[2], 0));
    if (rc == 0) rc = gj_invert_device(d_C, npad, 0);
    const int n = L * (q - 1);
    const long npairs = (long)L * (L - 1) / 2;
    if (rc == 0) {
        DCA_LAUNCH(negate_inplace_kernel, (unsigned)ceil_div<long>((long)npad * npad, 256), 256, 0, 0, d_C, (long)npad * npad);
        DCA_CUDA(cudaEventRecord(ev[3], 0));
        if (!dev_alloc_t(&d_fn, (size_t)npairs)) rc = -1;
    }
    if (rc == 0) rc = fn_matrix_device(d_C, npad, L, q, d_fn, 0);
    if (rc == 0) {
        DCA_CUDA(cudaEventRecord(ev[4], 0));
        if (di) {
            if (!dev_alloc_t(&d_di, (size_t)npairs)) rc = -1;
            if (rc == 0) rc = di_matrix_device(d_C, npad, W.d_rfi, L, q, d_di, 0);
        }
    }
    if (rc == 0) {
        DCA_CUDA(cudaEventRecord(ev[5], 0));
        cudaError_t e = cudaMemcpy(fn, d_fn, sizeof(double) * (size_t)npairs, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && di) e = cudaMemcpy(di, d_di, sizeof(double) * (size_t)npairs, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && reg_fi) e = cudaMemcpy(reg_fi, W.d_rfi, sizeof(double) * (size_t)L * q, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && couplings)
            e = cudaMemcpy2D(couplings, sizeof(double) * (size_t)n, d_C, sizeof(double) * (size_t)npad,
                             sizeof(double) * (size_t)n, (size_t)n, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { set_error("mf_scores: D2H failed: %s", cudaGetErrorString(e)); rc = -1; }
    }
    if (rc == 0) {
        if (fn_apc) apc_host(fn, L, fn_apc);
        if (stage_ms) for (int k = 0; k < 5; ++k) cudaEventElapsedTime(&stage_ms[k], ev[k], ev[k + 1]);
    }
    cudaDeviceSynchronize();
    dev_free(d_C); dev_free(d_fn); dev_free(d_di);
    for (auto &e : ev) cudaEventDestroy(e);
    return rc;
}

// ---- measured FP64 tensor-pipe peak (the roofline denominator of the inverse) --------------
// Register-only DMMA stream: every warp keeps 16 independent m8n8k4 accumulator pairs in
// flight (no memory traffic), 8 warps per CTA, enough CTAs to fill every SM several times.
__global__ void __launch_bounds__(256)
fp64_mma_peak_kernel(double *__restrict__ sink, int iters)
{
    double acc[16][2];
#pragma unroll
    for (int k = 0; k < 16; ++k) { acc[k][0] = 0.0; acc[k][1] = 0.0; }
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) dmma_m8n8k4(acc[k][0], acc[k][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += acc[k][0] + acc[k][1];
    if (s == -1.0) sink[0] = s;          // never true: keeps the loop alive
}

// best-of-reps TFLOP/s of mma.sync.m8n8k4.f64 (2*8*8*4 flops per warp instruction)
int fp64_mma_peak(int reps, double *tflops_out)
{
    if (require_device() != 0) return -1;
    if (reps < 1 || !tflops_out) { set_error("fp64_mma_peak: bad arguments"); return -1; }
    cudaStream_t st;
    DCA_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    DevBuf<double> sink;
    if (!sink.alloc(8)) { cudaStreamDestroy(st); return -1; }
    const int iters = 4096, ctas = sm_count() * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int r = -1; r < reps; ++r) {
        cudaEventRecord(e0, st);
        DCA_LAUNCH(fp64_mma_peak_kernel, ctas, 256, 0, st, sink.p, iters);
        cudaEventRecord(e1, st);
        cudaStreamSynchronize(st);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 512.0 * 16.0 * iters * 8.0 * ctas;
        if (r >= 0 && ms > 0.f) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaError_t e = cudaGetLastError();
    cudaStreamDestroy(st);
    if (e != cudaSuccess) { set_error("fp64_mma_peak: %s", cudaGetErrorString(e)); return -1; }
    *tflops_out = best;
    return 0;
}

// largest |A[r][c] - A[c][r]| and largest |A[r][c]| (as ordered bit patterns of non-negative doubles)
__global__ void asymmetry_kernel(const double *__restrict__ A, int ld, int n, unsigned long long *__restrict__ out2)
{
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    double da = 0.0, mx = 0.0;
    if (idx < (long)n * n) {
        const int r = (int)(idx / n), c = (int)(idx - (long)r * n);
        const double v = A[(size_t)r * ld + c];
        mx = fabs(v);
        if (c > r) da = fabs(v - A[(size_t)c * ld + r]);
        if (v != v) da = __longlong_as_double(0x7ff0000000000000LL);       // NaN input: report as infinitely asymmetric
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        da = fmax(da, __shfl_xor_sync(0xffffffffu, da, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(out2, (unsigned long long)__double_as_longlong(da));
        atomicMax(out2 + 1, (unsigned long long)__double_as_longlong(mx));
    }
}

int mf_couplings_host(const double *corr, int n, double *out)
{
    if (require_device() != 0) return -1;
    const int npad = ceil_div(n, GJ_B) * GJ_B;
    DevBuf<double> A_buf, out_buf;            // cached blocks, released on every return path
    if (!A_buf.alloc((size_t)npad * npad)) return -1;
    double *d_A = A_buf.p, *d_out = nullptr;
    DCA_CUDA(cudaMemset(d_A, 0, sizeof(double) * (size_t)npad * npad));
    DCA_CUDA(cudaMemcpy2D(d_A, sizeof(double) * (size_t)npad, corr, sizeof(double) * (size_t)n,
                          sizeof(double) * (size_t)n, (size_t)n, cudaMemcpyHostToDevice));
    if (npad > n) {   // identity on the padding diagonal
        std::vector<double> ones((size_t)(npad - n), 1.0);
        DCA_CUDA(cudaMemcpy2D(d_A + (size_t)n * npad + n, sizeof(double) * (size_t)(npad + 1), ones.data(),
                              sizeof(double), sizeof(double), (size_t)(npad - n), cudaMemcpyHostToDevice));
    }
    // The reference inverts with a general, pivoted LU (np.linalg.inv, msa_numerics.py:340); the kernel
    // here is a symmetric, non-pivoting elimination that reads the lower tiles only.  It is exact for what
    // construct_corr_mat produces (symmetric positive definite for pseudocount > 0); anything else is
    // refused instead of silently giving wrong couplings.
    int rc = 0;
    {
        DevBuf<unsigned long long> d_chk;
        unsigned long long h_chk[2] = {0, 0};
        if (!d_chk.alloc(2)) rc = -1;
        if (rc == 0) {
            cudaMemset(d_chk, 0, sizeof h_chk);
            DCA_LAUNCH(asymmetry_kernel, (unsigned)ceil_div<long>((long)n * n, 256), 256, 0, 0, d_A, npad, n, d_chk.p);
            if (cudaMemcpy(h_chk, d_chk, sizeof h_chk, cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("mf_couplings: symmetry check failed"); rc = -1; }
        }
        if (rc == 0) {
            double asym, mx;
            memcpy(&asym, &h_chk[0], 8); memcpy(&mx, &h_chk[1], 8);
            if (!(asym <= 1e-12 * mx)) {
                set_error("mf_couplings: corr_mat is not symmetric (max |C - C^T| = %g, max |C| = %g); only symmetric positive "
                          "definite matrices (construct_corr_mat's output) can be inverted here", asym, mx);
                rc = -1;
            }
        }
    }
    if (rc == 0) rc = gj_invert_device(d_A, npad, 0);
    if (rc == 0) {
        if (!out_buf.alloc((size_t)n * n)) rc = -1;
        d_out = out_buf.p;
    }
    if (rc == 0) {
        DCA_LAUNCH(negate_copy_kernel, (unsigned)ceil_div<long>((long)n * n, 256), 256, 0, 0, d_A, npad, n, d_out);
        cudaError_t e = cudaMemcpy(out, d_out, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { set_error("mf_couplings: D2H failed: %s", cudaGetErrorString(e)); rc = -1; }
    }
    cudaDeviceSynchronize();                  // nothing of this call is in flight when the blocks go back to the cache
    return rc;
}

}  // namespace dca
