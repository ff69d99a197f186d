// reweight.cu -- all-pairs Hamming sequence reweighting on sm_100a.
//
// Replaces  PlmDCA::computeSeqsWeight   pydca/plmdca/plmdca_numerics.cpp:611-671
//           compute_sequences_weight    pydca/meanfield_dca/msa_numerics.py:14-50
//                                       pydca/plmdca/msa_numerics.py:13-49
//
// c_i = #{ j : ident(i,j)/L > seqid }  (self included, strict >).  The float
// compare is turned into an INTEGER threshold on the host (the smallest ident
// for which the reference's float32 / float64 expression is true), so the
// device works on integers only and the counts are bit-exact.
//
// Layout: the u8 MSA is re-packed as rows of 32-bit words (4 residues per
// word), row stride Lw words (multiple of 4 -> 16-byte cp.async), pad bytes 0
// in every row (pads always match and add no mismatches).
//
// Kernel: CTA = 128 rows (i) x a range of j; 64-row j tiles and 128-byte k
// chunks are staged into shared memory with 16-byte cp.async (double
// buffered); each thread owns an 8x4 block of (i,j) pairs.  Per 4 residues:
//   t = a ^ b ; u = t + 0x1f1f1f1f ; m = u & 0x20202020 ; acc = dp4a(m, 1, acc)
// residues are < 32, so bit 5 of every byte of u is set iff that byte differs
// (no carries between bytes); dp4a sums the four flags into a 32-bit counter.
// Symmetric mode (sym = 1): ident(i,j) = ident(j,i), so only the tiles on and above the
// diagonal are compared (row tile I against j >= 128*I) and a hit outside the diagonal block
// counts for BOTH rows: half the compares.  Row tiles are dealt to `nparts` parts cyclically
// (multi-GPU: part = rank; the partial count vectors add up to the full counts).
// Roofline: integer-ALU bound (4 lane-ops per 4 residues), not HBM bound:
// algorithmic bytes are N*L + 4N, ops N^2*L byte compares.
#include "dca_common.cuh"
#include "dca_internal.h"

#include <algorithm>
#include <cstring>
#include <vector>

namespace dca {

constexpr int RW_TI = 128;      // i rows per CTA
constexpr int RW_TJ = 64;       // j rows per tile
constexpr int RW_KW = 32;       // words per k chunk (128 residues)
constexpr int RW_STRIDE = 36;   // smem row stride in words (16B aligned, conflict-free LDS.128)

__global__ void __launch_bounds__(256)
reweight_kernel(const uint32_t *__restrict__ msa_w, long N, long i_lo, long i_hi, int Lw,
                long j_per_split, unsigned max_mismatch32, unsigned *__restrict__ counts,
                int sym, int part, int nparts)
{
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t *Is = smem;                                   // [2][RW_TI][RW_STRIDE]
    uint32_t *Js = smem + 2 * RW_TI * RW_STRIDE;           // [2][RW_TJ][RW_STRIDE]
    __shared__ unsigned jc[RW_TJ];                         // symmetric mode: hits per j row of the current tile

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    long i0, jbeg, jend;
    if (sym) {
        // row tile `part + blockIdx.x * nparts` against j in [i0, N), that span cut into gridDim.y pieces
        i0 = ((long)part + (long)blockIdx.x * nparts) * RW_TI;
        i_lo = 0; i_hi = N;
        const long span = N - i0;
        const long per = ((span + gridDim.y - 1) / gridDim.y + RW_TJ - 1) / RW_TJ * RW_TJ;
        jbeg = i0 + (long)blockIdx.y * per;
        jend = min(N, jbeg + per);
        if (tid < RW_TJ) jc[tid] = 0;
    } else {
        i0 = i_lo + (long)blockIdx.x * RW_TI;
        jbeg = (long)blockIdx.y * j_per_split;
        jend = min(N, jbeg + j_per_split);
    }
    const int nk = (Lw + RW_KW - 1) / RW_KW;

    unsigned rowcount[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) rowcount[r] = 0;

    auto stage = [&](int buf, long j0, int kc) {
        // I tile: 128 rows x 8 x 16B ; J tile: 64 rows x 8 x 16B
        const int kw0 = kc * RW_KW;
        for (int t = tid; t < RW_TI * 8; t += 256) {
            const int row = t >> 3, seg = t & 7;
            uint32_t *dst = Is + ((size_t)buf * RW_TI + row) * RW_STRIDE + seg * 4;
            const long gi = i0 + row;
            const int kw = kw0 + seg * 4;
            if (gi < i_hi && kw < Lw) cp_async16(dst, msa_w + gi * Lw + kw);
            else *reinterpret_cast<uint4 *>(dst) = make_uint4(0, 0, 0, 0);
        }
        for (int t = tid; t < RW_TJ * 8; t += 256) {
            const int row = t >> 3, seg = t & 7;
            uint32_t *dst = Js + ((size_t)buf * RW_TJ + row) * RW_STRIDE + seg * 4;
            const long gj = j0 + row;
            const int kw = kw0 + seg * 4;
            if (gj < jend && kw < Lw) cp_async16(dst, msa_w + gj * Lw + kw);
            else *reinterpret_cast<uint4 *>(dst) = make_uint4(0, 0, 0, 0);
        }
        cp_async_commit();
    };

    for (long j0 = jbeg; j0 < jend; j0 += RW_TJ) {
        unsigned acc[8][4];
        unsigned pair_valid = 0;                               // bit 4 r + s: pair (row r, column s) is inside the ranges
        bool decided = false;
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                acc[r][s] = 0;
                if ((i0 + ty + 16 * r) < i_hi && (j0 + tx + 16 * s) < jend) pair_valid |= 1u << (4 * r + s);
            }

        stage(0, j0, 0);
        for (int kc = 0; kc < nk; ++kc) {
            const int buf = kc & 1;
            if (kc + 1 < nk) { stage(buf ^ 1, j0, kc + 1); cp_async_wait<1>(); }
            else             { cp_async_wait<0>(); }
            __syncthreads();
            const uint32_t *Ib = Is + (size_t)buf * RW_TI * RW_STRIDE;
            const uint32_t *Jb = Js + (size_t)buf * RW_TJ * RW_STRIDE;
            for (int k4 = 0; k4 < RW_KW / 4; ++k4) {
                uint4 a[8], b[4];
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    a[r] = *reinterpret_cast<const uint4 *>(Ib + (ty + 16 * r) * RW_STRIDE + k4 * 4);
#pragma unroll
                for (int s = 0; s < 4; ++s)
                    b[s] = *reinterpret_cast<const uint4 *>(Jb + (tx + 16 * s) * RW_STRIDE + k4 * 4);
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        unsigned c = acc[r][s];
                        c = __dp4a(((a[r].x ^ b[s].x) + 0x1f1f1f1fu) & 0x20202020u, 0x01010101u, c);
                        c = __dp4a(((a[r].y ^ b[s].y) + 0x1f1f1f1fu) & 0x20202020u, 0x01010101u, c);
                        c = __dp4a(((a[r].z ^ b[s].z) + 0x1f1f1f1fu) & 0x20202020u, 0x01010101u, c);
                        c = __dp4a(((a[r].w ^ b[s].w) + 0x1f1f1f1fu) & 0x20202020u, 0x01010101u, c);
                        acc[r][s] = c;
                    }
                // Early exit: mismatches only grow, so once EVERY pair of the tile is over the budget the tile is
                // decided (no hit).  On a similarity-sorted alignment almost all tile pairs are cross-cluster and are
                // decided after a fraction of the columns (budget 0.2 L at a mismatch rate of ~0.95).  One block
                // vote per 32 residues; pairs outside the ranges compare zeros with zeros and are masked out.
                if ((k4 & 1) == 1) {
                    bool open = false;
#pragma unroll
                    for (int r = 0; r < 8; ++r)
#pragma unroll
                        for (int s2 = 0; s2 < 4; ++s2)
                            open |= (acc[r][s2] <= max_mismatch32) && ((pair_valid >> (4 * r + s2)) & 1u);
                    if (!__syncthreads_or(open)) { decided = true; break; }
                }
            }
            if (decided) break;
            __syncthreads();
        }
        if (decided) { cp_async_wait<0>(); __syncthreads(); }        // the next chunk may already be in flight
        const bool mirror = sym && (j0 >= i0 + RW_TI);     // off the diagonal block: the hit counts for row j too
        unsigned colhit[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            unsigned c = 0;
            const bool vi = (i0 + ty + 16 * r) < i_hi;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const long gj = j0 + tx + 16 * s;
                const unsigned hit = (gj < jend && acc[r][s] <= max_mismatch32) ? 1u : 0u;
                c += hit;
                colhit[s] += vi ? hit : 0u;
            }
            rowcount[r] += c;
        }
        if (mirror) {                                      // block-uniform
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                unsigned c = colhit[s];
                c += __shfl_xor_sync(0xffffffffu, c, 16);  // the two ty rows of this warp
                if ((tid & 31) < 16 && c) atomicAdd(&jc[tx + 16 * s], c);
            }
            __syncthreads();
            if (tid < RW_TJ) {
                const unsigned c = jc[tid];
                if (c) atomicAdd(counts + j0 + tid, c);    // j0 + tid < jend whenever c != 0
                jc[tid] = 0;
            }
            __syncthreads();
        }
    }
    // reduce the 16 tx lanes that share a row (lanes 0-15 / 16-31 of a warp)
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        unsigned c = rowcount[r];
        c += __shfl_xor_sync(0xffffffffu, c, 8);
        c += __shfl_xor_sync(0xffffffffu, c, 4);
        c += __shfl_xor_sync(0xffffffffu, c, 2);
        c += __shfl_xor_sync(0xffffffffu, c, 1);
        const long gi = i0 + ty + 16 * r;
        if (tx == 0 && gi < i_hi) atomicAdd(counts + gi, c);   // integer: order-independent
    }
}

// host-side integer threshold: smallest ident for which the reference's
// compare holds (plmdca_numerics.cpp:642-643 in float32; msa_numerics.py:46 in
// float64).  Returns L+1 when no ident satisfies it (seqid >= 1).
int reweight_min_ident(int L, double seqid, int compare_f64)
{
    for (int k = 0; k <= L; ++k) {
        bool hit;
        if (compare_f64) hit = ((double)k / (double)L) > seqid;
        else {
            volatile float sim = (float)k / (float)L;
            hit = sim > (float)seqid;
        }
        if (hit) return k;
    }
    return L + 1;
}

// pack a u8 [N][L] device MSA into zero-padded word rows
__global__ void pack_rows_kernel(const uint8_t *__restrict__ msa, long N, int L, int Lw,
                                 uint32_t *__restrict__ out, unsigned *__restrict__ bad, const int *__restrict__ order)
{   // order (may be null): packed row n holds sequence order[n] (similarity-sorted rows)   // *bad is raised for residue codes >= 32: the mismatch test of reweight_kernel (bit 5 of
    // (a ^ b) + 0x1f per byte) is exact only for 5-bit codes
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * Lw) return;
    const long n = idx / Lw;
    const int kw = (int)(idx - n * Lw);
    uint32_t v = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const int s = kw * 4 + b;
        if (s < L) {
            const uint32_t c = msa[(order ? (long)order[n] : n) * L + s];
            if (c >= 32u) atomicOr(bad, 1u);
            v |= c << (8 * b);
        }
    }
    out[idx] = v;
}

// counts[order[k]] = sorted_counts[k]
__global__ void unsort_counts_kernel(const unsigned *__restrict__ sorted_counts, const int *__restrict__ order, long N,
                                     unsigned *__restrict__ counts)
{
    const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < N) counts[order[k]] = sorted_counts[k];
}

// counts for rows [row_lo,row_hi) against all N rows (multi-GPU: row shards,
// SURVEY 8e).  d_msa: u8 [N][L] on device; d_counts: u32 [N] (zeroed here for
// the shard range only).
// part >= 0: symmetric mode -- part `part` of `nparts` of the upper-triangular tile pairs;
// d_counts [N] receives this part's PARTIAL counts of every row (nparts = 1: the counts).
// part < 0: rows [row_lo, row_hi) against all N rows (full compares, counts of those rows only).
int reweight_counts_device(const uint8_t *d_msa, long N, int L, double seqid, int compare_f64,
                           long row_lo, long row_hi, unsigned *d_counts, cudaStream_t stream,
                           int part, int nparts)
{
    if (N <= 0 || L <= 0) { set_error("reweight: empty alignment (N=%ld L=%d)", N, L); return -1; }
    const bool sym = part >= 0;
    if (sym) {
        if (nparts < 1 || part >= nparts) { set_error("reweight: bad part %d of %d", part, nparts); return -1; }
        row_lo = 0; row_hi = N;
    }
    const int Lw = (int)(ceil_div(L, 4) + 3) / 4 * 4;
    DevBuf<uint32_t> d_words;            // released on every return path
    DevBuf<unsigned> d_bad, d_sorted_counts;
    DevBuf<int> d_order;
    if (!d_words.alloc((size_t)N * Lw) || !d_bad.alloc(1)) return -1;
    // Deep alignments (symmetric mode): compare the rows in similarity order.  Almost every tile pair then holds
    // unrelated sequences only, which the kernel's early exit decides after a fraction of the columns; the counts
    // come back through the permutation.  The order costs N^2/16 sampled compares -- worth it from ~1e10 on -- and
    // every part computes all of it (the parts must agree on the order): at N = 200 000 it is ~19 ms next to
    // 69 / nparts ms of compares against 167 / nparts ms in file order (measured per part: 4 parts 36.8 vs 42.7 ms,
    // 8 parts 28.1 vs 21.5 ms), so from five parts on the file order wins (DCA_B200_REWEIGHT_ORDER=sorted | file
    // overrides).
    const char *order_env = getenv("DCA_B200_REWEIGHT_ORDER");
    const bool order_file = order_env && strcmp(order_env, "file") == 0;
    const bool order_sorted = order_env && strcmp(order_env, "sorted") == 0;
    const bool use_sorted = sym && (double)N * (double)N * (double)L >= 2.0e10 && N < (1L << 24) && L < 2048 &&
                            !order_file && (nparts <= 4 || order_sorted);
    if (use_sorted) {
        std::vector<int> order, rowidx;
        if (!d_order.alloc((size_t)N) || !d_sorted_counts.alloc((size_t)N)) return -1;
        if (similarity_order(d_msa, N, L, stream, order, rowidx) != 0) return -1;
        DCA_CUDA(cudaMemcpyAsync(d_order.p, order.data(), sizeof(int) * (size_t)N, cudaMemcpyHostToDevice, stream));
        DCA_CUDA(cudaStreamSynchronize(stream));           // `order` goes out of scope
        DCA_CUDA(cudaMemsetAsync(d_sorted_counts.p, 0, sizeof(unsigned) * (size_t)N, stream));
    }
    unsigned *d_out = use_sorted ? d_sorted_counts.p : d_counts;
    DCA_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(unsigned), stream));
    {
        const long total = N * Lw;
        DCA_LAUNCH(pack_rows_kernel, (unsigned)ceil_div<long>(total, 256), 256, 0, stream,
                   d_msa, N, L, Lw, d_words, d_bad, use_sorted ? (const int *)d_order.p : (const int *)nullptr);
    }
    DCA_CUDA(cudaMemsetAsync(d_counts + row_lo, 0, sizeof(unsigned) * (size_t)(row_hi - row_lo), stream));
    const int T = reweight_min_ident(L, seqid, compare_f64);
    if (T <= L) {
        const unsigned max_mismatch32 = 32u * (unsigned)(L - T);
        const long rows = row_hi - row_lo;
        unsigned gx = (unsigned)ceil_div<long>(rows, RW_TI);
        if (sym) {
            const long ntile = ceil_div<long>(N, RW_TI);
            gx = (unsigned)((ntile - part + nparts - 1) / nparts);      // row tiles part, part + nparts, ...
        }
        // split j so that the grid covers the SMs at least twice (symmetric mode: row tile I has
        // (N - 128 I) rows of j, so CTA costs fall off linearly; more, smaller CTAs balance that)
        // (4 CTAs are resident per SM; symmetric CTAs get ~12 waves of work, at least 4 j tiles each)
        int js = (int)ceil_div<long>((sym ? 48L : 2L) * sm_count(), std::max(1u, gx));
        if (sym) js = (int)std::max<long>(1, std::min<long>(js, N / (4 * RW_TJ)));
        long jper = ceil_div<long>(ceil_div<long>(N, js), RW_TJ) * RW_TJ;
        js = (int)ceil_div<long>(N, jper);
        const size_t smem = sizeof(uint32_t) * 2 * (RW_TI + RW_TJ) * RW_STRIDE;
        // per-device attribute (a process may drive several GPUs): set on every call, it is cheap
        DCA_CUDA(cudaFuncSetAttribute(reweight_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid(gx, (unsigned)js);
        if (gx > 0)
            DCA_LAUNCH(reweight_kernel, grid, 256, smem, stream,
                       d_words, N, row_lo, row_hi, Lw, jper, max_mismatch32, d_out,
                       sym ? 1 : 0, sym ? part : 0, sym ? nparts : 1);
        if (use_sorted)
            DCA_LAUNCH(unsort_counts_kernel, (unsigned)ceil_div<long>(N, 256), 256, 0, stream, d_sorted_counts.p, d_order.p, N,
                       d_counts);
    }
    DCA_CUDA(cudaGetLastError());
    unsigned h_bad = 0;
    DCA_CUDA(cudaMemcpyAsync(&h_bad, d_bad, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
    DCA_CUDA(cudaStreamSynchronize(stream));
    if (h_bad) { set_error("reweight: residue code >= 32 in the alignment (states must be 0..q-1, q <= 32)"); return -1; }
    return 0;
}

}  // namespace dca
