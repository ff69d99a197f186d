// plm_eval.cu -- plmDCA objective + gradient on sm_100a.
//
// Replaces PlmDCA::gradient  (pydca/plmdca/plmdca_numerics.cpp:436-607) and
// PlmDCA::initFieldsAndCouplings (:207-249).  Same parameter vector layout:
//   x[a + q*i]                                   fields            (:226, :467)
//   x[L*q + pair(i,j)*q*q + a*q + b]   i<j       couplings J_ij(a,b) (:338-342)
//
// Mathematics (per evaluation, all float32 like the reference):
//   logit_n(i,a) = carry_{n-1}(i,a) + h_i(a) + sum_{j!=i} J_ij(a, s_nj)     (:499-517)
//   p_n(i,.)     = softmax(logit_n(i,.))                                    (:519-530)
//   fx          += -w_n log p_n(i, s_ni)                                    (:536)
//   r_n(i,a)     = w_n (p_n(i,a) - [a == s_ni])                             (:538-567)
//   dh_i(a)      = sum_n r_n(i,a)
//   dJ_ij(a,b)   = sum_n r_n(i,a)[s_nj == b] + sum_n [s_ni == a] r_n(j,b)   (:570-602)
// "carry" is the reference's quirk: prob_ni is allocated once per site (:492)
// and only += at :499, so sequence n starts from sequence n-1's probabilities.
// The chain p_n = softmax(p_{n-1} + S_n) contracts by >= 2x per step (the
// softmax Jacobian has norm <= 1/2), so a chain restarted `warm` = 32 steps
// early reproduces the serial chain to < 2^-32: the scan is chunked.
//
// Data layout in HBM (one PlmProblem):
//   msa   u8  [N][L]        row-major           (softmax: residue of (n,i))
//   msaT  u8  [L][Npad]     site-major          (logits: 16 residues per LDS.128)
//   W     f32 [L*q][ld]     full symmetric coupling matrix, row (j,b) = j*q+b,
//                           column (i,a) = i*q+a, zero diagonal blocks;
//                           rebuilt from x every evaluation (unpack)
//   S     f32 [Npad][ld]    logits
//   Rz    f32 [Npad][ld]    residuals r (separate buffer: see plm_softmax_kernel)
//   G     f32 [nsplit][L*q][ld]   G[(j,b)][(i,a)] = sum_n [s_nj==b] r_n(i,a)
//   perm  [site tile][nchunk][site][832 B]  per (site, 128-sequence chunk): bucket
//                           sizes + the chunk's rows as byte offsets grouped by
//                           residue (static for an MSA, built once) -> atomics-free
//                           scatter; one contiguous block per (CTA, chunk)
//
// Kernels per evaluation and what bounds them (DESIGN.md has the numbers):
//   unpack   x -> W                          HBM/L2 streaming, 12P bytes
//   logits   S = h + sum_j W[(j,s_nj), :]    shared-memory gather: 1 LDS word / add;
//                                            TMA tensor-map tiles + mbarrier ring
//   softmax  chunked carry scan, S -> r, fx  latency/L2, 8*N*L*q bytes
//   scatter  G = sum_n onehot^T r            shared-memory gather: 1 LDS word / add;
//                                            TMA tensor-map tiles + mbarrier ring
//   assemble G (+G^T) + 2*lambda*x -> g, and {fx, g.d, g.g, x.x} finished by the
//            last CTA                        HBM/L2 streaming
// The gather kernels do N*L*L*q adds each; they are bound by shared-memory
// bandwidth (128 B/clk/SM = 32 adds/clk/SM), not by HBM.
#include "dca_common.cuh"
#include "dca_internal.h"
#include "xrank.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <utility>
#include <vector>

namespace dca {

// ------------------------------------------------------------------------------
constexpr int LG_TN = 128;        // sequences per logits CTA (8 warps x 16)
constexpr int LG_TC = 128;        // columns per CTA (32 lanes x float4)
constexpr int LG_THREADS = 256;   // 2 CTAs per SM
template <int Q> struct LgCfg { static constexpr int JB = (Q > 8) ? 4 : 16; };   // sites per stage

constexpr int SC_CH = 128;        // sequences per scatter chunk
constexpr int SC_G4 = SC_CH / 4;  // groups of four rows per chunk: their sums ride along as 32 extra tile rows
constexpr int SC_TC = 128;
constexpr int SC_WARPS = 16;      // one site per warp
// perm record of one (site, chunk): 32 bytes of bucket sizes (u8[q]) followed by
// the chunk's rows as u32 BYTE OFFSETS into the staged residual tile
// (row * SC_TC * 4), grouped by residue; every bucket starts on a 16-byte
// boundary (padded to a multiple of 4 entries) so that a warp walks the record
// with one 16-byte broadcast load per 4 rows.
constexpr int PERM_CNT = 32;
constexpr int PERM_ENTRIES = SC_CH + 3 * 21 + 9;      // 200: worst-case padding + one prefetched quad
constexpr int PERM_REC = PERM_CNT + 4 * PERM_ENTRIES; // 832 bytes (52 x 16 B)

constexpr int AS_JG = 8;          // sites j per assemble / unpack CTA
constexpr int AS_UN = 4;          // elements per thread and step of the assemble loops

struct PlmProblem {
    long N = 0, Npad = 0, P = 0;
    int L = 0, q = 0, Lq = 0, ld = 0;
    float lh = 0.f, lJ = 0.f;
    int site_lo = 0, site_hi = 0;
    int carry = 1, chunk = 32, warm = 32;
    int fidelity = 0;        // 1: keep per-(scan chunk, site) fx partials for the reference-arithmetic optimiser (seqsum.cu)
    double *d_fxsite = nullptr;
    cudaStream_t stream = nullptr;
    uint8_t *d_msa = nullptr, *d_msaT = nullptr, *d_perm = nullptr;
    float *d_w = nullptr, *d_W = nullptr, *d_G = nullptr, *d_S = nullptr, *d_Rz = nullptr;
    double *d_fxpart = nullptr, *d_red = nullptr, *d_scal = nullptr, *h_scal = nullptr;
    float h_meff = 0.f;
    int nsplit = 1, chunks_per_split = 0, nchunk = 0;
    long nscan = 0;          // number of scan chunks
    int red_blocks = 0, n_active = 0;
    unsigned *d_ticket = nullptr;
    int col_tile0 = 0, col_tiles = 0;     // scatter: 128-column tiles covering my sites' columns
    int lg_tc = LG_TC, lg_tile0 = 0, lg_tiles = 0;   // logits: lg_tc-column tiles (128, or 64 for small grids)
    int site_tiles = 0, sites_per_tile = 0;
    CUtensorMap tmS;         // Rz (residuals) as [Npad][ld] f32, box 128 x 128 (scatter tiles)
    CUtensorMap tmW;         // W as [Lq][ld] f32, box lg_tc x JB*q (logits stage)
    CUtensorMap tmM;         // msaT as [L][Npad] u8, box 128 x JB (logits stage)
    // ---- owner-sharded multi-GPU mode (xrank.cuh): x, g and every L-BFGS vector hold only
    // the LOCAL slice [fields of my sites | my coupling blocks]; W lives in the peer arena
    bool sharded = false, connected = false;
    bool shared_process = false;   // peers are threads of this process (tests): see plm_read_scalars
    long Ploc = 0;           // local parameters = nq_loc + ncpl_loc
    long nq_loc = 0;         // (site_hi - site_lo) * q fields
    long pair0 = 0;          // pair index of my first block = pair(site_lo, site_lo + 1)
    long ncpl_loc = 0;       // floats in my coupling blocks
    int max_sites = 0;       // largest shard (rows of every rank's R buffer / q)
    unsigned char *arena = nullptr;
    XrLayout lay{};
    XrPeers peers{};
    float *d_R = nullptr;    // arena + off_R
    unsigned epoch = 0, red_seq = 0;
    unsigned *d_err = nullptr;
    double *d_scal_loc = nullptr;
    int n_active_own = 0;
    int own_jg = AS_JG;      // sites j per CTA of the owner kernels: 8, or 2 when the shard is small
    // ---- asymmetric (per-site) estimator: the sites [site_lo, site_hi) are independent problems, each
    // with its OWN copy of the couplings J^(i)_ij.  The parameter vector IS the coupling matrix in the
    // W layout, restricted to my columns:  x = [ h (ld floats: column (i - lo)*q + a) | W (Lq rows x ld) ],
    // so the logits kernel reads x through a tensor map and the gradient is the scatter's G plus the
    // regulariser: no unpack, no transposing assemble, no exchange between GPUs.
    // ---- similarity-sorted layout (single-GPU joint objective): rows of S / Rz and the columns of msaTs are in an
    // order that puts similar sequences next to each other, so that the logits kernel finds many groups of four
    // rows with one residue at a site (one shared-memory row instead of four; see plm_logits_kernel)
    bool sorted = false, want_sorted = false;
    int *d_rowidx = nullptr;          // sorted row of sequence n (the carry scan runs in file order)
    uint8_t *d_msaTs = nullptr;       // msaT in sorted order
    float *d_Rz4 = nullptr;           // [Npad / 4][ld]: sums of four consecutive sorted rows of Rz (scatter fast path)
    CUtensorMap tmS4;                 // Rz4 as [Npad / 4][ld] f32, box 128 x 32
    bool group4 = false;              // the perm records reference group sums where four rows share a residue
    double uniform4 = 0.0;            // fraction of the (site, 4-row group) pairs whose rows share one residue
    bool asym = false;
    int ncol = 0;            // (site_hi - site_lo) * q valid columns (ld = padded to 128)
    int fin_rows = 8;        // rows of x per finish CTA
    struct { const float *base; CUtensorMap tm; } wmap[4];   // tensor maps over x's W part, by base pointer
    int n_wmap = 0;
};

XrCall plm_xr_next(PlmProblem *p);

// ------------------------------------------------------------------------------
__global__ void transpose_msa_kernel(const uint8_t *__restrict__ msa, long N, int L, long Npad,
                                     uint8_t *__restrict__ msaT)
{
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)L * Npad) return;
    const int j = (int)(idx / Npad);
    const long n = idx - (long)j * Npad;
    msaT[idx] = (n < N) ? msa[n * L + j] : (uint8_t)0;
}

// ---- similarity-sorted layout ----------------------------------------------------------------------------------
// Sequences within the reweighting threshold of each other share most of their residues.  One pass groups them:
// every 16th sequence of the file is a leader, each sequence joins the leader it is most identical to, and the rows
// are sorted by leader (file order inside a leader's run).  Nothing downstream depends on the quality of the
// grouping for correctness -- the logits kernel tests every group of four rows for equality itself.
// thread = one sequence; leaders are staged through shared memory 8 at a time as packed words
constexpr int NL_TILE = 8;
constexpr int NL_ROWS = 128;                            // sequences per CTA

// packed[n][w] = residues at the sampled columns 4w .. 4w+3 of sequence n (column c -> site c * cstride), one byte
// each; words and bytes beyond the sample are 0 in every row (they always match: the same constant for every pair)
__global__ void pack_sampled_kernel(const uint8_t *__restrict__ msa, long N, int Lfull, int cstride, int Ls, int Lw,
                                    uint32_t *__restrict__ packed)
{
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * Lw) return;
    const long n = idx / Lw;
    const int w = (int)(idx - n * Lw);
    const uint8_t *row = msa + (size_t)n * Lfull;
    uint32_t v = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int c = 4 * w + e;
        if (c < Ls) v |= (uint32_t)row[(size_t)c * cstride] << (8 * e);
    }
    packed[idx] = v;
}

// rows classified: n * row_stride for n < nrows (row_stride 1: every sequence; 16: the leaders themselves).
// first_at > 0: not the most identical leader but the FIRST one with at least first_at identical residues (leaders
// classified against leaders: the first similar leader is the representative of a cluster of leaders).
// blockIdx.y takes a slice of the leaders (a grid of N / 128 CTAs alone leaves most SMs empty); the slices meet in
// `key`: most identical -> atomicMax of (identical residues << 20 | 0xfffff - leader), lowest leader on ties;
// first similar -> atomicMin of the leader index.  key is cleared to 0 / 0x7fffffff by the caller; L < 2048.
// Works on the packed sample (Lw words per sequence, a multiple of 4): the CTA's 128 rows sit in shared memory
// word-major (conflict-free), the leaders are read as broadcast 16-byte loads, and per 4 residues the inner loop
// is xor / add / and / add -- mismatch flags (bit 5 of (a ^ b) + 0x1f per byte: residue codes are below 32), four
// words of them added up unshifted in the byte lanes and folded by one dp4a.  The first version gathered the row's bytes
// from global memory again for every tile of leaders: 42 ms of a 111 ms reweighting at N = 200 000.
__global__ void __launch_bounds__(NL_ROWS)
nearest_leader_kernel(const uint32_t *__restrict__ packed, long nrows, int row_stride, int Lw, int nleaders,
                      int first_at, unsigned *__restrict__ key)
{
    extern __shared__ __align__(16) uint32_t nl_smem[];
    uint32_t *lead = nl_smem;                           // [NL_TILE][Lw]
    uint32_t *rows = nl_smem + NL_TILE * Lw;            // [Lw][NL_ROWS]
    const int tid = threadIdx.x;
    const long n = (long)blockIdx.x * NL_ROWS + tid;
    {
        const uint32_t *src = packed + (size_t)(min(n, nrows - 1) * row_stride) * Lw;
        for (int w = 0; w < Lw; w += 4) {
            const uint4 v = *reinterpret_cast<const uint4 *>(src + w);
            rows[(w + 0) * NL_ROWS + tid] = v.x; rows[(w + 1) * NL_ROWS + tid] = v.y;
            rows[(w + 2) * NL_ROWS + tid] = v.z; rows[(w + 3) * NL_ROWS + tid] = v.w;
        }
    }
    const int per = (((nleaders + (int)gridDim.y - 1) / (int)gridDim.y + NL_TILE - 1) / NL_TILE) * NL_TILE;
    const int lbeg = (int)blockIdx.y * per, lend = min(nleaders, lbeg + per);
    const int total = 4 * Lw;                           // compared bytes, padding included
    int best = 0, best_id = -1;
    for (int l0 = lbeg; l0 < lend; l0 += NL_TILE) {
        __syncthreads();
        for (int t = tid; t < NL_TILE * Lw; t += NL_ROWS) {
            const int k = t / Lw, w = t - k * Lw;
            lead[t] = (l0 + k < lend) ? packed[(size_t)((long)(l0 + k) * 16) * Lw + w] : 0u;
        }
        __syncthreads();
        uint32_t mism[NL_TILE];                         // 32 x mismatches (the flags are bit 5 of their byte)
#pragma unroll
        for (int k = 0; k < NL_TILE; ++k) mism[k] = 0;
        for (int w = 0; w < Lw; w += 4) {
            const uint32_t a0 = rows[(w + 0) * NL_ROWS + tid], a1 = rows[(w + 1) * NL_ROWS + tid];
            const uint32_t a2 = rows[(w + 2) * NL_ROWS + tid], a3 = rows[(w + 3) * NL_ROWS + tid];
#pragma unroll
            for (int k = 0; k < NL_TILE; ++k) {
                const uint4 b = *reinterpret_cast<const uint4 *>(lead + k * Lw + w);
                // four words of flags fit a byte lane unshifted (4 x 0x20); one dp4a folds the lanes
                uint32_t m = ((a0 ^ b.x) + 0x1f1f1f1fu) & 0x20202020u;
                m += ((a1 ^ b.y) + 0x1f1f1f1fu) & 0x20202020u;
                m += ((a2 ^ b.z) + 0x1f1f1f1fu) & 0x20202020u;
                m += ((a3 ^ b.w) + 0x1f1f1f1fu) & 0x20202020u;
                mism[k] = __dp4a(m, 0x01010101u, mism[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < NL_TILE; ++k) {
            if (l0 + k >= lend) continue;
            const int ident = total - (int)(mism[k] >> 5);
            if (first_at > 0) { if (best_id < first_at && ident >= first_at) { best_id = ident; best = l0 + k; } }
            else if (ident > best_id) { best_id = ident; best = l0 + k; }
        }
    }
    if (n < nrows && best_id >= 0) {
        if (first_at > 0) { if (best_id >= first_at) atomicMin(key + n, (unsigned)best); }
        else atomicMax(key + n, ((unsigned)best_id << 20) | (0xfffffu - (unsigned)best));
    }
}

// Row order that puts similar sequences next to each other (used by the plmDCA gather kernels' fast paths and by
// the reweighting kernel's early exit).  order[k] = file index of the sequence in sorted row k, rowidx = its inverse.
// Leaders = every 16th sequence of the file; identities are taken on a sample of ~64 columns.
int similarity_order(const uint8_t *d_msa, long N, int L, cudaStream_t st, std::vector<int> &order, std::vector<int> &rowidx)
{
    const int nleaders = (int)ceil_div<long>(N, 16);
    const int cstride = std::max(1, L / 64);
    const int Ls = ceil_div(L, cstride);
    const int Lw = (ceil_div(Ls, 4) + 3) / 4 * 4;       // words per packed sample row: 16-byte groups
    const int npad = 4 * Lw - Ls;                       // padding bytes: identical in every pair
    DevBuf<unsigned> d_key, d_rep;
    DevBuf<uint32_t> d_packed;
    if (!d_key.alloc((size_t)N) || !d_rep.alloc((size_t)nleaders) || !d_packed.alloc((size_t)N * Lw)) return -1;
    const size_t smem = sizeof(uint32_t) * (size_t)(NL_TILE + NL_ROWS) * Lw;
    const int slices = std::max(1, std::min(32, (int)((4L * sm_count()) / std::max<long>(1, ceil_div<long>(N, NL_ROWS)))));
    DCA_CUDA(cudaMemsetAsync(d_key.p, 0, sizeof(unsigned) * (size_t)N, st));
    DCA_CUDA(cudaMemsetAsync(d_rep.p, 0x7f, sizeof(unsigned) * (size_t)nleaders, st));
    DCA_LAUNCH(pack_sampled_kernel, (unsigned)ceil_div<long>(N * Lw, 256), 256, 0, st, d_msa, N, L, cstride, Ls, Lw, d_packed.p);
    DCA_LAUNCH(nearest_leader_kernel, dim3((unsigned)ceil_div<long>(N, NL_ROWS), (unsigned)slices), NL_ROWS, smem, st,
               d_packed.p, N, 1, Lw, nleaders, 0, d_key.p);
    // leaders against leaders: the first leader sharing half of the residues represents a cluster of leaders, so that
    // the runs of one true cluster end up next to each other (groups of four that straddle two runs stay similar)
    DCA_LAUNCH(nearest_leader_kernel, dim3((unsigned)ceil_div<long>(nleaders, NL_ROWS), 32), NL_ROWS, smem, st,
               d_packed.p, (long)nleaders, 16, Lw, nleaders, (Ls + 1) / 2 + npad, d_rep.p);
    std::vector<int> best((size_t)N), rep((size_t)nleaders);
    order.assign((size_t)N, 0);
    rowidx.assign((size_t)N, 0);
    DCA_CUDA(cudaMemcpyAsync(best.data(), d_key.p, sizeof(int) * (size_t)N, cudaMemcpyDeviceToHost, st));
    DCA_CUDA(cudaMemcpyAsync(rep.data(), d_rep.p, sizeof(int) * (size_t)nleaders, cudaMemcpyDeviceToHost, st));
    DCA_CUDA(cudaStreamSynchronize(st));
    for (long n = 0; n < N; ++n) {
        const unsigned l = 0xfffffu - ((unsigned)best[(size_t)n] & 0xfffffu);
        best[(size_t)n] = (l < (unsigned)nleaders) ? (int)l : 0;
    }
    for (int k = 0; k < nleaders; ++k)              // a leader always finds itself
        if (rep[(size_t)k] < 0 || rep[(size_t)k] > k) rep[(size_t)k] = k;
    // rank of every leader in (representative, leader) order, then a counting sort of the sequences by that
    // rank: file order inside a leader's run
    std::vector<int> by_rep((size_t)nleaders), rank((size_t)nleaders);
    for (int k = 0; k < nleaders; ++k) by_rep[(size_t)k] = k;
    std::stable_sort(by_rep.begin(), by_rep.end(), [&](int a, int b) { return rep[(size_t)a] < rep[(size_t)b]; });
    for (int r = 0; r < nleaders; ++r) rank[(size_t)by_rep[(size_t)r]] = r;
    std::vector<int> start((size_t)nleaders + 1, 0);
    for (long n = 0; n < N; ++n) ++start[(size_t)rank[(size_t)best[(size_t)n]] + 1];
    for (int k = 0; k < nleaders; ++k) start[(size_t)k + 1] += start[(size_t)k];
    for (long n = 0; n < N; ++n) {
        const int k = start[(size_t)rank[(size_t)best[(size_t)n]]]++;
        order[(size_t)k] = (int)n;
        rowidx[(size_t)n] = k;
    }
    return 0;
}

// how many (site, group of four rows) pairs carry ONE residue: what the fast paths of the logits and scatter kernels
// save (measurement only: bench.py turns it into executed shared-memory words per add)
__global__ void count_uniform4_kernel(const uint8_t *__restrict__ msaT, long N, int L, long Npad, unsigned long long *count)
{
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long ngroups = N / 4;
    if (idx >= (long)L * ngroups) return;
    const int j = (int)(idx / ngroups);
    const long g = idx - (long)j * ngroups;
    const unsigned word = *reinterpret_cast<const unsigned *>(msaT + (size_t)j * Npad + 4 * g);
    if (word == __funnelshift_r(word, word, 8)) atomicAdd(count, 1ull);
}

// msaTs[j][k] = residue of the sequence in sorted row k at site j (rows beyond N: 0)
__global__ void sorted_transpose_kernel(const uint8_t *__restrict__ msa, const int *__restrict__ order, long N, int L,
                                        long Npad, uint8_t *__restrict__ msaTs)
{
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)L * Npad) return;
    const int j = (int)(idx / Npad);
    const long k = idx - (long)j * Npad;
    msaTs[idx] = (k < N) ? msa[(size_t)order[k] * L + j] : (uint8_t)0;
}

// one thread per (site, chunk): bucket the chunk's local indices by residue
// records are laid out [site tile][chunk][site within tile] so that the records of
// the sites one scatter CTA works on are one contiguous block per chunk (one bulk copy)
template <int Q>
__global__ void build_perm_kernel(const uint8_t *__restrict__ msaT, long N, long Npad, int L,
                                  int nchunk, int sites_per_tile, uint8_t *__restrict__ perm, int group4)
{   // group4: four consecutive rows (4g .. 4g+3) with ONE residue become a single entry that points at row
    // SC_CH + g of the staged tile -- the sum of the four rows, which the scatter kernel loads next to them
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)L * nchunk) return;
    const int j = (int)(idx / nchunk);
    const int c = (int)(idx - (long)j * nchunk);
    const uint8_t *col = msaT + (long)j * Npad + (long)c * SC_CH;
    const int rows = (int)min((long)SC_CH, N - (long)c * SC_CH);
    const int tile = j / sites_per_tile, slot = j - tile * sites_per_tile;
    uint8_t *rec = perm + (((size_t)tile * nchunk + c) * sites_per_tile + slot) * PERM_REC;
    int cnt[Q], pos[Q];
#pragma unroll
    for (int b = 0; b < Q; ++b) cnt[b] = 0;
    auto uniform4 = [&](int r0) {
        return group4 && r0 + 3 < rows && col[r0] == col[r0 + 1] && col[r0] == col[r0 + 2] && col[r0] == col[r0 + 3];
    };
    for (int r0 = 0; r0 < rows; r0 += 4) {
        const bool u = uniform4(r0);
        for (int r = r0; r < min(rows, r0 + 4); ++r) {
            const int v = col[r];
            if (u && r != r0) continue;
#pragma unroll
            for (int b = 0; b < Q; ++b) cnt[b] += (v == b);
        }
    }
    int off = 0;
#pragma unroll
    for (int b = 0; b < Q; ++b) {
        rec[b] = (uint8_t)cnt[b];
        pos[b] = off;
        off += (cnt[b] + 3) & ~3;
    }
    for (int r0 = 0; r0 < rows; r0 += 4) {
        const bool u = uniform4(r0);
        for (int r = r0; r < min(rows, r0 + 4); ++r) {
            if (u && r != r0) continue;
            const int v = col[r];
            int where = 0;
#pragma unroll
            for (int b = 0; b < Q; ++b)
                if (v == b) { where = pos[b]; pos[b] = where + 1; }
            const int src = u ? (SC_CH + (r0 >> 2)) : r;
            reinterpret_cast<uint32_t *>(rec + PERM_CNT)[where] = (uint32_t)src * (SC_TC * 4);
        }
    }
}

// Rz4[g][:] = Rz[4g][:] + Rz[4g+1][:] + Rz[4g+2][:] + Rz[4g+3][:]   (float4 per thread)
__global__ void __launch_bounds__(256)
rowsum4_kernel(const float *__restrict__ Rz, float *__restrict__ Rz4, long ngroups, int ld)
{
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int ld4 = ld / 4;
    if (idx >= ngroups * ld4) return;
    const long g = idx / ld4;
    const int c4 = (int)(idx - g * ld4);
    const float4 *r = reinterpret_cast<const float4 *>(Rz + (size_t)g * 4 * ld) + c4;
    const float4 a = r[0], b = r[ld4], c = r[2 * (size_t)ld4], d = r[3 * (size_t)ld4];
    reinterpret_cast<float4 *>(Rz4 + (size_t)g * ld)[c4] =
        make_float4((a.x + b.x) + (c.x + d.x), (a.y + b.y) + (c.y + d.y), (a.z + b.z) + (c.z + d.z), (a.w + b.w) + (c.w + d.w));
}

// ------------------------------------------------------------------------------
// unpack: x (upper-triangular blocks) -> full symmetric W
template <int Q>
__global__ void __launch_bounds__(256)
plm_unpack_kernel(const float *__restrict__ x, int L, int ld, float *__restrict__ W)
{
    constexpr int QQ = Q * Q;
    __shared__ float tile[AS_JG * QQ];
    const int i = blockIdx.y;
    const int j0 = blockIdx.x * AS_JG;
    if (j0 + AS_JG - 1 <= i) return;
    // pairs (i, jlo..jhi-1) are consecutive blocks of x: one contiguous run, one flat loop over its elements
    // (a loop per pair left 185 of 256 threads idle in every second pass and paid its bookkeeping per pair)
    const int jlo = max(j0, i + 1), jhi = min(j0 + AS_JG, L);
    const int nel = (jhi - jlo) * QQ;
    const float *xb = x + (long)L * Q + pair_index(L, i, jlo) * QQ;
    float *Wi = W + (size_t)i * Q * ld + (size_t)jlo * Q;         // row (i, 0), column (jlo, 0)
    for (int u = threadIdx.x; u < nel; u += 256) {
        const float v = xb[u];
        tile[u] = v;
        const int jj = u / QQ, t = u - jj * QQ;
        const int a = t / Q, b = t - a * Q;
        Wi[(size_t)a * ld + jj * Q + b] = v;
    }
    __syncthreads();
    float *Wj = W + (size_t)jlo * Q * ld + (size_t)i * Q;         // row (jlo, 0), column (i, 0)
    for (int u = threadIdx.x; u < nel; u += 256) {
        const int jj = u / QQ, t = u - jj * QQ;
        const int b = t / Q, a = t - b * Q;
        Wj[(size_t)(jj * Q + b) * ld + a] = tile[jj * QQ + a * Q + b];
    }
}

// ------------------------------------------------------------------------------
// logits: S[n][col] = h[col] + sum_j W[j*Q + s_nj][col]
// CTA = 128 columns x 128 sequences (8 warps x 16 sequences, lane = 4 columns), two
// CTAs per SM.  Per stage the TMA engine fetches two 2-D tiles through tensor maps --
// the W rows of JB sites (JB*Q x 128 f32) and their residues (JB x 128 u8) -- into a
// two-deep ring guarded by mbarriers (full: transaction bytes, empty: one arrival per
// warp); warps are decoupled, there is no block-wide barrier in the loop.  Rows beyond
// L (last stage) are zero-filled by the hardware: residue 0 then selects a zero row.
// Inner loop: one 16-byte load brings a warp's 16 residues, then one conflict-free
// LDS.128 (a 512-byte row segment per warp) per 4 adds.
template <int Q, int TC>
__global__ void __launch_bounds__(LG_THREADS, (TC == 128) ? 2 : 3)
plm_logits_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmM,
                  int ld, long N, int L, const float *__restrict__ xf, int col_lo, int col_hi,
                  float *__restrict__ S, int col_tile0, XrWait wt, int quad)
{   // quad: use the equal-residues fast path for groups of four rows (reassociates the sums: off in the
    // reference-arithmetic mode, which keeps the order of every addition)   // wt: owner-sharded mode only -- wait until every rank's unpack_push has signalled (W complete)
    // xf[c - col_lo] = field of column c for col_lo <= c < col_hi (the columns this GPU owns)
    // TC = columns per CTA: 128 (lane = float4) or 64 (lane = float2).  The narrow variant moves
    // the same shared-memory bytes per add; it exists for small grids (multi-GPU shards), where
    // 128-column CTAs quantise badly over the 148 SMs.
    constexpr int JB = LgCfg<Q>::JB;
    constexpr int NWARP = LG_THREADS / 32;
    constexpr int VEC = TC / 32;
    constexpr unsigned STAGE_BYTES = JB * Q * TC * 4 + JB * LG_TN;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float *Wt = reinterpret_cast<float *>(smem_raw);                       // [2][JB*Q][TC]
    uint8_t *Mt = smem_raw + sizeof(float) * 2 * JB * Q * TC;              // [2][JB][128]
    uint64_t *bars = reinterpret_cast<uint64_t *>(Mt + 2 * JB * LG_TN);    // full[2], empty[2]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int col0 = (col_tile0 + blockIdx.x) * TC;
    const long n0 = (long)blockIdx.y * LG_TN;

    if (tid == 0) {
        mbar_init(&bars[0], 1); mbar_init(&bars[1], 1);
        mbar_init(&bars[2], NWARP); mbar_init(&bars[3], NWARP);
        mbar_fence_init();
    }
    __syncthreads();
    xr_wait_all(wt);

    float acc[16][VEC];
    float gacc[4][VEC];              // per group of four rows: what all four have in common
#pragma unroll
    for (int g4 = 0; g4 < 4; ++g4)
#pragma unroll
        for (int v = 0; v < VEC; ++v) gacc[g4][v] = 0.f;
    {
        const int c = col0 + VEC * lane;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const float h = (c + v >= col_lo && c + v < col_hi) ? xf[c + v - col_lo] : 0.f;
#pragma unroll
            for (int t = 0; t < 16; ++t) acc[t][v] = h;
        }
    }

    auto issue = [&](int buf, int st) {
        if (lane == 0) {
            mbar_expect_tx(&bars[buf], STAGE_BYTES);
            tma_load_2d(Wt + (size_t)buf * JB * Q * TC, &tmW, col0, st * JB * Q, &bars[buf]);
            tma_load_2d(Mt + (size_t)buf * JB * LG_TN, &tmM, (int)n0, st * JB, &bars[buf]);
        }
        __syncwarp();
    };

    const int nstage = (L + JB - 1) / JB;
    if (warp == 0) issue(0, 0);
    if (warp == 1 && nstage > 1) issue(1, 1);

    for (int st = 0; st < nstage; ++st) {
        const int buf = st & 1;
        const unsigned parity = (unsigned)(st >> 1) & 1u;
        mbar_wait(&bars[buf], parity);
        const float *Wb = Wt + (size_t)buf * JB * Q * TC + VEC * lane;
        const uint8_t *Mb = Mt + (size_t)buf * JB * LG_TN + warp * 16;
#pragma unroll
        for (int jj = 0; jj < JB; ++jj) {
            const uint4 r16 = *reinterpret_cast<const uint4 *>(Mb + jj * LG_TN);
            const float *Wj = Wb + jj * Q * TC;
            const unsigned rr[4] = {r16.x, r16.y, r16.z, r16.w};
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {
                const unsigned word = rr[g4];                       // the residues of rows 4 g4 .. 4 g4 + 3
                if constexpr (VEC == 4) {
                    // four rows with one residue (common on the similarity-sorted layout): ONE shared-memory row
                    // into the group's accumulator instead of four -- the kernel is bound by those loads.  The
                    // test is warp-uniform (lanes are columns).
                    if (quad && word == __funnelshift_r(word, word, 8)) {
                        const float4 v = *reinterpret_cast<const float4 *>(Wj + (word & 0xffu) * TC);
                        gacc[g4][0] += v.x; gacc[g4][1] += v.y; gacc[g4][2] += v.z; gacc[g4][3] += v.w;
                    } else {
#pragma unroll
                        for (int t4 = 0; t4 < 4; ++t4) {
                            const unsigned b = (word >> (8 * t4)) & 0xffu;
                            const float4 v = *reinterpret_cast<const float4 *>(Wj + b * TC);
                            acc[4 * g4 + t4][0] += v.x; acc[4 * g4 + t4][1] += v.y;
                            acc[4 * g4 + t4][2] += v.z; acc[4 * g4 + t4][3] += v.w;
                        }
                    }
                } else {       // (64 columns: the fast path does not pay -- measured 0.060 -> 0.067 ms on a 19-site shard)
#pragma unroll
                    for (int t4 = 0; t4 < 4; ++t4) {
                        const unsigned b = (word >> (8 * t4)) & 0xffu;
                        const float2 v = *reinterpret_cast<const float2 *>(Wj + b * TC);
                        acc[4 * g4 + t4][0] += v.x; acc[4 * g4 + t4][1] += v.y;
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[2 + buf]);
        if (st + 2 < nstage && warp == (st + 2) % NWARP) {
            mbar_wait(&bars[2 + buf], parity);
            issue(buf, st + 2);
        }
    }
#pragma unroll
    for (int t = 0; t < 16; ++t) {
        const long n = n0 + warp * 16 + t;
        if (n < N) {
            float *dst = S + (size_t)n * ld + col0 + VEC * lane;
            if constexpr (VEC == 4)
                *reinterpret_cast<float4 *>(dst) = make_float4(acc[t][0] + gacc[t >> 2][0], acc[t][1] + gacc[t >> 2][1],
                                                                acc[t][2] + gacc[t >> 2][2], acc[t][3] + gacc[t >> 2][3]);
            else *reinterpret_cast<float2 *>(dst) = make_float2(acc[t][0] + gacc[t >> 2][0], acc[t][1] + gacc[t >> 2][1]);
        }
    }
}

// ------------------------------------------------------------------------------
// softmax with the reference's carry, chunked with a warm-up overlap.
// warp = 32 consecutive sites of one chunk; lane = one (site, chunk) chain.
// Logits are read from S, residuals written to a SEPARATE buffer Rz: a chain's warm-up rows are
// the rows other chains emit, so writing the residuals in place (the first version did) is a race
// that only co-residency of all CTAs hid -- grids of more than one wave read residuals as logits.
template <int Q, int SM_PF, bool SORTED>
__global__ void __launch_bounds__(128)
plm_softmax_kernel(const float *__restrict__ S, float *__restrict__ Rz, int ld, const uint8_t *__restrict__ msa,
                   const float *__restrict__ w, long N, int L, int site_lo, int site_hi,
                   int chunk, int warm, int carry, double *__restrict__ fxpart, double *__restrict__ fxsite,
                   int col_shift, const int *__restrict__ rowidx)
{   // SORTED: rowidx[n] = row of S / Rz that belongs to sequence n (similarity-sorted layout); the chain itself
    // always runs in file order   // fxsite (may be null): [scan chunk][L] per-site partials of fx, for the reference-order sum of seqsum.cu
    // col_shift: column of S / Rz that holds state 0 of site 0 is -col_shift (asymmetric mode keeps only
    // its own sites' columns: col_shift = site_lo * Q; 0 otherwise)
    __shared__ float sbuf[4][32 * Q];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long k = (long)blockIdx.y * 4 + wib;
    const long nbeg = k * chunk;
    if (nbeg >= N) return;
    const long nend = min(N, nbeg + (long)chunk);
    const long nstart = carry ? max(0L, nbeg - (long)warm) : nbeg;
    const int s0 = site_lo + blockIdx.x * 32;
    const int site = s0 + lane;
    const bool active = site < site_hi;
    const int ncols = min(32, site_hi - s0) * Q;
    float *buf = sbuf[wib];
    const size_t cbase = (size_t)s0 * Q - (size_t)col_shift;

    // The chain is serial in n and every step needs one row segment of S from L2/HBM.  SM_PF rows
    // are kept in flight (register ring, statically indexed by unrolling the step loop SM_PF
    // times).  Large grids are issue-bound and want occupancy: SM_PF = 1 (96 registers).  Small
    // grids (multi-GPU shards) are bound by the memory round trip per step (~0.9 us x 48 steps
    // was the whole 45 us kernel): SM_PF = 4 (201 registers) hides it.
    // SORTED: the row of step n is rowidx[n]; the indices run one ring ahead of the rows (rcur: rows held in nxt,
    // rnxt: rows of the steps after those), so that no row load waits for its index
    float p[Q], nxt[SM_PF][Q];
    int rcur[SM_PF], rnxt[SM_PF];
#pragma unroll
    for (int a = 0; a < Q; ++a) p[a] = 0.f;
    double fx = 0.0;
#pragma unroll
    for (int u = 0; u < SM_PF; ++u) {
        const long nn = nstart + u;
        rcur[u] = SORTED ? ((nn < nend) ? rowidx[nn] : 0) : (int)nn;
        rnxt[u] = SORTED ? ((nn + SM_PF < nend) ? rowidx[nn + SM_PF] : 0) : (int)(nn + SM_PF);
#pragma unroll
        for (int t = 0; t < Q; ++t) {
            const int idx = lane + 32 * t;
            nxt[u][t] = (nn < nend && idx < ncols) ? S[(size_t)rcur[u] * ld + cbase + idx] : 0.f;
        }
    }
    for (long n0 = nstart; n0 < nend; n0 += SM_PF) {
#pragma unroll
        for (int u = 0; u < SM_PF; ++u) {
            const long n = n0 + u;
            if (n >= nend) break;                          // warp-uniform
#pragma unroll
            for (int t = 0; t < Q; ++t) buf[lane + 32 * t] = nxt[u][t];
            const int row_n = rcur[u];                     // row of this step (the residuals go there)
            if (n + SM_PF < nend) {
                const size_t rn = (size_t)rnxt[u] * ld + cbase;
#pragma unroll
                for (int t = 0; t < Q; ++t) {
                    const int idx = lane + 32 * t;
                    nxt[u][t] = (idx < ncols) ? S[rn + idx] : 0.f;
                }
                rcur[u] = rnxt[u];
                rnxt[u] = SORTED ? ((n + 2 * SM_PF < nend) ? rowidx[n + 2 * SM_PF] : 0) : (int)(n + 2 * SM_PF);
            }
            __syncwarp();
            float l[Q];
#pragma unroll
            for (int a = 0; a < Q; ++a) l[a] = carry ? (p[a] + buf[lane * Q + a]) : buf[lane * Q + a];
            float m = l[0];
#pragma unroll
            for (int a = 1; a < Q; ++a) m = fmaxf(m, l[a]);
            float z = 0.f;
#pragma unroll
            for (int a = 0; a < Q; ++a) { l[a] = expf(l[a] - m); z += l[a]; }
            const float zi = 1.f / z;
#pragma unroll
            for (int a = 0; a < Q; ++a) p[a] = l[a] * zi;

            if (n >= nbeg) {
                if (active) {
                    const int r = msa[(size_t)n * L + site];
                    const float wn = w[n];
                    float pr = p[0];
#pragma unroll
                    for (int a = 1; a < Q; ++a) pr = (a == r) ? p[a] : pr;
                    fx -= (double)(wn * logf(pr));
#pragma unroll
                    for (int a = 0; a < Q; ++a) buf[lane * Q + a] = wn * p[a] - ((a == r) ? wn : 0.f);
                }
                __syncwarp();
                const size_t ro = (size_t)row_n * ld + cbase;
#pragma unroll
                for (int t = 0; t < Q; ++t) {
                    const int idx = lane + 32 * t;
                    if (idx < ncols) Rz[ro + idx] = buf[idx];
                }
            }
            __syncwarp();
        }
    }
    if (fxsite && active) fxsite[k * L + site] = fx;
    fx = warp_sum(active ? fx : 0.0);
    if (lane == 0) fxpart[k * gridDim.x + blockIdx.x] = fx;
}

// ------------------------------------------------------------------------------
// scatter: G[(j,b)][col] = sum_{n : s_nj == b} r[n][col], warp = one site j,
// 21 float4 accumulators per lane, sequences visited in residue-bucket order
// (perm) -> no atomics, fixed order.
// Staging: the 128 x 128 f32 residual tile and the warps' perm records are
// fetched by the TMA engine (cp.async.bulk, one 512-byte row per copy) into a
// two-deep ring guarded by mbarriers: full[b] (transaction bytes) / empty[b]
// (one arrival per active warp).  There is no block-wide barrier in the loop:
// a warp that finishes chunk c moves on to c+1 at once, and the warp whose turn
// it is ((k+2) mod active warps) waits for empty[b] and re-arms the buffer with
// chunk c+2.
template <int Q>
__global__ void __launch_bounds__(SC_WARPS * 32, 1)
plm_scatter_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmS4, int group4,
                   int ld, const uint8_t *__restrict__ perm, int nchunk,
                   long N, int L, float *__restrict__ G, size_t g_split_stride,
                   int chunks_per_split, int col_tile0, int sites_per_tile)
{   // group4: the perm records also reference rows SC_CH .. SC_CH + SC_G4 - 1 of the staged tile = the sums of four
    // consecutive rows (Rz4, a second tensor map): where four rows share a residue one load replaces four
    constexpr int SC_ROWS = SC_CH + SC_G4;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float *Rs = reinterpret_cast<float *>(smem_raw);                                    // [2][SC_ROWS][SC_TC]
    uint8_t *recs = smem_raw + sizeof(float) * 2 * SC_ROWS * SC_TC;                     // [2][SC_WARPS][PERM_REC]
    uint64_t *bars = reinterpret_cast<uint64_t *>(recs + 2 * SC_WARPS * PERM_REC);      // full[2], empty[2]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int col0 = (col_tile0 + blockIdx.x) * SC_TC;
    // site tiles are balanced (e.g. L=150 -> 10 tiles of 15 sites); active warps are a prefix
    const int jbase = (int)blockIdx.y * sites_per_tile;
    const int nactive = max(0, min(sites_per_tile, L - jbase));
    const int j = jbase + warp;
    const int c0 = blockIdx.z * chunks_per_split;
    const int c1 = min(nchunk, c0 + chunks_per_split);
    const int nloc = max(0, c1 - c0);

    if (tid == 0) {
        mbar_init(&bars[0], 1); mbar_init(&bars[1], 1);
        mbar_init(&bars[2], nactive); mbar_init(&bars[3], nactive);
        mbar_fence_init();
    }
    __syncthreads();
    if (warp >= nactive) return;                         // idle warps: nothing to stage, nothing to sum

    float4 acc[Q];
#pragma unroll
    for (int b = 0; b < Q; ++b) acc[b] = make_float4(0.f, 0.f, 0.f, 0.f);

    // one lane arms buffer `buf` with chunk c: expect the bytes, then two TMA operations:
    // the 128 x 128 residual tile (2-D tensor map over S) and this tile's perm records
    auto issue = [&](int buf, int c) {
        if (lane == 0) {
            mbar_expect_tx(&bars[buf], (unsigned)((SC_CH + (group4 ? SC_G4 : 0)) * SC_TC * 4 + nactive * PERM_REC));
            tma_load_2d(Rs + (size_t)buf * SC_ROWS * SC_TC, &tmS, col0, c * SC_CH, &bars[buf]);
            if (group4) tma_load_2d(Rs + ((size_t)buf * SC_ROWS + SC_CH) * SC_TC, &tmS4, col0, c * SC_G4, &bars[buf]);
            tma_bulk_g2s(recs + (size_t)buf * SC_WARPS * PERM_REC,
                         perm + (((size_t)blockIdx.y * nchunk + c) * sites_per_tile) * PERM_REC,
                         (unsigned)(nactive * PERM_REC), &bars[buf]);
        }
        __syncwarp();
    };

    if (warp == 0 && nloc > 0) issue(0, c0);
    if (warp == (1 % nactive) && nloc > 1) issue(1, c0 + 1);

    for (int k = 0; k < nloc; ++k) {
        const int buf = k & 1;
        const unsigned parity = (unsigned)(k >> 1) & 1u;
        mbar_wait(&bars[buf], parity);                   // TMA bytes of chunk c0+k have landed
        {
            const uint8_t *rec = recs + ((size_t)buf * SC_WARPS + warp) * PERM_REC;
            const uint4 *qp = reinterpret_cast<const uint4 *>(rec + PERM_CNT);
            const char *Rb = reinterpret_cast<const char *>(Rs + (size_t)buf * SC_ROWS * SC_TC + 4 * lane);
            // bucket sizes: 32 bytes, broadcast once per chunk
            const uint4 h0 = *reinterpret_cast<const uint4 *>(rec);
            const uint4 h1 = *reinterpret_cast<const uint4 *>(rec + 16);
            const unsigned hw[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
            uint4 cur = qp[0];               // the quad stream is contiguous across buckets:
            int pos = 0;                     // always keep the next quad in flight
#define SC_ROW(off) (*reinterpret_cast<const float4 *>(Rb + (off)))
#pragma unroll
            for (int b = 0; b < Q; ++b) {
                const unsigned m = (hw[b >> 2] >> (8 * (b & 3))) & 0xffu;
                float4 a = acc[b];
                const unsigned nfull = m >> 2, tail = m & 3u;
                // Full quads: four unconditional LDS.128 + 16 adds, nothing else.  Deliberately not
                // unrolled: the 21 bucket copies must stay inside the 32 KB instruction cache
                // (ncu: "no instruction" stalls ~1 per issued instruction at 64 KB of SASS).
#pragma unroll 1
                for (unsigned t = 0; t < nfull; ++t) {
                    const uint4 o = cur;
                    cur = qp[++pos];
                    const float4 v0 = SC_ROW(o.x), v1 = SC_ROW(o.y), v2 = SC_ROW(o.z), v3 = SC_ROW(o.w);
                    a.x += (v0.x + v1.x) + (v2.x + v3.x);
                    a.y += (v0.y + v1.y) + (v2.y + v3.y);
                    a.z += (v0.z + v1.z) + (v2.z + v3.z);
                    a.w += (v0.w + v1.w) + (v2.w + v3.w);
                }
                // Ragged tail (1..3 rows of the bucket's last, padded quad): warp-uniform branches.
                // The earlier version predicated every row of every quad against the bucket size
                // (select-with-zero): 51 issued instructions per quad for 21 useful ones, and ncu
                // showed the kernel issue-bound (71 % issue-active) below the shared-memory roof.
                if (tail != 0u) {
                    const uint4 o = cur;
                    cur = qp[++pos];
                    const float4 v0 = SC_ROW(o.x);
                    a.x += v0.x; a.y += v0.y; a.z += v0.z; a.w += v0.w;
                    if (tail > 1u) {
                        const float4 v1 = SC_ROW(o.y);
                        a.x += v1.x; a.y += v1.y; a.z += v1.z; a.w += v1.w;
                        if (tail > 2u) {
                            const float4 v2 = SC_ROW(o.z);
                            a.x += v2.x; a.y += v2.y; a.z += v2.z; a.w += v2.w;
                        }
                    }
                }
                acc[b] = a;
            }
#undef SC_ROW
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[2 + buf]);      // this warp is done with buffer `buf`
        if (k + 2 < nloc && warp == (k + 2) % nactive) {
            mbar_wait(&bars[2 + buf], parity);           // every active warp released it
            issue(buf, c0 + k + 2);
        }
    }
    {
        float *g = G + (size_t)blockIdx.z * g_split_stride + (size_t)j * Q * ld + col0 + 4 * lane;
#pragma unroll
        for (int b = 0; b < Q; ++b) *reinterpret_cast<float4 *>(g + (size_t)b * ld) = acc[b];
    }
}

// ------------------------------------------------------------------------------
// tail shared by the assemble kernels: one fused block reduction of the four sums, the
// CTA's partials to `red` (SoA: red[c * nblk + blk]; slots of CTAs that own no pair stay
// zero), a ticket, and -- in the last CTA to finish, the only one that gets `true` back --
// the fixed-order reduction  scal = { fx = reg + sum fxpart, g.d, g.g, x.x }.
__device__ __forceinline__ bool assemble_finish(double s_reg, double s_dg, double s_gg, double s_xx,
                                                double *__restrict__ red, const double *__restrict__ fxpart,
                                                int nfx, unsigned *__restrict__ ticket, unsigned n_active,
                                                double *__restrict__ scal)
{
    __shared__ double wsum[8][4];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    s_reg = warp_sum(s_reg); s_dg = warp_sum(s_dg); s_gg = warp_sum(s_gg); s_xx = warp_sum(s_xx);
    if (lane == 0) { wsum[wid][0] = s_reg; wsum[wid][1] = s_dg; wsum[wid][2] = s_gg; wsum[wid][3] = s_xx; }
    __syncthreads();
    const int nblk = gridDim.x * gridDim.y;
    const int blk = blockIdx.y * gridDim.x + blockIdx.x;
    if (threadIdx.x < 4) {
        double r = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) r += wsum[w][threadIdx.x];
        red[(size_t)threadIdx.x * nblk + blk] = r;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicInc(ticket, n_active - 1u) == n_active - 1u);
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    // last CTA: everything is visible; reduce in a fixed order (4 independent loads in flight)
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int c = 0; c < 4; ++c) {
        const double *rc = red + (size_t)c * nblk;
        double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
        int t = threadIdx.x;
        for (; t + 768 < nblk; t += 1024) { p0 += rc[t]; p1 += rc[t + 256]; p2 += rc[t + 512]; p3 += rc[t + 768]; }
        for (; t < nblk; t += 256) p0 += rc[t];
        acc[c] = (p0 + p1) + (p2 + p3);
    }
    {
        double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
        int t = threadIdx.x;
        for (; t + 768 < nfx; t += 1024) { p0 += fxpart[t]; p1 += fxpart[t + 256]; p2 += fxpart[t + 512]; p3 += fxpart[t + 768]; }
        for (; t < nfx; t += 256) p0 += fxpart[t];
        acc[0] += (p0 + p1) + (p2 + p3);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[c] = warp_sum(acc[c]);
    __syncthreads();
    if (lane == 0) { wsum[wid][0] = acc[0]; wsum[wid][1] = acc[1]; wsum[wid][2] = acc[2]; wsum[wid][3] = acc[3]; }
    __syncthreads();
    if (threadIdx.x < 4) {
        double r = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) r += wsum[w][threadIdx.x];
        scal[threadIdx.x] = r;
    }
    __syncthreads();
    return true;               // only in the last CTA, with scal[] written and visible to the CTA
}

// ------------------------------------------------------------------------------
// assemble: g = 2*lambda*x + G^T-part + G-part in the reference's layout, plus
// the reductions every L-BFGS step needs, in one pass over x, and -- in the
// last CTA to finish (ticket counter, fixed summation order) -- the final
//   scal = { fx = reg + sum fxpart, g.d, g.g, x.x }.
// "owned" (multi-GPU site shards): fields and blocks whose first site i is in
// [site_lo,site_hi) carry the regulariser; the column-(i,a) data term needs i
// in the shard, the column-(j,b) term needs j in the shard.
// red is SoA: red[c * nblk + blk]; slots of CTAs that own no pair stay zero.
template <int Q>
__global__ void __launch_bounds__(256)
plm_assemble_kernel(const float *__restrict__ x, const float *__restrict__ G, size_t g_split_stride,
                    int nsplit, int L, int ld, float lh, float lJ, int site_lo, int site_hi,
                    const float *__restrict__ d, float *__restrict__ g, double *__restrict__ red,
                    const double *__restrict__ fxpart, int nfx, unsigned *__restrict__ ticket,
                    unsigned n_active, double *__restrict__ scal)
{
    constexpr int QQ = Q * Q;
    __shared__ float tile[AS_JG * QQ];
    const int i = blockIdx.y;
    const int j0 = blockIdx.x * AS_JG;
    const bool any = (j0 + AS_JG - 1 > i);
    if (!any && blockIdx.x != 0) return;               // owns neither a pair nor the fields
    const long Lq = (long)L * Q;
    const bool own_i = (i >= site_lo && i < site_hi);
    const int jlo = max(j0, i + 1), jhi = min(j0 + AS_JG, L);      // pairs (i, j), jlo <= j < jhi
    const int npair = max(0, jhi - jlo);
    double s_reg = 0.0, s_dg = 0.0, s_gg = 0.0, s_xx = 0.0;

    if (blockIdx.x == 0 && threadIdx.x < Q) {          // fields of site i
        const int a = threadIdx.x;
        const long k = (long)i * Q + a;
        const float h = x[k];
        float gv = 0.f;
        if (own_i) {
            gv = 2.f * lh * h;
            float col = 0.f;
            for (int s = 0; s < nsplit; ++s) {
                const float *Gs = G + (size_t)s * g_split_stride;
                for (int b = 0; b < Q; ++b) col += Gs[((long)i * Q + b) * ld + i * Q + a];
            }
            gv += col;
            s_reg += (double)lh * h * h;
            s_xx += (double)h * h;
        }
        g[k] = gv;
        if (d) s_dg += (double)gv * d[k];
        s_gg += (double)gv * gv;
    }

    // Both loops take AS_UN elements per thread and step, all their loads issued before the first use
    // (the kernel is bound by the latency of its strided reads, not by bandwidth); the summation
    // order over the splits is unchanged.
    const int nel = npair * QQ;
    // transposed part U[a][b] = G[(j,b)][(i,a)]: read with a fastest, park in smem
    if (own_i) {
        for (int t0 = threadIdx.x; t0 < nel; t0 += 256 * AS_UN) {
            float u[AS_UN];
            size_t off[AS_UN];
#pragma unroll
            for (int k = 0; k < AS_UN; ++k) {
                const int t = min(t0 + 256 * k, nel - 1);
                const int jj = t / QQ, e = t - jj * QQ;
                const int b = e / Q, a = e - b * Q;
                off[k] = ((size_t)(jlo + jj) * Q + b) * ld + i * Q + a;
                u[k] = 0.f;
            }
            for (int s = 0; s < nsplit; ++s) {
                const float *Gs = G + (size_t)s * g_split_stride;
#pragma unroll
                for (int k = 0; k < AS_UN; ++k) u[k] += Gs[off[k]];
            }
#pragma unroll
            for (int k = 0; k < AS_UN; ++k) {
                const int t = t0 + 256 * k;
                if (t < nel) {
                    const int jj = t / QQ, e = t - jj * QQ;
                    const int b = e / Q, a = e - b * Q;
                    tile[jj * QQ + a * Q + b] = u[k];
                }
            }
        }
    } else {
        for (int t = threadIdx.x; t < nel; t += 256) tile[t] = 0.f;
    }
    __syncthreads();
    // pairs (i, jlo..jhi-1) are consecutive blocks of x: one contiguous run
    const long base = Lq + pair_index(L, i, jlo) * QQ;
    for (int t0 = threadIdx.x; t0 < nel; t0 += 256 * AS_UN) {
        float xv[AS_UN], v[AS_UN], dv[AS_UN];
        size_t off[AS_UN];
        bool own_j[AS_UN];
#pragma unroll
        for (int k = 0; k < AS_UN; ++k) {
            const int t = min(t0 + 256 * k, nel - 1);
            const int jj = t / QQ, e = t - jj * QQ;
            const int a = e / Q, b = e - a * Q;
            const int j = jlo + jj;
            own_j[k] = (j >= site_lo && j < site_hi);
            off[k] = ((size_t)i * Q + a) * ld + (size_t)j * Q + b;
            xv[k] = x[base + t];
            dv[k] = d ? d[base + t] : 0.f;
            v[k] = 0.f;
        }
        for (int s = 0; s < nsplit; ++s) {
            const float *Gs = G + (size_t)s * g_split_stride;
#pragma unroll
            for (int k = 0; k < AS_UN; ++k) if (own_j[k]) v[k] += Gs[off[k]];
        }
#pragma unroll
        for (int k = 0; k < AS_UN; ++k) {
            const int t = t0 + 256 * k;
            if (t < nel) {
                float gv = own_i ? 2.f * lJ * xv[k] : 0.f;
                gv += tile[t];
                if (own_j[k]) gv += v[k];
                g[base + t] = gv;
                if (own_i) { s_reg += (double)lJ * xv[k] * xv[k]; s_xx += (double)xv[k] * xv[k]; }
                if (d) s_dg += (double)gv * dv[k];
                s_gg += (double)gv * gv;
            }
        }
    }

    assemble_finish(s_reg, s_dg, s_gg, s_xx, red, fxpart, nfx, ticket, n_active, scal);
}

// ------------------------------------------------------------------------------
// Asymmetric (per-site) estimator, SURVEY 8e last row: site i keeps its own couplings J^(i)_ij,
// the conditional of site i is the only term they appear in, and the L problems are independent.
// Objective of my sites (the relaxation of the joint one: each directed copy carries lambda_J / 2):
//   f = sum_i [ -sum_n w_n log p_n(i, s_ni) + lambda_h |h_i|^2 + (lambda_J / 2) sum_{j != i} |J^(i)_ij|^2 ]
// Parameters in the W layout (see PlmProblem::asym): x = [ h | W ],  W[(j,b)][(i,a)] = J^(i)_ij(a,b);
// the gradient of the data term with respect to W is the scatter's G, element by element.
//   g_W = sum_splits G + lambda_J * W      (diagonal blocks j == i and pad columns: 0)
//   g_h[(i,a)] = sum_b G[(i,b)][(i,a)] + 2 lambda_h h     (sum_b [s_ni == b] = 1: any row block sums
//                                                          r_n(i,a) over n; the diagonal one is used)
// grid (ceil(ld / 1024), row groups + field groups): CTAs with blockIdx.y < row_groups own `rows_per_cta` rows of
// W x 1024 columns (thread = 4 columns, float4); the remaining CTAs own the field row, 128 columns each, the
// Q x nsplit terms of a column spread over the 8 warps (as a tail of one row-group CTA that sum was a serial
// chain of Q * nsplit loads per thread: 0.04 ms of a 0.06 ms kernel on a 19-site shard).
template <int Q>
__global__ void __launch_bounds__(256)
plm_asym_finish_kernel(const float *__restrict__ x, const float *__restrict__ G, size_t g_split_stride, int nsplit,
                       int Lq, int ld, int ncol, int site_lo, float lh, float lJ, int rows_per_cta, int row_groups,
                       const float *__restrict__ d, float *__restrict__ g, double *__restrict__ red,
                       const double *__restrict__ fxpart, int nfx, unsigned *__restrict__ ticket,
                       unsigned n_active, double *__restrict__ scal)
{
    double s_reg = 0.0, s_dg = 0.0, s_gg = 0.0, s_xx = 0.0;
    if ((int)blockIdx.y < row_groups) {
        const int c = ((int)blockIdx.x * 256 + (int)threadIdx.x) * 4;           // my 4 columns
        const int r0 = (int)blockIdx.y * rows_per_cta;
        if (c < ld) {
            int site[4];
            bool valid[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) { valid[e] = (c + e) < ncol; site[e] = site_lo + (c + e) / Q; }
            const float *xw = x + ld;
            float *gw = g + ld;
            const float *dw = d ? d + ld : nullptr;
            for (int r = r0; r < min(r0 + rows_per_cta, Lq); ++r) {
                const size_t o = (size_t)r * ld + c;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int s = 0; s < nsplit; ++s) {
                    const float4 v = *reinterpret_cast<const float4 *>(G + (size_t)s * g_split_stride + o);
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                }
                const float4 xv = *reinterpret_cast<const float4 *>(xw + o);
                const int rs = r / Q;
                const float xe[4] = {xv.x, xv.y, xv.z, xv.w};
                const float ge[4] = {acc.x, acc.y, acc.z, acc.w};
                float out[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const bool on = valid[e] && rs != site[e];
                    out[e] = on ? (ge[e] + lJ * xe[e]) : 0.f;
                    if (on) { s_reg += 0.5 * (double)lJ * xe[e] * xe[e]; s_xx += (double)xe[e] * xe[e]; }
                    s_gg += (double)out[e] * out[e];
                }
                *reinterpret_cast<float4 *>(gw + o) = make_float4(out[0], out[1], out[2], out[3]);
                if (dw) {
                    const float4 dv = *reinterpret_cast<const float4 *>(dw + o);
                    s_dg += (double)out[0] * dv.x + (double)out[1] * dv.y + (double)out[2] * dv.z + (double)out[3] * dv.w;
                }
            }
        }
    } else {
        // the field row: 128 columns per CTA, warp w takes the states b = w, w + 8, w + 16
        __shared__ float fs[8][128];
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        const int c0 = (((int)blockIdx.y - row_groups) * (int)gridDim.x + (int)blockIdx.x) * 128;
        float part[4] = {0.f, 0.f, 0.f, 0.f};
        if (c0 < ld) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = c0 + 4 * lane + e;
                if (c < ncol) {
                    const size_t row0 = (size_t)(site_lo + c / Q) * Q;
                    for (int b = wid; b < Q; b += 8)
                        for (int s = 0; s < nsplit; ++s) part[e] += G[(size_t)s * g_split_stride + (row0 + b) * ld + c];
                }
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) fs[wid][4 * lane + e] = part[e];
        __syncthreads();
        if (threadIdx.x < 128 && c0 < ld) {
            const int c = c0 + (int)threadIdx.x;
            float out = 0.f;
            if (c < ncol) {
                float col = 0.f;
#pragma unroll
                for (int w8 = 0; w8 < 8; ++w8) col += fs[w8][threadIdx.x];
                const float h = x[c];
                out = 2.f * lh * h + col;
                s_reg += (double)lh * h * h; s_xx += (double)h * h;
                s_gg += (double)out * out;
                if (d) s_dg += (double)out * d[c];
            }
            g[c] = out;
        }
    }
    assemble_finish(s_reg, s_dg, s_gg, s_xx, red, fxpart, nfx, ticket, n_active, scal);
}

// reference layout from the directed copies: J_ij = (J^(i)_ij + (J^(j)_ji)^T) / 2, fields as they are.
// W: [Lq][ldw] with ALL sites' columns (column (i,a) = i*Q + a), h: [Lq].  One thread per output element.
template <int Q>
__global__ void plm_asym_pack_kernel(const float *__restrict__ W, int ldw, const float *__restrict__ h, int L,
                                     float *__restrict__ xref)
{
    constexpr int QQ = Q * Q;
    const long Lq = (long)L * Q;
    const long P = Lq + (long)L * (L - 1) / 2 * QQ;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    if (idx < Lq) { xref[idx] = h[idx]; return; }
    const long k = idx - Lq;
    const long pr = k / QQ;
    const int e = (int)(k - pr * QQ), a = e / Q, b = e - a * Q;
    // invert pair(i,j): rows of the strict upper triangle
    int i = (int)((2.0 * L - 1.0 - sqrt((2.0 * L - 1.0) * (2.0 * L - 1.0) - 8.0 * (double)pr)) * 0.5);
    while (pair_index(L, i + 1, i + 2) <= pr && i + 2 < L) ++i;
    while (pair_index(L, i, i + 1) > pr) --i;
    const int j = (int)(pr - pair_index(L, i, i + 1)) + i + 1;
    const float mine = W[((size_t)j * Q + b) * ldw + (size_t)i * Q + a];      // J^(i)_ij(a,b)
    const float theirs = W[((size_t)i * Q + a) * ldw + (size_t)j * Q + b];    // J^(j)_ji(b,a)
    xref[idx] = 0.5f * (mine + theirs);
}

// ==============================================================================
// Owner-sharded multi-GPU mode (xrank.cuh).  Rank r owns the
// sites [lo, hi) and the LOCAL parameter slice
//   xl[(i - lo)*Q + a]                               fields,    lo <= i < hi
//   xl[nq + (pair(i,j) - pair0)*QQ + a*Q + b]        couplings, lo <= i < hi, j > i
// (a contiguous run of the reference's vector: x[lo*Q .. hi*Q) and
//  x[Lq + pair0*QQ .. Lq + pair(hi,hi+1)*QQ)).  Per evaluation two exchanges over
// peer memory, both fused into the kernel that produces the data:
//   unpack_push   J_ij of my blocks -> my W (columns of site i) and, by P2P stores,
//                 the W of owner(j) (columns of site j)
//   export        my G rows (i,a), i < lo  (site-j conditional's half of dJ_ij, j mine)
//                 summed over the n-splits -> R of owner(i)
// each followed by a flag barrier (xr_barrier_kernel); the 4 objective scalars are
// summed across ranks in rank order by xr_reduce_kernel (identical bits on all ranks).

template <int Q, int JG>
__global__ void __launch_bounds__(256)
plm_unpack_push_kernel(const float *__restrict__ xl, long nq, long pair0, int L, int ld, int lo,
                       XrPeers P, size_t off_W, size_t off_flags, unsigned *__restrict__ ticket, unsigned epoch)
{
    constexpr int QQ = Q * Q;
    __shared__ float tile[JG * QQ];
    const int i = lo + blockIdx.y;
    const int j0 = blockIdx.x * JG;
    const int jlo = max(j0, i + 1), jhi = min(j0 + JG, L);
    const int npair = jhi - jlo;
    if (npair > 0) {
    const float *xb = xl + nq + (pair_index(L, i, jlo) - pair0) * QQ;    // npair consecutive blocks
    for (int t = threadIdx.x; t < npair * QQ; t += 256) tile[t] = xb[t];
    __syncthreads();
    // my columns (i,a), rows (j,b):  W_me[(j,b)][(i,a)] = J_ij(a,b)
    float *Wme = reinterpret_cast<float *>(P.arena[P.rank] + off_W);
    for (int t = threadIdx.x; t < npair * QQ; t += 256) {
        const int jj = t / QQ, e = t - jj * QQ;
        const int b = e / Q, a = e - b * Q;
        Wme[((long)(jlo + jj) * Q + b) * ld + i * Q + a] = tile[jj * QQ + a * Q + b];
    }
    // columns (j,b) belong to owner(j):  W_owner[(i,a)][(j,b)] = J_ij(a,b);
    // npair*Q contiguous floats per row (i,a), stored over NVLink when owner(j) != me
    const int run = npair * Q;
    for (int t = threadIdx.x; t < Q * run; t += 256) {
        const int a = t / run, c = t - a * run;
        const int jj = c / Q, b = c - jj * Q;
        float *Wo = reinterpret_cast<float *>(P.arena[xr_owner(P, jlo + jj)] + off_W);
        Wo[((long)i * Q + a) * ld + jlo * Q + c] = tile[jj * QQ + a * Q + b];
    }
    }
    // the barrier is part of this kernel: the last CTA tells every rank that my pushes are done
    // (plm_logits_kernel waits for all ranks' signals before it touches W)
    xr_grid_signal(ticket, P, off_flags, epoch);
}

// one CTA per G row (i,a) with i < lo: those rows are the site-j conditional's contribution
// to blocks (i, j) owned by a lower rank.  Sum the n-split partials, store into the owner's R.
template <int Q>
__global__ void __launch_bounds__(128)
plm_export_kernel(const float *__restrict__ G, size_t g_split_stride, int nsplit, int ld, int lo, int hi,
                  XrPeers P, size_t off_R, size_t off_flags, unsigned *__restrict__ ticket, unsigned epoch)
{
    const int row = blockIdx.x;
    if (row < lo * Q) {                 // rank 0 exports nothing: its single CTA only signals
        const int o = xr_owner(P, row / Q);
        float *R = reinterpret_cast<float *>(P.arena[o] + off_R) + (size_t)(row - P.bounds[o] * Q) * ld;
        const float *g = G + (size_t)row * ld;
        for (int c = lo * Q + threadIdx.x; c < hi * Q; c += 128) {
            float v = 0.f;
            for (int s = 0; s < nsplit; ++s) v += g[(size_t)s * g_split_stride + c];
            R[c] = v;
        }
    }
    xr_grid_signal(ticket, P, off_flags, epoch);      // plm_assemble_owner_kernel waits for all ranks
}

// gradient of my slice: 2*lambda*x + G[(j,b)][(i,a)] (my conditional, transposed through
// smem) + the site-j conditional's half, from my own G when I own j, else from R.
template <int Q, int JG>
__global__ void __launch_bounds__(256)
plm_assemble_owner_kernel(const float *__restrict__ xl, long nq, long pair0, const float *__restrict__ G,
                          size_t g_split_stride, int nsplit, const float *__restrict__ R, int L, int ld,
                          float lh, float lJ, int lo, int hi, const float *__restrict__ dl,
                          float *__restrict__ gl, double *__restrict__ red, const double *__restrict__ fxpart,
                          int nfx, unsigned *__restrict__ ticket, unsigned n_active, double *__restrict__ scal_loc,
                          XrWait wt, XrCall xc, double *__restrict__ scal)
{
    constexpr int QQ = Q * Q;
    __shared__ float tile[JG * QQ];
    xr_wait_all(wt);                     // every rank's export has landed in my R
    const int i = lo + blockIdx.y;
    const int j0 = blockIdx.x * JG;
    const bool any = (j0 + JG - 1 > i);
    if (!any && blockIdx.x != 0) return;
    const int jlo = max(j0, i + 1), jhi = min(j0 + JG, L);
    const int npair = max(0, jhi - jlo);
    double s_reg = 0.0, s_dg = 0.0, s_gg = 0.0, s_xx = 0.0;

    if (blockIdx.x == 0 && threadIdx.x < Q) {
        const int a = threadIdx.x;
        const long k = (long)(i - lo) * Q + a;
        const float h = xl[k];
        float gv = 2.f * lh * h;
        float col = 0.f;
        for (int s = 0; s < nsplit; ++s) {
            const float *Gs = G + (size_t)s * g_split_stride;
            for (int b = 0; b < Q; ++b) col += Gs[((long)i * Q + b) * ld + i * Q + a];
        }
        gv += col;
        s_reg += (double)lh * h * h;
        s_xx += (double)h * h;
        gl[k] = gv;
        if (dl) s_dg += (double)gv * dl[k];
        s_gg += (double)gv * gv;
    }
    for (int t = threadIdx.x; t < npair * QQ; t += 256) {
        const int jj = t / QQ, e = t - jj * QQ;
        const int b = e / Q, a = e - b * Q;
        const int j = jlo + jj;
        float u = 0.f;
        for (int s = 0; s < nsplit; ++s)
            u += G[(size_t)s * g_split_stride + ((long)j * Q + b) * ld + i * Q + a];
        tile[jj * QQ + a * Q + b] = u;
    }
    __syncthreads();
    const long base = nq + (pair_index(L, i, jlo) - pair0) * QQ;
    for (int t = threadIdx.x; t < npair * QQ; t += 256) {
        const int jj = t / QQ, e = t - jj * QQ;
        const int a = e / Q, b = e - a * Q;
        const int j = jlo + jj;
        const float xv = xl[base + t];
        float gv = 2.f * lJ * xv;
        gv += tile[t];
        float v = 0.f;
        if (j < hi) {
            for (int s = 0; s < nsplit; ++s)
                v += G[(size_t)s * g_split_stride + ((long)i * Q + a) * ld + j * Q + b];
        } else {
            v = R[((long)(i - lo) * Q + a) * ld + j * Q + b];
        }
        gv += v;
        gl[base + t] = gv;
        s_reg += (double)lJ * xv * xv; s_xx += (double)xv * xv;
        if (dl) s_dg += (double)gv * dl[base + t];
        s_gg += (double)gv * gv;
    }
    // the last CTA holds this rank's four partial scalars: sum them over the ranks right here
    if (assemble_finish(s_reg, s_dg, s_gg, s_xx, red, fxpart, nfx, ticket, n_active, scal_loc))
        xr_reduce_body(xc.P, xc.lay, xc.epoch, xc.seq, scal_loc, (int)PS_COUNT, scal, xc.err);
}

__global__ void xr_barrier_kernel(XrPeers P, size_t off_flags, unsigned epoch, unsigned *err)
{
    if ((int)threadIdx.x < P.world) xr_signal_and_wait(P, off_flags, epoch, threadIdx.x, err);
}

// ------------------------------------------------------------------------------
// initial point (plmdca_numerics.cpp:51-81, :207-249): sequences accumulated in file order in float32
// exactly like the reference.  One thread per (site, state): it adds w_n for the sequences that carry its
// state and an exact 0.0f for the others, i.e. the reference's sum for that state, 16 sequences per load
// (one thread per site with 21 predicated adds per sequence took 0.54 ms at N = 5000, 21 ms at N = 200 000).
constexpr int INIT_SITES = 6;       // sites per CTA (6 x 21 = 126 threads of 128)
template <int Q>
__global__ void __launch_bounds__(128)
plm_init_fields_kernel(const uint8_t *__restrict__ msaT, const float *__restrict__ w,
                       long N, long Npad, int site_lo, int site_hi, float meff,
                       float *__restrict__ x)
{   // x[a + Q*(i - site_lo)]: the whole vector (site_lo = 0) or an owner's local slice
    constexpr int SPB = (Q <= 21) ? INIT_SITES : 1;
    __shared__ float fs[SPB][Q];
    const int sl = threadIdx.x / Q, a = threadIdx.x - sl * Q;
    const int i = site_lo + (int)blockIdx.x * SPB + sl;
    const bool on = sl < SPB && i < site_hi;
    if (on) {
        float f = 0.f;
        const uint8_t *col = msaT + (size_t)i * Npad;          // Npad is a multiple of 128: 16-byte loads
        for (long n0 = 0; n0 < N; n0 += 16) {
            const uint4 r16 = *reinterpret_cast<const uint4 *>(col + n0);
            const unsigned rr[4] = {r16.x, r16.y, r16.z, r16.w};
            float wv[16];
#pragma unroll
            for (int t = 0; t < 16; ++t) wv[t] = (n0 + t < N) ? w[n0 + t] : 0.f;
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                const int r = (int)((rr[t >> 2] >> (8 * (t & 3))) & 0xffu);
                if (n0 + t < N) f += (a == r) ? wv[t] : 0.f;
            }
        }
        const float fr = __fdiv_rn(f, meff);
        fs[sl][a] = logf(__fadd_rn(__fmul_rn(fr, meff), 1.f));
    }
    __syncthreads();
    if (on) {
        float s = 0.f;
#pragma unroll
        for (int b = 0; b < Q; ++b) s += fs[sl][b];
        const float av = __fdiv_rn(s, (float)Q);
        x[a + Q * (i - site_lo)] = fs[sl][a] - av;
    }
}

// ------------------------------------------------------------------------------
// (re)build everything that depends on the order of the rows of S / Rz: the transposed alignment the logits kernel
// stages, the softmax's row table, the scatter's perm records and the tensor map over the alignment.
// use_sorted = false: file order (reference-arithmetic mode, sharded and asymmetric objectives).
template <int Q>
static int plm_layout_t(PlmProblem *p, bool use_sorted)
{
    cudaStream_t st = p->stream;
    const long N = p->N, Npad = p->Npad;
    const int L = p->L;
    {   // file-order transposed alignment: the initial point accumulates in file order
        const long total = (long)L * Npad;
        DCA_LAUNCH(transpose_msa_kernel, (unsigned)ceil_div<long>(total, 256), 256, 0, st, p->d_msa, N, L, Npad, p->d_msaT);
    }
    p->sorted = false;
    if (use_sorted && p->d_msaTs && p->d_rowidx) {
        DevBuf<int> d_order;
        if (!d_order.alloc((size_t)N)) return -1;
        std::vector<int> order, rowidx;
        if (similarity_order(p->d_msa, N, L, st, order, rowidx) != 0) return -1;
        DCA_CUDA(cudaMemcpyAsync(d_order.p, order.data(), sizeof(int) * (size_t)N, cudaMemcpyHostToDevice, st));
        DCA_CUDA(cudaMemcpyAsync(p->d_rowidx, rowidx.data(), sizeof(int) * (size_t)N, cudaMemcpyHostToDevice, st));
        const long total = (long)L * Npad;
        DCA_LAUNCH(sorted_transpose_kernel, (unsigned)ceil_div<long>(total, 256), 256, 0, st, p->d_msa, d_order.p, N, L, Npad,
                   p->d_msaTs);
        DCA_CUDA(cudaStreamSynchronize(st));          // the host vectors and the scratch blocks go out of scope
        p->sorted = true;
    }
    const uint8_t *mT = p->sorted ? p->d_msaTs : p->d_msaT;
    {
        DevBuf<unsigned long long> d_cnt;
        unsigned long long h_cnt = 0;
        const long total = (long)L * (N / 4);
        if (total > 0 && d_cnt.alloc(1)) {
            DCA_CUDA(cudaMemsetAsync(d_cnt.p, 0, sizeof(unsigned long long), st));
            DCA_LAUNCH(count_uniform4_kernel, (unsigned)ceil_div<long>(total, 256), 256, 0, st, mT, N, L, Npad, d_cnt.p);
            DCA_CUDA(cudaMemcpyAsync(&h_cnt, d_cnt.p, sizeof h_cnt, cudaMemcpyDeviceToHost, st));
            DCA_CUDA(cudaStreamSynchronize(st));
            p->uniform4 = (double)h_cnt / (double)total;
        }
    }
    {
        const long total = (long)L * p->nchunk;
        DCA_CUDA(cudaMemsetAsync(p->d_perm, 0, (size_t)p->site_tiles * p->sites_per_tile * p->nchunk * PERM_REC, st));
        p->group4 = p->sorted && p->d_Rz4 != nullptr;
        DCA_LAUNCH((build_perm_kernel<Q>), (unsigned)ceil_div<long>(total, 128), 128, 0, st,
                   mT, N, Npad, L, p->nchunk, p->sites_per_tile, p->d_perm, p->group4 ? 1 : 0);
    }
    if (make_tensor_map_2d(&p->tmM, 1, mT, (unsigned long long)Npad, (unsigned long long)L, (unsigned long long)Npad, LG_TN,
                           (unsigned)LgCfg<Q>::JB) != 0)
        return -1;
    DCA_CUDA(cudaGetLastError());
    return 0;
}

template <int Q>
static int plm_setup_t(PlmProblem *p)
{
    if (plm_layout_t<Q>(p, p->want_sorted) != 0) return -1;
    const size_t smemA = sizeof(float) * 2 * LgCfg<Q>::JB * Q * LG_TC + 2 * LgCfg<Q>::JB * LG_TN + 64 + 1024;
    const size_t smemC = sizeof(float) * 2 * (SC_CH + SC_G4) * SC_TC + 2 * SC_WARPS * PERM_REC + 64 + 1024;
    DCA_CUDA(cudaFuncSetAttribute((plm_logits_kernel<Q, 128>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemA));
    DCA_CUDA(cudaFuncSetAttribute((plm_logits_kernel<Q, 64>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemA));
    DCA_CUDA(cudaFuncSetAttribute(plm_scatter_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemC));
    if (p->sharded) {
        // CUDA loads kernels lazily, and loading one synchronises the context: a rank that
        // first-launches a kernel while a peer thread of the same process spins in a barrier
        // would dead-lock.  Load everything the collective path launches up front.
        cudaFuncAttributes fa;
        DCA_CUDA(cudaFuncGetAttributes(&fa, (plm_unpack_push_kernel<Q, 8>)));
        DCA_CUDA(cudaFuncGetAttributes(&fa, (plm_unpack_push_kernel<Q, 2>)));
        DCA_CUDA(cudaFuncGetAttributes(&fa, plm_export_kernel<Q>));
        DCA_CUDA(cudaFuncGetAttributes(&fa, (plm_assemble_owner_kernel<Q, 8>)));
        DCA_CUDA(cudaFuncGetAttributes(&fa, (plm_assemble_owner_kernel<Q, 2>)));
        DCA_CUDA(cudaFuncGetAttributes(&fa, (plm_logits_kernel<Q, 128>)));
        DCA_CUDA(cudaFuncGetAttributes(&fa, (plm_logits_kernel<Q, 64>)));
        DCA_CUDA(cudaFuncGetAttributes(&fa, (plm_softmax_kernel<Q, 1, false>)));
        DCA_CUDA(cudaFuncGetAttributes(&fa, (plm_softmax_kernel<Q, 4, false>)));
        DCA_CUDA(cudaFuncGetAttributes(&fa, (plm_softmax_kernel<Q, 1, true>)));
        DCA_CUDA(cudaFuncGetAttributes(&fa, (plm_softmax_kernel<Q, 4, true>)));
        DCA_CUDA(cudaFuncGetAttributes(&fa, nearest_leader_kernel));
        DCA_CUDA(cudaFuncGetAttributes(&fa, sorted_transpose_kernel));
        DCA_CUDA(cudaFuncGetAttributes(&fa, plm_scatter_kernel<Q>));
        DCA_CUDA(cudaFuncGetAttributes(&fa, plm_init_fields_kernel<Q>));
        DCA_CUDA(cudaFuncGetAttributes(&fa, xr_barrier_kernel));
        if (lbfgs_preload_kernels() != 0) return -1;
    }
    DCA_CUDA(cudaGetLastError());
    return 0;
}

// the carry scan: PF rows in flight; the sorted instantiation follows the row table
#define SM_LAUNCH(PF, FXSITE, COLSHIFT)                                                                         \
    do {                                                                                                        \
        if (p->sorted)                                                                                          \
            DCA_LAUNCH((plm_softmax_kernel<Q, PF, true>), gridB, 128, 0, st, p->d_S, p->d_Rz, ld, p->d_msa,    \
                       p->d_w, p->N, L, lo, hi, p->chunk, p->warm, p->carry, p->d_fxpart, FXSITE, COLSHIFT,     \
                       (const int *)p->d_rowidx);                                                               \
        else                                                                                                    \
            DCA_LAUNCH((plm_softmax_kernel<Q, PF, false>), gridB, 128, 0, st, p->d_S, p->d_Rz, ld, p->d_msa,   \
                       p->d_w, p->N, L, lo, hi, p->chunk, p->warm, p->carry, p->d_fxpart, FXSITE, COLSHIFT,     \
                       (const int *)nullptr);                                                                   \
    } while (0)

// asymmetric mode: logits straight from x (tensor map over its W part), softmax, scatter, finish
template <int Q>
static int plm_eval_asym_t(PlmProblem *p, const float *d_x, float *d_g, const float *d_dir, cudaEvent_t *ev)
{
    cudaStream_t st = p->stream;
    const int L = p->L, ld = p->ld, lo = p->site_lo, hi = p->site_hi, nsites = hi - lo;
    int evi = 0;
#define PLM_MARK() do { if (ev) cudaEventRecord(ev[evi++], st); } while (0)
    const CUtensorMap *tm = nullptr;
    for (int k = 0; k < p->n_wmap; ++k)
        if (p->wmap[k].base == d_x) tm = &p->wmap[k].tm;
    if (!tm) {
        const int k = (p->n_wmap < 4) ? p->n_wmap++ : 3;          // the optimiser alternates between two buffers
        const unsigned jb = (Q > 8) ? 4 : 16;
        if (make_tensor_map_2d(&p->wmap[k].tm, 4, d_x + ld, (unsigned long long)ld, (unsigned long long)p->Lq,
                               sizeof(float) * (unsigned long long)ld, (unsigned)p->lg_tc, jb * (unsigned)Q) != 0) return -1;
        p->wmap[k].base = d_x;
        tm = &p->wmap[k].tm;
    }
    PLM_MARK(); PLM_MARK();                                        // (no unpack)
    const size_t smemA = sizeof(float) * 2 * LgCfg<Q>::JB * Q * p->lg_tc + 2 * LgCfg<Q>::JB * LG_TN + 64 + 1024;
    dim3 gridA(p->lg_tiles, (unsigned)(p->Npad / LG_TN));
    const XrWait wt{nullptr, nullptr, 0, 0u};
    if (p->lg_tc == 128)
        DCA_LAUNCH((plm_logits_kernel<Q, 128>), gridA, LG_THREADS, smemA, st,
                   *tm, p->tmM, ld, p->N, L, d_x, 0, p->ncol, p->d_S, 0, wt, 1);
    else
        DCA_LAUNCH((plm_logits_kernel<Q, 64>), gridA, LG_THREADS, smemA, st,
                   *tm, p->tmM, ld, p->N, L, d_x, 0, p->ncol, p->d_S, 0, wt, 1);
    PLM_MARK();
    dim3 gridB(ceil_div(nsites, 32), (unsigned)ceil_div<long>(p->nscan, 4));
    if ((long)gridB.x * gridB.y > 2L * sm_count())
        SM_LAUNCH(1, (double *)nullptr, lo * Q);
    else
        SM_LAUNCH(4, (double *)nullptr, lo * Q);
    PLM_MARK();
    const size_t smemC = sizeof(float) * 2 * (SC_CH + SC_G4) * SC_TC + 2 * SC_WARPS * PERM_REC + 64 + 1024;
    const size_t gstride = (size_t)p->Lq * ld;
    dim3 gridC(p->col_tiles, p->site_tiles, p->nsplit);
    if (p->group4)
        DCA_LAUNCH(rowsum4_kernel, (unsigned)ceil_div<long>((p->Npad / 4) * (long)(ld / 4), 256), 256, 0, st, p->d_Rz, p->d_Rz4,
                   p->Npad / 4, ld);
    DCA_LAUNCH((plm_scatter_kernel<Q>), gridC, SC_WARPS * 32, smemC, st,
               p->tmS, p->group4 ? p->tmS4 : p->tmS, p->group4 ? 1 : 0, ld, p->d_perm, p->nchunk, p->N, L, p->d_G, gstride,
               p->chunks_per_split, 0, p->sites_per_tile);
    PLM_MARK(); PLM_MARK();                                        // (no export)
    const int nfx = (int)(gridB.x * p->nscan);
    const int row_groups = ceil_div(p->Lq, p->fin_rows), field_groups = ceil_div(ceil_div(ld, 128), ceil_div(ld, 1024));
    dim3 gridF(ceil_div(ld, 1024), row_groups + field_groups);
    DCA_LAUNCH((plm_asym_finish_kernel<Q>), gridF, 256, 0, st, d_x, p->d_G, gstride, p->nsplit, p->Lq, ld, p->ncol, lo,
               p->lh, p->lJ, p->fin_rows, row_groups, d_dir, d_g, p->d_red, p->d_fxpart, nfx, p->d_ticket,
               gridF.x * gridF.y, p->d_scal);
    PLM_MARK();
#undef PLM_MARK
    DCA_CUDA(cudaGetLastError());
    return 0;
}

template <int Q>
static int plm_eval_t(PlmProblem *p, const float *d_x, float *d_g, const float *d_dir,
                      cudaEvent_t *ev = nullptr)
{
    if (p->asym) return plm_eval_asym_t<Q>(p, d_x, d_g, d_dir, ev);
    cudaStream_t st = p->stream;
    const int L = p->L, ld = p->ld;
    const int lo = p->site_lo, hi = p->site_hi, nsites = hi - lo;
    int evi = 0;
#define PLM_MARK() do { if (ev) cudaEventRecord(ev[evi++], st); } while (0)
    if (p->sharded && !p->connected) { set_error("plm: sharded objective used before its peers were connected"); return -1; }
    PLM_MARK();
    dim3 gridU(ceil_div(L, p->sharded ? p->own_jg : AS_JG), p->sharded ? nsites : L);
    if (p->sharded) {
        // fused unpack + exchange: my blocks go to my W and, by P2P stores, to the W of owner(j)
        if (p->own_jg == 8)
            DCA_LAUNCH((plm_unpack_push_kernel<Q, 8>), gridU, 256, 0, st, d_x, p->nq_loc, p->pair0, L, ld, lo,
                       p->peers, p->lay.off_W, p->lay.off_flags, p->d_ticket + 1, ++p->epoch);
        else
            DCA_LAUNCH((plm_unpack_push_kernel<Q, 2>), gridU, 256, 0, st, d_x, p->nq_loc, p->pair0, L, ld, lo,
                       p->peers, p->lay.off_W, p->lay.off_flags, p->d_ticket + 1, ++p->epoch);
    } else {
        DCA_LAUNCH((plm_unpack_kernel<Q>), gridU, 256, 0, st, d_x, L, ld, p->d_W);
    }
    PLM_MARK();

    const size_t smemA = sizeof(float) * 2 * LgCfg<Q>::JB * Q * p->lg_tc + 2 * LgCfg<Q>::JB * LG_TN + 64 + 1024;
    dim3 gridA(p->lg_tiles, (unsigned)(p->Npad / LG_TN));
    // fields: the head of the local slice (sharded) or of the whole vector
    const int col_lo = p->sharded ? lo * Q : 0, col_hi = p->sharded ? hi * Q : L * Q;
    XrWait wt{nullptr, nullptr, 0, 0u};
    if (p->sharded && p->peers.world > 1)
        wt = XrWait{reinterpret_cast<const unsigned *>(p->arena + p->lay.off_flags), p->d_err, p->peers.world, p->epoch};
    if (p->lg_tc == 128)
        DCA_LAUNCH((plm_logits_kernel<Q, 128>), gridA, LG_THREADS, smemA, st,
                   p->tmW, p->tmM, ld, p->N, L, d_x, col_lo, col_hi, p->d_S, p->lg_tile0, wt, p->fidelity ? 0 : 1);
    else
        DCA_LAUNCH((plm_logits_kernel<Q, 64>), gridA, LG_THREADS, smemA, st,
                   p->tmW, p->tmM, ld, p->N, L, d_x, col_lo, col_hi, p->d_S, p->lg_tile0, wt, p->fidelity ? 0 : 1);
    PLM_MARK();

    dim3 gridB(ceil_div(nsites, 32), (unsigned)ceil_div<long>(p->nscan, 4));
    const int col_shift = 0;
    if ((long)gridB.x * gridB.y > 2L * sm_count())
        SM_LAUNCH(1, p->fidelity ? p->d_fxsite : (double *)nullptr, col_shift);
    else
        SM_LAUNCH(4, p->fidelity ? p->d_fxsite : (double *)nullptr, col_shift);
    PLM_MARK();

    const size_t smemC = sizeof(float) * 2 * (SC_CH + SC_G4) * SC_TC + 2 * SC_WARPS * PERM_REC + 64 + 1024;
    const size_t gstride = (size_t)p->Lq * ld;
    dim3 gridC(p->col_tiles, p->site_tiles, p->nsplit);
    if (p->group4)
        DCA_LAUNCH(rowsum4_kernel, (unsigned)ceil_div<long>((p->Npad / 4) * (long)(ld / 4), 256), 256, 0, st, p->d_Rz, p->d_Rz4,
                   p->Npad / 4, ld);
    DCA_LAUNCH((plm_scatter_kernel<Q>), gridC, SC_WARPS * 32, smemC, st,
               p->tmS, p->group4 ? p->tmS4 : p->tmS, p->group4 ? 1 : 0, ld, p->d_perm, p->nchunk, p->N, L, p->d_G, gstride,
               p->chunks_per_split, p->col_tile0, p->sites_per_tile);
    PLM_MARK();

    const int nfx = (int)(gridB.x * p->nscan);
    if (p->sharded) {
        // export (+ signal) -> assemble (waits; its last CTA sums the scalars over the ranks)
        DCA_LAUNCH((plm_export_kernel<Q>), std::max(1, lo * Q), 128, 0, st, p->d_G, gstride, p->nsplit, ld, lo, hi,
                   p->peers, p->lay.off_R, p->lay.off_flags, p->d_ticket + 2, ++p->epoch);
        PLM_MARK();
        wt.epoch = p->epoch;
        const XrCall xc = plm_xr_next(p);
        if (p->own_jg == 8)
            DCA_LAUNCH((plm_assemble_owner_kernel<Q, 8>), gridU, 256, 0, st,
                       d_x, p->nq_loc, p->pair0, p->d_G, gstride, p->nsplit, p->d_R, L, ld, p->lh, p->lJ, lo, hi,
                       d_dir, d_g, p->d_red, p->d_fxpart, nfx, p->d_ticket, (unsigned)p->n_active_own, p->d_scal_loc,
                       wt, xc, p->d_scal);
        else
            DCA_LAUNCH((plm_assemble_owner_kernel<Q, 2>), gridU, 256, 0, st,
                       d_x, p->nq_loc, p->pair0, p->d_G, gstride, p->nsplit, p->d_R, L, ld, p->lh, p->lJ, lo, hi,
                       d_dir, d_g, p->d_red, p->d_fxpart, nfx, p->d_ticket, (unsigned)p->n_active_own, p->d_scal_loc,
                       wt, xc, p->d_scal);
    } else {
        PLM_MARK();
        DCA_LAUNCH((plm_assemble_kernel<Q>), gridU, 256, 0, st,
                   d_x, p->d_G, gstride, p->nsplit, L, ld, p->lh, p->lJ, lo, hi,
                   d_dir, d_g, p->d_red, p->d_fxpart, nfx, p->d_ticket, (unsigned)p->n_active, p->d_scal);
    }
    PLM_MARK();
#undef PLM_MARK
    DCA_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------
// Peer arenas are recycled for the life of the process: cudaMalloc / cudaFree of an IPC-exported
// block and, on the peers, cudaIpcOpenMemHandle / CloseMemHandle cost 10-50 ms per objective at
// 8 GPUs (measured: the whole end-to-end call is ~30 ms).  A recycled arena keeps its address, so
// its IPC handle and every peer's mapping of it stay valid (capi.cu caches those by handle).
static std::mutex g_arena_mu;
static std::vector<std::pair<unsigned char *, size_t>> g_arena_pool;   // .second = bytes ^ (device << 56)

static size_t arena_key(size_t bytes)
{
    int dev = 0;
    cudaGetDevice(&dev);
    return bytes ^ ((size_t)dev << 56);
}

static unsigned char *arena_acquire(size_t bytes)
{
    {
        std::lock_guard<std::mutex> lk(g_arena_mu);
        for (size_t k = 0; k < g_arena_pool.size(); ++k)
            if (g_arena_pool[k].second == arena_key(bytes)) {
                unsigned char *ptr = g_arena_pool[k].first;
                g_arena_pool.erase(g_arena_pool.begin() + (long)k);
                return ptr;
            }
    }
    void *ptr = nullptr;
    if (cudaMalloc(&ptr, bytes) != cudaSuccess) {
        set_error("plm: cudaMalloc of the %zu-byte peer arena failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return static_cast<unsigned char *>(ptr);
}

static void arena_release(unsigned char *ptr, size_t bytes)
{
    std::lock_guard<std::mutex> lk(g_arena_mu);
    g_arena_pool.emplace_back(ptr, arena_key(bytes));
}

static void plm_free(PlmProblem *p)
{
    if (!p) return;
    if (p->stream) cudaStreamSynchronize(p->stream);
    dev_free(p->d_msa); dev_free(p->d_msaT); dev_free(p->d_perm); dev_free(p->d_w);
    dev_free(p->d_msaTs); dev_free(p->d_rowidx); dev_free(p->d_Rz4);
    if (!p->sharded) dev_free(p->d_W);
    dev_free(p->d_G); dev_free(p->d_S); dev_free(p->d_Rz); dev_free(p->d_fxpart);
    dev_free(p->d_fxsite); dev_free(p->d_red); dev_free(p->d_scal); dev_free(p->d_ticket); dev_free(p->d_scal_loc);
    if (p->arena) arena_release(p->arena, p->lay.bytes);   // no peer pushes any more: every rank synchronised before destroy
    if (p->h_scal) host_pinned_free(p->h_scal);
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
}

static void plm_plan_scan(PlmProblem *p)
{
    p->nscan = ceil_div<long>(p->N, p->chunk);
}

// byte offsets of the peer arena; identical on every rank (depends on L, q, the padded
// row length and the largest shard only)
static XrLayout xr_layout(int Lq, int ld, int max_sites, int q)
{
    auto up = [](size_t v) { return (v + 4095) & ~(size_t)4095; };
    XrLayout l;
    l.off_flags = 0;
    l.off_slots = 1024;
    l.off_W = up(l.off_slots + sizeof(double) * 2 * XR_MAX_RANKS * XR_SLOT);
    l.off_R = up(l.off_W + sizeof(float) * (size_t)Lq * ld);
    l.bytes = up(l.off_R + sizeof(float) * (size_t)max_sites * q * ld);
    return l;
}

static PlmProblem *plm_create_impl(const uint8_t *h_msa, const float *h_w, long N, int L, int q,
                                   float lambda_h, float lambda_J, int site_lo, int site_hi,
                                   int rank, int world, const int *bounds, bool asym = false)
{
    if (require_device() != 0) return nullptr;
    if (q != 21 && q != 5) { set_error("plm: num_site_states must be 21 or 5 (got %d)", q); return nullptr; }
    if (N <= 0 || L < 2) { set_error("plm: empty alignment (N=%ld L=%d)", N, L); return nullptr; }
    const bool sharded = (bounds != nullptr);
    if (sharded) {
        if (world < 1 || world > XR_MAX_RANKS || rank < 0 || rank >= world) {
            set_error("plm: bad rank %d / world %d (at most %d ranks)", rank, world, XR_MAX_RANKS); return nullptr;
        }
        if (bounds[0] != 0 || bounds[world] != L) { set_error("plm: shard bounds must run from 0 to L"); return nullptr; }
        for (int r = 0; r < world; ++r)
            if (bounds[r + 1] <= bounds[r]) { set_error("plm: every rank must own at least one site (rank %d owns none)", r); return nullptr; }
        site_lo = bounds[rank]; site_hi = bounds[rank + 1];
    }
    if (site_lo < 0 || site_hi > L || site_lo >= site_hi) {
        set_error("plm: bad site shard [%d,%d) for L=%d", site_lo, site_hi, L); return nullptr;
    }
    {
        const long k = first_bad_residue(h_msa, N * (long)L, q);
        if (k >= 0) { set_error("plm: residue code %d >= q=%d at flat index %ld", h_msa[k], q, k); return nullptr; }
    }

    PlmProblem *p = new PlmProblem();
    p->N = N; p->L = L; p->q = q; p->Lq = L * q;
    p->Npad = ceil_div<long>(N, LG_TN) * LG_TN;
    p->asym = asym;
    p->ncol = (site_hi - site_lo) * q;
    p->ld = asym ? ceil_div(p->ncol, LG_TC) * LG_TC : ceil_div(p->Lq, LG_TC) * LG_TC;
    p->P = (long)L * q + (long)L * (L - 1) / 2 * q * q;
    p->lh = lambda_h; p->lJ = lambda_J;
    p->site_lo = site_lo; p->site_hi = site_hi;
    p->nchunk = (int)ceil_div<long>(N, SC_CH);
    // column space of S / Rz / G / W: all sites' columns, or (asymmetric) my sites' columns from 0
    const int col_lo = asym ? 0 : site_lo * q, col_hi = asym ? p->ncol : site_hi * q;
    p->col_tile0 = col_lo / SC_TC;
    p->col_tiles = ceil_div(col_hi, SC_TC) - p->col_tile0;
    {   // logits tiling: every SM streams (CTAs it hosts) x (columns per CTA) through shared memory;
        // take the narrow 64-column CTAs only when they shorten that makespan (small grids)
        const long seq_tiles = p->Npad / LG_TN;
        const int sms = sm_count();
        long best = -1;
        for (int tc : {128, 64}) {
            const int t0 = col_lo / tc, nt = ceil_div(col_hi, tc) - t0;
            long cost = ceil_div<long>((long)nt * seq_tiles, sms) * tc;
            // grids of several waves: the wave model says little (the last partial wave spreads one CTA per SM), the
            // 64-column kernel is ~7 % slower per add and has no equal-residue fast path: only small grids take it
            // (measured on the sorted layout, 75-site shard = 520 CTAs of 128 columns: 0.152 ms against 0.185 ms)
            if (tc == 64 && (long)(ceil_div(col_hi, 128) - col_lo / 128) * seq_tiles >= 3L * sms) cost = best + 1;
            if (best < 0 || cost < best) { best = cost; p->lg_tc = tc; p->lg_tile0 = t0; p->lg_tiles = nt; }
        }
        if (const char *e = getenv("DCA_B200_LOGITS_TC")) {        // measurement override
            const int tc = atoi(e);
            if (tc == 64 || tc == 128) {
                p->lg_tc = tc; p->lg_tile0 = col_lo / tc; p->lg_tiles = ceil_div(col_hi, tc) - p->lg_tile0;
            }
        }
    }
    p->red_blocks = ceil_div(L, AS_JG) * L;
    p->n_active = 0;                 // assemble CTAs that own a pair or the fields of a site
    for (int i = 0; i < L; ++i)
        for (int bx = 0; bx < ceil_div(L, AS_JG); ++bx)
            p->n_active += (bx == 0 || bx * AS_JG + AS_JG - 1 > i) ? 1 : 0;
    // owner kernels (sharded): 8 sites j per CTA unless that leaves the GPU short of CTAs
    const int forced_jg = getenv("DCA_B200_OWN_JG") ? atoi(getenv("DCA_B200_OWN_JG")) : 0;    // tests / measurements
    for (int jg : {8, 2}) {
        if (forced_jg == 2 && jg == 8) continue;
        p->own_jg = jg;
        p->n_active_own = 0;
        for (int i = site_lo; i < site_hi; ++i)
            for (int bx = 0; bx < ceil_div(L, jg); ++bx)
                p->n_active_own += (bx == 0 || bx * jg + jg - 1 > i) ? 1 : 0;
        if (p->n_active_own >= 4 * sm_count() || forced_jg == 8) break;
    }
    p->red_blocks = std::max(p->red_blocks, ceil_div(L, 2) * (site_hi - site_lo));
    p->Ploc = p->P;
    if (asym) {
        p->Ploc = (long)p->ld * (p->Lq + 1);
        // finish kernel: enough CTAs to fill the GPU a few times over
        p->fin_rows = 8;
        while (p->fin_rows > 1 && (long)ceil_div(p->ld, 1024) * ceil_div(p->Lq, p->fin_rows) < 4L * sm_count()) p->fin_rows >>= 1;
        p->red_blocks = std::max(p->red_blocks, ceil_div(p->ld, 1024) * (ceil_div(p->Lq, p->fin_rows) + 8));
    }
    if (sharded) {
        p->sharded = true;
        p->peers.rank = rank; p->peers.world = world;
        p->max_sites = 0;
        for (int r = 0; r <= XR_MAX_RANKS; ++r) p->peers.bounds[r] = (r <= world) ? bounds[r] : L;
        for (int r = 0; r < world; ++r) p->max_sites = std::max(p->max_sites, bounds[r + 1] - bounds[r]);
        p->nq_loc = (long)(site_hi - site_lo) * q;
        p->pair0 = pair_index(L, site_lo, site_lo + 1);
        p->ncpl_loc = (pair_index(L, site_hi, site_hi + 1) - p->pair0) * q * q;
        p->Ploc = p->nq_loc + p->ncpl_loc;
        p->lay = xr_layout(p->Lq, p->ld, p->max_sites, q);
    }
    {   // scatter plan: balanced site tiles, and the n-split count that minimises the
        // makespan ceil(CTAs / SMs) / nsplit of equal-cost CTAs (1 CTA per SM), with a
        // small penalty per split for the extra G traffic the assemble kernel reads
        p->site_tiles = ceil_div(L, SC_WARPS);
        p->sites_per_tile = ceil_div(L, p->site_tiles);
        const long tiles = (long)p->col_tiles * p->site_tiles;
        const int sms = sm_count();
        double best = 1e300;
        int best_s = 1;
        for (int s = 1; s <= (sharded ? 16 : 8); ++s) {
            if (s > 1 && p->nchunk / s < 4) break;
            // every extra split is one more copy of G for export / assemble to sum; on the small
            // grids of a multi-GPU shard that costs more than a slightly longer scatter (measured
            // at 8 GPUs: 3 splits 2756 evaluations/s, 7 splits 2700)
            const double cost = (double)ceil_div<long>(tiles * s, sms) / s * (1.0 + (sharded ? 0.06 : 0.02) * s);
            if (cost < best - 1e-12) { best = cost; best_s = s; }
        }
        p->nsplit = best_s;
        if (const char *e = getenv("DCA_B200_NSPLIT")) {            // measurement override
            const int v = atoi(e);
            if (v >= 1 && v <= 16 && v <= p->nchunk) p->nsplit = v;
        }
        p->chunks_per_split = ceil_div(p->nchunk, p->nsplit);
        p->nsplit = ceil_div(p->nchunk, p->chunks_per_split);
    }
    float meff = 0.f;
    for (long n = 0; n < N; ++n) meff += h_w[n];     // plmdca_numerics.cpp:46
    p->h_meff = meff;
    // the carry scan is latency bound on small problems: shorter chains = more of them
    if ((long)(site_hi - site_lo) * (N / 32) < 50000) p->chunk = 16;
    plm_plan_scan(p);

#define PLM_TRY(call)                                                                 \
    do { cudaError_t e__ = (call); if (e__ != cudaSuccess) {                          \
        set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
        plm_free(p); return nullptr; } } while (0)

#define PLM_ALLOC(ptr, bytes)                                                         \
    do { *reinterpret_cast<void **>(&(ptr)) = dev_alloc(bytes);                       \
         if (!(ptr)) { plm_free(p); return nullptr; } } while (0)

    PLM_TRY(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    PLM_ALLOC(p->d_msa, (size_t)N * L);
    PLM_ALLOC(p->d_msaT, (size_t)L * p->Npad);
    {   // similarity-sorted layout.  Every rank of a multi-GPU objective sorts for itself: the order only concerns
        // its own S / Rz rows
        const char *e = getenv("DCA_B200_LAYOUT");           // file | sorted (default)
        p->want_sorted = N >= 256 && N < (1L << 24) && L < 2048 && !(e && strcmp(e, "file") == 0);
        if (p->want_sorted) {
            PLM_ALLOC(p->d_msaTs, (size_t)L * p->Npad);
            PLM_ALLOC(p->d_rowidx, sizeof(int) * (size_t)N);
            if (!(getenv("DCA_B200_GROUP4") && atoi(getenv("DCA_B200_GROUP4")) == 0)) {      // measurement switch
                PLM_ALLOC(p->d_Rz4, sizeof(float) * (size_t)(p->Npad / 4) * p->ld);
                PLM_TRY(cudaMemsetAsync(p->d_Rz4, 0, sizeof(float) * (size_t)(p->Npad / 4) * p->ld, p->stream));
                if (make_tensor_map_2d(&p->tmS4, 4, p->d_Rz4, (unsigned long long)p->ld, (unsigned long long)(p->Npad / 4),
                                       sizeof(float) * (unsigned long long)p->ld, SC_TC, SC_G4) != 0) { plm_free(p); return nullptr; }
            }
        }
    }
    PLM_ALLOC(p->d_perm, (size_t)p->site_tiles * p->sites_per_tile * p->nchunk * PERM_REC);
    PLM_ALLOC(p->d_w, sizeof(float) * (size_t)N);
    if (sharded) {
        // the peer arena is exported through CUDA IPC: a plain cudaMalloc of its own, never
        // a sub-allocation of the caching allocator
        p->arena = arena_acquire(p->lay.bytes);
        if (!p->arena) { plm_free(p); return nullptr; }
        PLM_TRY(cudaMemsetAsync(p->arena, 0, p->lay.bytes, p->stream));
        p->d_W = reinterpret_cast<float *>(p->arena + p->lay.off_W);
        p->d_R = reinterpret_cast<float *>(p->arena + p->lay.off_R);
        for (int r = 0; r < XR_MAX_RANKS; ++r) p->peers.arena[r] = nullptr;
        p->peers.arena[rank] = p->arena;
        p->connected = (world == 1);
        PLM_ALLOC(p->d_scal_loc, sizeof(double) * 8);
    } else if (!asym) {
        PLM_ALLOC(p->d_W, sizeof(float) * (size_t)p->Lq * p->ld);
        PLM_TRY(cudaMemsetAsync(p->d_W, 0, sizeof(float) * (size_t)p->Lq * p->ld, p->stream));
    }
    PLM_ALLOC(p->d_G, sizeof(float) * (size_t)p->nsplit * p->Lq * p->ld);
    PLM_ALLOC(p->d_S, sizeof(float) * (size_t)p->Npad * p->ld);
    PLM_ALLOC(p->d_Rz, sizeof(float) * (size_t)p->Npad * p->ld);
    PLM_ALLOC(p->d_fxpart, sizeof(double) * (size_t)(N / 8 + 1) * L);   // chunk >= 8
    PLM_ALLOC(p->d_red, sizeof(double) * 4 * (size_t)p->red_blocks);
    PLM_TRY(cudaMemsetAsync(p->d_red, 0, sizeof(double) * 4 * (size_t)p->red_blocks, p->stream));
    PLM_ALLOC(p->d_ticket, sizeof(unsigned) * 4);      // [0] assemble, [1] unpack_push, [2] export
    PLM_TRY(cudaMemsetAsync(p->d_ticket, 0, sizeof(unsigned) * 4, p->stream));
    PLM_ALLOC(p->d_scal, sizeof(double) * 16);
    PLM_TRY(cudaMemsetAsync(p->d_scal, 0, sizeof(double) * 16, p->stream));
    p->d_err = reinterpret_cast<unsigned *>(p->d_scal + PS_COUNT);       // read back with the scalars
    p->h_scal = static_cast<double *>(host_pinned_alloc(sizeof(double) * 16));
    if (!p->h_scal) { plm_free(p); return nullptr; }
    PLM_TRY(cudaMemcpyAsync(p->d_msa, h_msa, (size_t)N * L, cudaMemcpyHostToDevice, p->stream));
    PLM_TRY(cudaMemcpyAsync(p->d_w, h_w, sizeof(float) * (size_t)N, cudaMemcpyHostToDevice, p->stream));
    PLM_TRY(cudaMemsetAsync(p->d_S, 0, sizeof(float) * (size_t)p->Npad * p->ld, p->stream));
    PLM_TRY(cudaMemsetAsync(p->d_Rz, 0, sizeof(float) * (size_t)p->Npad * p->ld, p->stream));
    PLM_TRY(cudaMemsetAsync(p->d_fxpart, 0, sizeof(double) * (size_t)(N / 8 + 1) * L, p->stream));
    if (make_tensor_map_2d(&p->tmS, 4, p->d_Rz, (unsigned long long)p->ld, (unsigned long long)p->Npad,
                           sizeof(float) * (unsigned long long)p->ld, SC_TC, SC_CH) != 0) { plm_free(p); return nullptr; }
    {
        const unsigned jb = (q > 8) ? 4 : 16;         // LgCfg<Q>::JB
        if ((!asym && make_tensor_map_2d(&p->tmW, 4, p->d_W, (unsigned long long)p->ld, (unsigned long long)p->Lq,
                                         sizeof(float) * (unsigned long long)p->ld, (unsigned)p->lg_tc, jb * (unsigned)q) != 0) ||
            make_tensor_map_2d(&p->tmM, 1, p->d_msaT, (unsigned long long)p->Npad, (unsigned long long)L,
                               (unsigned long long)p->Npad, LG_TN, jb) != 0) { plm_free(p); return nullptr; }
    }
    const int rc = (q == 21) ? plm_setup_t<21>(p) : plm_setup_t<5>(p);
    if (rc != 0) { plm_free(p); return nullptr; }
    PLM_TRY(cudaStreamSynchronize(p->stream));
#undef PLM_TRY
#undef PLM_ALLOC
    return p;
}

PlmProblem *plm_create(const uint8_t *h_msa, const float *h_w, long N, int L, int q,
                       float lambda_h, float lambda_J, int site_lo, int site_hi)
{
    return plm_create_impl(h_msa, h_w, N, L, q, lambda_h, lambda_J, site_lo, site_hi, 0, 1, nullptr);
}

// asymmetric (per-site) estimator on the sites [site_lo, site_hi): see plm_asym_finish_kernel
PlmProblem *plm_create_asym(const uint8_t *h_msa, const float *h_w, long N, int L, int q,
                            float lambda_h, float lambda_J, int site_lo, int site_hi)
{
    return plm_create_impl(h_msa, h_w, N, L, q, lambda_h, lambda_J, site_lo, site_hi, 0, 1, nullptr, true);
}

bool plm_is_asym(const PlmProblem *p) { return p->asym; }

// my columns of the directed couplings and my fields into host arrays that hold ALL sites:
// W_full [Lq][ld_full] (column (i,a) = i*q + a), h_full [Lq]
int plm_asym_export(PlmProblem *p, const float *d_x, float *h_W_full, int ld_full, float *h_fields_full)
{
    if (!p->asym) { set_error("plm_asym_export: not an asymmetric objective"); return -1; }
    const size_t c0 = (size_t)p->site_lo * p->q;
    DCA_CUDA(cudaMemcpy2DAsync(h_W_full + c0, sizeof(float) * (size_t)ld_full, d_x + p->ld, sizeof(float) * (size_t)p->ld,
                               sizeof(float) * (size_t)p->ncol, (size_t)p->Lq, cudaMemcpyDeviceToHost, p->stream));
    DCA_CUDA(cudaMemcpyAsync(h_fields_full + c0, d_x, sizeof(float) * (size_t)p->ncol, cudaMemcpyDeviceToHost, p->stream));
    DCA_CUDA(cudaStreamSynchronize(p->stream));
    return 0;
}

// directed copies of all sites -> the reference's parameter vector (symmetrised couplings)
int plm_asym_pack_host(const float *h_W_full, int ld_full, const float *h_fields_full, int L, int q, float *h_xref)
{
    if (require_device() != 0) return -1;
    if (q != 21 && q != 5) { set_error("plm_asym_pack: num_site_states must be 21 or 5 (got %d)", q); return -1; }
    const long Lq = (long)L * q, P = Lq + (long)L * (L - 1) / 2 * q * q;
    DevBuf<float> W, h, x;
    if (!W.alloc((size_t)Lq * ld_full) || !h.alloc((size_t)Lq) || !x.alloc((size_t)P)) return -1;
    DCA_CUDA(cudaMemcpy(W, h_W_full, sizeof(float) * (size_t)Lq * ld_full, cudaMemcpyHostToDevice));
    DCA_CUDA(cudaMemcpy(h, h_fields_full, sizeof(float) * (size_t)Lq, cudaMemcpyHostToDevice));
    const unsigned blocks = (unsigned)ceil_div<long>(P, 256);
    if (q == 21) DCA_LAUNCH((plm_asym_pack_kernel<21>), blocks, 256, 0, 0, W.p, ld_full, h.p, L, x.p);
    else DCA_LAUNCH((plm_asym_pack_kernel<5>), blocks, 256, 0, 0, W.p, ld_full, h.p, L, x.p);
    DCA_CUDA(cudaGetLastError());
    DCA_CUDA(cudaMemcpy(h_xref, x, sizeof(float) * (size_t)P, cudaMemcpyDeviceToHost));
    return 0;
}

PlmProblem *plm_create_sharded(const uint8_t *h_msa, const float *h_w, long N, int L, int q,
                               float lambda_h, float lambda_J, int rank, int world, const int *bounds)
{
    if (!bounds) { set_error("plm: shard bounds missing"); return nullptr; }
    return plm_create_impl(h_msa, h_w, N, L, q, lambda_h, lambda_J, 0, 0, rank, world, bounds);
}

// peers: device pointers of every rank's arena valid in THIS process (IPC-opened, or raw
// when the ranks share a process); the own entry is ignored
int plm_connect(PlmProblem *p, void *const *arenas, bool shared_process)
{
    p->shared_process = shared_process;
    if (!p->sharded) { set_error("plm_connect: not a sharded objective"); return -1; }
    for (int r = 0; r < p->peers.world; ++r) {
        if (r == p->peers.rank) continue;
        if (!arenas[r]) { set_error("plm_connect: arena of rank %d missing", r); return -1; }
        p->peers.arena[r] = static_cast<unsigned char *>(arenas[r]);
    }
    p->connected = true;
    return 0;
}

void plm_disconnect(PlmProblem *p)
{
    if (!p->sharded) return;
    cudaStreamSynchronize(p->stream);
    for (int r = 0; r < p->peers.world; ++r)
        if (r != p->peers.rank) p->peers.arena[r] = nullptr;
    p->connected = (p->peers.world == 1);
}

void *plm_arena(const PlmProblem *p) { return p->arena; }
size_t plm_arena_bytes(const PlmProblem *p) { return p->sharded ? p->lay.bytes : 0; }
bool plm_is_sharded(const PlmProblem *p) { return p->sharded; }
long plm_local_params(const PlmProblem *p) { return p->Ploc; }

// where the local slice sits in the reference's parameter vector:
// out = {field offset, field count, coupling offset, coupling count}
void plm_local_layout(const PlmProblem *p, long *out)
{
    if (!p->sharded) { out[0] = 0; out[1] = p->Lq; out[2] = p->Lq; out[3] = p->P - p->Lq; return; }
    out[0] = (long)p->site_lo * p->q; out[1] = p->nq_loc;
    out[2] = (long)p->Lq + p->pair0 * p->q * p->q; out[3] = p->ncpl_loc;
}

// align the ranks (flag barrier on the stream); no-op unless sharded
int plm_xr_barrier(PlmProblem *p)
{
    if (!p->sharded) return 0;
    if (!p->connected) { set_error("plm: sharded objective used before its peers were connected"); return -1; }
    DCA_LAUNCH(xr_barrier_kernel, 1, 32, 0, p->stream, p->peers, p->lay.off_flags, ++p->epoch, p->d_err);
    DCA_CUDA(cudaGetLastError());
    return 0;
}

XrCall plm_xr_next(PlmProblem *p)
{
    XrCall c;
    c.P = p->peers; c.lay = p->lay; c.epoch = ++p->epoch; c.seq = p->red_seq++; c.err = p->d_err;
    return c;
}

void plm_destroy(PlmProblem *p) { plm_free(p); }
long plm_num_params(const PlmProblem *p) { return p->P; }
cudaStream_t plm_stream(const PlmProblem *p) { return p->stream; }
double *plm_scalar_buffer(PlmProblem *p) { return p->d_scal; }

void plm_set_carry(PlmProblem *p, int carry, int chunk, int warm)
{
    p->carry = carry ? 1 : 0;
    if (chunk > 0) p->chunk = std::max(8, chunk);
    if (warm >= 0) p->warm = warm;
    plm_plan_scan(p);
}

// reference-arithmetic mode: the softmax kernel also keeps the per-site
// partials of fx, which seqsum.cu adds in the reference's order
int plm_set_fidelity(PlmProblem *p, int on)
{
    if (on && p->sharded) { set_error("plm: the reference-arithmetic mode runs on one GPU (sequential sums span the whole vector)"); return -1; }
    if (on && p->asym) { set_error("plm: the reference-arithmetic mode reproduces the reference's joint fit; the asymmetric estimator has no reference"); return -1; }
    if ((on != 0) == p->sorted && p->want_sorted) {
        // the reference-arithmetic fit reproduces the reference's trajectory: keep the order of the sums inside the
        // scatter (rows of a chunk in file order), i.e. the file-order layout; back to the sorted one when it is
        // switched off again
        if (((p->q == 21) ? plm_layout_t<21>(p, !on) : plm_layout_t<5>(p, !on)) != 0) return -1;
        DCA_CUDA(cudaStreamSynchronize(p->stream));
    }
    if (on && !p->d_fxsite) {
        const size_t count = (size_t)(p->N / 8 + 1) * p->L;          // chunk >= 8
        if (!dev_alloc_t(&p->d_fxsite, count)) return -1;
        DCA_CUDA(cudaMemsetAsync(p->d_fxsite, 0, sizeof(double) * count, p->stream));
    }
    p->fidelity = on ? 1 : 0;
    return 0;
}
int plm_fidelity(const PlmProblem *p) { return p->fidelity; }
void plm_fx_layout(const PlmProblem *p, const double **fxsite, long *nscan, int *L, int *q, float *lh, float *lJ)
{
    *fxsite = p->d_fxsite; *nscan = p->nscan; *L = p->L; *q = p->q; *lh = p->lh; *lJ = p->lJ;
}

int plm_eval_device(PlmProblem *p, const float *d_x, float *d_g, const float *d_dir)
{
    return (p->q == 21) ? plm_eval_t<21>(p, d_x, d_g, d_dir) : plm_eval_t<5>(p, d_x, d_g, d_dir);
}

// one evaluation with CUDA events between the kernels, on the stream they are
// launched on: ms[0..5] = unpack (+push), logits (waits for the peers' pushes), softmax, scatter,
// export (sharded only), assemble + finalize (waits for the peers' exports; cross-rank scalar sum)
int plm_eval_timed(PlmProblem *p, const float *d_x, float *d_g, float *ms)
{
    cudaEvent_t ev[PLM_NUM_KERNEL_TIMES + 1];
    for (auto &e : ev) DCA_CUDA(cudaEventCreate(&e));
    const int rc = (p->q == 21) ? plm_eval_t<21>(p, d_x, d_g, nullptr, ev) : plm_eval_t<5>(p, d_x, d_g, nullptr, ev);
    DCA_CUDA(cudaStreamSynchronize(p->stream));
    for (int k = 0; k < PLM_NUM_KERNEL_TIMES; ++k) cudaEventElapsedTime(&ms[k], ev[k], ev[k + 1]);
    for (auto &e : ev) cudaEventDestroy(e);
    return rc;
}

// layout statistics, for the bench's roofline arithmetic: {sorted layout in use, scatter uses group sums, fraction of
// (site, 4-row group) pairs with one residue, logits kernel columns}
void plm_layout_stats(const PlmProblem *p, double *out4)
{
    out4[0] = p->sorted ? 1.0 : 0.0; out4[1] = p->group4 ? 1.0 : 0.0; out4[2] = p->uniform4; out4[3] = (double)p->lg_tc;
}

// static launch geometry, for the bench's roofline arithmetic
void plm_describe(const PlmProblem *p, long *out)
{
    out[0] = p->N; out[1] = p->L; out[2] = p->q; out[3] = p->ld; out[4] = p->P;
    out[5] = p->nsplit; out[6] = p->nchunk; out[7] = p->col_tiles; out[8] = p->Npad; out[9] = p->nscan;
}

// the PS_COUNT scalars plus the peer-timeout flag behind them: one small D2H + stream sync
int plm_read_scalars(PlmProblem *p, double *out, double *dginit)
{
    // Ranks sharing one process (and one GPU) share its copy-engine queues: a D2H copy parked
    // behind my spinning barrier kernel would block the peer's H2D copies, i.e. the very work
    // the barrier waits for.  Drain the stream first in that (test-only) configuration.
    if (p->shared_process) DCA_CUDA(cudaStreamSynchronize(p->stream));
    DCA_CUDA(cudaMemcpyAsync(p->h_scal, p->d_scal, sizeof(double) * (PS_DGINIT + 1), cudaMemcpyDeviceToHost, p->stream));
    DCA_CUDA(cudaStreamSynchronize(p->stream));
    for (int k = 0; k < PS_COUNT; ++k) out[k] = p->h_scal[k];
    if (dginit) *dginit = p->h_scal[PS_DGINIT];
    if (*reinterpret_cast<const unsigned *>(p->h_scal + PS_COUNT) != 0u) {
        set_error("plm: a peer GPU did not reach the exchange within %.0f s (rank %d of %d)",
                  XR_TIMEOUT_NS * 1e-9, p->peers.rank, p->peers.world);
        return -1;
    }
    return 0;
}

// initial point of the whole vector, or of the local slice of a sharded objective
int plm_init_x_device(PlmProblem *p, float *d_x)
{
    DCA_CUDA(cudaMemsetAsync(d_x, 0, sizeof(float) * (size_t)p->Ploc, p->stream));
    const bool local = p->sharded || p->asym;
    const int lo = local ? p->site_lo : 0, hi = local ? p->site_hi : p->L;
    if (p->q == 21)
        DCA_LAUNCH((plm_init_fields_kernel<21>), ceil_div(hi - lo, INIT_SITES), 128, 0, p->stream,
                   p->d_msaT, p->d_w, p->N, p->Npad, lo, hi, p->h_meff, d_x);
    else
        DCA_LAUNCH((plm_init_fields_kernel<5>), ceil_div(hi - lo, INIT_SITES), 128, 0, p->stream,
                   p->d_msaT, p->d_w, p->N, p->Npad, lo, hi, p->h_meff, d_x);
    DCA_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace dca
