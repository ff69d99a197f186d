// align.cu -- reference-sequence matching: local alignment SCORE of one reference sequence
// against every sequence of the alignment (gaps removed), all N in parallel.
//
// Replaces the search loop of SequenceBackmapper.find_matching_seqs_from_alignment
// (pydca/sequence_backmapper/sequence_backmapper.py:228-283), which calls Biopython's
// pairwise2.align.localds(ref, seq, matrix, open, extend, score_only=True) once per sequence
// (:188-226): BLOSUM62 with gap open -10 / extend -1 for proteins, the module's NUC44 table
// (scoring_matrix.py: +5 match, -4 mismatch on ACGU) with -8 / 0 for RNA.  Biopython is a
// third-party dependency absent from /root/reference and from this image (setup.py pins
// biopython==1.74): the score is the published Smith-Waterman recurrence with Gotoh's affine
// gaps, a gap of length k costing open + (k-1)*extend (pairwise2's default,
// penalize_extend_when_opening=False):
//   E(i,j) = max(H(i,j-1) + open, E(i,j-1) + extend)       gap in the reference
//   F(i,j) = max(H(i-1,j) + open, F(i-1,j) + extend)       gap in the sequence
//   H(i,j) = max(0, H(i-1,j-1) + s(a_i,b_j), E(i,j), F(i,j)) ;  score = max H
// Integer arithmetic: bit-exact by construction.  One thread per alignment row; the two DP
// rows over the reference live in per-thread local memory (int16: |score| <= 11 * 1024).
#include "dca_common.cuh"
#include "dca_internal.h"

namespace dca {

namespace {

constexpr int SW_MAX_REF = 1024;

// BLOSUM62 restricted to the 20 standard residues, in the alphabet order of the readers
// (ACDEFGHIKLMNPQRSTVWY, plmdca_numerics.cpp:708-717).  Public NCBI table.
__constant__ signed char c_sub[21 * 21];

const signed char BLOSUM62_ACDEFGHIKLMNPQRSTVWY[20][20] = {
    /*A*/ { 4,  0, -2, -1, -2,  0, -2, -1, -1, -1, -1, -2, -1, -1, -1,  1,  0,  0, -3, -2},
    /*C*/ { 0,  9, -3, -4, -2, -3, -3, -1, -3, -1, -1, -3, -3, -3, -3, -1, -1, -1, -2, -2},
    /*D*/ {-2, -3,  6,  2, -3, -1, -1, -3, -1, -4, -3,  1, -1,  0, -2,  0, -1, -3, -4, -3},
    /*E*/ {-1, -4,  2,  5, -3, -2,  0, -3,  1, -3, -2,  0, -1,  2,  0,  0, -1, -2, -3, -2},
    /*F*/ {-2, -2, -3, -3,  6, -3, -1,  0, -3,  0,  0, -3, -4, -3, -3, -2, -2, -1,  1,  3},
    /*G*/ { 0, -3, -1, -2, -3,  6, -2, -4, -2, -4, -3,  0, -2, -2, -2,  0, -2, -3, -2, -3},
    /*H*/ {-2, -3, -1,  0, -1, -2,  8, -3, -1, -3, -2,  1, -2,  0,  0, -1, -2, -3, -2,  2},
    /*I*/ {-1, -1, -3, -3,  0, -4, -3,  4, -3,  2,  1, -3, -3, -3, -3, -2, -1,  3, -3, -1},
    /*K*/ {-1, -3, -1,  1, -3, -2, -1, -3,  5, -2, -1,  0, -1,  1,  2,  0, -1, -2, -3, -2},
    /*L*/ {-1, -1, -4, -3,  0, -4, -3,  2, -2,  4,  2, -3, -3, -2, -2, -2, -1,  1, -2, -1},
    /*M*/ {-1, -1, -3, -2,  0, -3, -2,  1, -1,  2,  5, -2, -2,  0, -1, -1, -1,  1, -1, -1},
    /*N*/ {-2, -3,  1,  0, -3,  0,  1, -3,  0, -3, -2,  6, -2,  0,  0,  1,  0, -3, -4, -2},
    /*P*/ {-1, -3, -1, -1, -4, -2, -2, -3, -1, -3, -2, -2,  7, -1, -2, -1, -1, -2, -4, -3},
    /*Q*/ {-1, -3,  0,  2, -3, -2,  0, -3,  1, -2,  0,  0, -1,  5,  1,  0, -1, -2, -2, -1},
    /*R*/ {-1, -3, -2,  0, -3, -2,  0, -3,  2, -2, -1,  0, -2,  1,  5, -1, -1, -3, -3, -2},
    /*S*/ { 1, -1,  0,  0, -2,  0, -1, -2,  0, -2, -1,  1, -1,  0, -1,  4,  1, -2, -3, -2},
    /*T*/ { 0, -1, -1, -1, -2, -2, -2, -1, -1, -1, -1,  0, -1, -1, -1,  1,  5,  0, -2, -2},
    /*V*/ { 0, -1, -3, -2, -1, -3, -3,  3, -2,  1,  1, -3, -2, -2, -3, -2,  0,  4, -3, -1},
    /*W*/ {-3, -2, -4, -3,  1, -2, -2, -3, -3, -2, -1, -4, -4, -2, -3, -3, -2, -3, 11,  2},
    /*Y*/ {-2, -2, -3, -2,  3, -3,  2, -1, -2, -1, -1, -2, -3, -1, -2, -2, -2, -1,  2,  7},
};

__global__ void __launch_bounds__(128)
sw_score_kernel(const uint8_t *__restrict__ msa, long N, int L, int q, const uint8_t *__restrict__ ref, int ref_len,
                int gap_open, int gap_extend, int *__restrict__ scores)
{
    __shared__ uint8_t sref[SW_MAX_REF];
    for (int k = threadIdx.x; k < ref_len; k += blockDim.x) sref[k] = ref[k];
    __syncthreads();
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    short H[SW_MAX_REF + 1], E[SW_MAX_REF + 1];
    for (int i = 0; i <= ref_len; ++i) { H[i] = 0; E[i] = -30000; }
    const int gap = q - 1;
    int best = 0;
    const uint8_t *row = msa + n * (long)L;
    for (int j = 0; j < L; ++j) {
        const int b = row[j];
        if (b == gap) continue;                             // seq.replace('-', '') (:262)
        const signed char *srow = c_sub + b * 21;
        int diag = 0, left = 0, f = -30000;
        for (int i = 1; i <= ref_len; ++i) {
            const int up = H[i];
            const int e = max(up + gap_open, (int)E[i] + gap_extend);
            f = max(left + gap_open, f + gap_extend);
            int h = diag + srow[sref[i - 1]];
            h = max(max(h, 0), max(e, f));
            diag = up;
            left = h;
            H[i] = (short)h;
            E[i] = (short)max(e, -30000);
            best = max(best, h);
        }
    }
    scores[n] = best;
}

}  // namespace

// scores_out[n] = local alignment score of ref (codes 0..q-2) against row n of the alignment
// with its gaps removed.  msa / ref are HOST buffers.
int sw_scores_host(const uint8_t *msa, long N, int L, int q, const uint8_t *ref, int ref_len, int *scores_out)
{
    if (require_device() != 0) return -1;
    if (!msa || !ref || !scores_out || N <= 0 || L <= 0 || ref_len <= 0) { set_error("sw_scores: bad arguments"); return -1; }
    if (q != 21 && q != 5) { set_error("sw_scores: num_site_states must be 21 or 5"); return -1; }
    if (ref_len > SW_MAX_REF) { set_error("sw_scores: reference sequence longer than %d residues", SW_MAX_REF); return -1; }
    for (int k = 0; k < ref_len; ++k)
        if (ref[k] >= q - 1) { set_error("sw_scores: the reference sequence must consist of standard residues"); return -1; }
    signed char sub[21 * 21];
    memset(sub, 0, sizeof sub);
    for (int a = 0; a < q - 1; ++a)
        for (int b = 0; b < q - 1; ++b)
            sub[a * 21 + b] = (q == 21) ? BLOSUM62_ACDEFGHIKLMNPQRSTVWY[a][b] : (signed char)(a == b ? 5 : -4);
    DCA_CUDA(cudaMemcpyToSymbol(c_sub, sub, sizeof sub));
    const int gap_open = (q == 21) ? -10 : -8, gap_extend = (q == 21) ? -1 : 0;      // sequence_backmapper.py:207-214
    DevBuf<uint8_t> d_msa, d_ref;
    DevBuf<int> d_scores;
    if (!d_msa.alloc((size_t)N * L) || !d_ref.alloc((size_t)ref_len) || !d_scores.alloc((size_t)N)) return -1;
    DCA_CUDA(cudaMemcpy(d_msa, msa, (size_t)N * L, cudaMemcpyHostToDevice));
    DCA_CUDA(cudaMemcpy(d_ref, ref, (size_t)ref_len, cudaMemcpyHostToDevice));
    DCA_LAUNCH(sw_score_kernel, (unsigned)ceil_div<long>(N, 128), 128, 0, 0, d_msa.p, N, L, q, d_ref.p, ref_len,
               gap_open, gap_extend, d_scores.p);
    DCA_CUDA(cudaGetLastError());
    DCA_CUDA(cudaMemcpy(scores_out, d_scores.p, sizeof(int) * (size_t)N, cudaMemcpyDeviceToHost));
    return 0;
}

}  // namespace dca
