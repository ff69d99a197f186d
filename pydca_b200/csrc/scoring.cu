// scoring.cu -- Frobenius-norm + average-product-correction scoring.
//
// Replaces  PlmDCA.get_couplings_no_gap_state / compute_sorted_FN / _FN_APC
//                      pydca/plmdca/plmdca.py:246-268, :437-524
//           MeanFieldDCA.compute_sorted_FN / _FN_APC
//                      pydca/meanfield_dca/meanfield_dca.py:902-988
//
// Per site pair (i<j): take the (q-1)x(q-1) non-gap block of J_ij, subtract row
// and column means, add the grand mean, Frobenius norm.  One warp per pair, the
// block parked in shared memory, sums in float64 in a fixed order.  APC
// s_ij - s_i. s_.j / s_.. needs only L(L-1)/2 numbers and is finished on the
// host in float64 (the reference's pinned numpy also promotes to float64 there).
// HBM-bound: 4*q*q (plm, f32) or 8*(q-1)^2 (mf, f64) bytes per pair, read once.
#include "dca_common.cuh"
#include "dca_internal.h"

#include <vector>

namespace dca {

constexpr int FN_WARPS = 4;

// SRC = float : block at x[L*q + pair*q*q + a*q + b]          (plm layout)
// SRC = double: block at J[(i*(q-1)+a)*n + j*(q-1)+b], n=L(q-1) (mf layout)
template <typename SRC>
__global__ void __launch_bounds__(FN_WARPS * 32)
fn_kernel(const SRC *__restrict__ src, int L, int q, int is_matrix, long ldm, double *__restrict__ fn)
{
    extern __shared__ double sm[];
    const int qm1 = q - 1, nb = qm1 * qm1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *blk = sm + (size_t)warp * (nb + 2 * qm1);
    double *rmean = blk + nb, *cmean = rmean + qm1;
    const long npairs = (long)L * (L - 1) / 2;
    const long pair = (long)blockIdx.x * FN_WARPS + warp;
    if (pair >= npairs) return;
    // invert pair -> (i,j)
    int i = 0;
    {
        long rem = pair;
        while (rem >= L - 1 - i) { rem -= L - 1 - i; ++i; }
        const int j = i + 1 + (int)rem;
        for (int t = lane; t < nb; t += 32) {
            const int a = t / qm1, b = t - a * qm1;
            double v;
            if (is_matrix) v = (double)src[((long)i * qm1 + a) * ldm + (long)j * qm1 + b];
            else v = (double)src[(long)L * q + pair * q * q + a * q + b];
            blk[t] = v;
        }
    }
    __syncwarp();
    for (int a = lane; a < qm1; a += 32) {
        double r = 0.0, c = 0.0;
        for (int b = 0; b < qm1; ++b) { r += blk[a * qm1 + b]; c += blk[b * qm1 + a]; }
        rmean[a] = r / qm1;
        cmean[a] = c / qm1;
    }
    __syncwarp();
    double tot = 0.0;
    for (int a = 0; a < qm1; ++a) tot += rmean[a];
    const double gmean = tot / qm1;
    double acc = 0.0;
    for (int t = lane; t < nb; t += 32) {
        const int a = t / qm1, b = t - a * qm1;
        const double z = blk[t] - cmean[b] - rmean[a] + gmean;
        acc += z * z;
    }
    acc = warp_sum(acc);
    if (lane == 0) fn[pair] = sqrt(acc);
}

// APC on the host: s_ij - mean_i * mean_j / mean_all (plmdca.py:503-519)
void apc_host(const double *fn, int L, double *out)
{
    std::vector<double> site((size_t)L, 0.0);
    long k = 0;
    for (int i = 0; i < L - 1; ++i)
        for (int j = i + 1; j < L; ++j, ++k) { site[i] += fn[k]; site[j] += fn[k]; }
    double all = 0.0;
    for (int i = 0; i < L; ++i) { site[i] /= (double)(L - 1); all += site[i]; }
    all /= (double)L;
    k = 0;
    for (int i = 0; i < L - 1; ++i)
        for (int j = i + 1; j < L; ++j, ++k) out[k] = fn[k] - site[i] * (site[j] / all);
}

template <typename SRC>
static int fn_apc_impl(const SRC *h_src, size_t count, int L, int q, int is_matrix,
                       double *fn, double *fn_apc)
{
    if (require_device() != 0) return -1;
    if (L < 2 || q < 2) { set_error("fn_apc: bad shape L=%d q=%d", L, q); return -1; }
    const long npairs = (long)L * (L - 1) / 2;
    DevBuf<SRC> src_buf;                       // cached blocks, released on every return path
    DevBuf<double> fn_buf;
    if (!src_buf.alloc(count) || !fn_buf.alloc((size_t)npairs)) return -1;
    SRC *d_src = src_buf.p;
    double *d_fn = fn_buf.p;
    DCA_CUDA(cudaMemcpy(d_src, h_src, sizeof(SRC) * count, cudaMemcpyHostToDevice));
    const int qm1 = q - 1;
    const size_t smem = sizeof(double) * FN_WARPS * (size_t)(qm1 * qm1 + 2 * qm1);
    DCA_LAUNCH((fn_kernel<SRC>), (unsigned)ceil_div<long>(npairs, FN_WARPS), FN_WARPS * 32, smem, 0,
               d_src, L, q, is_matrix, (long)L * (q - 1), d_fn);
    DCA_CUDA(cudaGetLastError());
    DCA_CUDA(cudaMemcpy(fn, d_fn, sizeof(double) * (size_t)npairs, cudaMemcpyDeviceToHost));
    if (fn_apc) apc_host(fn, L, fn_apc);
    return 0;
}

// ------------------------------------------------------------------------------
// Direct information: two-site-model fixed point + DI per pair.
//   compute_two_site_model_fields  pydca/meanfield_dca/msa_numerics.py:378-442
//                                  pydca/plmdca/msa_numerics.py:156-220
//   compute_direct_info            :446-533 / :224-311
// One warp per pair, lane a owns state a (q <= 32); W = exp(J_ij) with the gap
// row/column at exp(0) = 1 (slice_couplings zero-fills them, :346-374).
constexpr int DI_WARPS = 4;

template <typename SRC>
__global__ void __launch_bounds__(DI_WARPS * 32)
di_kernel(const SRC *__restrict__ src, const double *__restrict__ reg_fi, int L, int q, int is_matrix, long ldm,
          double *__restrict__ di, double *__restrict__ fields)
{
    extern __shared__ double sm[];
    const int qm1 = q - 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *Wm = sm + (size_t)warp * (q * q + 2 * q);
    double *hi_s = Wm + q * q, *hj_s = hi_s + q;
    const long npairs = (long)L * (L - 1) / 2;
    const long pair = (long)blockIdx.x * DI_WARPS + warp;
    if (pair >= npairs) return;
    int i = 0;
    long rem = pair;
    while (rem >= L - 1 - i) { rem -= L - 1 - i; ++i; }
    const int j = i + 1 + (int)rem;
    for (int t = lane; t < q * q; t += 32) {
        const int a = t / q, b = t - a * q;
        double v = 0.0;
        if (a < qm1 && b < qm1) {
            if (is_matrix) v = (double)src[((long)i * qm1 + a) * ldm + (long)j * qm1 + b];
            else v = (double)src[(long)L * q + pair * q * q + a * q + b];
        }
        Wm[t] = exp(v);
    }
    const bool on = lane < q;
    const double fi = on ? reg_fi[i * q + lane] : 0.0;
    const double fj = on ? reg_fi[j * q + lane] : 0.0;
    double hi = 1.0 / (double)q, hj = 1.0 / (double)q;
    if (on) { hi_s[lane] = hi; hj_s[lane] = hj; }
    __syncwarp();
    double change = 10.0;
    while (change > 1.0e-4) {                                   // TOLERANCE, :411
        double xi = 0.0, xj = 0.0;
        if (on) {
            for (int b = 0; b < q; ++b) { xi += Wm[lane * q + b] * hj_s[b]; xj += Wm[b * q + lane] * hi_s[b]; }
        }
        double ni = on ? fi / xi : 0.0, nj = on ? fj / xj : 0.0;
        const double si = warp_sum(ni), sj = warp_sum(nj);
        ni /= si; nj /= sj;
        double c = on ? fmax(fabs(ni - hi), fabs(nj - hj)) : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c = fmax(c, __shfl_xor_sync(0xffffffffu, c, o));
        change = c;
        hi = ni; hj = nj;
        __syncwarp();
        if (on) { hi_s[lane] = hi; hj_s[lane] = hj; }
        __syncwarp();
    }
    if (fields && on) {
        fields[(pair * 2 + 0) * q + lane] = hi;
        fields[(pair * 2 + 1) * q + lane] = hj;
    }
    // pdir = W * hi hj^T / sum ; DI over residue-residue entries with EPSILON 1e-20 (:499-529)
    double z = 0.0;
    if (on) for (int b = 0; b < q; ++b) z += Wm[lane * q + b] * hi * hj_s[b];
    z = warp_sum(z);
    double acc = 0.0;
    if (lane < qm1) {
        for (int b = 0; b < qm1; ++b) {
            const double pd = Wm[lane * q + b] * hi * hj_s[b] / z + 1.0e-20;
            const double ff = fi * reg_fi[j * q + b] + 1.0e-20;
            acc += pd * log(pd / ff);
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) di[pair] = acc;
}

template <typename SRC>
static int di_impl(const SRC *h_src, size_t count, const double *h_reg_fi, int L, int q, int is_matrix,
                   double *di, double *fields)
{
    if (require_device() != 0) return -1;
    if (L < 2 || q < 2 || q > 32) { set_error("direct_info: bad shape L=%d q=%d", L, q); return -1; }
    const long npairs = (long)L * (L - 1) / 2;
    DevBuf<SRC> src_buf;                       // cached blocks, released on every return path
    DevBuf<double> rfi_buf, di_buf, fields_buf;
    if (!src_buf.alloc(count) || !rfi_buf.alloc((size_t)L * q) || !di_buf.alloc((size_t)npairs) ||
        (fields && !fields_buf.alloc((size_t)npairs * 2 * q)))
        return -1;
    SRC *d_src = src_buf.p;
    double *d_rfi = rfi_buf.p, *d_di = di_buf.p, *d_fields = fields_buf.p;
    DCA_CUDA(cudaMemcpy(d_src, h_src, sizeof(SRC) * count, cudaMemcpyHostToDevice));
    DCA_CUDA(cudaMemcpy(d_rfi, h_reg_fi, sizeof(double) * (size_t)L * q, cudaMemcpyHostToDevice));
    const size_t smem = sizeof(double) * DI_WARPS * (size_t)(q * q + 2 * q);
    DCA_LAUNCH((di_kernel<SRC>), (unsigned)ceil_div<long>(npairs, DI_WARPS), DI_WARPS * 32, smem, 0,
               d_src, d_rfi, L, q, is_matrix, (long)L * (q - 1), d_di, d_fields);
    DCA_CUDA(cudaGetLastError());
    DCA_CUDA(cudaMemcpy(di, d_di, sizeof(double) * (size_t)npairs, cudaMemcpyDeviceToHost));
    if (fields) DCA_CUDA(cudaMemcpy(fields, d_fields, sizeof(double) * (size_t)npairs * 2 * q, cudaMemcpyDeviceToHost));
    return 0;
}

int plm_di_host(const float *h_x, const double *h_reg_fi, int L, int q, double *di, double *fields)
{
    const size_t P = (size_t)L * q + (size_t)L * (L - 1) / 2 * q * q;
    return di_impl<float>(h_x, P, h_reg_fi, L, q, 0, di, fields);
}

int mf_di_host(const double *h_J, const double *h_reg_fi, int L, int q, double *di, double *fields)
{
    const size_t n = (size_t)L * (q - 1);
    return di_impl<double>(h_J, n * n, h_reg_fi, L, q, 1, di, fields);
}

// device-resident variants used by the fused mean-field pipeline (mf.cu): J is a
// row-major matrix with leading dimension ldm on the device
int fn_matrix_device(const double *d_J, long ldm, int L, int q, double *d_fn, cudaStream_t st)
{
    const long npairs = (long)L * (L - 1) / 2;
    const int qm1 = q - 1;
    const size_t smem = sizeof(double) * FN_WARPS * (size_t)(qm1 * qm1 + 2 * qm1);
    DCA_LAUNCH((fn_kernel<double>), (unsigned)ceil_div<long>(npairs, FN_WARPS), FN_WARPS * 32, smem, st,
               d_J, L, q, 1, ldm, d_fn);
    DCA_CUDA(cudaGetLastError());
    return 0;
}

int di_matrix_device(const double *d_J, long ldm, const double *d_rfi, int L, int q, double *d_di, cudaStream_t st)
{
    const long npairs = (long)L * (L - 1) / 2;
    const size_t smem = sizeof(double) * DI_WARPS * (size_t)(q * q + 2 * q);
    DCA_LAUNCH((di_kernel<double>), (unsigned)ceil_div<long>(npairs, DI_WARPS), DI_WARPS * 32, smem, st,
               d_J, d_rfi, L, q, 1, ldm, d_di, (double *)nullptr);
    DCA_CUDA(cudaGetLastError());
    return 0;
}

int plm_fn_apc_host(const float *h_x, int L, int q, double *fn, double *fn_apc)
{
    const size_t P = (size_t)L * q + (size_t)L * (L - 1) / 2 * q * q;
    return fn_apc_impl<float>(h_x, P, L, q, 0, fn, fn_apc);
}

int mf_fn_apc_host(const double *h_J, int L, int q, double *fn, double *fn_apc)
{
    const size_t n = (size_t)L * (q - 1);
    return fn_apc_impl<double>(h_J, n * n, L, q, 1, fn, fn_apc);
}

}  // namespace dca
