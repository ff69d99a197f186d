// dca_common.cuh -- shared helpers for the sm_100a DCA kernels.
#pragma once
#include <cuda.h>
#include <atomic>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

namespace dca {

// ---- error channel ---------------------------------------------------------
// The reference has none (exceptions cross the C ABI and abort); we
// never throw across the ABI: every entry point records a message and returns
// a non-zero status / NULL.
void set_error(const char *fmt, ...);
const char *last_error();

#define DCA_CUDA(call)                                                          \
    do {                                                                        \
        cudaError_t e__ = (call);                                               \
        if (e__ != cudaSuccess) {                                               \
            dca::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,          \
                           cudaGetErrorString(e__));                            \
            return -1;                                                          \
        }                                                                       \
    } while (0)

#define DCA_CUDA_NULL(call)                                                     \
    do {                                                                        \
        cudaError_t e__ = (call);                                               \
        if (e__ != cudaSuccess) {                                               \
            dca::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,          \
                           cudaGetErrorString(e__));                            \
            return nullptr;                                                     \
        }                                                                       \
    } while (0)

// true when a usable CUDA device is present (no CPU fallback exists: callers
// fail loudly otherwise)
int require_device();
int sm_count();

// launch counter: every kernel launch of ours goes through DCA_LAUNCH so that
// bench.py can report gpu_launches honestly
extern std::atomic<unsigned long long> g_launches;   // threads driving different GPUs launch concurrently
#define DCA_LAUNCH(kernel, grid, block, smem, stream, ...)                      \
    do {                                                                        \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);             \
        dca::g_launches.fetch_add(1, std::memory_order_relaxed);                                                    \
    } while (0)

template <typename T>
static inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

#ifdef __CUDACC__
__host__ __device__
#endif
static inline long pair_index(long L, long i, long j)
{   // pydca/plmdca/plmdca_numerics.cpp:340 ; meanfield_dca/msa_numerics.py:220
    return (L * (L - 1)) / 2 - ((L - i) * (L - i - 1)) / 2 + j - i - 1;
}

// ---- device helpers ----------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned warp_sum(unsigned v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum of doubles in a fixed order (deterministic); result valid in
// thread 0.  scratch must hold >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double *scratch)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    double r = 0.0;
    if (wid == 0) {
        r = (lane < nw) ? scratch[lane] : 0.0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
}
// ---- mbarrier + TMA bulk copy (cp.async.bulk, SASS UBLKCP) ---------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy by the TMA engine; bytes % 16 == 0, both sides 16-byte aligned;
// completion is signalled on the mbarrier as transaction bytes
__device__ __forceinline__ void tma_bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// 2-D tile load through a tensor map (cp.async.bulk.tensor, SASS UTMALDG); x = inner
// (contiguous) coordinate, y = row coordinate, both in elements; out-of-range parts of
// the box are zero-filled by the hardware
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tmap, int x, int y, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
#endif

}  // namespace dca
