// alloc.cu -- caching device allocator for the library's work buffers.
//
// A plmDCA fit needs ~0.6 GB of device buffers at the BASELINE config-2 size
// (W, S, G, perm, 16 L-BFGS vectors).  cudaMalloc / cudaFree cost milliseconds
// each and cudaFree synchronises the device, which dominated the end-to-end time
// of repeated plmdcaBackend() calls.  Freed blocks are kept (per device, exact
// size) and handed back on the next request; the cache is capped per device and is
// released when an allocation fails.  Callers free a block only after the
// stream that used it has been synchronised, so reuse needs no stream tracking.
#include "dca_common.cuh"
#include "dca_internal.h"

#include <map>
#include <mutex>
#include <unordered_map>

namespace dca {

namespace {
constexpr int MAX_DEV = 64;
struct Cache {
    std::mutex mu;
    std::multimap<std::pair<int, size_t>, void *> free_blocks;
    std::unordered_map<void *, std::pair<int, size_t>> live;
    size_t cached[MAX_DEV] = {0};      // per device: a process may drive several GPUs
    size_t cap = 0;                    // per device
};
Cache &cache()
{
    static Cache c;
    if (c.cap == 0) {
        const char *e = getenv("DCA_B200_CACHE_MB");
        // config 4 (N=50000, L=500) needs ~11 GB of work buffers on one B200 (180 GB of HBM)
        c.cap = (e ? (size_t)atol(e) : (size_t)24576) << 20;
    }
    return c;
}
void release_locked(Cache &c)
{
    for (auto &kv : c.free_blocks) cudaFree(kv.second);
    c.free_blocks.clear();
    for (auto &v : c.cached) v = 0;
}
}  // namespace

void *dev_alloc(size_t bytes)
{
    if (bytes == 0) bytes = 1;
    bytes = (bytes + 511) & ~(size_t)511;
    int dev = 0;
    cudaGetDevice(&dev);
    Cache &c = cache();
    std::lock_guard<std::mutex> lk(c.mu);
    auto it = c.free_blocks.find({dev, bytes});
    void *p = nullptr;
    if (it != c.free_blocks.end()) {
        p = it->second;
        c.free_blocks.erase(it);
        c.cached[dev & (MAX_DEV - 1)] -= bytes;
    } else {
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            release_locked(c);
            e = cudaMalloc(&p, bytes);
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
            return nullptr;
        }
    }
    c.live[p] = {dev, bytes};
    return p;
}

void dev_free(void *p)
{
    if (!p) return;
    Cache &c = cache();
    std::lock_guard<std::mutex> lk(c.mu);
    auto it = c.live.find(p);
    if (it == c.live.end()) { cudaFree(p); return; }
    const auto key = it->second;
    c.live.erase(it);
    size_t &cached = c.cached[key.first & (MAX_DEV - 1)];
    if (cached + key.second <= c.cap) {
        c.free_blocks.insert({key, p});
        cached += key.second;
    } else {
        cudaFree(p);
    }
}

// pinned host buffers (result vectors handed to the caller): same exact-size cache
namespace {
struct HostCache {
    std::mutex mu;
    std::multimap<size_t, void *> free_blocks;
    std::unordered_map<void *, size_t> live;
    size_t cached = 0;
};
HostCache &host_cache() { static HostCache c; return c; }
}  // namespace

void *host_pinned_alloc(size_t bytes)
{
    bytes = (bytes + 4095) & ~(size_t)4095;
    HostCache &c = host_cache();
    std::lock_guard<std::mutex> lk(c.mu);
    void *p = nullptr;
    auto it = c.free_blocks.find(bytes);
    if (it != c.free_blocks.end()) {
        p = it->second;
        c.free_blocks.erase(it);
        c.cached -= bytes;
    } else if (cudaMallocHost(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        set_error("pinned host allocation of %zu bytes failed", bytes);
        return nullptr;
    }
    c.live[p] = bytes;
    return p;
}

// returns false when p was not allocated here (caller falls back to free())
bool host_pinned_free(void *p)
{
    HostCache &c = host_cache();
    std::lock_guard<std::mutex> lk(c.mu);
    auto it = c.live.find(p);
    if (it == c.live.end()) return false;
    const size_t bytes = it->second;
    c.live.erase(it);
    if (c.cached + bytes <= ((size_t)2 << 30)) { c.free_blocks.insert({bytes, p}); c.cached += bytes; }
    else cudaFreeHost(p);
    return true;
}

// ---- tensor maps for TMA tile loads ---------------------------------------------
// cuTensorMapEncodeTiled is a driver entry point; it is looked up through the runtime
// so that the library does not link against libcuda.
int make_tensor_map_2d(void *out, int elem_bytes, const void *base, unsigned long long dim0,
                       unsigned long long dim1, unsigned long long stride1_bytes, unsigned box0, unsigned box1)
{
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess || !fn) {
            cudaGetLastError();
            set_error("cuTensorMapEncodeTiled is not available from this driver");
            return -1;
        }
        encode = reinterpret_cast<encode_fn>(fn);
    }
    const cuuint64_t gdim[2] = {dim0, dim1};
    const cuuint64_t gstride[1] = {stride1_bytes};
    const cuuint32_t box[2] = {box0, box1};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = (elem_bytes == 4) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8;
    const CUresult r = encode(static_cast<CUtensorMap *>(out), dt, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return -1; }
    return 0;
}

void dev_cache_release()
{
    Cache &c = cache();
    std::lock_guard<std::mutex> lk(c.mu);
    release_locked(c);
}

}  // namespace dca
