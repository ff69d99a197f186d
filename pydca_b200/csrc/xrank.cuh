// xrank.cuh -- cross-GPU exchange over peer memory (NVLink 5 / NVSwitch P2P stores).
//
// Multi-GPU plmDCA is owner-sharded: rank r owns the sites
// [bounds[r], bounds[r+1]) and the slice of the parameter vector whose blocks J_ij have
// their first site i in that range.  Every rank exposes ONE device arena to its peers
// (CUDA IPC handle, or a raw pointer when several ranks share a process):
//
//   flags   u32 [XR_MAX_RANKS]            flags[s] = last epoch rank s signalled to me
//   slots   f64 [2][XR_MAX_RANKS][XR_SLOT] small all-reduce payloads, by source rank,
//                                         double-buffered on the reduction sequence number
//   W       f32 [L*q][ld]                 coupling matrix: my columns, all rows; the
//                                         rows of blocks owned by lower ranks are STORED
//                                         HERE BY THOSE RANKS (plm_unpack_push_kernel)
//   R       f32 [max_sites*q][ld]         gradient halves pushed by higher ranks
//                                         (plm_export_kernel)
//
// The data path never reads remote memory: producers push with plain P2P stores,
// then signal with st.release.sys; consumers poll their OWN flags with
// ld.acquire.sys.  Epochs increase monotonically and are identical on all ranks
// because every rank runs the same host control flow on bit-identical scalars.
// Every spin is bounded by XR_TIMEOUT_NS: a dead peer yields an error flag, not a hang.
#pragma once
#include <cstddef>
#include <cstdint>

namespace dca {

constexpr int XR_MAX_RANKS = 16;
constexpr int XR_SLOT = 32;                               // doubles per (parity, source rank)
constexpr unsigned long long XR_TIMEOUT_NS = 4000000000ull;   // 4 s

struct XrPeers {                                          // passed to kernels by value
    int rank, world;
    unsigned char *arena[XR_MAX_RANKS];                   // base of every rank's arena (own included)
    int bounds[XR_MAX_RANKS + 1];                         // site ownership: rank r owns [bounds[r], bounds[r+1])
};

struct XrLayout {                                         // byte offsets inside an arena (same on all ranks)
    size_t off_flags, off_slots, off_W, off_R, bytes;
};

struct XrWait { const unsigned *flags; unsigned *err; int world; unsigned epoch; };   // see xr_wait_all

#ifdef __CUDACC__
__device__ __forceinline__ void xr_st_release_sys(unsigned *p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned xr_ld_acquire_sys(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long xr_now_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;\n" : "=l"(t));
    return t;
}
__device__ __forceinline__ int xr_owner(const XrPeers &P, int site)
{
    int o = 0;
#pragma unroll 1
    for (int r = 1; r < P.world; ++r) o += (site >= P.bounds[r]) ? 1 : 0;
    return o;
}

// thread `t` (< world): tell rank t that I reached `epoch`, then wait until rank t told me.
// All earlier writes of this CTA (after a __syncthreads) and of earlier kernels on the
// stream are ordered before the signal (release at system scope is cumulative).
__device__ __forceinline__ void xr_signal_and_wait(const XrPeers &P, size_t off_flags, unsigned epoch,
                                                   int t, unsigned *err)
{
    __threadfence_system();
    unsigned *remote = reinterpret_cast<unsigned *>(P.arena[t] + off_flags) + P.rank;
    xr_st_release_sys(remote, epoch);
    const unsigned *mine = reinterpret_cast<const unsigned *>(P.arena[P.rank] + off_flags) + t;
    const unsigned long long t0 = xr_now_ns();
    while ((int)(xr_ld_acquire_sys(mine) - epoch) < 0) {
        if (xr_now_ns() - t0 > XR_TIMEOUT_NS) { atomicExch(err, 1u); break; }
        __nanosleep(64);
    }
}

// ---- barriers fused into the producing / consuming kernels -----------------------------
// Producer side: every thread calls this after its (possibly remote) stores; the last CTA of
// the grid to arrive tells every rank that this rank reached `epoch`.  `ticket` must be a
// zero-initialised counter used by this kernel only (atomicInc wraps it back to zero).
__device__ __forceinline__ void xr_grid_signal(unsigned *ticket, const XrPeers &P, size_t off_flags, unsigned epoch)
{
    __shared__ bool xr_last;
    __threadfence_system();                       // my stores are performed before the ticket
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned n = gridDim.x * gridDim.y * gridDim.z;
        xr_last = (atomicInc(ticket, n - 1u) == n - 1u);
    }
    __syncthreads();
    if (xr_last && (int)threadIdx.x < P.world) {
        __threadfence_system();                   // cumulative: every CTA's stores precede the flag
        xr_st_release_sys(reinterpret_cast<unsigned *>(P.arena[threadIdx.x] + off_flags) + P.rank, epoch);
    }
}

// Consumer side: every CTA of the consuming kernel waits (on this rank's OWN flags) until all
// ranks have signalled `epoch`; world == 0 disables the wait.  Ends with a block barrier and a
// proxy fence, so that TMA (async proxy) loads issued afterwards see the peers' stores.
__device__ __forceinline__ void xr_wait_all(const XrWait &w)
{
    if (w.world == 0) return;
    if ((int)threadIdx.x < w.world) {
        const unsigned long long t0 = xr_now_ns();
        while ((int)(xr_ld_acquire_sys(w.flags + threadIdx.x) - w.epoch) < 0) {
            if (xr_now_ns() - t0 > XR_TIMEOUT_NS) { atomicExch(w.err, 1u); break; }
            __nanosleep(32);
        }
    }
    __syncthreads();
    asm volatile("fence.proxy.async.global;\n" ::: "memory");
}

// Sum n (<= XR_SLOT) doubles over all ranks, in rank order (bit-identical result everywhere).
// Called by ONE CTA of >= max(world, n) threads on every rank with the same (epoch, seq).
// Slots are double-buffered on seq: a peer can only start reduction seq+1 after it has seen
// my signal of reduction seq, which I send before reading, so it may overwrite buffer
// (seq+1)&1 while I still read buffer seq&1 -- never the same one.  Ends with a block barrier.
__device__ __forceinline__ void xr_reduce_body(const XrPeers &P, const XrLayout &lay, unsigned epoch, unsigned seq,
                                               const double *__restrict__ local, int n, double *__restrict__ out,
                                               unsigned *err)
{
    const size_t par = (size_t)(seq & 1u) * XR_MAX_RANKS * XR_SLOT;
    for (int idx = threadIdx.x; idx < P.world * n; idx += blockDim.x) {
        const int peer = idx / n, k = idx - peer * n;
        double *dst = reinterpret_cast<double *>(P.arena[peer] + lay.off_slots) + par + (size_t)P.rank * XR_SLOT + k;
        *dst = local[k];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < P.world) xr_signal_and_wait(P, lay.off_flags, epoch, threadIdx.x, err);
    __syncthreads();
    if ((int)threadIdx.x < n) {
        const double *mine = reinterpret_cast<const double *>(P.arena[P.rank] + lay.off_slots) + par + threadIdx.x;
        double s = 0.0;
        for (int r = 0; r < P.world; ++r) s += __ldcv(mine + (size_t)r * XR_SLOT);
        out[threadIdx.x] = s;
    }
    __syncthreads();
}
#endif

}  // namespace dca
