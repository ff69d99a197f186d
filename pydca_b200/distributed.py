"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink) for
the collectives, nothing else from torch.

How the hot path shards:

* reweighting  -- the all-pairs comparison is symmetric: rank r takes every world-th 128-row
  tile of the upper-triangular tile pairs against the replicated MSA (half the compares of
  the reference's row decomposition); one all-reduce of N integer counts.
* plmDCA       -- the loop-carried softmax chain is per site, so SITES are sharded,
  never sequences.  Default (``ShardedPlm``): rank r OWNS a contiguous site range and
  the slice of the parameter vector whose blocks J_ij start in it; the optimiser's
  vectors are sharded the same way.  Per evaluation only the cross blocks move, by
  P2P stores over NVLink fused into the unpack / export kernels (csrc/xrank.cuh),
  plus a few float64 scalars summed in rank order.  torch.distributed carries the
  64-byte CUDA IPC handles at start-up and the one NCCL all-gather of the final
  parameters; nothing else.
  Legacy (``make_allreduce_callback``): x replicated, one NCCL all-reduce of g (P floats)
  per evaluation -- kept as the fallback when peer memory cannot be mapped.
* mfDCA        -- replicas only at the named sizes: the 6000^2 FP64 inverse is
  tens of milliseconds on one GPU.
"""
import ctypes

import numpy as np


def shard_range(total, rank, world):
    """Contiguous balanced shard [lo, hi) of range(total) for `rank` of `world`."""
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class _CudaView:
    """Zero-copy view of raw device memory for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 2}


def make_allreduce_callback(stream_ptr, group=None, device=None):
    """ctypes callback of type dca_allreduce_fn (include/dca_b200.h): sums the gradient
    buffer and the scalar block over all ranks with NCCL.  Both arguments are DEVICE
    pointers owned by the C library.  The collectives are enqueued in stream order on the
    library's own stream (``stream_ptr`` = PlmObjective.cuda_stream(), wrapped as a
    torch ExternalStream): no host synchronisation on either side of the call."""
    import torch
    import torch.distributed as dist
    from ._lib import ALLREDUCE_FN

    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    ext = torch.cuda.ExternalStream(int(stream_ptr), device=dev)
    views = {}

    def view(ptr, count, typestr):
        key = (ptr, count, typestr)
        t = views.get(key)
        if t is None:
            t = views[key] = torch.as_tensor(_CudaView(ptr, count, typestr), device=dev)
        return t

    def _cb(_ctx, g_ptr, n, s_ptr, ns):
        try:
            with torch.cuda.stream(ext):
                dist.all_reduce(view(s_ptr, ns, "<f8"), group=group)
                dist.all_reduce(view(g_ptr, n, "<f4"), group=group)
            return 0
        except Exception as exc:          # never let an exception cross the C ABI
            print("dca all-reduce callback failed: %r" % (exc,))
            return 1

    return ALLREDUCE_FN(_cb)


def allreduce_counts(partial_counts, group=None, device=None):
    """Sum the ranks' partial neighbour counts (symmetric tile parts, or disjoint row shards)
    into the complete uint32[N] on every rank.  Works on CPU tensors (gloo) and CUDA (nccl)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(partial_counts, dtype=np.int64).copy())
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, group=group)
    return t.cpu().numpy().astype(np.uint32)


def allgather_counts(local_counts, rows, num_seqs, group=None, device=None):
    """Row-sharded variant: counts are zero outside this rank's rows, so the sum is a gather."""
    return allreduce_counts(local_counts, group=group, device=device)


#: below this many residue compares (N * N * L) every rank simply computes all counts itself:
#: the whole reweighting is then under ~10 ms on one B200 and a collective would cost more
REPLICATE_REWEIGHT_BELOW = 1.0e11


def sharded_counts(msa, seqid, compare_f64=False, group=None, device=None, force_shard=False,
                   symmetric=True):
    """c_i on every rank.  Deep alignments: this rank's share of the all-pairs comparison on its
    GPU + one all-reduce of N integers.  symmetric=True (default): rank r takes the row tiles
    r, r + world, ... of the UPPER-triangular tile pairs and every hit counts for both rows
    (half the compares, balanced by the cyclic deal); symmetric=False: the reference's
    decomposition, rows [rN/G, (r+1)N/G) against all rows (plmdca_numerics.cpp:633).
    Shallow alignments: computed redundantly, no collective."""
    import torch.distributed as dist
    from . import _lib
    n, l = msa.shape
    if not force_shard and float(n) * n * l < REPLICATE_REWEIGHT_BELOW:
        return _lib.sequence_counts(msa, seqid, compare_f64=compare_f64)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if symmetric:
        local = _lib.sequence_counts_part(msa, seqid, rank, world, compare_f64=compare_f64)
        return allreduce_counts(local, group=group, device=device)
    lo, hi = shard_range(n, rank, world)
    local = _lib.sequence_counts(msa, seqid, compare_f64=compare_f64, rows=(lo, hi))
    return allgather_counts(local, (lo, hi), n, group=group, device=device)


# ---- owner-sharded plmDCA (peer-memory exchange) ------------------------------------------
def balanced_site_bounds(seqs_len, world, rho=0.0):
    """Contiguous site ranges, bounds[r] .. bounds[r+1], that balance the per-rank cost.
    Site i costs 1 (its column of the gather / scatter kernels) + rho * (L-1-i)/(L-1) (the
    L-1-i coupling blocks it owns: unpack, assemble and every L-BFGS vector pass), so rho > 0
    gives low ranks fewer sites.  Measured at the bench workload the gather / scatter kernels and
    their 128-column tile quantisation dominate: the plain balanced split (rho = 0, the default)
    was fastest at 2 and 8 GPUs (2756 vs 2590 evaluations/s at rho = 0.15 on 8).  Every rank owns
    at least one site."""
    L, world = int(seqs_len), int(world)
    if world > L:
        raise ValueError("more ranks (%d) than sites (%d)" % (world, L))
    cost = np.array([1.0 + rho * (L - 1 - i) / max(1, L - 1) for i in range(L)])
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    bounds = [0]
    for r in range(1, world):
        target = cum[-1] * r / world
        k = int(np.searchsorted(cum, target))
        if k > 0 and abs(cum[k - 1] - target) <= abs(cum[k] - target):
            k -= 1
        k = max(k, bounds[-1] + 1)
        k = min(k, L - (world - r))
        bounds.append(k)
    bounds.append(L)
    return bounds


def scatter_full_vector(slices, layouts, num_params):
    """Inverse of PlmObjective.local_slice over all ranks: slices[r] (fields then couplings
    of rank r) placed at layouts[r] = (field off, count, coupling off, count)."""
    x = np.zeros(int(num_params), dtype=np.float32)
    for sl, (fo, fc, co, cc) in zip(slices, layouts):
        x[fo:fo + fc] = sl[:fc]
        x[co:co + cc] = sl[fc:fc + cc]
    return x


class ShardedPlm:
    """One rank of an owner-sharded plmDCA objective: create, wire the peers, fit, gather.

        sp = ShardedPlm(msa, w, q, lam_h, lam_J)        # inside an initialised process group
        x_local, res = sp.fit(max_iterations)
        x = sp.gather(x_local)                           # NCCL all-gather -> full vector on every rank
        sp.close()
    """

    def __init__(self, msa, weights, num_site_states, lambda_h, lambda_J, group=None, bounds=None):
        import time
        import torch.distributed as dist
        from . import _lib
        t0 = time.perf_counter()
        self.trace = {}                          # wall-clock ms of the set-up stages (diagnostics)

        def mark(name):
            nonlocal t0
            t1 = time.perf_counter()
            self.trace[name] = self.trace.get(name, 0.0) + 1e3 * (t1 - t0)
            t0 = t1
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        L = msa.shape[1]
        self.bounds = list(bounds) if bounds is not None else balanced_site_bounds(L, self.world)
        self.obj = _lib.PlmObjective(msa, weights, num_site_states, lambda_h, lambda_J,
                                     shard=(self.rank, self.world, self.bounds))
        mark("create")
        handles = [None] * self.world
        dist.all_gather_object(handles, self.obj.ipc_handle(), group=group)
        mark("exchange_handles")
        err = None
        try:
            self.obj.connect_ipc(handles)
        except Exception as exc:                 # e.g. no P2P between two devices
            err = repr(exc)
        mark("ipc_open")
        errs = [None] * self.world
        dist.all_gather_object(errs, err, group=group)       # all ranks succeed or all give up
        mark("agree")
        if any(e is not None for e in errs):
            self.obj.disconnect()
            dist.barrier(group=group)
            self.obj.close()
            self.obj = None
            raise RuntimeError("peer-memory wiring failed: %s" % [e for e in errs if e is not None][0])
        self.layouts = [_lib.shard_layout(L, num_site_states, self.bounds[r], self.bounds[r + 1])
                        for r in range(self.world)]
        dist.barrier(group=group)               # every arena is mapped everywhere before the first push
        mark("barrier")

    def init(self):
        return self.obj.init()

    def fit(self, max_iterations, x0_local=None, verbose=False):
        """L-BFGS on the local slices (collective).  Without ``x0_local`` the reference's initial
        point is built on the device and the result stays there for gather(): no host copies."""
        if x0_local is None:
            return None, self.obj.fit_from_init(max_iterations, verbose=verbose)
        return self.obj.fit(x0_local, max_iterations, verbose=verbose)

    def gather(self, x_local=None, device=None):
        """All ranks' slices -> the full parameter vector (reference layout) on every rank:
        ONE NCCL all-gather of the padded device slices (north_star: a single all-gather of
        the final J blocks before FN + APC)."""
        import torch
        import torch.distributed as dist
        nmax = max(fc + cc for _, fc, _, cc in self.layouts)
        backend = dist.get_backend(self.group)
        if backend == "nccl":
            dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
            mine = torch.zeros(nmax, dtype=torch.float32, device=dev)
            src = torch.as_tensor(_CudaView(self.obj.x_device(), self.obj.Ploc, "<f4"), device=dev)
            mine[:self.obj.Ploc].copy_(src)
        else:
            mine = torch.zeros(nmax, dtype=torch.float32)
            mine[:self.obj.Ploc] = torch.from_numpy(np.ascontiguousarray(x_local, dtype=np.float32))
        out = torch.empty(self.world * nmax, dtype=torch.float32, device=mine.device)
        dist.all_gather_into_tensor(out, mine, group=self.group)
        if out.is_cuda:
            host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
            host.copy_(out, non_blocking=True)
            torch.cuda.current_stream(out.device).synchronize()
            out = host
        allx = out.numpy().reshape(self.world, nmax)
        return scatter_full_vector([allx[r] for r in range(self.world)], self.layouts, self.obj.P)

    def close(self):
        import torch.distributed as dist
        if self.obj is not None:
            self.obj.disconnect()
            dist.barrier(group=self.group)       # nobody frees an arena a peer still maps
            self.obj.close()
            self.obj = None


# ---- asymmetric (per-site) estimator: independent site ranges, one all-gather ----------------------------
def assemble_asym_slabs(slabs, bounds, seqs_len, num_site_states):
    """slabs[r]: float32 [(L*q + 1), ld_max] of rank r (row 0 = fields, rows 1.. = its columns of the
    directed couplings, first (bounds[r+1] - bounds[r]) * q columns valid) -> (W_full [L*q, L*q],
    h_full [L*q]) holding every site's columns."""
    lq = int(seqs_len) * int(num_site_states)
    W = np.zeros((lq, lq), dtype=np.float32)
    h = np.zeros(lq, dtype=np.float32)
    for r, slab in enumerate(slabs):
        c0, c1 = bounds[r] * num_site_states, bounds[r + 1] * num_site_states
        h[c0:c1] = slab[0, :c1 - c0]
        W[:, c0:c1] = slab[1:, :c1 - c0]
    return W, h


class AsymmetricPlm:
    """One rank of the asymmetric estimator (include/dca_b200.h: dca_plm_create_asym): the rank fits the
    sites [bounds[rank], bounds[rank+1]) on its GPU with NO communication, then ONE all-gather of the
    column slabs gives every rank all directed couplings, which it symmetrises into the reference's
    parameter vector -- north_star's "independent per-site subproblems ... a single NCCL all-gather
    of the final J blocks".

        ap = AsymmetricPlm(msa, w, q, lam_h, lam_J)      # inside an initialised process group
        res = ap.fit(max_iterations)
        x = ap.gather()                                   # reference layout, on every rank
        ap.close()
    """

    def __init__(self, msa, weights, num_site_states, lambda_h, lambda_J, group=None, bounds=None):
        import torch.distributed as dist
        from . import _lib
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.L, self.q = int(msa.shape[1]), int(num_site_states)
        if bounds is None:          # every site costs the same (L - 1 blocks): an even split is balanced
            bounds = [shard_range(self.L, r, self.world)[0] for r in range(self.world)] + [self.L]
        self.bounds = [int(b) for b in bounds]
        self.obj = _lib.PlmObjective(msa, weights, num_site_states, lambda_h, lambda_J,
                                     sites=(self.bounds[self.rank], self.bounds[self.rank + 1]), asymmetric=True)

    def fit(self, max_iterations, verbose=False):
        return self.obj.fit_from_init(max_iterations, verbose=verbose)

    def _ld(self, r):
        return -(-((self.bounds[r + 1] - self.bounds[r]) * self.q) // 128) * 128

    def gather(self, x_local=None, device=None, pack=None, only_rank=None):
        """x_local: host copy of my vector (gloo / tests); None = the device-resident fit result (nccl).
        pack: symmetrisation callable (tests substitute the oracle's); default the CUDA one.
        only_rank: symmetrise on that rank only (the others return None)."""
        import torch
        import torch.distributed as dist
        from . import _lib
        lq = self.L * self.q
        ld, ld_max = self._ld(self.rank), max(self._ld(r) for r in range(self.world))
        backend = dist.get_backend(self.group)
        if backend == "nccl" and x_local is None:
            dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
            src = torch.as_tensor(_CudaView(self.obj.x_device(), (lq + 1) * ld, "<f4"), device=dev).view(lq + 1, ld)
            mine = torch.zeros((lq + 1, ld_max), dtype=torch.float32, device=dev)
            mine[:, :ld].copy_(src)
        else:
            mine = torch.zeros((lq + 1, ld_max), dtype=torch.float32)
            mine[:, :ld] = torch.from_numpy(np.ascontiguousarray(x_local, dtype=np.float32).reshape(lq + 1, ld))
            if backend == "nccl":
                mine = mine.cuda()
        out = torch.empty(self.world * (lq + 1) * ld_max, dtype=torch.float32, device=mine.device)
        dist.all_gather_into_tensor(out, mine.reshape(-1), group=self.group)
        if only_rank is not None and self.rank != only_rank:
            return None
        W, h = assemble_asym_slabs(out.cpu().numpy().reshape(self.world, lq + 1, ld_max), self.bounds, self.L, self.q)
        return (pack or _lib.asym_pack)(W, h, self.L, self.q)

    def close(self):
        if self.obj is not None:
            self.obj.close()
            self.obj = None
