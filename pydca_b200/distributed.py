"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink) for
the collectives, nothing else from torch.

How the hot path shards:

* reweighting  -- rows are independent: rank r computes c_i for its row range against
  the replicated MSA; one all-gather of N uint32 counts.
* plmDCA       -- the loop-carried softmax chain is per site, so SITES are sharded,
  never sequences.  Rank r evaluates the data term of its site range [lo, hi); the
  parameter vector x is replicated and the gradient needs one exchange per
  evaluation: an all-reduce (sum) of g (P floats) plus the 4 objective scalars.
  The L-BFGS vector algebra is replicated (identical on every rank, deterministic).
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


def allgather_counts(local_counts, rows, num_seqs, group=None, device=None):
    """Row-sharded reweighting: every rank passes its counts array (zeros outside its
    rows); returns the complete uint32[N].  Works on CPU tensors (gloo) and CUDA (nccl)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(local_counts, dtype=np.int64).copy())
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, group=group)           # disjoint row ranges: sum == gather
    return t.cpu().numpy().astype(np.uint32)


def sharded_counts(msa, seqid, compare_f64=False, group=None, device=None):
    """c_i over all ranks: row shard on this rank's GPU + one collective."""
    import torch.distributed as dist
    from . import _lib
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_range(msa.shape[0], rank, world)
    local = _lib.sequence_counts(msa, seqid, compare_f64=compare_f64, rows=(lo, hi))
    return allgather_counts(local, (lo, hi), msa.shape[0], group=group, device=device)
