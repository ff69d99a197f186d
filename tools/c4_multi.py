#!/usr/bin/env python
"""BASELINE config 4 (plmdca N=50k, L=500, q=21) owner-sharded over the GPUs of one box
(DCA_B200_EXCHANGE=allreduce: the older replicated-x + NCCL all-reduce path).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
        --master-port 29533 tools/c4_multi.py [N L iters]

Prints per-kernel times of one sharded evaluation, the time of a short fit and the
evaluations/s; rank 0 also reports FN+APC top pairs (final scoring).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from pydca_b200 import _lib, distributed as dd  # noqa: E402
from pydca_b200.synthetic import synthetic_msa  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 500
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    q = 21
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    _lib.check(_lib.lib().dca_set_device(local), "set_device")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    msa = synthetic_msa(N, L, q)
    dist.barrier()
    t = time.perf_counter()
    counts = dd.sharded_counts(msa, 0.8, device=torch.device("cuda", local))
    dist.barrier()
    t_rw = time.perf_counter() - t
    w = (1.0 / counts.astype(np.float32)).astype(np.float32)
    lam = 0.2 * (L - 1)
    peer = os.environ.get("DCA_B200_EXCHANGE", "peer") == "peer"
    if peer:
        sp = dd.ShardedPlm(msa, w, q, lam, lam)
        obj, sites = sp.obj, (sp.bounds[rank], sp.bounds[rank + 1])
        x0 = obj.init()
        ms = obj.eval_timed(x0)
        ms = obj.eval_timed(x0)
        dist.barrier()
        t = time.perf_counter()
        _, res = sp.fit(iters)
        dist.barrier()
        dt = time.perf_counter() - t
        x = sp.gather()
    else:
        sites = dd.shard_range(L, rank, world)
        obj = _lib.PlmObjective(msa, w, q, lam, lam, sites=sites)
        x0 = obj.init()
        ms = obj.eval_timed(x0)
        ms = obj.eval_timed(x0)
        cb = dd.make_allreduce_callback(obj.cuda_stream())
        dist.barrier()
        t = time.perf_counter()
        x, res = obj.fit(x0, iters, allreduce=cb)
        dist.barrier()
        dt = time.perf_counter() - t
    tt = torch.tensor([res.device_ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("[C4 x%d] N=%d L=%d  Meff %.1f  sharded reweight %.3fs" % (world, N, L, float((1.0 / counts).sum()), t_rw))
        print("   rank0 sites %s kernels ms %s" % (sites, {k: round(v, 3) for k, v in ms.items()}))
        print("   fit: %d iterations %d evaluations status %d fx %.2f  device %.1f ms (max over ranks) wall %.2fs => %.2f evals/s"
              % (res.iterations, res.evaluations, res.status, res.fx, float(tt[0]), dt, res.evaluations / (float(tt[0]) * 1e-3)))
        fn, apc = _lib.plm_fn_apc(x, L, q)
        order = np.argsort(-apc, kind="stable")[:5]
        iu, ju = np.triu_indices(L, k=1)
        print("   FN-APC top5:", [(int(iu[k]), int(ju[k]), round(float(apc[k]), 5)) for k in order], "finite:", bool(np.isfinite(x).all()))
    if peer:
        sp.close()
    else:
        obj.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
