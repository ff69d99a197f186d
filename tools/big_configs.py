#!/usr/bin/env python
"""BASELINE configs[3] (C4: plmdca N=50000, L=500) and configs[4] (C5: deep MSA N=200000,
L=250: all-pairs reweighting + plmdca) on the G GPUs of one box, owner-sharded.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \\
        --master-port 29533 tools/big_configs.py c4|c5 [iters] [--e2e]

Rank 0 prints ONE JSON line per run: sharded symmetric reweighting (ms, max over ranks), the
CUDA-event time of a short fit (max over ranks) -> evaluations/s, per-kernel ms of one
evaluation on rank 0 and on the last rank, and a numerical check of the gathered parameters
against a single-GPU fit of the same input made in the same run on rank 0's GPU.
--e2e additionally times the reference's entry point plmdcaBackend(file) with DCA_B200_GPUS=G.
--asym additionally runs the asymmetric (per-site) estimator on the same G GPUs: independent fits of equal
site ranges (no exchange), one NCCL all-gather of the column slabs, symmetrisation on rank 0.
"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from pydca_b200 import _lib, distributed as dd  # noqa: E402
from pydca_b200.synthetic import synthetic_msa, write_fasta  # noqa: E402

CONFIGS = {"c4": (50000, 500), "c5": (200000, 250), "c2": (5000, 150)}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c4"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else 10
    want_e2e = "--e2e" in sys.argv
    want_asym = "--asym" in sys.argv
    N, L = CONFIGS[name]
    q = 21
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    saved = os.dup(1)
    os.dup2(2, 1)                                   # NCCL banners etc. go to stderr
    torch.cuda.set_device(local)
    _lib.check(_lib.lib().dca_set_device(local), "set_device")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cpu = dist.new_group(backend="gloo")
    dev = torch.device("cuda", local)
    msa = synthetic_msa(N, L, q)
    lam = 0.2 * (L - 1)

    def tmax(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # ---- reweighting: symmetric tile parts + one all-reduce of N integers ---------------------------
    dd.sharded_counts(msa[:4096], 0.8, device=dev, force_shard=True)          # warm (NCCL channel, kernels)
    dist.barrier()
    t = time.perf_counter()
    part, ms_part = _lib.sequence_counts_part(msa, 0.8, rank, world, want_ms=True)
    counts = dd.allreduce_counts(part, device=dev)
    torch.cuda.synchronize()
    t_rw = tmax(time.perf_counter() - t)
    ms_part = tmax(ms_part)
    w = (1.0 / counts.astype(np.float32)).astype(np.float32)

    # ---- owner-sharded fit ----------------------------------------------------------------------
    sp = dd.ShardedPlm(msa, w, q, lam, lam)
    obj = sp.obj
    x0 = obj.init()
    obj.eval_timed(x0)
    ms = obj.eval_timed(x0)
    sp.fit(2)                                                                  # warm
    dist.barrier()
    _, res = sp.fit(iters)
    fit_ms = tmax(res.device_ms)
    x = sp.gather()
    all_ms = [None] * world
    dist.all_gather_object(all_ms, {k: round(v, 3) for k, v in ms.items()}, group=cpu)

    line = None
    if rank == 0:
        adds = float(N) * L * L * q / world
        line = {"config": name, "N": N, "L": L, "q": q, "n_gpus": world, "iterations": res.iterations,
                "evaluations": res.evaluations, "status": res.status, "fx": res.fx,
                "fit_device_ms_max_over_ranks": fit_ms, "evals_per_s": res.evaluations / (fit_ms * 1e-3),
                "reweight": {"wall_ms_incl_h2d_and_allreduce": 1e3 * t_rw, "device_ms_max_over_ranks": ms_part,
                             "byte_compares_per_s": float(N) * N * L / (ms_part * 1e-3),
                             "packed_lane_floor_ms_per_gpu": float(N) * N * L / 4 / (148 * 128 * 1.965e9) * 1e3 / world},
                "kernel_ms_rank0": all_ms[0], "kernel_ms_last_rank": all_ms[-1], "sites_per_rank": sp.bounds,
                "gather_adds_per_rank_per_kernel": adds,
                "layout_rank0": obj.layout_stats(),
                "logits_gadds_per_s_rank0": adds / (ms["logits"] * 1e-3) / 1e9,
                "scatter_gadds_per_s_rank0": adds / (ms["scatter"] * 1e-3) / 1e9,
                "smem_roof_gwords_per_s": 148 * 32 * 1.965,
                "note": "the gather kernels move adds x (1 - 3/4 uniform4) shared-memory words where the fast path applies "
                        "(logits: 128-column CTAs; scatter: group sums)"}
        # numerical check: the same short fit on ONE GPU, same run
        try:
            torch.cuda.synchronize()
            solo = _lib.PlmObjective(msa, w, q, lam, lam)
            solo.fit(solo.init(), 1)
            xs, rs = solo.fit(solo.init(), iters)
            solo.close()
            fn_m, _ = _lib.plm_fn_apc(x, L, q)
            fn_s, _ = _lib.plm_fn_apc(xs, L, q)
            line["single_gpu"] = {"evaluations": rs.evaluations, "fit_device_ms": rs.device_ms,
                                  "evals_per_s": rs.evaluations / (rs.device_ms * 1e-3)}
            line["speedup_vs_single_gpu"] = rs.device_ms / fit_ms
            line["scaling_efficiency"] = rs.device_ms / fit_ms / world
            line["multi_gpu_check"] = {"fn_max_rel_dev_vs_single_gpu": float(np.max(np.abs(fn_m - fn_s) / np.abs(fn_s))),
                                       "x_max_abs_dev": float(np.max(np.abs(x - xs))), "x_max_abs": float(np.max(np.abs(xs))),
                                       "same_evaluations": bool(rs.evaluations == res.evaluations)}
        except Exception as exc:
            line["single_gpu"] = {"error": repr(exc)}
    torch.cuda.synchronize()
    dist.barrier(group=cpu)
    sp.close()

    if want_asym:
        torch.cuda.synchronize()
        dist.barrier(group=cpu)
        ap = dd.AsymmetricPlm(msa, w, q, lam, lam)
        ap.fit(2)                                                              # warm
        ams = ap.obj.eval_timed(ap.obj.init())
        dist.barrier()
        ra = ap.fit(iters)
        a_ms = tmax(ra.device_ms)
        a_ev = int(tmax(float(ra.evaluations)))
        t = time.perf_counter()
        xa = ap.gather(only_rank=0)
        torch.cuda.synchronize()
        t_gather = time.perf_counter() - t
        ap.close()
        if rank == 0:
            fn_a, _ = _lib.plm_fn_apc(xa, L, q)
            fn_j, _ = _lib.plm_fn_apc(x, L, q)
            top = L
            ra_, rj_ = np.argsort(-fn_a, kind="stable")[:top], np.argsort(-fn_j, kind="stable")[:top]
            line["asymmetric"] = {"iterations": ra.iterations, "evaluations_max_over_ranks": a_ev,
                                  "fit_device_ms_max_over_ranks": a_ms, "evals_per_s": a_ev / (a_ms * 1e-3),
                                  "gather_and_symmetrise_seconds_rank0": t_gather,
                                  "kernel_ms_rank0": {k: round(v, 3) for k, v in ams.items()},
                                  "fn_topL_overlap_with_joint_fit_after_%d_iterations" % iters:
                                      int(len(set(ra_.tolist()) & set(rj_.tolist()))), "topL": top}
        dist.barrier(group=cpu)

    if want_e2e:
        torch.cuda.synchronize()
        dist.barrier(group=cpu)
        if rank == 0:
            tmp = tempfile.mkdtemp(prefix="dca_big_")
            path = write_fasta(msa, os.path.join(tmp, "%s.fa" % name), q)
            _lib.set_backend_gpus(world)
            os.environ["DCA_B200_TRACE"] = "1"
            walls = []
            for _ in range(2):
                t = time.perf_counter()
                xe = _lib.plmdca_backend(path, "protein", L, q, 0.8, lam, lam, iters)
                walls.append(time.perf_counter() - t)
            fit = _lib.last_fit()
            line["e2e_plmdcaBackend"] = {"seconds": walls[-1], "first_call_seconds": walls[0], "evaluations": fit.evaluations,
                                         "evals_per_s": fit.evaluations / walls[-1], "gpus": world,
                                         "x_max_abs_dev_vs_torchrun_path": float(np.max(np.abs(xe - x)))}
        dist.barrier(group=cpu)
    if rank == 0:
        os.dup2(saved, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
