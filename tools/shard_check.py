"""Multi-process check of the owner-sharded plmDCA path (one process per GPU, CUDA IPC):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/shard_check.py [N_seq L rho]

Every rank evaluates its slice at a perturbed point; rank 0 gathers the gradient and
compares it with an unsharded evaluation on its own GPU.  Then per-kernel times of one
sharded evaluation (max over ranks) and the evaluation rate of a full fit."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from pydca_b200 import _lib
    from pydca_b200 import distributed as dd
    from pydca_b200.synthetic import synthetic_msa

    N = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 150
    rhos = [float(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0.0]
    nsplits = [int(v) for v in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0]
    q = 21
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    _lib.check(_lib.lib().dca_set_device(local), "set_device")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    msa = synthetic_msa(N, L, q)
    counts = dd.sharded_counts(msa, 0.8, device=torch.device("cuda", local))
    w = (1.0 / counts.astype(np.float32)).astype(np.float32)
    lam = 0.2 * (L - 1)
    for rho, ns in [(r, n) for r in rhos for n in nsplits]:
        if ns > 0:
            os.environ["DCA_B200_NSPLIT"] = str(ns)       # read by the scatter planner at create time
        bounds = dd.balanced_site_bounds(L, world, rho)
        sp = dd.ShardedPlm(msa, w, q, lam, lam, bounds=bounds)
        P = sp.obj.P
        rng = np.random.default_rng(1)
        xl0 = sp.init()
        # parity at a perturbed point
        pert = rng.normal(0, 0.05, P).astype(np.float32)
        xl = (xl0 + sp.obj.local_slice(pert)).astype(np.float32)
        fx, gl = sp.obj.eval(xl)
        nmax = max(fc + cc for _, fc, _, cc in sp.layouts)
        buf = torch.zeros(nmax, dtype=torch.float32, device="cuda")
        buf[:gl.size] = torch.from_numpy(gl).cuda()
        out = torch.empty(world * nmax, dtype=torch.float32, device="cuda")
        dist.all_gather_into_tensor(out, buf)
        g = dd.scatter_full_vector(list(out.cpu().numpy().reshape(world, nmax)), sp.layouts, P)
        bufx = torch.zeros(nmax, dtype=torch.float32, device="cuda")
        bufx[:xl.size] = torch.from_numpy(xl).cuda()
        dist.all_gather_into_tensor(out, bufx)
        xfull = dd.scatter_full_vector(list(out.cpu().numpy().reshape(world, nmax)), sp.layouts, P)
        if rank == 0:
            ref = _lib.PlmObjective(msa, w, q, lam, lam)
            f1, g1 = ref.eval(xfull)
            ref.close()
            print("rho %.2f nsplit override %d bounds %s" % (rho, ns, bounds))
            print("parity: fx %.9g vs %.9g (rel %.2e)  g relmax %.2e" %
                  (fx, f1, abs(fx - f1) / abs(f1), np.abs(g - g1).max() / np.abs(g1).max()), flush=True)
        # per-kernel times (collective call), max over ranks
        sp.obj.eval_timed(xl)
        acc = np.zeros(6)
        for _ in range(5):
            t = sp.obj.eval_timed(xl)
            acc += np.array([t[k] for k in sp.obj.KERNELS])
        tt = torch.tensor(acc / 5, device="cuda")
        tmax = tt.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        alls = [torch.zeros_like(tt) for _ in range(world)]
        dist.all_gather(alls, tt)
        if rank == 0:
            print("kernel ms (max over ranks) unpack+push %.3f logits %.3f softmax %.3f scatter %.3f export %.3f assemble+reduce %.3f  sum %.3f"
                  % (*tmax.tolist(), float(tmax.sum())))
            for r, a in enumerate(alls):
                print("   rank %d sites %d: %s" % (r, bounds[r + 1] - bounds[r], " ".join("%.3f" % v for v in a.tolist())))
        d = sp.obj.describe()
        print("   rank %d nsplit %d col_tiles %d nscan %d Ploc %d" % (rank, d["nsplit"], d["col_tiles"], d["nscan"], sp.obj.Ploc), flush=True)
        # fit rate
        for rep in range(3):
            dist.barrier()
            t0 = time.perf_counter()
            _, res = sp.fit(500)
            dist.barrier()
            dt = time.perf_counter() - t0
            if rank == 0:
                print("fit: %d evals %d iters status %d fx %.4f  device %.2f ms  wall %.2f ms -> %.0f evals/s (device)"
                      % (res.evaluations, res.iterations, res.status, res.fx, res.device_ms, dt * 1e3,
                         res.evaluations / (res.device_ms * 1e-3)), flush=True)
        xg = sp.gather()
        if rank == 0:
            print("gathered x: %d floats, |x| %.4f" % (xg.size, float(np.linalg.norm(xg))))
        sp.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
