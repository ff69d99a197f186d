#!/usr/bin/env python
"""External yardstick for the mfDCA inverse: cuSOLVER's potrf + potri through
torch.linalg on the same 6000 x 6000 float64 covariance-like matrix, next to the library's symmetric
Gauss-Jordan.  A measurement only -- nothing in pydca_b200 calls a library solver.  Run under gpurun:

    python tools/inv_yardstick.py [n]

NOT RUN YET: written after this round's GPU budget was spent."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from pydca_b200 import msa_numerics as mfn
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
    rng = np.random.default_rng(0)
    B = rng.normal(size=(n, n))
    A = B @ B.T / n + np.eye(n)
    mfn.compute_couplings(corr_mat=A)                                  # warm
    t = time.perf_counter()
    J = mfn.compute_couplings(corr_mat=A)
    ours_wall = time.perf_counter() - t
    dA = torch.from_numpy(A).cuda()
    for _ in range(2):
        torch.cholesky_inverse(torch.linalg.cholesky(dA))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    inv = torch.cholesky_inverse(torch.linalg.cholesky(dA))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ref = -inv.cpu().numpy()
    print("n=%d  cuSOLVER potrf+potri (torch.linalg, device time) %.2f ms = %.1f TFLOP/s (n^3);  "
          "pydca_b200 compute_couplings wall incl. H2D/D2H %.1f ms;  max|J - ref|/max|ref| = %.2e"
          % (n, ms, n ** 3 / ms / 1e9, ours_wall * 1e3, np.abs(J - ref).max() / np.abs(ref).max()))


if __name__ == "__main__":
    main()
