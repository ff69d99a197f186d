#!/usr/bin/env python
"""Symmetric Gauss-Jordan inverse: correctness vs numpy and stage time (developer tool, run under gpurun).

    python tools/inv_check.py [n ...]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pydca_b200 import _lib, msa_numerics as mfn  # noqa: E402
from pydca_b200.synthetic import synthetic_msa  # noqa: E402


def check(n):
    rng = np.random.default_rng(n)
    B = rng.normal(size=(n, n))
    A = B @ B.T / n + np.eye(n)
    J = mfn.compute_couplings(corr_mat=A)
    t = time.perf_counter(); J = mfn.compute_couplings(corr_mat=A); dt = time.perf_counter() - t
    ref = -np.linalg.inv(A)
    print("n=%d  max|J - ref|/max|ref| = %.2e  asym %.2e  wall (H2D+inv+D2H) %.1f ms"
          % (n, np.abs(J - ref).max() / np.abs(ref).max(), np.abs(J - J.T).max(), dt * 1e3), flush=True)


if __name__ == "__main__":
    for n in [int(a) for a in sys.argv[1:]] or [100, 128, 129, 1000, 2000]:
        check(n)
    msa = synthetic_msa(20000, 300, 21)
    c = _lib.sequence_counts(msa, 0.8, compare_f64=True)
    w = 1.0 / c.astype(np.float64)
    for rep in range(3):
        res = mfn.mean_field_scores(msa, 21, w, 0.5)
        print("C3 stage ms {pairfreq+H2D, corr, inverse, fn, di} =", np.round(res["stage_ms"], 3).tolist(), flush=True)
    tf = _lib.fp64_mma_peak() if hasattr(_lib, "fp64_mma_peak") else None
    inv_ms = res["stage_ms"][2]
    print("inverse: %.2f TFLOP/s (n^3)%s" % (6000.0 ** 3 / inv_ms / 1e9, "" if tf is None else " of %.1f measured peak" % tf))
