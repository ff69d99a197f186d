#!/usr/bin/env python
"""Timings at the larger BASELINE shapes (developer tool, run under gpurun).

    python tools/scale_check.py [mf|rw|plm|all]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pydca_b200 import _lib, msa_numerics as mfn  # noqa: E402
from pydca_b200.synthetic import synthetic_msa  # noqa: E402


def tic():
    return time.perf_counter()


def mf_c3(N=20000, L=300, q=21):
    msa = synthetic_msa(N, L, q)
    t = tic(); c = _lib.sequence_counts(msa, 0.8, compare_f64=True); t_w = tic() - t
    w = 1.0 / c.astype(np.float64)
    for rep in range(2):
        t = tic(); rfi, corr = mfn.corr_mat_from_alignment(msa, q, w, 0.5); t_c = tic() - t
        t = tic(); J = mfn.compute_couplings(corr_mat=corr); t_i = tic() - t
        t = tic(); fn, apc = mfn.frobenius_norm_scores(J, L, q); t_f = tic() - t
        print("[mf C3 N=%d L=%d] rep %d weights %.3fs corr(+D2H %d MB) %.3fs inverse(+H2D/D2H) %.3fs fn %.3fs  total %.3fs"
              % (N, L, rep, t_w, corr.nbytes >> 20, t_c, t_i, t_f, t_w + t_c + t_i + t_f))
    for rep in range(3):
        t = tic()
        res = mfn.mean_field_scores(msa, q, w, 0.5, want_couplings=False, want_di=(rep == 2))
        dt = tic() - t
        print("[mf C3 fused] rep %d wall %.3fs  stage ms {pairfreq+H2D, corr, inverse, fn, di} = %s  => %.0f site-pairs/s"
              % (rep, dt, np.round(res["stage_ms"], 2).tolist(), L * (L - 1) / 2 / dt))
    print("   fused FN == staged FN: max rel diff %.2e" % (np.abs(res["fn"] - fn).max() / fn.max()))
    n = corr.shape[0]
    rng = np.random.default_rng(0)
    v = rng.normal(size=(n, 4))
    res = corr @ (-(J @ v)) - v
    print("   |C * inv(C) v - v|_max = %.3e ; symmetric J: %.3e ; Meff %.2f" % (np.abs(res).max(), np.abs(J - J.T).max(), w.sum()))


def rw(N=50000, L=500, q=21):
    msa = synthetic_msa(N, L, q)
    for rep in range(2):
        t = tic(); c = _lib.sequence_counts(msa, 0.8); dt = tic() - t
        print("[reweight N=%d L=%d] %.3fs => %.3e residue compares/s ; Meff %.1f" % (N, L, dt, N * N * L / dt, (1.0 / c).sum()))


def plm_big(N=50000, L=500, q=21):
    msa = synthetic_msa(N, L, q)
    c = _lib.sequence_counts(msa, 0.8)
    w = (1.0 / c.astype(np.float32)).astype(np.float32)
    lam = 0.2 * (L - 1)
    t = tic(); obj = _lib.PlmObjective(msa, w, q, lam, lam); print("[plm N=%d L=%d] create %.2fs %s" % (N, L, tic() - t, obj.describe()))
    x0 = obj.init()
    for _ in range(2):
        ms = obj.eval_timed(x0)
        print("   kernels ms:", {k: round(v, 3) for k, v in ms.items()}, "sum %.2f" % sum(ms.values()))
    d = obj.describe()
    adds = d["N"] * d["L"] * d["L"] * d["q"]
    print("   logits %.0f Gadd/s scatter %.0f Gadd/s (smem roof 9306)" % (adds / ms["logits"] / 1e6, adds / ms["scatter"] / 1e6))
    obj.close()


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("mf", "all"):
        mf_c3()
    if what in ("rw", "all"):
        rw()
    if what in ("plm", "all"):
        plm_big()
