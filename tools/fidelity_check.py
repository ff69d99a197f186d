#!/usr/bin/env python
"""GPU check of the reference-arithmetic mode (csrc/seqsum.cu + the fidelity path of lbfgs.cu):
the sequential-sum kernel against the naive loop, then the config-2 fit in both arithmetic
modes against the frozen reference fit (tests/golden/plm_c2_fit.npz).  Development tool."""
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pydca_b200 import _lib  # noqa: E402
from pydca_b200.synthetic import synthetic_msa  # noqa: E402


def naive_lib():
    so = os.path.join(ROOT, "oracle", "_build", "libseqsum_model.so")
    L = ctypes.CDLL(so)
    f32p = ctypes.POINTER(ctypes.c_float)
    L.seq_naive_dot.argtypes = [f32p, f32p, ctypes.c_long]
    L.seq_naive_dot.restype = ctypes.c_float
    L.seq_naive_reg.argtypes = [f32p, ctypes.c_long, ctypes.c_long, ctypes.c_float, ctypes.c_float]
    L.seq_naive_reg.restype = ctypes.c_float
    return L, f32p


def fn_scores(x, L, q):
    J = x[L * q:].reshape(-1, q, q)[:, :q - 1, :q - 1].astype(np.float64)
    C = J - J.mean(axis=2, keepdims=True) - J.mean(axis=1, keepdims=True) + J.mean(axis=(1, 2), keepdims=True)
    return np.sqrt((C * C).sum(axis=(1, 2)))


def main():
    NL, f32p = naive_lib()
    rng = np.random.default_rng(0)
    n = 4931325
    g = (rng.standard_normal(n) * 0.02).astype(np.float32)
    g[:3150] = (rng.standard_normal(3150) * 40).astype(np.float32)
    d = (-g * rng.uniform(0.5, 1.5, n)).astype(np.float32)
    r = rng.standard_normal(n).astype(np.float32)
    for name, a, b in (("g.g", g, g), ("g.d", g, d), ("g.noise", g, r), ("noise.noise", r, rng.standard_normal(n).astype(np.float32))):
        t0 = time.time()
        got, kms, stats = _lib.seq_dot_stats(a, b)
        t1 = time.time()
        print("   kernel %.3f ms, pipeline %s" % (kms, stats))
        want = np.float32(NL.seq_naive_dot(a.ctypes.data_as(f32p), b.ctypes.data_as(f32p), n))
        print("seq_dot %-12s gpu %.9g naive %.9g %s (%.1f ms incl. H2D)" %
              (name, got, want, "OK" if got.view(np.uint32) == want.view(np.uint32) else "MISMATCH", 1e3 * (t1 - t0)))
    got = _lib.seq_reg(g, 3150, 29.8, 29.8)
    want = np.float32(NL.seq_naive_reg(g.ctypes.data_as(f32p), n, 3150, 29.8, 29.8))
    print("seq_reg gpu %.9g naive %.9g %s" % (got, want, "OK" if got.view(np.uint32) == want.view(np.uint32) else "MISMATCH"))

    N, L, q = 5000, 150, 21
    lam = 0.2 * (L - 1)
    msa = synthetic_msa(N, L, q)
    counts = _lib.sequence_counts(msa, 0.8)
    w = (np.float32(1.0) / counts.astype(np.float32)).astype(np.float32)
    gold = None
    gp = os.path.join(ROOT, "tests", "golden", "plm_c2_fit.npz")
    if os.path.exists(gp):
        gold = np.load(gp)
        print("reference nt=1: evaluations %d fx %.4f ; nt=N: evaluations %d fx %.4f" %
              (gold["nt1_evaluations"], gold["nt1_fx"], gold["ntN_evaluations"], gold["ntN_fx"]))
    for mode in ("reference", "accurate"):
        obj = _lib.PlmObjective(msa, w, q, lam, lam)
        obj.set_arithmetic(mode)
        x0 = obj.init()
        obj.fit(x0, 3)                        # warm
        t0 = time.time()
        x, res = obj.fit(x0, 500, verbose=bool(os.environ.get("FID_VERBOSE")))
        dt = time.time() - t0
        print("mode %-9s status %d iterations %d evaluations %d fx %.4f device %.1f ms wall %.1f ms (%.0f evals/s)" %
              (mode, res.status, res.iterations, res.evaluations, res.fx, res.device_ms, 1e3 * dt,
               res.evaluations / (res.device_ms * 1e-3)))
        if gold is not None:
            fn = fn_scores(x, L, q)
            ref = gold["nt1_fn"].astype(np.float64)
            refN = gold["ntN_fn"].astype(np.float64)
            print("   max rel FN deviation vs reference nt=1: %.3e (reference's own nt=1 vs nt=N band %.3e)" %
                  (np.max(np.abs(fn - ref) / np.abs(ref)), np.max(np.abs(refN - ref) / np.abs(ref))))
            top = np.argsort(-ref, kind="stable")[:L]
            print("   top-L overlap %d / %d" % (len(set(top) & set(np.argsort(-fn, kind="stable")[:L])), L))
        obj.close()


if __name__ == "__main__":
    main()
