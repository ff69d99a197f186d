#!/usr/bin/env python
"""Developer smoke/diagnostic run on a GPU box: parity numbers + rough timings.

    python tools/gpu_check.py [quick|full]

Prints error magnitudes of every CUDA entry point against the CPU oracle
(tests/ assert the same things; this prints them so a gpurun log shows the margins).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from oracle import mf_oracle  # noqa: E402
from pydca_b200 import _lib  # noqa: E402
from pydca_b200.synthetic import synthetic_msa  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def relmax(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def check_plm_case(name):
    d = np.load(os.path.join(GOLD, name))
    msa, q = d["msa"], int(d["q"])
    N, L = msa.shape
    P = oracle.port()
    c = _lib.sequence_counts(msa, float(d["seqid"]))
    c_or = P.reweight_counts(msa, float(d["seqid"]))
    print("[%s] N=%d L=%d q=%d counts exact: %s" % (name, N, L, q, bool((c == c_or).all())))
    w = P.plm_weights(c)
    assert (w == d["w"]).all()
    obj = _lib.PlmObjective(msa, w, q, float(d["lambda_h"]), float(d["lambda_J"]))
    x0 = obj.init()
    print("   init: max abs diff vs ref", float(np.abs(x0 - d["x0"]).max()))
    for tag in ("0", "1"):
        x = d["x" + tag]
        fx, g = obj.eval(x)
        f64, g64 = P.plm_gradient_f64(msa, w, q, float(d["lambda_h"]), float(d["lambda_J"]), x)
        print("   x%s: fx gpu %.6f ref %.6f f64 %.6f | rel fx vs ref %.2e vs f64 %.2e | g relmax vs ref %.2e vs f64 %.2e (ref vs f64 %.2e)"
              % (tag, fx, float(d["fx" + tag]), f64, abs(fx - float(d["fx" + tag])) / abs(f64), abs(fx - f64) / abs(f64),
                 relmax(g, d["g" + tag]), relmax(g, g64), relmax(d["g" + tag], g64)))
    # no-carry
    obj.set_carry(False)
    fx, g = obj.eval(d["x1"])
    fo, go = P.plm_gradient(msa, w, q, float(d["lambda_h"]), float(d["lambda_J"]), d["x1"], carry=False)
    print("   no-carry: rel fx %.2e g relmax %.2e" % (abs(fx - fo) / abs(fo), relmax(g, go)))
    obj.set_carry(True)
    fn, apc = _lib.plm_fn_apc(d["x1"], L, q)
    print("   FN relmax vs ref %.2e ; APC order equal: %s" % (
        relmax(fn, d["x1_fn"]), bool((np.argsort(-apc, kind="stable")[:10] ==
                                      np.argsort(-mf_oracle.apc(d["x1_fn"].astype(np.float64), L), kind="stable")[:10]).all())))
    obj.close()


def check_fit(name, iters):
    d = np.load(os.path.join(GOLD, name))
    msa, q = d["msa"], int(d["q"])
    N, L = msa.shape
    P = oracle.port()
    w = P.plm_weights(P.reweight_counts(msa, float(d["seqid"])))
    obj = _lib.PlmObjective(msa, w, q, float(d["lambda_h"]), float(d["lambda_J"]))
    x0 = obj.init()
    t = time.time()
    x, res = obj.fit(x0, iters)
    dt = time.time() - t
    fn, apc = _lib.plm_fn_apc(x, L, q)
    fn_ref = mf_oracle.plm_fn_scores(d["x_fit_nt1"], L, q)
    print("[fit %s] status %d iters %d evals %d fx %.4f  %.3fs (device %.1f ms) | FN relmax vs ref nt=1 fit: %.2e"
          % (name, res.status, res.iterations, res.evaluations, res.fx, dt, res.device_ms, relmax(fn, fn_ref)))
    fxr, _ = P.plm_gradient(msa, w, q, float(d["lambda_h"]), float(d["lambda_J"]), d["x_fit_nt1"])
    print("   ref fit fx %.4f" % fxr)
    obj.close()


def check_mf(name):
    d = np.load(os.path.join(GOLD, name))
    msa, q = d["msa"], int(d["q"])
    N, L = msa.shape
    c = _lib.sequence_counts(msa, 0.8, compare_f64=True)
    w = 1.0 / c.astype(np.float64)
    print("[%s] weights exact: %s" % (name, bool((w == d["w"]).all())))
    lib = _lib.lib()
    fi = np.zeros((L, q))
    fij = np.zeros((L * (L - 1) // 2, q - 1, q - 1))
    _lib.check(lib.dca_mf_freqs(_lib.ptr(msa, _lib._u8p), _lib.ptr(w, _lib._f64p), N, L, q,
                                _lib.ptr(fi, _lib._f64p), _lib.ptr(fij, _lib._f64p)), "mf_freqs")
    print("   fi relmax %.2e fij relmax %.2e" % (relmax(fi, d["fi"]), relmax(fij, d["fij"])))
    n = L * (q - 1)
    C = np.zeros((n, n))
    rfi = np.zeros((L, q))
    _lib.check(lib.dca_mf_corr_mat(_lib.ptr(msa, _lib._u8p), _lib.ptr(w, _lib._f64p), N, L, q, 0.5,
                                   _lib.ptr(rfi, _lib._f64p), _lib.ptr(C, _lib._f64p)), "mf_corr")
    print("   corr relmax %.2e reg_fi relmax %.2e symmetric: %s" % (relmax(C, d["corr"]), relmax(rfi, d["reg_fi"]),
                                                                   bool((C == C.T).all())))
    J = np.zeros((n, n))
    _lib.check(lib.dca_mf_couplings(_lib.ptr(C, _lib._f64p), n, _lib.ptr(J, _lib._f64p)), "mf_couplings")
    print("   couplings relmax %.2e" % relmax(J, d["couplings"]))
    npairs = L * (L - 1) // 2
    fn = np.zeros(npairs)
    apc = np.zeros(npairs)
    _lib.check(lib.dca_mf_fn_apc(_lib.ptr(J, _lib._f64p), L, q, _lib.ptr(fn, _lib._f64p), _lib.ptr(apc, _lib._f64p)), "mf_fn")
    print("   FN relmax %.2e" % relmax(fn, d["fn"]))


def time_c2(N=5000, L=150, q=21, iters=30):
    msa = synthetic_msa(N, L, q)
    t = time.time()
    c = _lib.sequence_counts(msa, 0.8)
    print("[C2] reweight %.3fs  Meff %.2f" % (time.time() - t, float((1.0 / c).sum())))
    w = (1.0 / c.astype(np.float32)).astype(np.float32)
    lam = 0.2 * (L - 1)
    t = time.time()
    obj = _lib.PlmObjective(msa, w, q, lam, lam)
    print("   create %.3fs" % (time.time() - t))
    x0 = obj.init()
    g = np.empty_like(x0)
    for _ in range(3):
        t = time.time()
        fx, g = obj.eval(x0, out=g)
        print("   host eval %.4fs fx %.3f" % (time.time() - t, fx))
    t = time.time()
    x, res = obj.fit(x0, iters, warmup_iterations=5)
    dt = time.time() - t
    print("   fit: status %d iters %d evals %d fx %.3f wall %.3fs device %.1f ms  timed %.1f ms / %d evals => %.1f evals/s"
          % (res.status, res.iterations, res.evaluations, res.fx, dt, res.device_ms, res.timed_ms,
             res.timed_evaluations, 1e3 * res.timed_evaluations / max(res.timed_ms, 1e-9)))
    obj.close()


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "quick"
    print("devices:", _lib.lib().dca_device_count())
    check_plm_case("plm_toy_rna.npz")
    check_plm_case("plm_toy_protein.npz")
    check_fit("plm_toy_rna.npz", 30)
    check_fit("plm_toy_protein.npz", 20)
    check_mf("mf_toy_rna.npz")
    check_mf("mf_toy_protein.npz")
    if mode == "full":
        check_plm_case("plm_mid_protein.npz")
        time_c2()
    print("launches:", _lib.lib().dca_kernel_launches())
