#!/usr/bin/env python
"""Freeze the UNMODIFIED reference's complete plmDCA fit on the bench workload
(BASELINE config 2: synthetic MSA N=5000, L=150, q=21, lambda = 0.2(L-1), 500 iterations)
into tests/golden/plm_c2_fit.npz -- the end-to-end golden for the GPU optimiser.

Runs only in the build container (needs oracle/_ref, i.e. /root/reference):

    python tests/golden/make_c2_fit.py          # ~6 min on 8 vCPU

The reference's own lbfgs() (lbfgs/lib/lbfgs.cpp:248) drives its own PlmDCA::gradient
through oracle/ref_shim.cpp, once at num_threads=1 and once at all cores: the distance
between the two is the reference's own noise band, stored next to the scores.
The 20 MB parameter vector is not stored; FN / FN-APC scores (plmdca.py:437-524 executed on
the reference's vector), the per-iteration trace and the evaluation count are.
"""
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import oracle  # noqa: E402
from pydca_b200.synthetic import synthetic_msa, write_fasta  # noqa: E402

N, L, Q = 5000, 150, 21
LAMBDA = 0.2 * (L - 1)


def fn_scores(x):
    """plmdca.py:246-268 + :437-481 vectorised (float32 block arithmetic as numpy does on
    float32 arrays); pinned against make_golden.ref_plm_scores' literal loops below."""
    J = x[L * Q:].reshape(-1, Q, Q)[:, :Q - 1, :Q - 1]
    avx = J.mean(axis=2, keepdims=True)
    avy = J.mean(axis=1, keepdims=True)
    C = J - avx - avy + J.mean(axis=(1, 2), keepdims=True)
    return np.sqrt((C * C).sum(axis=(1, 2)))


def main():
    tmp = tempfile.mkdtemp()
    msa = synthetic_msa(N, L, Q)
    path = write_fasta(msa, os.path.join(tmp, "c2.fa"), Q)
    R = oracle.ref()
    out = dict(msa_shape=np.array(msa.shape), lambda_h=LAMBDA, lambda_J=LAMBDA, seqid=0.8)
    for nt in (1, os.cpu_count() or 8):
        rp = R.plm(path, "protein", L, Q, 0.8, LAMBDA, LAMBDA, nt)
        assert (rp.seqs() == msa).all()
        t0 = time.time()
        x, status, evals, fx, trace = rp.fit_trace(500)
        dt = time.time() - t0
        rp.close()
        tag = "nt1" if nt == 1 else "ntN"
        fn = fn_scores(x)
        print("%s: threads %d status %d evaluations %d iterations %d fx %.4f (%.0f s)"
              % (tag, nt, status, evals, len(trace), fx, dt), flush=True)
        out.update({tag + "_status": status, tag + "_evaluations": evals, tag + "_fx": np.float32(fx),
                    tag + "_trace": trace, tag + "_fn": fn.astype(np.float32), tag + "_threads": nt,
                    tag + "_xnorm2": float(np.dot(x.astype(np.float64), x.astype(np.float64)))})
        if nt == 1:
            import make_golden
            sc = make_golden.ref_plm_scores(x, L, Q)            # the reference's literal loops
            assert np.allclose(sc["fn"], fn, rtol=2e-6, atol=0)
            out["nt1_fn_apc_sorted_pairs"] = sc["fn_apc_sorted_pairs"][:600]
            out["nt1_fn_apc_sorted_scores"] = sc["fn_apc_sorted_scores"][:600]
            out["nt1_fn_sorted_pairs"] = sc["fn_sorted_pairs"][:600]
            out["nt1_fields"] = x[:L * Q].copy()
    a, b = out["nt1_fn"].astype(np.float64), out["ntN_fn"].astype(np.float64)
    print("reference self band (nt=1 vs nt=N): max rel FN deviation %.3e" % np.max(np.abs(a - b) / np.abs(a)))
    np.savez_compressed(os.path.join(HERE, "plm_c2_fit.npz"), **out)


if __name__ == "__main__":
    main()
