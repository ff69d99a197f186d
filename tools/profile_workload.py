#!/usr/bin/env python
"""Short workloads for ncu captures (development tool):

    python tools/profile_workload.py seqsum     # reference-arithmetic fit of C2, 3 iterations
    python tools/profile_workload.py reweight   # symmetric reweighting, N=60000 L=250
    python tools/profile_workload.py mf         # mfDCA C3 pipeline once
    python tools/profile_workload.py asym       # asymmetric estimator: a few evaluations of C2 (all sites, then 19 sites)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pydca_b200 import _lib, msa_numerics  # noqa: E402
from pydca_b200.synthetic import synthetic_msa  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "seqsum"
if what == "seqsum":
    N, L, q = 5000, 150, 21
    lam = 0.2 * (L - 1)
    msa = synthetic_msa(N, L, q)
    w = (1.0 / _lib.sequence_counts(msa, 0.8).astype(np.float32)).astype(np.float32)
    obj = _lib.PlmObjective(msa, w, q, lam, lam)
    obj.set_arithmetic("reference")
    x, res = obj.fit(obj.init(), 12)
    print("reference-arithmetic fit:", res.iterations, res.evaluations, res.fx, res.device_ms)
    obj.close()
elif what == "reweight":
    msa = synthetic_msa(60000, 250, 21)
    c, ms = _lib.sequence_counts_part(msa, 0.8, 0, 1, want_ms=True)
    print("reweight N=60000 L=250:", ms, "ms", int(c.sum()))
elif what == "mf":
    msa = synthetic_msa(20000, 300, 21)
    w = 1.0 / _lib.sequence_counts(msa, 0.8, compare_f64=True).astype(np.float64)
    r = msa_numerics.mean_field_scores(msa, 21, w, 0.5)
    print("mf C3 stages:", r["stage_ms"])
elif what == "asym":
    N, L, q = 5000, 150, 21
    lam = 0.2 * (L - 1)
    msa = synthetic_msa(N, L, q)
    w = (1.0 / _lib.sequence_counts(msa, 0.8).astype(np.float32)).astype(np.float32)
    for sites in ((0, L), (0, 19)):
        obj = _lib.PlmObjective(msa, w, q, lam, lam, sites=sites, asymmetric=True)
        x = obj.init()
        for _ in range(3):
            fx, g = obj.eval(x)
        print("asymmetric objective, sites", sites, "fx", fx)
        obj.close()
