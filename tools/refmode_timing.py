#!/usr/bin/env python
"""Reference-arithmetic fit of config 2 with per-launch timing of the sequential sums
(DCA_B200_SEQSUM_TIMING) and the pipeline counters (DCA_B200_TRACE).  Development tool."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pydca_b200 import _lib  # noqa: E402
from pydca_b200.synthetic import synthetic_msa  # noqa: E402

N, L, q = 5000, 150, 21
lam = 0.2 * (L - 1)
msa = synthetic_msa(N, L, q)
counts = _lib.sequence_counts(msa, 0.8)
w = (np.float32(1.0) / counts.astype(np.float32)).astype(np.float32)
obj = _lib.PlmObjective(msa, w, q, lam, lam)
obj.set_arithmetic("reference")
x0 = obj.init()
obj.fit(x0, 3)
for rep in range(2):
    t0 = time.time()
    x, res = obj.fit(x0, 500)
    print("reference mode: status %d iterations %d evaluations %d fx %.4f device %.1f ms wall %.1f ms" %
          (res.status, res.iterations, res.evaluations, res.fx, res.device_ms, 1e3 * (time.time() - t0)), flush=True)
obj.close()
