import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pydca_b200 import _lib
from pydca_b200.synthetic import synthetic_msa
N, L, q = 5000, 150, 21
msa = synthetic_msa(N, L, q)
c = _lib.sequence_counts(msa, 0.8)
w = (1.0 / c.astype(np.float32)).astype(np.float32)
lam = 0.2 * (L - 1)
t = time.perf_counter(); obj = _lib.PlmObjective(msa, w, q, lam, lam); print("create %.1f ms" % (1e3 * (time.perf_counter() - t)))
t = time.perf_counter(); x0 = obj.init(); print("init %.1f ms" % (1e3 * (time.perf_counter() - t)))
for it in (1, 1, 2, 5, 10, 29, 500, 500):
    t = time.perf_counter(); x, res = obj.fit(x0, it); dt = 1e3 * (time.perf_counter() - t)
    print("fit(max_it=%d): wall %.1f ms device %.1f ms evals %d status %d" % (it, dt, res.device_ms, res.evaluations, res.status))
g = np.empty_like(x0)
for _ in range(3):
    t = time.perf_counter(); obj.eval(x0, out=g); print("host eval %.2f ms" % (1e3 * (time.perf_counter() - t)))
