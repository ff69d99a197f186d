import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pydca_b200 import _lib
from pydca_b200.synthetic import synthetic_msa
N, L, q = 5000, 150, 21
msa = synthetic_msa(N, L, q)
c = _lib.sequence_counts(msa, 0.8)
w = (1.0 / c.astype(np.float32)).astype(np.float32)
lam = 0.2 * (L - 1)
obj = _lib.PlmObjective(msa, w, q, lam, lam)
x0 = obj.init()
for chunk, warm in ((16, 32), (32, 32), (48, 32), (64, 32), (128, 32), (32, 24), (64, 24)):
    obj.set_carry(True, chunk, warm)
    obj.eval_timed(x0)
    t = [obj.eval_timed(x0)["softmax_scan"] for _ in range(5)]
    print("chunk %3d warm %2d: softmax_scan %.4f ms" % (chunk, warm, float(np.median(t))))
