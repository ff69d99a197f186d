"""Per-kernel times of one asymmetric-estimator evaluation at shard widths (developer tool, run under gpurun):
the site range [0, width) of the C2 workload is what one of 150 / width GPUs evaluates."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pydca_b200 import _lib  # noqa: E402
from pydca_b200.synthetic import synthetic_msa  # noqa: E402

N, L = 5000, 150
msa = synthetic_msa(N, L, 21)
w = (1.0 / _lib.sequence_counts(msa, 0.8).astype(np.float32)).astype(np.float32)
lam = 0.2 * (L - 1)
for width in (150, 75, 38, 19):
    obj = _lib.PlmObjective(msa, w, 21, lam, lam, sites=(0, width), asymmetric=True)
    x = obj.init()
    acc = None
    for _ in range(6):
        t = obj.eval_timed(x)
        acc = t if acc is None else {k: acc[k] + t[k] for k in t}
    print("width %3d" % width, {k: round(v / 6, 4) for k, v in acc.items() if v > 0.005}, obj.describe())
    obj.close()
