"""A handful of plmDCA evaluations at the bench workload (C2 by default): the short command
to wrap in ncu (tools: `ncu --set full -k regex:plm_scatter -s 2 -c 1 python tools/eval_once.py`)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pydca_b200 import _lib  # noqa: E402
from pydca_b200.synthetic import synthetic_msa  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 150
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
msa = synthetic_msa(N, L, 21)
w = (1.0 / _lib.sequence_counts(msa, 0.8).astype(np.float32)).astype(np.float32)
lam = 0.2 * (L - 1)
obj = _lib.PlmObjective(msa, w, 21, lam, lam)
x = obj.init()
acc = None
for _ in range(reps):
    t = obj.eval_timed(x)
    acc = t if acc is None else {k: acc[k] + t[k] for k in t}
print({k: round(v / reps, 4) for k, v in acc.items()}, obj.describe())
obj.close()
