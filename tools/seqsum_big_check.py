"""Sequential-sum pipeline at config 4's vector length (55 M terms, 13 434 chunks) against the naive loop:
bit-exactness and kernel time at a length the test suite does not reach.  Development tool (gpurun)."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pydca_b200 import _lib  # noqa: E402

so = os.path.join(ROOT, "oracle", "_build", "libseqsum_model.so")
NL = ctypes.CDLL(so)
f32p = ctypes.POINTER(ctypes.c_float)
NL.seq_naive_dot.argtypes = [f32p, f32p, ctypes.c_long]
NL.seq_naive_dot.restype = ctypes.c_float
rng = np.random.default_rng(4)
n = 55_025_250
g = (rng.standard_normal(n) * 0.01).astype(np.float32)
g[:10500] = (rng.standard_normal(10500) * 30).astype(np.float32)
d = (-g * rng.uniform(0.5, 1.5, n).astype(np.float32)).astype(np.float32)
s = (rng.standard_normal(n) * 1e-3).astype(np.float32)
ok = True
for name, a, b in (("g.g", g, g), ("g.d", g, d), ("s.d", s, d), ("s.g", s, g)):
    got, kms, stats = _lib.seq_dot_stats(a, b)
    want = np.float32(NL.seq_naive_dot(a.ctypes.data_as(f32p), b.ctypes.data_as(f32p), n))
    same = got.view(np.uint32) == want.view(np.uint32)
    ok &= bool(same)
    print("%-4s gpu %.9g naive %.9g %s kernel %.3f ms %s" % (name, got, want, "OK" if same else "MISMATCH", kms, stats), flush=True)
sys.exit(0 if ok else 1)
