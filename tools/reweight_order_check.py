"""Reweighting in similarity order (early exit) against file order: device ms and identical counts
(developer tool, run under gpurun)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pydca_b200 import _lib  # noqa: E402
from pydca_b200.synthetic import synthetic_msa  # noqa: E402

for (N, L) in ((200000, 250), (50000, 500), (20000, 300)):
    msa = synthetic_msa(N, L, 21)
    ref = None
    for mode in ("sorted", "file"):
        os.environ["DCA_B200_REWEIGHT_ORDER"] = mode
        _lib.sequence_counts_part(msa, 0.8, 0, 1, want_ms=True)
        t = time.perf_counter()
        c, ms = _lib.sequence_counts_part(msa, 0.8, 0, 1, want_ms=True)
        wall = time.perf_counter() - t
        same = "" if ref is None else (" identical to sorted: %s" % bool((c == ref).all()))
        ref = c if ref is None else ref
        print(N, L, mode, "device ms %.2f wall ms %.1f sum %d%s" % (ms, wall * 1e3, int(c.sum()), same), flush=True)
