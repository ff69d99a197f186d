"""Symmetric reweighting at config 5, one part of nparts, in similarity order and in file order: device ms
(developer tool, run under gpurun; decides from how many parts on the file order wins)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pydca_b200 import _lib  # noqa: E402
from pydca_b200.synthetic import synthetic_msa  # noqa: E402

N, L = 200000, 250
msa = synthetic_msa(N, L, 21)
for nparts in (1, 2, 4, 8):
    for mode in ("sorted", "file", None):
        if mode:
            os.environ["DCA_B200_REWEIGHT_ORDER"] = mode
        else:
            os.environ.pop("DCA_B200_REWEIGHT_ORDER", None)
        _lib.sequence_counts_part(msa, 0.8, 0, nparts, want_ms=True)
        c, ms = _lib.sequence_counts_part(msa, 0.8, nparts - 1, nparts, want_ms=True)
        print("nparts", nparts, "mode", mode, "device ms of one part %.1f" % ms, "sum", int(c.sum()), flush=True)
