import os, sys
sys.path.insert(0, "/root/repo")
from pydca_b200 import _lib
from pydca_b200.synthetic import synthetic_msa
N, L = 200000, 250
msa = synthetic_msa(N, L, 21)
for nparts in (1, 2, 4, 8):
    for mode in ("sorted", "file", None):
        if mode: os.environ["DCA_B200_REWEIGHT_ORDER"] = mode
        else: os.environ.pop("DCA_B200_REWEIGHT_ORDER", None)
        _lib.sequence_counts_part(msa, 0.8, 0, nparts, want_ms=True)
        c, ms = _lib.sequence_counts_part(msa, 0.8, nparts - 1, nparts, want_ms=True)
        print("nparts", nparts, "mode", mode, "device ms of one part %.1f" % ms, "sum", int(c.sum()), flush=True)
