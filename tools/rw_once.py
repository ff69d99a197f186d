import sys
sys.path.insert(0, "/root/repo")
from pydca_b200 import _lib
from pydca_b200.synthetic import synthetic_msa
msa = synthetic_msa(200000, 250, 21)
c, ms = _lib.sequence_counts_part(msa, 0.8, 0, 1, want_ms=True)
c, ms = _lib.sequence_counts_part(msa, 0.8, 0, 1, want_ms=True)
print("device ms", ms)
