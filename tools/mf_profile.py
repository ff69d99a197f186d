"""cProfile of the mfDCA class path at config 3 (FASTA file in, ranked FN-APC list out): where the
host time of the end-to-end call goes.  Development tool (run under gpurun)."""
import cProfile
import os
import pstats
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pydca_b200.meanfield_dca import MeanFieldDCA  # noqa: E402
from pydca_b200.synthetic import synthetic_msa, write_fasta  # noqa: E402

N, L, q = 20000, 300, 21
path = write_fasta(synthetic_msa(N, L, q), os.path.join(tempfile.mkdtemp(), "c3.fa"), q)


def call():
    m = MeanFieldDCA(path, "protein", pseudocount=0.5, seqid=0.8)
    out = m.compute_sorted_FN_APC()
    return m, out


for _ in range(3):
    t0 = time.perf_counter()
    m, out = call()
    print("call %.2f ms, stages %s" % (1e3 * (time.perf_counter() - t0), m.last_stage_ms), file=sys.stderr)
pr = cProfile.Profile()
pr.enable()
call()
pr.disable()
pstats.Stats(pr, stream=sys.stderr).sort_stats("cumulative").print_stats(28)
