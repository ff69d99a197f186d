"""Wall-clock breakdown of the reference entry point plmdcaBackend(file, ...) at the bench
workload (DCA_B200_TRACE=1 prints the stages from inside the library)."""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["DCA_B200_TRACE"] = "1"
from pydca_b200 import _lib  # noqa: E402
from pydca_b200.synthetic import synthetic_msa, write_fasta  # noqa: E402

N, L, q = 5000, 150, 21
tmp = tempfile.mkdtemp()
path = write_fasta(synthetic_msa(N, L, q), os.path.join(tmp, "c2.fa"), q)
lam = 0.2 * (L - 1)
for rep in range(3):
    t0 = time.perf_counter()
    _lib.plmdca_backend(path, "protein", L, q, 0.8, lam, lam, 500)
    print("call %d: %.2f ms, fit device %.2f ms" % (rep, 1e3 * (time.perf_counter() - t0), _lib.last_fit().device_ms),
          file=sys.stderr)
