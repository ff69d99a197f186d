"""Local pairwise alignment score -- TEST INFRASTRUCTURE ONLY.

The reference calls Biopython's ``pairwise2.align.localds(ref, seq, matrix, open, extend,
score_only=True)`` (pydca/sequence_backmapper/sequence_backmapper.py:188-226) for every sequence
of the alignment.  Biopython (setup.py pins biopython==1.74) is absent from /root/reference and
from this image, so this file restates the published algorithm it implements -- Smith-Waterman
with Gotoh's affine gaps, a gap of length k costing open + (k-1)*extend -- with the reference's
parameters (:207-214): BLOSUM62 / -10 / -1 for proteins, scoring_matrix.py's +5 / -4 with -8 / 0
for RNA.  Parity status: UNPINNED against Biopython itself (it cannot run here); pinned only to
hand-checkable cases in tests/test_backmapper_cpu.py.
"""
import numpy as np

PROTEIN = "ACDEFGHIKLMNPQRSTVWY"
RNA = "ACGU"

# BLOSUM62, NCBI order ARNDCQEGHILKMFPSTWYV, upper triangle written out pair by pair so that this
# table is typed independently of the product's
_ORDER = "ARNDCQEGHILKMFPSTWYV"
_ROWS = """
A 4 -1 -2 -2 0 -1 -1 0 -2 -1 -1 -1 -1 -2 -1 1 0 -3 -2 0
R -1 5 0 -2 -3 1 0 -2 0 -3 -2 2 -1 -3 -2 -1 -1 -3 -2 -3
N -2 0 6 1 -3 0 0 0 1 -3 -3 0 -2 -3 -2 1 0 -4 -2 -3
D -2 -2 1 6 -3 0 2 -1 -1 -3 -4 -1 -3 -3 -1 0 -1 -4 -3 -3
C 0 -3 -3 -3 9 -3 -4 -3 -3 -1 -1 -3 -1 -2 -3 -1 -1 -2 -2 -1
Q -1 1 0 0 -3 5 2 -2 0 -3 -2 1 0 -3 -1 0 -1 -2 -1 -2
E -1 0 0 2 -4 2 5 -2 0 -3 -3 1 -2 -3 -1 0 -1 -3 -2 -2
G 0 -2 0 -1 -3 -2 -2 6 -2 -4 -4 -2 -3 -3 -2 0 -2 -2 -3 -3
H -2 0 1 -1 -3 0 0 -2 8 -3 -3 -1 -2 -1 -2 -1 -2 -2 2 -3
I -1 -3 -3 -3 -1 -3 -3 -4 -3 4 2 -3 1 0 -3 -2 -1 -3 -1 3
L -1 -2 -3 -4 -1 -2 -3 -4 -3 2 4 -2 2 0 -3 -2 -1 -2 -1 1
K -1 2 0 -1 -3 1 1 -2 -1 -3 -2 5 -1 -3 -1 0 -1 -3 -2 -2
M -1 -1 -2 -3 -1 0 -2 -3 -2 1 2 -1 5 0 -2 -1 -1 -1 -1 1
F -2 -3 -3 -3 -2 -3 -3 -3 -1 0 0 -3 0 6 -4 -2 -2 1 3 -1
P -1 -2 -2 -1 -3 -1 -1 -2 -2 -3 -3 -1 -2 -4 7 -1 -1 -4 -3 -2
S 1 -1 1 0 -1 0 0 0 -1 -2 -2 0 -1 -2 -1 4 1 -3 -2 -2
T 0 -1 0 -1 -1 -1 -1 -2 -2 -1 -1 -1 -1 -2 -1 1 5 -2 -2 0
W -3 -3 -4 -4 -2 -2 -3 -2 -2 -3 -2 -3 -1 1 -4 -3 -2 11 2 -3
Y -2 -2 -2 -3 -2 -1 -2 -3 2 -1 -1 -2 -1 3 -3 -2 -2 2 7 -1
V 0 -3 -3 -3 -1 -2 -2 -3 -3 3 1 -2 1 -1 -2 -2 0 -3 -1 4
"""


def score_table(biomolecule):
    """{(a, b): score} on letters, plus (open, extend)."""
    if biomolecule.strip().upper() == "PROTEIN":
        tab = {}
        for line in _ROWS.strip().splitlines():
            f = line.split()
            for c, v in zip(_ORDER, f[1:]):
                tab[(f[0], c)] = int(v)
        return tab, (-10, -1)
    return {(a, b): (5 if a == b else -4) for a in RNA for b in RNA}, (-8, 0)


def local_score(ref, seq, biomolecule):
    """Best local alignment score of two letter strings (no gaps inside), row by row in numpy."""
    tab, (go, ge) = score_table(biomolecule)
    n, m = len(ref), len(seq)
    if n == 0 or m == 0:
        return 0
    S = np.array([[tab[(a, b)] for b in seq] for a in ref], dtype=np.int64)
    NEG = -10 ** 9
    Hprev = np.zeros(m + 1, dtype=np.int64)
    Fprev = np.full(m + 1, NEG, dtype=np.int64)
    best = 0
    for i in range(1, n + 1):
        F = np.maximum(Hprev + go, Fprev + ge)            # gap in seq: comes from the row above
        H = np.zeros(m + 1, dtype=np.int64)
        e = NEG
        for j in range(1, m + 1):
            e = max(H[j - 1] + go, e + ge)                # gap in ref: along the row
            H[j] = max(0, Hprev[j - 1] + S[i - 1, j - 1], e, F[j])
        best = max(best, int(H.max()))
        Hprev, Fprev = H, F
    return best
