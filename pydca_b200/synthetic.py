"""Seeded synthetic MSAs of a named (N, L, q) -- the bench / parity inputs.

Generator of SURVEY.md section 8(d): K = N/50 cluster centres uniform over the q-1
residues; each sequence is a random centre with per-site mutation probability
0.10 (to a uniform residue) and then per-site gap probability 0.05.  Clusters
straddle the 0.8 identity threshold, so neighbour counts are non-trivial.

States are 0..q-1 with gap = q-1, i.e. the encoding of the reference's C++
reader (pydca/plmdca/plmdca_numerics.cpp:708-731).
"""
import numpy as np

PROTEIN_ALPHABET = "ACDEFGHIKLMNPQRSTVWY-"
RNA_ALPHABET = "ACGU-"


def synthetic_msa(num_seqs, seqs_len, num_site_states=21, seed=0xDCA, unique=True):
    """uint8[N, L] alignment, rows unique in first-occurrence order."""
    rng = np.random.default_rng(seed)
    q = int(num_site_states)
    k = max(1, num_seqs // 50)
    centres = rng.integers(0, q - 1, size=(k, seqs_len), dtype=np.uint8)
    rows = []
    seen = set()
    need = num_seqs
    while need > 0:
        m = need
        which = rng.integers(0, k, size=m)
        seqs = centres[which].copy()
        mut = rng.random((m, seqs_len)) < 0.10
        seqs[mut] = rng.integers(0, q - 1, size=int(mut.sum()), dtype=np.uint8)
        gap = rng.random((m, seqs_len)) < 0.05
        seqs[gap] = q - 1
        if not unique:
            rows.append(seqs)
            break
        for r in seqs:
            key = r.tobytes()
            if key not in seen:
                seen.add(key)
                rows.append(r[None, :])
        need = num_seqs - len(seen)
    return np.ascontiguousarray(np.concatenate(rows, axis=0)[:num_seqs])


def write_fasta(msa, path, num_site_states=21):
    """Unwrapped FASTA (one line per record) as both reference readers expect."""
    alpha = PROTEIN_ALPHABET if num_site_states == 21 else RNA_ALPHABET
    lut = np.frombuffer(alpha.encode(), dtype=np.uint8)
    chars = lut[np.asarray(msa)]
    with open(path, "wb") as f:
        for n, row in enumerate(chars):
            f.write(b">seq%06d\n" % n)
            f.write(row.tobytes())
            f.write(b"\n")
    return path
