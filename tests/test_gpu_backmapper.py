"""Reference-sequence path on the GPU (-m gpu): the all-sequences alignment-score
kernel against the oracle, and the reference's own example end to end -- trimming the RF00167
alignment by its reference sequence must give the alignment the demo notebook works on
(examples/pydca_demo.ipynb:43: 71 columns, 2448 unique sequences) and `mfdca compute_fn --apc
--refseq_file` the notebook's top-5 (ipynb:162-166), now from the command line."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# examples/pydca_demo.ipynb:162-166
NOTEBOOK_MF = [((44, 56), 4.290005464965937), ((16, 28), 4.210860173400806), ((17, 27), 4.207758141824402),
               ((45, 55), 4.190480172499375), ((43, 57), 4.039065434604404)]


@pytest.mark.parametrize("bio,q", [("protein", 21), ("rna", 5)])
def test_alignment_score_kernel_matches_the_oracle(bio, q):
    from oracle import sw_oracle
    from pydca_b200 import _lib
    from pydca_b200.synthetic import synthetic_msa
    alphabet = sw_oracle.PROTEIN if q == 21 else sw_oracle.RNA
    N, L = 200, 48
    msa = synthetic_msa(N, L, q, seed=31)
    rng = np.random.default_rng(8)
    # a mutated, truncated member of the alignment as reference
    ref_codes = msa[17][msa[17] != q - 1][3:-2].copy()
    for k in rng.choice(ref_codes.size, 4, replace=False):
        ref_codes[k] = rng.integers(0, q - 1)
    got = _lib.sw_scores(msa, q, ref_codes)
    ref = "".join(alphabet[c] for c in ref_codes)
    for n in list(range(0, N, 9)) + [17]:
        seq = "".join(alphabet[c] for c in msa[n] if c != q - 1)
        assert got[n] == sw_oracle.local_score(ref, seq, bio), n
    assert int(np.argmax(got)) == 17


def _write_examples(tmp_path, golden):
    d = golden("backmapper.npz")
    msa = tmp_path / "MSA_RF00167.fa"
    ref = tmp_path / "ref_RF00167.fa"
    msa.write_bytes(bytes(d["msa_fasta"]))
    ref.write_bytes(bytes(d["ref_fasta"]))
    return str(msa), str(ref)


def test_rf00167_trim_by_refseq_gives_the_notebook_alignment(tmp_path, golden):
    from pydca_b200 import fasta_reader
    from pydca_b200.msa_trimmer import MSATrimmer
    from pydca_b200.sequence_backmapper import SequenceBackmapper
    msa_file, ref_file = _write_examples(tmp_path, golden)
    trimmer = MSATrimmer(msa_file, biomolecule="rna", refseq_file=ref_file)
    trimmed = trimmer.get_msa_trimmed_by_refseq(remove_all_gaps=True)       # ipynb:43
    assert len(trimmed[0][1]) == 71
    out = tmp_path / "trimmed.fa"
    out.write_text("".join(">%s\n%s\n" % (sid, seq) for sid, seq in trimmed))
    msa0 = fasta_reader.get_alignment_zero_based(str(out), "rna")
    want = golden("rf00167_trimmed.npz")["msa"]
    assert msa0.shape == want.shape == (2448, 71) and (msa0 == want).all()
    # the reference sequence is the ungapped record 18: every trimmed column maps to its own position
    mapping = SequenceBackmapper(msa_file=str(out), refseq_file=ref_file, biomolecule="rna").map_to_reference_sequence()
    assert mapping == {k: k for k in range(71)}
    # on the untrimmed alignment the mapping skips the template's gap columns
    full = SequenceBackmapper(msa_file=msa_file, refseq_file=ref_file, biomolecule="rna").map_to_reference_sequence()
    assert sorted(full.values()) == list(range(71)) and len(full) == 71
    cols = sorted(full)
    assert cols == sorted(set(range(102)) - set(trimmer.trim_by_refseq(remove_all_gaps=True)))


def test_mfdca_command_line_with_refseq_reproduces_the_notebook(tmp_path, golden):
    from pydca_b200 import mfdca_main
    from pydca_b200.msa_trimmer import MSATrimmer
    msa_file, ref_file = _write_examples(tmp_path, golden)
    trimmed = MSATrimmer(msa_file, biomolecule="rna", refseq_file=ref_file).get_msa_trimmed_by_refseq(remove_all_gaps=True)
    tfile = tmp_path / "MSA_RF00167_trimmed.fa"
    tfile.write_text("".join(">%s\n%s\n" % (sid, seq) for sid, seq in trimmed))
    out = str(tmp_path / "out")
    mfdca_main.run_meanfield_dca(["compute_fn", "rna", str(tfile), "--pseudocount", "0.5", "--seqid", "0.8", "--apc",
                                  "--refseq_file", ref_file, "--output_dir", out])
    lines = [l.split() for l in open(os.path.join(out, "MFDCA_apc_fn_scores_MSA_RF00167_trimmed.txt"))
             if not l.startswith("#")]
    for (pair, score), f in zip(NOTEBOOK_MF, lines):
        assert (int(f[0]) - 1, int(f[1]) - 1) == pair
        assert abs(float(f[2]) - score) <= 1e-9
