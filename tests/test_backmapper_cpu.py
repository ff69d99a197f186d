"""Host logic of the reference-sequence path, CPU only: the traceback alignment of
pydca_b200/sequence_backmapper.py against the independent score oracle (oracle/sw_oracle.py), its
align_subsequences against outputs of the UNMODIFIED reference's static method
(tests/golden/backmapper.npz), the substitution tables, and MSATrimmer's column arithmetic."""
import numpy as np
import pytest

from oracle import sw_oracle
from pydca_b200 import sequence_backmapper as sb


def _random_pair(rng, alphabet, n, m, related):
    a = "".join(rng.choice(list(alphabet), n))
    if not related:
        return a, "".join(rng.choice(list(alphabet), m))
    b = list(a)
    for _ in range(max(1, n // 6)):                      # point mutations, an insertion, a deletion
        b[int(rng.integers(0, len(b)))] = str(rng.choice(list(alphabet)))
    k = int(rng.integers(1, len(b) - 1))
    b[k:k] = list(rng.choice(list(alphabet), int(rng.integers(1, 4))))
    del b[int(rng.integers(0, len(b) - 2))]
    return a, "".join(b)


@pytest.mark.parametrize("bio,alphabet", [("protein", sw_oracle.PROTEIN), ("rna", sw_oracle.RNA)])
def test_traceback_alignment_scores_match_the_oracle(bio, alphabet):
    rng = np.random.default_rng(3)
    sub, (go, ge) = sb.substitution_matrix(bio)
    tab, pen = sw_oracle.score_table(bio)
    assert pen == (go, ge)
    for a in range(len(alphabet)):                       # the two independently typed tables agree
        for b in range(len(alphabet)):
            assert sub[a, b] == tab[(alphabet[a], alphabet[b])]
    for trial in range(30):
        ref, seq = _random_pair(rng, alphabet, int(rng.integers(5, 60)), int(rng.integers(5, 60)), trial % 2 == 0)
        ca = np.array([alphabet.index(c) for c in ref])
        cb = np.array([alphabet.index(c) for c in seq])
        score, pairs = sb.local_alignment(ca, cb, sub, go, ge)
        assert score == sw_oracle.local_score(ref, seq, bio)
        # the traceback is a valid alignment whose own score is the reported one
        total, k = 0, 0
        while k < len(pairs):
            i, j = pairs[k]
            if i is not None and j is not None:
                total += int(sub[ca[i], cb[j]])
                k += 1
            else:
                side = 0 if i is None else 1
                run = 0
                while k < len(pairs) and pairs[k][side] is None:
                    run += 1
                    k += 1
                total += go + (run - 1) * ge
        assert total == score
        ii = [p[0] for p in pairs if p[0] is not None]
        jj = [p[1] for p in pairs if p[1] is not None]
        assert ii == sorted(ii) and jj == sorted(jj)
    # an identical pair aligns end to end without gaps
    s = "".join(rng.choice(list(alphabet), 40))
    c = np.array([alphabet.index(x) for x in s])
    score, pairs = sb.local_alignment(c, c, sub, go, ge)
    assert pairs == [(k, k) for k in range(40)] and score == int(sub[c, c].sum())


def test_align_subsequences_matches_the_reference_static_method(golden):
    d = golden("backmapper.npz")
    for case, want in zip(d["align_subsequences_cases"].tolist(), d["align_subsequences_outputs"].tolist()):
        ref_mid, tmpl, k = case.split("|")
        try:
            got = sb.SequenceBackmapper.align_subsequences(ref_middle_subseq=ref_mid, template_subseq_in_msa=tmpl,
                                                           num_res_middle_template=int(k))
        except IndexError:
            got = "!IndexError"
        assert got == want, case


def test_pairwise_layout_and_counts():
    """align_pairs_local lays the pair out like pairwise2's local alignments: full sequences, gap
    padded, [begin, end) delimiting the aligned region -- the reference only counts residues in it
    (sequence_backmapper.py:405-431)."""
    bm = sb.SequenceBackmapper.__new__(sb.SequenceBackmapper)
    bm._SequenceBackmapper__biomolecule, bm._SequenceBackmapper__q = "RNA", 5
    ref, other = "GGACGUACGUAACC", "UUUACGUACGUAAGG"
    (ra, oa, score, b, e), = bm.align_pairs_local(ref, other)
    assert len(ra) == len(oa) and ra.replace("-", "") == ref and oa.replace("-", "") == other
    assert ra[b:e] == "ACGUACGUAA" == oa[b:e] and score == 50
    assert len(ra[:b].replace("-", "")) == 2 and len(oa[:b].replace("-", "")) == 3
    assert bm.align_pairs_local(ref, other, score_only=True) == 50


def test_msa_trimmer_columns(tmp_path):
    from pydca_b200.msa_trimmer import MSATrimmer
    seqs = ["AC-GU", "A--GU", ".C-GU", "AC-G-"]
    path = tmp_path / "t.fa"
    path.write_text("".join(">s%d\n%s\n" % (k, s) for k, s in enumerate(seqs)))
    t = MSATrimmer(str(path), biomolecule="rna", max_gap=0.4)
    assert t.compute_msa_columns_gap_size() == (0.25, 0.25, 1.0, 0.0, 0.25)
    assert t.trim_by_gap_size() == (2,)
    assert [s for _, s in t.get_msa_trimmed_by_gap_size()] == ["ACGU", "A-GU", ".CGU", "ACG-"]
