"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every
symbol include/dca_b200.h declares, the product never touches the oracle, compute
calls fail loudly without a GPU, and the host-side FASTA logic matches the oracle."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "dca_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"typedef\s+int\s*\(\*[^;]*;", "", text)          # function-pointer typedefs
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}()]*\)\s*;", text)
    return sorted(set(n for n in names if n not in ("defined", "int", "long", "void", "float", "double")))


def test_library_exports_every_declared_symbol():
    from pydca_b200 import _lib
    lib = ctypes.CDLL(_lib.library_path())
    names = _declared_symbols()
    assert "plmdcaBackend" in names and "freeFieldsAndCouplings" in names and len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "include/dca_b200.h declares %s but the library does not export it" % n


def test_library_name_matches_reference_glob():
    """pydca/plmdca/plmdca.py:33-38 finds its backend with glob('_plmdcaBackend*')."""
    from pydca_b200 import _lib
    assert os.path.basename(_lib.library_path()).startswith("_plmdcaBackend")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pydca_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(base, f), errors="replace").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
                assert "dca_oracle" not in src and "libpydca_ref" not in src, f
    code = "import sys; import pydca_b200, pydca_b200.msa_numerics; assert 'oracle' not in sys.modules"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)


def test_compute_fails_loudly_without_gpu():
    from pydca_b200 import _lib
    if _lib.lib().dca_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.DcaBackendError, match="no CPU fallback"):
        _lib.sequence_counts(np.zeros((4, 8), dtype=np.uint8), 0.8)
    with pytest.raises(_lib.DcaBackendError):
        _lib.PlmObjective(np.zeros((4, 8), dtype=np.uint8), np.ones(4, dtype=np.float32), 21, 1.0, 1.0)
    from pydca_b200 import msa_numerics
    with pytest.raises(_lib.DcaBackendError):
        msa_numerics.compute_couplings(corr_mat=np.eye(8))


FASTA = """>a desc
ACDEFGHIKL
>b
acdefghikl
>c
AC-EF.HI~L
>d wrapped record
ACDEF
GHIKX

>e
BZJOUXACDE
>f duplicate of c after gap folding
AC.EF-HI-L
"""


def test_fasta_reader_semantics(tmp_path, port):
    """Upper-casing, unknown letters -> gap, first-occurrence dedup in file order,
    blank lines and wrapped records (plmdca_numerics.cpp:699-767, fasta_reader.py:122-163)."""
    from pydca_b200 import _lib, fasta_reader
    p = tmp_path / "t.fa"
    p.write_text(FASTA)
    msa = _lib.read_fasta(str(p), "protein", 10)
    want = port.read_fasta(str(p), "protein", 10)
    assert (msa == want).all()
    # a == b (case), f == c (gap aliases): 4 unique rows a, c, d, e
    assert msa.shape == (4, 10)
    assert msa[0].tolist() == list(range(10))
    assert msa[1].tolist() == [0, 1, 20, 3, 4, 20, 6, 7, 20, 9]
    assert msa[2].tolist() == [0, 1, 2, 3, 4, 5, 6, 7, 8, 20]          # X -> gap
    assert msa[3].tolist() == [20, 20, 20, 20, 20, 20, 0, 1, 2, 3]
    assert fasta_reader.first_record_length(str(p)) == (6, 10)
    one_based = fasta_reader.get_alignment_int_form(str(p), "protein")
    assert one_based[1] == [1, 2, 21, 4, 5, 21, 7, 8, 21, 10]
    seqs = fasta_reader.get_alignment_from_fasta_file(str(p))
    assert (fasta_reader.alignment_zero_based(seqs, "protein") == msa).all()
    # RNA: T is not in the C++ map (it would abort there); the Python reader's gap wins
    r = tmp_path / "r.fa"
    r.write_text(">x\nACGUT-N\n>y\nacgut.n\n")
    rna = _lib.read_fasta(str(r), "rna", 7)
    assert rna.tolist() == [[0, 1, 2, 3, 4, 4, 4]]


def test_fasta_reader_line_endings_whitespace_and_many_duplicates(tmp_path, port):
    """The single-pass reader (csrc/fasta.cpp): CRLF files, blanks/tabs inside a line, no
    trailing newline, residues beyond seqs_len ignored, a headerless file is one record, and
    hash-then-verify dedup keeps first occurrences in file order on a duplicate-heavy file."""
    from pydca_b200 import _lib
    p = tmp_path / "crlf.fa"
    p.write_bytes(b">a\r\nAC DE\tF\r\n>b\r\nACDEFGG\r\n>c\r\nWWWWW")        # b: 2 extra residues; c: no newline
    msa = _lib.read_fasta(str(p), "protein", 5)
    assert msa.tolist() == [[0, 1, 2, 3, 4], [18, 18, 18, 18, 18]]           # b duplicates a after truncation
    h = tmp_path / "headerless.fa"
    h.write_text("ACD\nEF\n")
    assert _lib.read_fasta(str(h), "protein", 5).tolist() == [[0, 1, 2, 3, 4]]
    rng = np.random.default_rng(3)
    base = rng.integers(0, 21, size=(40, 17), dtype=np.uint8)
    order = rng.integers(0, 40, size=3000)
    letters = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY-", dtype=np.uint8)
    d = tmp_path / "dups.fa"
    with open(d, "wb") as f:
        for k, r in enumerate(order):
            f.write(b">s%d\n" % k + letters[base[r]].tobytes() + b"\n")
    got = _lib.read_fasta(str(d), "protein", 17)
    first = list(dict.fromkeys(order.tolist()))
    assert got.shape == (len(first), 17) and (got == base[first]).all()
    assert (got == port.read_fasta(str(d), "protein", 17)).all()


def test_fasta_reader_pieces_on_several_threads(tmp_path, port, monkeypatch):
    """The reader cuts large files in front of header lines and parses the pieces on host threads
    (csrc/fasta.cpp); first-occurrence order across pieces, wrapped / CRLF / lower-case records,
    blank lines, and the line number of the first malformed record must not depend on the cut."""
    from pydca_b200 import _lib
    rng = np.random.default_rng(8)
    L, N = 61, 6000
    letters = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY-", dtype=np.uint8)
    msa = rng.integers(0, 21, size=(N, L), dtype=np.uint8)
    msa[1500:2500] = msa[0:1000]                    # duplicates that straddle pieces
    msa[5500:] = msa[700:1200]
    f = tmp_path / "pieces.fa"
    with open(f, "wb") as fh:
        for i, row in enumerate(msa):
            s = letters[row].tobytes()
            fh.write(b">s%d\n" % i)
            if i % 3 == 0:
                fh.write(s[:40] + b"\n" + s[40:] + b"\n")
            elif i % 7 == 0:
                fh.write(s.lower() + b"\r\n")
            else:
                fh.write(s + b"\n")
            if i % 500 == 0:
                fh.write(b"\n")
    want = port.read_fasta(str(f), "protein", L)
    assert want.shape == (N - 1500, L)
    for threads in ("1", "2", "5", "16"):
        monkeypatch.setenv("DCA_B200_READER_THREADS", threads)
        assert (_lib.read_fasta(str(f), "protein", L) == want).all(), threads
    with open(f, "ab") as fh:
        fh.write(b">short\nACD\n>fine\n" + b"A" * L + b"\n")
    msgs = []
    for threads in ("1", "5"):
        monkeypatch.setenv("DCA_B200_READER_THREADS", threads)
        with pytest.raises(Exception) as e:
            _lib.read_fasta(str(f), "protein", L)
        msgs.append(str(e.value))
    assert msgs[0] == msgs[1] and "has 3 residues, expected 61" in msgs[0]


def test_char_form_helpers_and_slice_couplings(tmp_path):
    """Host helpers of the reference surface: fasta_reader.res_to_char / sequences_to_char_form /
    get_alignment_char_form (fasta_reader.py:53-76, :191-249) and msa_numerics.slice_couplings
    (meanfield_dca/msa_numerics.py:346-375)."""
    from pydca_b200 import fasta_reader, msa_numerics
    assert fasta_reader.res_to_char("rna") == {1: "A", 2: "C", 3: "G", 4: "U", 5: "-"}
    prot = fasta_reader.res_to_char("protein")
    assert len(prot) == 21 and prot[1] == "A" and prot[20] == "Y" and prot[21] == "-"
    assert fasta_reader.sequences_to_char_form([[1, 2, 21, 20], [5, 5, 5, 5]], "PROTEIN") == ["AC-Y", "FFFF"]
    p = tmp_path / "t.fa"
    p.write_text(FASTA)
    assert fasta_reader.get_alignment_char_form(str(p), "protein") == [
        "ACDEFGHIKL", "AC-EF-HI-L", "ACDEFGHIK-", "------ACDE"]
    q, L = 5, 3
    J = np.arange((L * (q - 1)) ** 2, dtype=np.float64).reshape(L * (q - 1), L * (q - 1))
    blk = msa_numerics.slice_couplings(couplings=J, site_pair=(0, 2), num_site_states=q)
    assert blk.shape == (q, q) and (blk[q - 1] == 0).all() and (blk[:, q - 1] == 0).all()
    assert (blk[:q - 1, :q - 1] == J[0:4, 8:12]).all()


def test_fasta_reader_errors(tmp_path):
    from pydca_b200 import _lib
    with pytest.raises(_lib.DcaBackendError, match="Unable to open"):
        _lib.read_fasta(str(tmp_path / "missing.fa"), "protein", 5)
    p = tmp_path / "short.fa"
    p.write_text(">a\nACD\n")
    with pytest.raises(_lib.DcaBackendError, match="expected 5"):
        _lib.read_fasta(str(p), "protein", 5)
    e = tmp_path / "empty.fa"
    e.write_text("\n\n")
    with pytest.raises(_lib.DcaBackendError, match="no sequences"):
        _lib.read_fasta(str(e), "protein", 5)


def test_golden_rf00167_roundtrip_through_reader(tmp_path, golden):
    """Write the golden integer MSA back to FASTA; the reader must return it unchanged."""
    from pydca_b200 import _lib
    from pydca_b200.synthetic import write_fasta
    msa = golden("rf00167_full.npz")["msa"]
    path = write_fasta(msa, str(tmp_path / "rf.fa"), 5)
    assert (_lib.read_fasta(path, "rna", msa.shape[1]) == msa).all()


def test_class_validation_and_defaults(tmp_path):
    """Constructor defaults and validation of the reference classes (plmdca.py:47-73,
    meanfield_dca.py:73-95) without needing a device."""
    from pydca_b200 import PlmDCA, PlmDCAException
    from pydca_b200.synthetic import synthetic_msa, write_fasta
    path = write_fasta(synthetic_msa(20, 12, 21, seed=3), str(tmp_path / "m.fa"), 21)
    inst = PlmDCA(path, "protein")
    assert inst.sequences_len == 12 and inst.num_sequences == 20
    assert inst.sequence_identity == 0.8 and inst.max_iterations == 100
    assert abs(inst.lambda_h - 0.2 * 11) < 1e-12 and abs(inst.lambda_J - 0.2 * 11) < 1e-12
    assert inst.map_index_couplings(0, 1, 0, 0) == 12 * 21
    assert inst.map_index_couplings(1, 2, 3, 4) == 12 * 21 + 11 * 441 + 3 * 21 + 4
    with pytest.raises(PlmDCAException):
        PlmDCA(path, "dna")
    with pytest.raises(PlmDCAException):
        PlmDCA(path, "protein", seqid=1.5)
    with pytest.raises(PlmDCAException):
        PlmDCA(path, "protein", lambda_J=-1.0)
    x = np.arange(12 * 21 + 66 * 441, dtype=np.float32)
    nogap = inst.get_couplings_no_gap_state(x)
    assert nogap.shape == (66 * 400,)
    assert nogap[0] == x[12 * 21] and nogap[20] == x[12 * 21 + 21]        # row 1 of block 0 skips b = 20
    assert len(inst.get_fields_no_gap_state(x)) == 12 * 20


def test_synthetic_generator_is_deterministic():
    from pydca_b200.synthetic import synthetic_msa
    a = synthetic_msa(500, 30, 21)
    b = synthetic_msa(500, 30, 21)
    assert (a == b).all() and a.shape == (500, 30) and a.max() == 20
    assert len({r.tobytes() for r in a}) == 500


def test_cli_surface_and_writer_format(tmp_path):
    """Flags of pydca/plmdca_main.py:265-328 and pydca/mfdca_main.py:316-394 parse, and
    the score table keeps the reference's row format (dca_utilities.py:255-266)."""
    from pydca_b200 import dca_utilities, mfdca_main, plmdca_main
    a = vars(plmdca_main.build_parser().parse_args(
        ["compute_fn", "protein", "x.fa", "--seqid", "0.7", "--lambda_h", "1.0", "--lambda_J", "20", "--max_iterations",
         "500", "--num_threads", "10", "--apc", "--verbose", "--output_dir", "o"]))
    assert a["subcommand_name"] == "compute_fn" and a["max_iterations"] == 500 and a["apc"] and a["lambda_J"] == 20.0
    a = vars(plmdca_main.build_parser().parse_args(["compute_params", "rna", "x.fa", "--ranked_by", "fn_apc",
                                                     "--linear_dist", "5", "--num_site_pairs", "10"]))
    assert a["ranked_by"] == "fn_apc" and a["linear_dist"] == 5
    for cmd in ("compute_di", "compute_fn", "compute_fi", "compute_fij"):
        a = vars(mfdca_main.build_parser().parse_args([cmd, "rna", "x.fa", "--pseudocount", "0.5", "--seqid", "0.8", "--apc"]))
        assert a["subcommand_name"] == cmd and a["pseudocount"] == 0.5
    out = tmp_path / "scores.txt"
    dca_utilities.write_sorted_dca_scores(str(out), [((60, 74), 4.030616483327582), ((0, 1), 0.5)],
                                          metadata=["# meta"], score_type="X")
    lines = out.read_text().splitlines()
    assert lines[0] == "#" + "=" * 70 and lines[1] == "# meta"
    assert lines[-2] == "{0:<7} {1:<14} {2:<35}".format(61, 75, 4.030616483327582)
    assert dca_utilities.get_dca_output_file_path("d", "/a/b/MSA_RF00167.fa", prefix="MFDCA_apc_fn_scores_",
                                                  postfix=".txt") == os.path.join("d", "MFDCA_apc_fn_scores_MSA_RF00167.txt")
