"""The class and command-line layer against the reference's own (-m gpu).

tests/golden/class_layer.npz holds every output file the UNMODIFIED reference's `mfdca` and
`plmdca` commands write on two toy alignments (tests/golden/make_class_goldens.py ran
pydca.mfdca_main / pydca.plmdca_main.execute_from_command_line and MeanFieldDCA.compute_fields).
Here the same commands run through pydca_b200's command lines and the files are compared line
by line: header lines and integer columns exactly, floating-point columns to the stated tolerance,
ranked site pairs in the same order.

plmDCA: the reference's fitted parameter vector is injected in place of the backend call, so that
compute_params / FN / APC / DI / the writers are compared on identical parameters (the optimiser
has its own tests).  The reference's files were produced under numpy 2.x, where its float32 FN
scores stay float32 through APC (under the numpy <= 1.15 it pins they are promoted to float64, which
is what pydca_b200 does), and APC subtracts nearly equal numbers: hence 1e-5 on plmDCA scores,
relative to at least a tenth of the file's largest score; 1e-9 on mean-field ones.
"""
import os
import re

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NUM = re.compile(r"^[-+]?(\d+\.?\d*|\.\d+)([eE][-+]?\d+)?$")


def _tokens(line):
    return [t for t in re.split(r"[,\s]+", line.strip()) if t]


def _compare_files(got_text, want_text, rtol, what, floor_frac=0.0):
    """floor_frac: errors are taken relative to max(|reference value|, floor_frac * largest value of
    the file) -- for float32 results of a subtraction (APC) the rounding error scales with the
    operands, not with the difference."""
    got, want = got_text.splitlines(), want_text.splitlines()
    assert len(got) == len(want), "%s: %d lines, reference has %d" % (what, len(got), len(want))
    worst = 0.0
    vals = [abs(float(t)) for w in want if not w.startswith("#") for t in _tokens(w)
            if NUM.match(t) and ("." in t or "e" in t.lower())]
    floor = max(1e-3, floor_frac * (max(vals) if vals else 0.0))
    for k, (g, w) in enumerate(zip(got, want)):
        if g == w:
            continue
        tg, tw = _tokens(g), _tokens(w)
        assert len(tg) == len(tw), "%s line %d: %r vs %r" % (what, k + 1, g, w)
        for a, b in zip(tg, tw):
            if a == b:
                continue
            assert NUM.match(a) and NUM.match(b), "%s line %d: %r vs %r" % (what, k + 1, g, w)
            assert ("." in b or "e" in b.lower()), "%s line %d: integer column differs: %r vs %r" % (what, k + 1, g, w)
            fa, fb = float(a), float(b)
            err = abs(fa - fb) / max(abs(fb), floor)
            worst = max(worst, err)
            assert err <= rtol, "%s line %d: %r vs %r (rel %.2e > %.1e)" % (what, k + 1, a, b, err, rtol)
    return worst


def _golden_files(d, prefix):
    return {k[len(prefix) + 1:]: bytes(d[k]).decode() for k in d.files if k.startswith(prefix + "/")}


MF_RUNS = [("compute_fn", dict(apc=True), "fn_apc"), ("compute_fn", dict(apc=False), "fn_raw"),
           ("compute_di", dict(apc=True), "di_apc"), ("compute_di", dict(apc=False), "di_raw"),
           ("compute_params", dict(), "params_default"),
           ("compute_params", dict(ranked_by="di", linear_dist=2, num_site_pairs=7), "params_di"),
           ("compute_fi", dict(), "fi"), ("compute_fij", dict(), "fij")]


@pytest.mark.parametrize("bio,q", [("protein", 21), ("rna", 5)])
def test_mfdca_command_line_files_match_the_reference(tmp_path, golden, bio, q):
    from pydca_b200 import mfdca_main
    from pydca_b200.meanfield_dca import MeanFieldDCA
    from pydca_b200.synthetic import write_fasta
    d = golden("class_layer.npz")
    fasta = write_fasta(d["%s/msa" % bio], str(tmp_path / ("toy_%s.fa" % bio)), q)
    for cmd, kw, tag in MF_RUNS:
        want = _golden_files(d, "mf/%s/%s" % (bio, tag))
        if not want:
            continue                                   # compute_fij is stored for the RNA case only
        out = str(tmp_path / ("mf_" + tag))
        mfdca_main.execute_from_command_line(msa_file=fasta, biomolecule=bio, seqid=0.8, pseudocount=0.5,
                                             the_command=cmd, output_dir=out, **kw)
        assert sorted(os.listdir(out)) == sorted(want), (cmd, tag)
        for name, text in want.items():
            _compare_files(open(os.path.join(out, name)).read(), text, 1e-9, "mfdca %s %s %s" % (bio, tag, name))
    inst = MeanFieldDCA(fasta, bio, pseudocount=0.5, seqid=0.8)
    fields = inst.compute_fields()                     # meanfield_dca.py:588-633
    got = np.array([fields[i] for i in range(inst.sequences_len)])
    assert np.abs(got - d["mf/%s/compute_fields" % bio]).max() <= 1e-9 * np.abs(got).max()


PLM_RUNS = [("compute_fn", dict(apc=True), "fn_apc"), ("compute_fn", dict(apc=False), "fn_raw"),
            ("compute_di", dict(apc=True), "di_apc"), ("compute_di", dict(apc=False), "di_raw"),
            ("compute_params", dict(), "params_default"),
            ("compute_params", dict(ranked_by="fn", linear_dist=1, num_site_pairs=9), "params_fn")]


@pytest.mark.parametrize("bio,q", [("protein", 21), ("rna", 5)])
def test_plmdca_command_line_files_match_the_reference(tmp_path, golden, monkeypatch, bio, q):
    from pydca_b200 import plmdca as plm_mod, plmdca_main
    from pydca_b200.synthetic import write_fasta
    d = golden("class_layer.npz")
    fasta = write_fasta(d["%s/msa" % bio], str(tmp_path / ("toy_%s.fa" % bio)), q)
    x_ref = d["plm/%s/x_fit" % bio]
    lam = float(d["plm/%s/lambda" % bio])
    monkeypatch.setattr(plm_mod.PlmDCA, "get_fields_and_couplings_from_backend", lambda self: x_ref.copy())
    for cmd, kw, tag in PLM_RUNS:
        want = _golden_files(d, "plm/%s/%s" % (bio, tag))
        out = str(tmp_path / ("plm_" + tag))
        plmdca_main.execute_from_command_line(bio, fasta, the_command=cmd, seqid=0.8, lambda_h=lam, lambda_J=lam,
                                              max_iterations=int(d["plm_iters"]), num_threads=1, output_dir=out, **kw)
        assert sorted(os.listdir(out)) == sorted(want), (cmd, tag)
        for name, text in want.items():
            _compare_files(open(os.path.join(out, name)).read(), text, 1e-5, "plmdca %s %s %s" % (bio, tag, name),
                           floor_frac=0.1)


@pytest.mark.parametrize("bio,q", [("protein", 21), ("rna", 5)])
def test_plmdca_class_end_to_end_near_the_reference_fit(tmp_path, golden, bio, q):
    """Without the injection: the GPU fit in reference arithmetic, 20 iterations like the golden."""
    from pydca_b200.plmdca import PlmDCA
    from pydca_b200.synthetic import write_fasta
    d = golden("class_layer.npz")
    fasta = write_fasta(d["%s/msa" % bio], str(tmp_path / ("toy_%s.fa" % bio)), q)
    lam = float(d["plm/%s/lambda" % bio])
    inst = PlmDCA(fasta, bio, seqid=0.8, lambda_h=lam, lambda_J=lam, max_iterations=int(d["plm_iters"]),
                  arithmetic="reference")
    x = np.asarray(inst.get_fields_and_couplings_from_backend())
    x_ref = d["plm/%s/x_fit" % bio]
    assert np.abs(x - x_ref).max() <= 5e-4 * np.abs(x_ref).max()
