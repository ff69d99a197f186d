#!/usr/bin/env python
"""Goldens of the reference's CLASS and COMMAND-LINE layer: compute_params,
compute_fields, and every output file the `mfdca` / `plmdca` commands write, produced by the
UNMODIFIED reference code itself.

Runs only in the build container:

    python tests/golden/make_class_goldens.py

* the reference package is copied to a scratch directory (its PlmDCA class wants its shared
  object next to plmdca.py; /root/reference is read-only) and the C++ backend is compiled there
  with the reference's own flags (setup.py:17-34);
* Biopython is absent: tests/golden/bio_stub/ supplies the three names the DCA classes touch
  (AlignIO.read, Align.MultipleSeqAlignment, an importable pairwise2) -- see its docstring;
* pydca.mfdca_main.execute_from_command_line / pydca.plmdca_main.execute_from_command_line are
  then called exactly as the console scripts do, and the files they write are stored verbatim in
  tests/golden/class_layer.npz together with the input alignments and the plmDCA parameter
  vector of the reference fit (so that the test can take the optimiser out of the comparison).
"""
import glob
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from pydca_b200.synthetic import synthetic_msa, write_fasta  # noqa: E402

CASES = {"protein": dict(N=300, L=20, q=21, seed=2027), "rna": dict(N=150, L=12, q=5, seed=2028)}
PLM_ITERS = 20


def stage_reference(scratch):
    dst = os.path.join(scratch, "pydca")
    shutil.copytree(os.path.join(REF, "pydca"), dst)
    src = os.path.join(dst, "plmdca")
    so = os.path.join(src, "_plmdcaBackend.so")
    subprocess.run(["g++", "-fopenmp", "-std=c++11", "-O3", "-fPIC", "-shared", "-I" + src, "-o", so,
                    os.path.join(src, "lbfgs/lib/lbfgs.cpp"), os.path.join(src, "plmdca_numerics.cpp"),
                    os.path.join(src, "plmdcaBackend.cpp")], check=True)
    sys.path.insert(0, os.path.join(HERE, "bio_stub"))
    sys.path.insert(0, scratch)


def collect(out_dir, tag, store):
    for path in sorted(glob.glob(os.path.join(out_dir, "*"))):
        key = "%s/%s" % (tag, os.path.basename(path))
        store[key] = np.frombuffer(open(path, "rb").read(), dtype=np.uint8)
        print("   ", key, os.path.getsize(path), "bytes")


def main():
    scratch = tempfile.mkdtemp(prefix="pydca_ref_")
    stage_reference(scratch)
    from pydca import mfdca_main, plmdca_main                    # the reference's own modules
    from pydca.meanfield_dca.meanfield_dca import MeanFieldDCA
    from pydca.plmdca.plmdca import PlmDCA

    store = {}
    for bio, c in CASES.items():
        msa = synthetic_msa(c["N"], c["L"], c["q"], seed=c["seed"])
        fasta = write_fasta(msa, os.path.join(scratch, "toy_%s.fa" % bio), c["q"])
        store["%s/msa" % bio] = msa
        # ---- mean-field: every sub-command of the console script (mfdca_main.py:130-296) ----------
        for cmd, kw, tag in [("compute_fn", dict(apc=True), "fn_apc"), ("compute_fn", dict(apc=False), "fn_raw"),
                             ("compute_di", dict(apc=True), "di_apc"), ("compute_di", dict(apc=False), "di_raw"),
                             ("compute_params", dict(), "params_default"),
                             ("compute_params", dict(ranked_by="di", linear_dist=2, num_site_pairs=7), "params_di"),
                             ("compute_fi", dict(), "fi")] + ([("compute_fij", dict(), "fij")] if bio == "rna" else []):
            out = os.path.join(scratch, "mf_%s_%s" % (bio, tag))
            mfdca_main.execute_from_command_line(msa_file=fasta, biomolecule=bio, seqid=0.8, pseudocount=0.5,
                                                 the_command=cmd, output_dir=out, **kw)
            collect(out, "mf/%s/%s" % (bio, tag), store)
        inst = MeanFieldDCA(fasta, bio, pseudocount=0.5, seqid=0.8)
        fields = inst.compute_fields()                             # meanfield_dca.py:588-633
        store["mf/%s/compute_fields" % bio] = np.array([fields[i] for i in range(c["L"])])
        # ---- plmDCA: the class on its own backend (plmdca.py:202-434), then the console script -------
        lam = 0.2 * (c["L"] - 1)
        pinst = PlmDCA(fasta, bio, seqid=0.8, lambda_h=lam, lambda_J=lam, max_iterations=PLM_ITERS, num_threads=1)
        x = np.array(pinst.get_fields_and_couplings_from_backend(), dtype=np.float32)
        store["plm/%s/x_fit" % bio] = x
        store["plm/%s/lambda" % bio] = np.float64(lam)
        for cmd, kw, tag in [("compute_fn", dict(apc=True), "fn_apc"), ("compute_fn", dict(apc=False), "fn_raw"),
                             ("compute_di", dict(apc=True), "di_apc"), ("compute_di", dict(apc=False), "di_raw"),
                             ("compute_params", dict(), "params_default"),
                             ("compute_params", dict(ranked_by="fn", linear_dist=1, num_site_pairs=9), "params_fn")]:
            out = os.path.join(scratch, "plm_%s_%s" % (bio, tag))
            plmdca_main.execute_from_command_line(bio, fasta, the_command=cmd, seqid=0.8, lambda_h=lam, lambda_J=lam,
                                                  max_iterations=PLM_ITERS, num_threads=1, output_dir=out, **kw)
            collect(out, "plm/%s/%s" % (bio, tag), store)
    store["plm_iters"] = np.int64(PLM_ITERS)
    np.savez_compressed(os.path.join(HERE, "class_layer.npz"), **store)
    print("wrote class_layer.npz with", len(store), "entries")
    shutil.rmtree(scratch, ignore_errors=True)


if __name__ == "__main__":
    main()
