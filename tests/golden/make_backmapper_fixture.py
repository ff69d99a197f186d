#!/usr/bin/env python
"""Fixture for the reference-sequence path: the reference's own example inputs
(examples/MSA_RF00167.fa, examples/ref_RF00167.fa -- data files, stored verbatim) and outputs
of the parts of pydca/sequence_backmapper/sequence_backmapper.py that run without Biopython
(the static align_subsequences, :286-336), produced by the UNMODIFIED reference module imported
with tests/golden/bio_stub on the path.  Runs only in the build container:

    python tests/golden/make_backmapper_fixture.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, os.path.join(HERE, "bio_stub"))
sys.path.insert(0, REF)

from pydca.sequence_backmapper.sequence_backmapper import SequenceBackmapper  # noqa: E402


def main():
    store = {"msa_fasta": np.frombuffer(open(os.path.join(REF, "examples/MSA_RF00167.fa"), "rb").read(), dtype=np.uint8),
             "ref_fasta": np.frombuffer(open(os.path.join(REF, "examples/ref_RF00167.fa"), "rb").read(), dtype=np.uint8)}
    rng = np.random.default_rng(7)
    cases, outs = [], []
    # the docstring example of the module (:480-488) and random gap patterns
    cases.append(("AAAAAAAA---AA", "---BBB--BB-BBBB----", 9))
    for _ in range(40):
        nres = int(rng.integers(3, 25))
        ref_mid = "".join(rng.choice(list("ACGU-"), p=[.22, .22, .22, .22, .12]) for _ in range(nres + int(rng.integers(0, 6))))
        ref_mid = "A" + ref_mid + "C"                     # an aligned region starts and ends with residues
        ntemp = sum(1 for c in ref_mid if c != "-") + int(rng.integers(-2, 3))
        ntemp = max(1, ntemp)
        tmpl = []
        left = ntemp + int(rng.integers(0, 4))
        while left > 0:
            if rng.random() < 0.25:
                tmpl.append("-")
            else:
                tmpl.append("G")
                left -= 1
        tmpl += ["-"] * int(rng.integers(0, 4))
        cases.append((ref_mid, "".join(tmpl), ntemp))
    for ref_mid, tmpl, k in cases:
        try:
            outs.append(SequenceBackmapper.align_subsequences(ref_middle_subseq=ref_mid, template_subseq_in_msa=tmpl,
                                                              num_res_middle_template=k))
        except IndexError:
            outs.append("!IndexError")
    store["align_subsequences_cases"] = np.array(["%s|%s|%d" % c for c in cases])
    store["align_subsequences_outputs"] = np.array(outs)
    np.savez_compressed(os.path.join(HERE, "backmapper.npz"), **store)
    print("wrote backmapper.npz:", len(cases), "align_subsequences cases;", outs[0])


if __name__ == "__main__":
    main()
