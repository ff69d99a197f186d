"""pydca_b200 -- B200-native direct coupling analysis behind pydca's own surface.

Hot path only: plmDCA objective/gradient + L-BFGS, all-pairs
Hamming reweighting, mean-field frequencies / covariance / inverse, FN + APC + DI
scoring, as hand-written sm_100a CUDA in ``_plmdcaBackend_b200.so`` (C ABI:
include/dca_b200.h), plus the host-side mirror of the reference's interface:

    PlmDCA, MeanFieldDCA           pydca/plmdca/plmdca.py, pydca/meanfield_dca/meanfield_dca.py
    msa_numerics                   pydca/meanfield_dca/msa_numerics.py operator functions
    fasta_reader                   pydca/fasta_reader/fasta_reader.py (Biopython-free)
    plmdca_main, mfdca_main        the `plmdca` / `mfdca` command lines

Importing the package does not need a GPU; computing does (no CPU fallback).
"""
from .plmdca import PlmDCA, PlmDCAException  # noqa: F401
from .meanfield_dca import MeanFieldDCA, MeanFieldDCAException  # noqa: F401
from ._lib import DcaBackendError  # noqa: F401

__version__ = "0.1.0"
