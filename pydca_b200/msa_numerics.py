"""The mean-field DCA operator API of pydca/meanfield_dca/msa_numerics.py -- same
function names, keyword arguments, shapes and dtypes -- computed by the sm_100a
kernels behind include/dca_b200.h.  ``alignment_data`` uses the reference's Python
encoding (states 1..q, gap = q); it is shifted to 0..q-1 for the device.

No function here computes on the CPU beyond trivial element-wise reshaping: the
O(N^2 L), O(N L^2 q) and O(n^3) work is CUDA, and without a device every call
raises ``DcaBackendError``.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import _f32p, _f64p, _u8p, check, lib, ptr

__all__ = ["compute_sequences_weight", "compute_single_site_freqs", "get_reg_single_site_freqs",
           "compute_pair_site_freqs", "get_reg_pair_site_freqs", "construct_corr_mat",
           "compute_couplings", "compute_two_site_model_fields", "compute_direct_info",
           "corr_mat_from_alignment", "frobenius_norm_scores", "mean_field_scores",
           "compute_pair_site_freqs_serial", "slice_couplings"]


def _zero_based(alignment_data, num_site_states=None):
    a = np.asarray(alignment_data)
    if a.ndim != 2 or a.size == 0:
        raise ValueError("alignment_data must be a non-empty 2d integer array")
    if a.min() < 1:
        raise ValueError("alignment_data must use states 1..q (gap = q)")
    if num_site_states is not None and a.max() > num_site_states:
        raise ValueError("alignment_data contains state %d > num_site_states %d" % (a.max(), num_site_states))
    if a.max() > 32:        # the uint8 cast below would wrap; the reweighting kernel packs 5-bit codes
        raise ValueError("alignment_data contains state %d; at most 32 states are supported" % a.max())
    return np.ascontiguousarray(a - 1, dtype=np.uint8)


def compute_sequences_weight(alignment_data=None, seqid=None):
    """msa_numerics.py:14-50 -> float64[N] = 1 / #{j : ident(i,j)/L > seqid}."""
    msa = _zero_based(alignment_data)
    counts = _lib.sequence_counts(msa, seqid, compare_f64=True)
    return 1.0 / counts.astype(np.float64)


def _freqs(msa, num_site_states, seqs_weight, want_pairs):
    n, l = msa.shape
    q = int(num_site_states)
    w = np.ascontiguousarray(seqs_weight, dtype=np.float64)
    if w.shape != (n,):
        raise ValueError("seqs_weight must have one entry per sequence")
    fi = np.zeros((l, q), dtype=np.float64)
    fij = np.zeros((l * (l - 1) // 2, q - 1, q - 1), dtype=np.float64) if want_pairs else None
    check(lib().dca_mf_freqs(ptr(msa, _u8p), ptr(w, _f64p), n, l, q, ptr(fi, _f64p),
                             ptr(fij, _f64p) if want_pairs else None), "dca_mf_freqs")
    return fi, fij


def compute_single_site_freqs(alignment_data=None, num_site_states=None, seqs_weight=None):
    """msa_numerics.py:54-89 -> float64[L, q] (gap column kept)."""
    return _freqs(_zero_based(alignment_data, num_site_states), num_site_states, seqs_weight, False)[0]


def get_reg_single_site_freqs(single_site_freqs=None, seqs_len=None, num_site_states=None, pseudocount=None):
    """msa_numerics.py:93-125: theta/q + (1-theta) f, in place like the reference (:119)."""
    f = single_site_freqs
    f[:seqs_len, :num_site_states] = (np.float64(pseudocount) / np.float64(num_site_states)
                                      + (1.0 - pseudocount) * f[:seqs_len, :num_site_states])
    return f


def compute_pair_site_freqs(alignment_data=None, num_site_states=None, seqs_weight=None):
    """msa_numerics.py:183-229 -> float64[L(L-1)/2, q-1, q-1] in pair(i,j) order."""
    return _freqs(_zero_based(alignment_data, num_site_states), num_site_states, seqs_weight, True)[1]


def compute_pair_site_freqs_serial(alignment_data=None, num_site_states=None, seqs_weight=None):
    """msa_numerics.py:130-179, the reference's serial twin of compute_pair_site_freqs: same
    result, same device kernel here."""
    return compute_pair_site_freqs(alignment_data=alignment_data, num_site_states=num_site_states,
                                   seqs_weight=seqs_weight)


def slice_couplings(couplings=None, site_pair=None, num_site_states=None):
    """msa_numerics.py:346-375 -> float64[q, q]: the (q-1) x (q-1) block of the pair, gap row and
    column zero."""
    q = int(num_site_states)
    i, j = int(site_pair[0]), int(site_pair[1])
    out = np.zeros((q, q), dtype=np.float64)
    out[:q - 1, :q - 1] = np.asarray(couplings)[i * (q - 1):(i + 1) * (q - 1), j * (q - 1):(j + 1) * (q - 1)]
    return out


def get_reg_pair_site_freqs(pair_site_freqs=None, seqs_len=None, num_site_states=None, pseudocount=None):
    """msa_numerics.py:232-267: theta/q^2 + (1-theta) f, in place (:257)."""
    f = pair_site_freqs
    f[...] = pseudocount / float(num_site_states * num_site_states) + (1.0 - pseudocount) * f
    return f


def construct_corr_mat(reg_fi=None, reg_fij=None, seqs_len=None, num_site_states=None):
    """msa_numerics.py:271-318 -> float64[n, n], n = L(q-1)."""
    l, q = int(seqs_len), int(num_site_states)
    rfi = np.ascontiguousarray(reg_fi, dtype=np.float64)
    rfij = np.ascontiguousarray(reg_fij, dtype=np.float64)
    n = l * (q - 1)
    corr = np.zeros((n, n), dtype=np.float64)
    check(lib().dca_mf_corr_from_freqs(ptr(rfi, _f64p), ptr(rfij, _f64p), l, q, ptr(corr, _f64p)),
          "dca_mf_corr_from_freqs")
    return corr


def compute_couplings(corr_mat=None):
    """msa_numerics.py:322-342 -> -inv(corr_mat).  A singular matrix raises
    (the reference surfaces a LAPACK error, meanfield_dca.py:544-548)."""
    c = np.ascontiguousarray(corr_mat, dtype=np.float64)
    if c.ndim != 2 or c.shape[0] != c.shape[1]:
        raise ValueError("corr_mat must be square")
    out = np.zeros_like(c)
    check(lib().dca_mf_couplings(ptr(c, _f64p), c.shape[0], ptr(out, _f64p)), "dca_mf_couplings")
    return out


def _direct_info(couplings, reg_fi, seqs_len, num_site_states, want_fields):
    l, q = int(seqs_len), int(num_site_states)
    j = np.ascontiguousarray(couplings, dtype=np.float64)
    rfi = np.ascontiguousarray(reg_fi, dtype=np.float64)
    npairs = l * (l - 1) // 2
    di = np.zeros(npairs, dtype=np.float64)
    fields = np.zeros((npairs, 2, q), dtype=np.float64) if want_fields else None
    check(lib().dca_mf_direct_info(ptr(j, _f64p), ptr(rfi, _f64p), l, q, ptr(di, _f64p),
                                   ptr(fields, _f64p) if want_fields else None), "dca_mf_direct_info")
    return di, fields


def compute_two_site_model_fields(couplings=None, reg_fi=None, seqs_len=None, num_site_states=None):
    """msa_numerics.py:378-442 -> float64[P, 2, q]."""
    return _direct_info(couplings, reg_fi, seqs_len, num_site_states, True)[1]


def compute_direct_info(couplings=None, fields_ij=None, reg_fi=None, seqs_len=None, num_site_states=None):
    """msa_numerics.py:446-533 -> float64[P].  ``fields_ij`` is accepted for
    signature compatibility; the kernel recomputes the (deterministic) fixed
    point and the DI in one pass."""
    return _direct_info(couplings, reg_fi, seqs_len, num_site_states, False)[0]


# ---- fused paths used by MeanFieldDCA (one H2D of the MSA, no fij round trip) ----
def corr_mat_from_alignment(msa_zero_based, num_site_states, seqs_weight, pseudocount):
    """(reg_fi[L,q], corr[n,n]) straight from the alignment: frequencies,
    pseudocount and covariance assembly stay on the device."""
    msa = np.ascontiguousarray(msa_zero_based, dtype=np.uint8)
    n, l = msa.shape
    q = int(num_site_states)
    w = np.ascontiguousarray(seqs_weight, dtype=np.float64)
    dim = l * (q - 1)
    rfi = np.zeros((l, q), dtype=np.float64)
    corr = np.zeros((dim, dim), dtype=np.float64)
    check(lib().dca_mf_corr_mat(ptr(msa, _u8p), ptr(w, _f64p), n, l, q, float(pseudocount),
                                ptr(rfi, _f64p), ptr(corr, _f64p)), "dca_mf_corr_mat")
    return rfi, corr


def mean_field_scores(msa_zero_based, num_site_states, seqs_weight, pseudocount, want_couplings=False,
                      want_di=False):
    """The whole pipeline in one device-resident pass (dca_mf_scores).  Returns a dict with
    reg_fi, fn, fn_apc, stage_ms and, on request, couplings and di."""
    msa = np.ascontiguousarray(msa_zero_based, dtype=np.uint8)
    n, l = msa.shape
    q = int(num_site_states)
    w = np.ascontiguousarray(seqs_weight, dtype=np.float64)
    npairs = l * (l - 1) // 2
    dim = l * (q - 1)
    out = {"reg_fi": np.zeros((l, q)), "fn": np.zeros(npairs), "fn_apc": np.zeros(npairs),
           "stage_ms": np.zeros(5, dtype=np.float32)}
    if want_couplings:
        out["couplings"] = np.zeros((dim, dim))
    if want_di:
        out["di"] = np.zeros(npairs)
    check(lib().dca_mf_scores(ptr(msa, _u8p), ptr(w, _f64p), n, l, q, float(pseudocount),
                              ptr(out["reg_fi"], _f64p),
                              ptr(out["couplings"], _f64p) if want_couplings else None,
                              ptr(out["fn"], _f64p), ptr(out["fn_apc"], _f64p),
                              ptr(out["di"], _f64p) if want_di else None,
                              ptr(out["stage_ms"], _f32p)), "dca_mf_scores")
    return out


def frobenius_norm_scores(couplings, seqs_len, num_site_states):
    """(FN, FN_APC) float64[P] of an n x n couplings matrix (meanfield_dca.py:902-988)."""
    j = np.ascontiguousarray(couplings, dtype=np.float64)
    npairs = seqs_len * (seqs_len - 1) // 2
    fn = np.zeros(npairs, dtype=np.float64)
    apc = np.zeros(npairs, dtype=np.float64)
    check(lib().dca_mf_fn_apc(ptr(j, _f64p), int(seqs_len), int(num_site_states), ptr(fn, _f64p),
                              ptr(apc, _f64p)), "dca_mf_fn_apc")
    return fn, apc


def plm_direct_info(x, reg_fi, seqs_len, num_site_states, want_fields=False):
    """DI from a plmDCA parameter vector (pydca/plmdca/msa_numerics.py:156-311)."""
    l, q = int(seqs_len), int(num_site_states)
    xv = np.ascontiguousarray(x, dtype=np.float32)
    rfi = np.ascontiguousarray(reg_fi, dtype=np.float64)
    npairs = l * (l - 1) // 2
    di = np.zeros(npairs, dtype=np.float64)
    fields = np.zeros((npairs, 2, q), dtype=np.float64) if want_fields else None
    check(lib().dca_plm_direct_info(ptr(xv, _f32p), ptr(rfi, _f64p), l, q, ptr(di, _f64p),
                                    ptr(fields, _f64p) if want_fields else None), "dca_plm_direct_info")
    return (di, fields) if want_fields else di
