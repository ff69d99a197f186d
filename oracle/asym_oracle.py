"""Asymmetric (per-site) plmDCA objective -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference has no asymmetric mode (it fits one symmetric J, plmdca_numerics.cpp:436-607), so
this numpy restatement cannot be pinned to a reference output of its own.  It is pinned through
the joint objective instead: the conditional of site i is the reference's (logits :499-517 with the
carry of :492, softmax :519-530, -w log p :536, residuals :538-567), and at a symmetric point
W[(j,b),(i,a)] = W[(i,a),(j,b)] the data term and its gradient must reduce to those of the joint
objective (oracle port, itself bit-exact against the reference): tests/test_oracle_golden.py.

Parameters: h [L, q] and W [L*q, L*q] with W[j*q + b, i*q + a] = J^(i)_ij(a, b), the copy of the
couplings that only site i's conditional sees (diagonal blocks are identically zero).
    f = sum_i [ -sum_n w_n log p_n(i, s_ni) + lambda_h |h_i|^2 + (lambda_J / 2) sum_{j != i} |J^(i)_ij|^2 ]
"""
import numpy as np


def offdiag_mask(L, q):
    m = np.ones((L, q, L, q), dtype=bool)
    for i in range(L):
        m[i, :, i, :] = False
    return m.reshape(L * q, L * q)


def asym_objective(msa, w, q, lambda_h, lambda_J, h, W, carry=True):
    """float64 (f, gh [L, q], gW [L*q, L*q]) over all sites; restrict by slicing columns."""
    msa = np.asarray(msa)
    N, L = msa.shape
    w = np.asarray(w, dtype=np.float64)
    h = np.asarray(h, dtype=np.float64).reshape(L, q)
    mask = offdiag_mask(L, q)
    W = np.where(mask, np.asarray(W, dtype=np.float64), 0.0)
    rows = np.arange(L)[None, :] * q + msa                      # row (j, s_nj) of W for every n
    f = 0.0
    gh = np.zeros((L, q))
    gW = np.zeros((L * q, L * q))
    p_prev = np.zeros((L, q))
    sites = np.arange(L)
    for n in range(N):
        logits = h + W[rows[n]].sum(axis=0).reshape(L, q)
        if carry:
            logits = logits + p_prev                             # plmdca_numerics.cpp:492, :499
        logits -= logits.max(axis=1, keepdims=True)
        p = np.exp(logits)
        p /= p.sum(axis=1, keepdims=True)
        p_prev = p
        f -= w[n] * np.log(p[sites, msa[n]]).sum()
        r = w[n] * p
        r[sites, msa[n]] -= w[n]
        gh += r
        gW[rows[n], :] += r.reshape(-1)[None, :]
    gW = np.where(mask, gW, 0.0)
    f += lambda_h * (h ** 2).sum() + 0.5 * lambda_J * (W ** 2).sum()
    return f, gh + 2.0 * lambda_h * h, gW + lambda_J * W


def symmetrise(h, W, L, q):
    """Reference parameter vector: fields, then J_ij = (J^(i)_ij + J^(j)_ji^T) / 2 for i < j in pair order."""
    W4 = np.asarray(W, dtype=np.float64).reshape(L, q, L, q)     # [j, b, i, a]
    blocks = []
    for i in range(L):
        for j in range(i + 1, L):
            mine = W4[j, :, i, :].T                              # [a, b] = J^(i)_ij(a, b)
            theirs = W4[i, :, j, :]                              # [a, b] = J^(j)_ji(b, a) read as (a, b)
            blocks.append(0.5 * (mine + theirs))
    return np.concatenate([np.asarray(h, dtype=np.float64).reshape(-1)] + [b.reshape(-1) for b in blocks])


def from_reference_vector(x, L, q):
    """(h [L, q], symmetric W [L*q, L*q]) of a parameter vector in the reference layout."""
    x = np.asarray(x, dtype=np.float64)
    h = x[:L * q].reshape(L, q)
    W4 = np.zeros((L, q, L, q))
    k = L * q
    for i in range(L):
        for j in range(i + 1, L):
            J = x[k:k + q * q].reshape(q, q)                     # [a, b]
            k += q * q
            W4[j, :, i, :] = J.T                                 # W[(j,b),(i,a)] = J_ij(a,b)
            W4[i, :, j, :] = J                                   # W[(i,a),(j,b)] = J_ij(a,b)
    return h, W4.reshape(L * q, L * q)
