"""numpy restatement of pydca's mean-field DCA numerics and of the FN / APC / DI
scoring of both algorithms.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Citations are relative to /root/reference/.  States here are 0..q-1 with gap =
q-1 (the C++ encoding); the reference's Python side uses 1..q, which only shifts
indices.  Everything is float64 as in the numba code.

Parity status: pinned -- tests/test_oracle_vs_ref.py checks these functions
against the reference's own numba module (imported from /root/reference in the
build container; results frozen in tests/golden/) and against the notebook
known answers examples/pydca_demo.ipynb:100-104, :162-166.
"""
import numpy as np


def pair_index(L, i, j):
    """meanfield_dca/msa_numerics.py:220, plmdca_numerics.cpp:340."""
    return (L * (L - 1)) // 2 - ((L - i) * (L - i - 1)) // 2 + j - i - 1


def sequences_weight(counts):
    """msa_numerics.py:49 -- w = 1/count in float64."""
    return 1.0 / np.asarray(counts, dtype=np.float64)


def single_site_freqs(msa, q, w):
    """msa_numerics.py:54-89: f_i(a) = sum_n w_n [s_ni == a] / Meff, gap kept."""
    msa = np.asarray(msa)
    N, L = msa.shape
    meff = np.sum(w)
    fi = np.zeros((L, q), dtype=np.float64)
    for a in range(q):
        fi[:, a] = ((msa == a) * w[:, None]).sum(axis=0) / meff
    return fi


def reg_single_site_freqs(fi, q, pseudocount):
    """msa_numerics.py:93-125: theta/q + (1-theta) f."""
    return pseudocount / float(q) + (1.0 - pseudocount) * fi


def pair_site_freqs(msa, q, w):
    """msa_numerics.py:183-229: f_ij(a,b) for non-gap a,b, pair(i,j) order."""
    msa = np.asarray(msa)
    N, L = msa.shape
    qm1 = q - 1
    meff = np.sum(w)
    onehot = np.zeros((N, L, qm1), dtype=np.float64)
    n_idx, l_idx = np.nonzero(msa < qm1)
    onehot[n_idx, l_idx, msa[n_idx, l_idx]] = 1.0
    X = onehot.reshape(N, L * qm1)
    F = (X * w[:, None]).T @ X / meff
    F = F.reshape(L, qm1, L, qm1)
    iu, ju = np.triu_indices(L, k=1)
    return np.ascontiguousarray(F[iu, :, ju, :])  # [(pairs), a, b]


def reg_pair_site_freqs(fij, q, pseudocount):
    """msa_numerics.py:232-267: theta/q^2 + (1-theta) f."""
    return pseudocount / float(q * q) + (1.0 - pseudocount) * fij


def construct_corr_mat(reg_fi, reg_fij, L, q):
    """msa_numerics.py:271-318."""
    qm1 = q - 1
    n = L * qm1
    C = np.zeros((n, n), dtype=np.float64)
    pc = 0
    for i in range(L):
        fi = reg_fi[i, :qm1]
        blk = -np.outer(fi, fi)
        blk[np.diag_indices(qm1)] = fi * (1.0 - fi)
        C[i * qm1:(i + 1) * qm1, i * qm1:(i + 1) * qm1] = blk
        for j in range(i + 1, L):
            blk = reg_fij[pc] - np.outer(fi, reg_fi[j, :qm1])
            C[i * qm1:(i + 1) * qm1, j * qm1:(j + 1) * qm1] = blk
            C[j * qm1:(j + 1) * qm1, i * qm1:(i + 1) * qm1] = blk.T
            pc += 1
    return C


def compute_couplings(corr_mat):
    """msa_numerics.py:322-342: J = -inv(C) (LAPACK getrf/getri)."""
    return -1.0 * np.linalg.inv(corr_mat)


def _double_centre_norm(block):
    r = block.mean(axis=1, keepdims=True)
    c = block.mean(axis=0, keepdims=True)
    m = block.mean()
    z = block - c - r + m
    return np.sqrt(np.sum(z * z))


def mf_fn_scores(couplings, L, q):
    """meanfield_dca.py:902-940: Frobenius norm of the double-centred (q-1)^2
    block, pair order (0,1),(0,2)...; returns float64[L(L-1)/2]."""
    qm1 = q - 1
    out = np.zeros(L * (L - 1) // 2, dtype=np.float64)
    k = 0
    for i in range(L):
        for j in range(i + 1, L):
            out[k] = _double_centre_norm(couplings[i * qm1:(i + 1) * qm1, j * qm1:(j + 1) * qm1])
            k += 1
    return out


def plm_fn_scores(x, L, q):
    """plmdca.py:246-268 (gap strip) + :437-481: float32 arithmetic on the
    (q-1)^2 sub-block of each J_ij, as numpy does for float32 inputs."""
    x = np.asarray(x, dtype=np.float32)
    qm1 = q - 1
    J = x[L * q:].reshape(L * (L - 1) // 2, q, q)[:, :qm1, :qm1]
    out = np.zeros(J.shape[0], dtype=np.float32)
    for k in range(J.shape[0]):
        b = J[k]
        avx = np.mean(b, axis=1).reshape(qm1, 1)
        avy = np.mean(b, axis=0).reshape(1, qm1)
        av = np.mean(b)
        z = b - avx - avy + av
        out[k] = np.sqrt(np.sum(z * z))
    return out


def sorted_pairs(scores, L):
    """sorted(..., key=score, reverse=True): stable, ties keep (i,j) order
    (plmdca.py:478, meanfield_dca.py:940)."""
    iu, ju = np.triu_indices(L, k=1)
    order = np.argsort(-np.asarray(scores, dtype=np.float64), kind="stable")
    return [((int(iu[k]), int(ju[k])), scores[k]) for k in order]


def apc(scores, L):
    """plmdca.py:503-519 / meanfield_dca.py:969-984: s_ij - s_i. s_j. / s_..
    (float64 sums; the reference's pinned numpy promotes to float64 here)."""
    s = np.asarray(scores, dtype=np.float64)
    iu, ju = np.triu_indices(L, k=1)
    site_sum = np.zeros(L, dtype=np.float64)
    np.add.at(site_sum, iu, s)
    np.add.at(site_sum, ju, s)
    av_site = site_sum / float(L - 1)
    av_all = av_site.sum() / float(L)
    return s - av_site[iu] * (av_site[ju] / av_all)


def slice_couplings(couplings, i, j, q):
    """msa_numerics.py:346-374: (q,q) block with zero gap row/column."""
    qm1 = q - 1
    out = np.zeros((q, q), dtype=np.float64)
    out[:qm1, :qm1] = couplings[i * qm1:(i + 1) * qm1, j * qm1:(j + 1) * qm1]
    return out


def two_site_model_fields(couplings, reg_fi, L, q, tol=1.0e-4):
    """msa_numerics.py:378-442: per-pair fixed point, stop when the largest
    field change <= 1e-4; returns float64[P, 2, q]."""
    P = L * (L - 1) // 2
    out = np.zeros((P, 2, q), dtype=np.float64)
    k = 0
    for i in range(L - 1):
        fi = reg_fi[i]
        for j in range(i + 1, L):
            fj = reg_fi[j]
            W = np.exp(slice_couplings(couplings, i, j, q))
            hi = np.full(q, 1.0 / q)
            hj = np.full(q, 1.0 / q)
            change = 10.0
            while change > tol:
                xi = W @ hj
                xj = W.T @ hi
                ni = fi / xi
                ni /= ni.sum()
                nj = fj / xj
                nj /= nj.sum()
                change = max(np.abs(ni - hi).max(), np.abs(nj - hj).max())
                hi, hj = ni, nj
            out[k, 0], out[k, 1] = hi, hj
            k += 1
    return out


def direct_info(couplings, fields_ij, reg_fi, L, q):
    """msa_numerics.py:446-533."""
    eps = 1.0e-20
    qm1 = q - 1
    P = L * (L - 1) // 2
    out = np.zeros(P, dtype=np.float64)
    k = 0
    for i in range(L - 1):
        for j in range(i + 1, L):
            W = np.exp(slice_couplings(couplings, i, j, q))
            pdir = W * np.outer(fields_ij[k, 0], fields_ij[k, 1])
            pdir /= pdir.sum()
            fij = np.outer(reg_fi[i], reg_fi[j])
            pr = pdir[:qm1, :qm1] + eps
            fr = fij[:qm1, :qm1] + eps
            out[k] = np.sum(pr * np.log(pr / fr))
            k += 1
    return out


def plm_couplings_as_matrix(x, L, q):
    """Couplings of a plmDCA parameter vector (layout plmdca_numerics.cpp:338-342)
    with the gap state stripped (plmdca.py:246-268), arranged as the
    (L(q-1))^2 matrix the DI code of plmdca/msa_numerics.py:127-311 indexes."""
    x = np.asarray(x)
    qm1 = q - 1
    J = x[L * q:].reshape(L * (L - 1) // 2, q, q)[:, :qm1, :qm1].astype(np.float64)
    M = np.zeros((L * qm1, L * qm1), dtype=np.float64)
    k = 0
    for i in range(L - 1):
        for j in range(i + 1, L):
            M[i * qm1:(i + 1) * qm1, j * qm1:(j + 1) * qm1] = J[k]
            M[j * qm1:(j + 1) * qm1, i * qm1:(i + 1) * qm1] = J[k].T
            k += 1
    return M


def mf_pipeline(msa, q, counts, pseudocount=0.5):
    """meanfield_dca.py:902-988 end to end -> dict of every intermediate."""
    msa = np.asarray(msa)
    N, L = msa.shape
    w = sequences_weight(counts)
    fi = single_site_freqs(msa, q, w)
    rfi = reg_single_site_freqs(fi, q, pseudocount)
    fij = pair_site_freqs(msa, q, w)
    rfij = reg_pair_site_freqs(fij, q, pseudocount)
    C = construct_corr_mat(rfi, rfij, L, q)
    J = compute_couplings(C)
    fn = mf_fn_scores(J, L, q)
    return dict(w=w, meff=w.sum(), fi=fi, reg_fi=rfi, fij=fij, reg_fij=rfij, corr=C,
                couplings=J, fn=fn, fn_apc=apc(fn, L))


# ---- model of the device inverse (csrc/mf.cu) -------------------------------------------
def symmetric_block_gauss_jordan(a, block):
    """Numpy model of gj_invert_device (pydca_b200/csrc/mf.cu): blocked Gauss-Jordan on a
    symmetric positive definite matrix that only ever reads and updates the tiles on and below
    the diagonal.  With K the pivoted and U the unpivoted block indices the intermediates
    satisfy  B_KK, B_UU symmetric and  B_KU = -B_UK^T,  so column block k and row block k are
    rebuilt from lower tiles with signed transposes (gj_gather_kernel) and the rank-`block`
    update touches the tiles j <= i only.  The upper tiles are poisoned with NaN here: a model
    that passed would have failed loudly had it read one.  Returns inv(a), like
    numpy.linalg.inv (msa_numerics.py:340 of the reference negates it afterwards).
    Test infrastructure only."""
    a = np.array(a, dtype=np.float64, copy=True)
    n = a.shape[0]
    if n % block:
        raise ValueError("the model wants n to be a multiple of the block size")
    nb = n // block

    def tile(i, j):
        return a[i * block:(i + 1) * block, j * block:(j + 1) * block]

    for i in range(nb):
        for j in range(i + 1, nb):
            tile(i, j)[:] = np.nan
    eye = np.eye(block)
    for k in range(nb):
        dinv = unblocked_gauss_jordan(tile(k, k))
        ck = np.empty((n, block))
        arow = np.empty((block, n))
        for i in range(nb):
            ck[i * block:(i + 1) * block] = tile(i, k) if i >= k else -tile(k, i).T
        for j in range(nb):
            arow[:, j * block:(j + 1) * block] = tile(k, j) if j < k else (eye if j == k else tile(j, k).T)
        for i in range(k, nb):
            tile(i, k)[:] = eye if i == k else 0.0
        rk = dinv @ arow
        for i in range(nb):
            if i == k:
                continue
            for j in range(i + 1):
                tile(i, j)[:] -= ck[i * block:(i + 1) * block] @ rk[:, j * block:(j + 1) * block]
        for j in range(k + 1):
            tile(k, j)[:] = rk[:, j * block:(j + 1) * block]
    # gj_mirror_kernel: every element above the diagonal is copied from below it
    return np.tril(a) + np.tril(a, -1).T


def unblocked_gauss_jordan(a):
    """Numpy model of gj_diag_inv_kernel: in-place Gauss-Jordan without pivoting, every pivot
    step computed from the OLD pivot row and column (what the register-resident kernel
    publishes through shared memory)."""
    a = np.array(a, dtype=np.float64, copy=True)
    for p in range(a.shape[0]):
        prow, pcol = a[p, :].copy(), a[:, p].copy()
        piv = 1.0 / prow[p]
        rowv = prow * piv
        new = a - np.outer(pcol, rowv)
        new[p, :] = rowv
        new[:, p] = -pcol * piv
        new[p, p] = piv
        a = new
    return a
