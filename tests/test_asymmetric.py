"""The asymmetric (per-site) estimator -- SURVEY 8e last row.

The reference has no such mode, so the chain of evidence is: the numpy oracle
(oracle/asym_oracle.py) reduces to the joint objective of the reference-pinned port at
symmetric points (CPU test below); the CUDA path matches the oracle per evaluation
(<= 1e-5, like the joint path); the fit reaches the optimum a float64 L-BFGS finds on the
oracle; the result does not depend on how the sites are spread over GPUs."""
import numpy as np
import pytest

from conftest import relmax
from oracle import asym_oracle as A


def _problem(port, N, L, q, seed):
    from pydca_b200.synthetic import synthetic_msa
    msa = synthetic_msa(N, L, q, seed=seed)
    w = port.plm_weights(port.reweight_counts(msa, 0.8))
    return msa, w


# ---- CPU: the oracle is the reference's objective with the symmetry constraint dropped --------
@pytest.mark.parametrize("N,L,q", [(40, 6, 5), (25, 4, 21)])
def test_asym_oracle_reduces_to_the_joint_objective(port, N, L, q):
    msa, w = _problem(port, N, L, q, seed=3)
    lam_h, lam_J = 0.7, 1.3
    rng = np.random.default_rng(0)
    x = rng.normal(0, 0.2, L * q + L * (L - 1) // 2 * q * q).astype(np.float32)
    fj, gj = port.plm_gradient_f64(msa, w, q, lam_h, lam_J, x)
    h, W = A.from_reference_vector(x, L, q)
    f, gh, gW = A.asym_objective(msa, w, q, lam_h, lam_J, h, W)
    assert abs(f - fj) <= 1e-12 * abs(fj)
    # dJ_ij(a,b) of the joint fit = the two directed halves added (plmdca_numerics.cpp:570-602)
    g = A.symmetrise(gh, gW, L, q)
    g[L * q:] *= 2.0
    assert relmax(g, gj) <= 1e-12
    assert np.allclose(A.symmetrise(h, W, L, q), x.astype(np.float64), atol=0, rtol=0)
    # diagonal blocks never move
    assert not gW[~A.offdiag_mask(L, q)].any()


# ---- GPU ------------------------------------------------------------------------------------------
def lib():
    from pydca_b200 import _lib
    return _lib


@pytest.mark.gpu
@pytest.mark.parametrize("N,L,q,sites", [(60, 7, 5, None), (300, 24, 21, None), (131, 13, 21, (4, 9)),
                                         (257, 40, 21, (33, 40)), (500, 9, 5, (0, 3))])
def test_asym_eval_vs_oracle(port, N, L, q, sites):
    msa, w = _problem(port, N, L, q, seed=N + L)
    lam_h, lam_J = 0.2 * (L - 1), 0.3 * (L - 1)
    obj = lib().PlmObjective(msa, w, q, lam_h, lam_J, sites=sites, asymmetric=True)
    try:
        lo, hi = obj.sites
        rng = np.random.default_rng(1)
        h = rng.normal(0, 0.3, (L, q))
        W = np.where(A.offdiag_mask(L, q), rng.normal(0, 0.1, (L * q, L * q)), 0.0)
        x = obj.asym_join(h[lo:hi], W[:, lo * q:hi * q])
        fx, g = obj.eval(x)
        f, gh, gW = A.asym_objective(msa, w, q, lam_h, lam_J, h, W)
        g_h, g_W = obj.asym_split(g)
        assert relmax(g_W, gW[:, lo * q:hi * q]) <= 1e-5
        assert relmax(g_h, gh[lo:hi]) <= 1e-5
        ncol, ld = obj.asym_shape()
        assert not g[ncol:ld].any() and not g[ld:].reshape(L * q, ld)[:, ncol:].any()      # pad columns stay 0
        if sites is None:
            assert abs(fx - f) <= 1e-5 * abs(f)
        else:       # objective of my sites only: the others' columns contribute nothing
            Wm = np.zeros_like(W); Wm[:, lo * q:hi * q] = W[:, lo * q:hi * q]
            hm = np.zeros_like(h); hm[lo:hi] = h[lo:hi]
            # data terms of the other sites at zero parameters are not part of this handle: subtract them
            f_all = A.asym_objective(msa, w, q, lam_h, lam_J, hm, Wm)[0]
            keep = np.ones(L, dtype=bool); keep[lo:hi] = False
            f_others = _data_terms(msa, w, q, hm, Wm)[keep].sum()
            assert abs(fx - (f_all - f_others)) <= 1e-5 * abs(f_all)
    finally:
        obj.close()


def _data_terms(msa, w, q, h, W):
    """Per-site data terms -sum_n w_n log p_n(i, s_ni) of the oracle (with the carry)."""
    N, L = msa.shape
    rows = np.arange(L)[None, :] * q + msa
    out = np.zeros(L)
    p_prev = np.zeros((L, q))
    for n in range(N):
        logits = h + W[rows[n]].sum(axis=0).reshape(L, q) + p_prev
        logits -= logits.max(axis=1, keepdims=True)
        p = np.exp(logits); p /= p.sum(axis=1, keepdims=True)
        p_prev = p
        out -= w[n] * np.log(p[np.arange(L), msa[n]])
    return out


@pytest.mark.gpu
def test_asym_pack_is_the_symmetrisation(port):
    L, q = 9, 21
    rng = np.random.default_rng(2)
    W = rng.normal(size=(L * q, L * q)).astype(np.float32)
    h = rng.normal(size=L * q).astype(np.float32)
    got = lib().asym_pack(W, h, L, q)
    want = A.symmetrise(h.astype(np.float64), W.astype(np.float64), L, q)
    assert relmax(got, want) <= 1e-7


@pytest.mark.gpu
def test_asym_fit_reaches_the_oracle_optimum_and_ignores_the_split(port):
    from scipy.optimize import minimize
    N, L, q = 200, 8, 5
    msa, w = _problem(port, N, L, q, seed=11)
    lam_h, lam_J = 0.5, 2.0
    mask = A.offdiag_mask(L, q)

    def fun(z):       # textbook objective (no carry): a true gradient for the float64 optimiser
        h, W = z[:L * q].reshape(L, q), z[L * q:].reshape(L * q, L * q)
        f, gh, gW = A.asym_objective(msa, w, q, lam_h, lam_J, h, W, carry=False)
        return f, np.concatenate([gh.ravel(), np.where(mask, gW, 0.0).ravel()])

    opt = minimize(fun, np.zeros(L * q + (L * q) ** 2), jac=True, method="L-BFGS-B",
                   options={"maxiter": 2000, "ftol": 1e-15, "gtol": 1e-9})
    x_ref = A.symmetrise(opt.x[:L * q], opt.x[L * q:], L, q)

    # one handle over all sites, no carry -> same objective
    obj = lib().PlmObjective(msa, w, q, lam_h, lam_J, asymmetric=True)
    obj.set_carry(False)
    x, res = obj.fit(obj.init(), 500)
    assert res.evaluations > 5
    Wf = np.zeros((L * q, L * q), dtype=np.float32); hf = np.zeros(L * q, dtype=np.float32)
    obj.asym_export(Wf, hf, x)
    obj.close()
    x_one = lib().asym_pack(Wf, hf, L, q)
    scale = np.abs(x_ref[L * q:]).max()
    assert np.abs(x_one[L * q:] - x_ref[L * q:]).max() <= 2e-2 * scale

    # two handles with disjoint site ranges, fitted separately, assembled: the same answer
    Wf2 = np.zeros_like(Wf); hf2 = np.zeros_like(hf)
    for sites in ((0, 3), (3, L)):
        o = lib().PlmObjective(msa, w, q, lam_h, lam_J, sites=sites, asymmetric=True)
        o.set_carry(False)
        xs, _ = o.fit(o.init(), 500)
        o.asym_export(Wf2, hf2, xs)
        o.close()
    x_two = lib().asym_pack(Wf2, hf2, L, q)
    assert np.abs(x_two[L * q:] - x_ref[L * q:]).max() <= 2e-2 * scale
    fn_one, _ = lib().plm_fn_apc(x_one, L, q)
    fn_two, _ = lib().plm_fn_apc(x_two, L, q)
    fn_ref, _ = lib().plm_fn_apc(x_ref.astype(np.float32), L, q)
    assert relmax(fn_one, fn_ref) <= 2e-2 and relmax(fn_two, fn_ref) <= 2e-2


@pytest.mark.gpu
def test_asym_driver_and_class(port, tmp_path):
    """dca_plm_fit_asymmetric (all GPUs of the box) and PlmDCA(estimator='asymmetric') through plmdcaBackend;
    the ranking stays close to the joint (reference) estimator's."""
    from pydca_b200 import PlmDCA
    from pydca_b200.synthetic import write_fasta
    N, L, q = 400, 30, 21
    msa, w = _problem(port, N, L, q, seed=5)
    lam = 0.2 * (L - 1)
    x1, r1 = lib().fit_asymmetric(msa, w, q, lam, lam, 200, gpus=1)
    xa, ra = lib().fit_asymmetric(msa, w, q, lam, lam, 200, gpus=0)
    assert r1.evaluations > 5 and np.isfinite(x1).all()
    fn1, _ = lib().plm_fn_apc(x1, L, q)
    fna, _ = lib().plm_fn_apc(xa, L, q)
    assert relmax(fna, fn1) <= 2e-2
    path = write_fasta(msa, str(tmp_path / "a.fa"), q)
    joint = PlmDCA(path, "protein", seqid=0.8, max_iterations=200).compute_sorted_FN_APC()
    inst = PlmDCA(path, "protein", seqid=0.8, max_iterations=200, estimator="asymmetric")
    asym = inst.compute_sorted_FN_APC()
    assert inst.last_fit.evaluations > 5
    top = L
    assert len(set(p for p, _ in joint[:top]) & set(p for p, _ in asym[:top])) >= int(0.7 * top)
    with pytest.raises(Exception):
        PlmDCA(path, "protein", estimator="per-site")
    lib().set_estimator(None)


@pytest.mark.gpu
def test_asym_gradient_at_bench_size_vs_the_reference_port(port):
    """C2 size (N=5000, L=150, q=21): at a symmetric point the data gradient of the conditionals of a site
    range R, scattered into the reference layout -- J_ij(a,b) gets [i in R] D[(j,b),(i,a)] + [j in R] D[(i,a),(j,b)]
    -- is what the reference's own gradient restricted to those sites computes (plmdca_numerics.cpp:488-602;
    the C port is bit-exact against it and linear in sites)."""
    from pydca_b200.synthetic import synthetic_msa
    N, L, q = 5000, 150, 21
    msa = synthetic_msa(N, L, q)
    w = port.plm_weights(port.reweight_counts(msa, 0.8))
    lam = 0.2 * (L - 1)
    rng = np.random.default_rng(8)
    nP = L * q + L * (L - 1) // 2 * q * q
    x = (port.plm_init(msa, w, q) + rng.normal(0, 0.05, nP)).astype(np.float32)
    h, W = A.from_reference_vector(x, L, q)
    _, g_reg = port.plm_gradient(msa, w, q, lam, lam, x, sites=(0, 0))
    for lo, hi in ((0, 19), (70, 75), (131, 150)):
        obj = lib().PlmObjective(msa, w, q, lam, lam, sites=(lo, hi), asymmetric=True)
        fx, g = obj.eval(obj.asym_join(h[lo:hi], W[:, lo * q:hi * q]))
        g_h, g_W = obj.asym_split(g)
        obj.close()
        D = np.zeros((L * q, L * q))
        D[:, lo * q:hi * q] = g_W.astype(np.float64) - lam * W[:, lo * q:hi * q]      # data part of my columns
        got = A.symmetrise(np.zeros(L * q), D, L, q)[L * q:] * 2.0                     # the two halves added
        _, g_port = port.plm_gradient(msa, w, q, lam, lam, x, sites=(lo, hi))
        want = (g_port.astype(np.float64) - g_reg)[L * q:]
        assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
        want_h = (g_port.astype(np.float64) - g_reg)[lo * q:hi * q]
        got_h = g_h.astype(np.float64).ravel() - 2.0 * lam * h[lo:hi].ravel()
        assert np.abs(got_h - want_h).max() <= 1e-5 * max(np.abs(want_h).max(), 1e-30)
