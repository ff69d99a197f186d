"""GPU tests (-m gpu) of the owner-sharded multi-GPU plmDCA path (csrc/xrank.cuh): the
ranks run as THREADS of this process on one GPU and exchange through raw arena pointers
(dca_plm_connect_ptrs), which exercises exactly the kernels, flags and slot buffers the
one-process-per-GPU runs use over CUDA IPC -- P2P stores simply land in the same device.

Parity bar: the gathered gradient equals the single-GPU gradient to float32 rounding of
the n-split partial sums (<= 2e-6 of max|g|), fx to 1e-9 relative (float64 sums in a
different order), and both match the oracle within the 1e-5 bar of test_gpu_parity.py."""
import threading

import numpy as np
import pytest

from conftest import relmax

pytestmark = pytest.mark.gpu


def _run_ranks(world, fn):
    """fn(rank) on `world` threads; re-raises the first failure."""
    errs, outs = [None] * world, [None] * world

    def body(r):
        try:
            outs[r] = fn(r)
        except BaseException as exc:          # noqa: BLE001 -- reported below
            errs[r] = exc

    ts = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=120)
    for e in errs:
        if e is not None:
            raise e
    assert all(not t.is_alive() for t in ts), "a rank thread did not finish"
    return outs


def _make_ranks(msa, w, q, lam, bounds):
    from pydca_b200 import _lib
    world = len(bounds) - 1
    objs = [_lib.PlmObjective(msa, w, q, lam, lam, shard=(r, world, bounds)) for r in range(world)]
    arenas = [o.arena() for o in objs]
    for o in objs:
        o.connect_ptrs(arenas)
    return objs


def _close(objs):
    for o in objs:
        o.disconnect()
    for o in objs:
        o.close()


@pytest.mark.parametrize("N,L,q,bounds", [
    (257, 17, 21, [0, 9, 17]),
    (640, 37, 21, [0, 10, 24, 37]),
    (129, 30, 5, [0, 1, 29, 30]),            # single-site shards at both ends (last one owns no block)
    (300, 12, 21, [0, 3, 6, 9, 12]),
    (500, 16, 21, [0, 2, 4, 6, 8, 10, 12, 14, 16]),
])
def test_sharded_eval_matches_single_gpu_and_oracle(port, N, L, q, bounds):
    from pydca_b200 import _lib
    from pydca_b200.distributed import scatter_full_vector
    from pydca_b200.synthetic import synthetic_msa
    msa = synthetic_msa(N, L, q, seed=5 * N + L)
    w = port.plm_weights(port.reweight_counts(msa, 0.8))
    lam = 0.2 * (L - 1)
    ref = _lib.PlmObjective(msa, w, q, lam, lam)
    rng = np.random.default_rng(L)
    x = (ref.init() + rng.normal(0, 0.2, ref.P)).astype(np.float32)
    fx1, g1 = ref.eval(x)
    ref.close()

    objs = _make_ranks(msa, w, q, lam, bounds)
    world = len(objs)
    layouts = [o.local_layout() for o in objs]
    assert sum(o.Ploc for o in objs) == objs[0].P
    x0 = scatter_full_vector([o.init() for o in objs], layouts, objs[0].P)
    assert np.abs(x0 - port.plm_init(msa, w, q)).max() <= 1e-6

    for rep in range(3):                      # repeated collective evaluations: epochs, slot parity
        outs = _run_ranks(world, lambda r: objs[r].eval(objs[r].local_slice(x)))
        fxs = [o[0] for o in outs]
        assert all(f == fxs[0] for f in fxs), "ranks disagree on fx bits"
        g = scatter_full_vector([o[1] for o in outs], layouts, objs[0].P)
        # (the single-GPU objective may run the 128-column logits kernel with its equal-residue fast path where a
        # shard runs the 64-column one: float32 sums in a different association, not bit-identical any more)
        assert abs(fxs[0] - fx1) <= 2e-7 * abs(fx1)
        assert relmax(g, g1) <= 2e-6
    f64, g64 = port.plm_gradient_f64(msa, w, q, lam, lam, x)
    assert abs(fxs[0] - f64) <= 1e-5 * abs(f64) and relmax(g, g64) <= 1e-5
    _close(objs)


def test_sharded_fit_follows_single_gpu_fit(port):
    """Full L-BFGS on 3 ranks vs 1 GPU: same objective value path within the float32 noise
    band of the line search, every rank takes the same decisions (same evaluation count)."""
    from pydca_b200 import _lib
    from pydca_b200.distributed import balanced_site_bounds, scatter_full_vector
    from pydca_b200.synthetic import synthetic_msa
    N, L, q = 400, 25, 21
    msa = synthetic_msa(N, L, q, seed=77)
    w = port.plm_weights(port.reweight_counts(msa, 0.8))
    lam = 0.2 * (L - 1)
    ref = _lib.PlmObjective(msa, w, q, lam, lam)
    x1, r1 = ref.fit(ref.init(), 30)
    ref.close()

    bounds = balanced_site_bounds(L, 3)
    objs = _make_ranks(msa, w, q, lam, bounds)
    outs = _run_ranks(3, lambda r: objs[r].fit(objs[r].init(), 30))
    res = [o[1] for o in outs]
    assert len({(r.status, r.iterations, r.evaluations) for r in res}) == 1
    assert all(r.fx == res[0].fx for r in res)
    x = scatter_full_vector([o[0] for o in outs], [o.local_layout() for o in objs], objs[0].P)
    assert abs(res[0].fx - r1.fx) <= 2e-3 * abs(r1.fx)
    assert res[0].evaluations >= 10
    fn1, _ = _lib.plm_fn_apc(x1, L, q)
    fn, _ = _lib.plm_fn_apc(x, L, q)
    assert relmax(fn, fn1) <= 3e-2
    # the sharded optimum is a stationary point of the full objective too
    full = _lib.PlmObjective(msa, w, q, lam, lam)
    fx_full, g_full = full.eval(x)
    full.close()
    assert abs(fx_full - res[0].fx) <= 1e-5 * abs(fx_full)
    _close(objs)


def test_sharded_eval_wide_owner_tiles(port, monkeypatch):
    """The owner kernels come in two tile widths (8 or 2 sites j per CTA, picked by grid size);
    small test problems take the narrow one, so force the wide one here."""
    from pydca_b200 import _lib
    from pydca_b200.distributed import scatter_full_vector
    from pydca_b200.synthetic import synthetic_msa
    monkeypatch.setenv("DCA_B200_OWN_JG", "8")
    N, L, q, bounds = 300, 21, 21, [0, 5, 13, 21]
    msa = synthetic_msa(N, L, q, seed=9)
    w = port.plm_weights(port.reweight_counts(msa, 0.8))
    lam = 0.2 * (L - 1)
    objs = _make_ranks(msa, w, q, lam, bounds)
    monkeypatch.delenv("DCA_B200_OWN_JG")
    ref = _lib.PlmObjective(msa, w, q, lam, lam)
    x = (ref.init() + np.random.default_rng(3).normal(0, 0.2, ref.P)).astype(np.float32)
    fx1, g1 = ref.eval(x)
    ref.close()
    outs = _run_ranks(3, lambda r: objs[r].eval(objs[r].local_slice(x)))
    g = scatter_full_vector([o[1] for o in outs], [o.local_layout() for o in objs], objs[0].P)
    assert abs(outs[0][0] - fx1) <= 1e-9 * abs(fx1) and relmax(g, g1) <= 2e-6
    _close(objs)


def test_sharded_world1_is_the_plain_objective(port):
    from pydca_b200 import _lib
    from pydca_b200.synthetic import synthetic_msa
    msa = synthetic_msa(200, 11, 5, seed=2)
    w = port.plm_weights(port.reweight_counts(msa, 0.8))
    ref = _lib.PlmObjective(msa, w, 5, 2.0, 2.0)
    one = _lib.PlmObjective(msa, w, 5, 2.0, 2.0, shard=(0, 1, [0, 11]))
    assert one.Ploc == ref.P and one.local_layout() == (0, 55, 55, ref.P - 55)
    x = (ref.init() + 0.1).astype(np.float32)
    f1, g1 = ref.eval(x)
    f2, g2 = one.eval(x)
    assert f1 == f2 and (g1 == g2).all()
    ref.close()
    one.disconnect()
    one.close()


def test_unconnected_shard_fails_loudly(port):
    from pydca_b200 import _lib
    from pydca_b200.synthetic import synthetic_msa
    msa = synthetic_msa(50, 6, 5, seed=1)
    w = port.plm_weights(port.reweight_counts(msa, 0.8))
    o = _lib.PlmObjective(msa, w, 5, 1.0, 1.0, shard=(0, 2, [0, 3, 6]))
    with pytest.raises(_lib.DcaBackendError, match="peers"):
        o.eval(o.init())
    o.close()
    with pytest.raises(_lib.DcaBackendError):
        _lib.PlmObjective(msa, w, 5, 1.0, 1.0, shard=(0, 2, [0, 6, 6]))      # a rank without sites


def test_plmdca_backend_drives_several_gpus(tmp_path):
    """The reference's own entry point with DCA_B200_GPUS > 1 (plmdcaBackend.cpp:151: one synchronous
    call; here one host thread per GPU inside it, peer-memory exchange over NVLink, every GPU
    copying its slice into the returned vector).  Needs >= 2 devices: skipped on a 1-GPU box."""
    from pydca_b200 import _lib
    from pydca_b200.synthetic import synthetic_msa, write_fasta
    ngpu = _lib.lib().dca_device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    N, L, q = 1500, 41, 21
    lam = 0.2 * (L - 1)
    msa = synthetic_msa(N, L, q, seed=123)
    path = write_fasta(msa, str(tmp_path / "multi.fa"), q)
    try:
        _lib.set_backend_gpus(1)
        x1 = _lib.plmdca_backend(path, "protein", L, q, 0.8, lam, lam, 40, copy=True)
        r1 = _lib.last_fit()
        for g in sorted({2, min(ngpu, 4), ngpu}):
            _lib.set_backend_gpus(g)
            xg = _lib.plmdca_backend(path, "protein", L, q, 0.8, lam, lam, 40, copy=True)
            rg = _lib.last_fit()
            assert (rg.status, rg.iterations, rg.evaluations) == (r1.status, r1.iterations, r1.evaluations), g
            fn1, _ = _lib.plm_fn_apc(x1, L, q)
            fng, _ = _lib.plm_fn_apc(xg, L, q)
            assert relmax(fng, fn1) <= 5e-6, g
            assert relmax(xg, x1) <= 5e-5, g
    finally:
        _lib.set_backend_gpus(0)


def test_mean_field_pair_histograms_on_several_gpus():
    """the pair-frequency histograms split by site ranges over the GPUs, every GPU
    storing its columns straight into the solver GPU's matrix.  Each (site, tile) is computed by
    the same code in the same order wherever it runs, so the results are bit-identical."""
    from pydca_b200 import _lib, msa_numerics
    from pydca_b200.synthetic import synthetic_msa
    ngpu = _lib.lib().dca_device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    N, L, q = 6000, 130, 21
    msa = synthetic_msa(N, L, q, seed=9)
    w = 1.0 / _lib.sequence_counts(msa, 0.8, compare_f64=True).astype(np.float64)
    try:
        _lib.set_backend_gpus(1)
        one = msa_numerics.mean_field_scores(msa, q, w, 0.5, want_couplings=True)
        for g in sorted({2, ngpu}):
            _lib.set_backend_gpus(g)
            many = msa_numerics.mean_field_scores(msa, q, w, 0.5, want_couplings=True)
            assert (many["fn"] == one["fn"]).all() and (many["couplings"] == one["couplings"]).all(), g
            assert (many["reg_fi"] == one["reg_fi"]).all()
    finally:
        _lib.set_backend_gpus(0)
