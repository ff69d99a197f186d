"""The block-parallel evaluation of a SEQUENTIAL float32 sum (pydca_b200/csrc/seqsum_core.h,
used by csrc/seqsum.cu for the reference-arithmetic L-BFGS mode) must equal the naive loop of
the reference (arithmetic_ansi.h:114-121, plmdca_numerics.cpp:463-486) bit for bit.  CPU only:
oracle/seqsum_model.cpp drives the product's own arithmetic header the way the kernel does."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE = os.path.join(os.path.dirname(HERE), "oracle")
SO = os.path.join(ORACLE, "_build", "libseqsum_model.so")

_f32p = ctypes.POINTER(ctypes.c_float)
_i64p = ctypes.POINTER(ctypes.c_long)


@pytest.fixture(scope="module")
def model():
    subprocess.run(["make", "-s", "-C", ORACLE, "seqmodel"], check=True)
    L = ctypes.CDLL(SO)
    for name, args in [("seq_naive_dot", [_f32p, _f32p, ctypes.c_long]),
                       ("seq_model_dot", [_f32p, _f32p, ctypes.c_long, _i64p]),
                       ("seq_naive_reg", [_f32p, ctypes.c_long, ctypes.c_long, ctypes.c_float, ctypes.c_float]),
                       ("seq_model_reg", [_f32p, ctypes.c_long, ctypes.c_long, ctypes.c_float, ctypes.c_float, _i64p]),
                       ("seq_naive_sum", [_f32p, ctypes.c_long]),
                       ("seq_model_sum", [_f32p, ctypes.c_long, _i64p])]:
        f = getattr(L, name)
        f.argtypes = args
        f.restype = ctypes.c_float
    return L


def _p(a):
    return a.ctypes.data_as(_f32p)


def _bits(v):
    return np.float32(v).view(np.uint32)


def _check_sum(model, t):
    t = np.ascontiguousarray(t, dtype=np.float32)
    stats = (ctypes.c_long * 3)()
    a = model.seq_naive_sum(_p(t), t.size)
    b = model.seq_model_sum(_p(t), t.size, stats)
    assert _bits(a) == _bits(b) or (np.isnan(a) and np.isnan(b)), (a, b, list(stats))
    return list(stats)


def test_dot_matches_naive_on_gradient_like_vectors(model):
    """Few large entries (fields) followed by millions of small ones (couplings): the shape that
    makes the reference's |g|^2 stagnate.  Most groups must take the one-step path."""
    rng = np.random.default_rng(1)
    n = 1_500_000
    g = (rng.standard_normal(n) * 0.02).astype(np.float32)
    g[:3150] = (rng.standard_normal(3150) * 40).astype(np.float32)
    stats = (ctypes.c_long * 3)()
    a = model.seq_naive_dot(_p(g), _p(g), n)
    b = model.seq_model_dot(_p(g), _p(g), n, stats)
    assert _bits(a) == _bits(b)
    exact = float(np.dot(g.astype(np.float64), g.astype(np.float64)))
    assert abs(a - exact) / exact > 1e-4            # the float32 chain really is biased here
    assert stats[0] > 0.95 * (n // 4096)            # block path, not the serial fallback
    d = (-g * rng.uniform(0.5, 1.5, n)).astype(np.float32)
    assert _bits(model.seq_naive_dot(_p(g), _p(d), n)) == _bits(model.seq_model_dot(_p(g), _p(d), n, stats))


def test_signed_cancelling_dots(model):
    rng = np.random.default_rng(2)
    for n in (1, 5, 127, 128, 129, 4095, 4096, 4097, 70_001, 400_000):
        x = rng.standard_normal(n).astype(np.float32)
        y = rng.standard_normal(n).astype(np.float32)
        stats = (ctypes.c_long * 3)()
        assert _bits(model.seq_naive_dot(_p(x), _p(y), n)) == _bits(model.seq_model_dot(_p(x), _p(y), n, stats)), n


def test_regulariser_sum(model):
    rng = np.random.default_rng(3)
    n, n0 = 900_000, 3150
    x = (rng.standard_normal(n) * 0.05).astype(np.float32)
    x[:n0] = rng.standard_normal(n0).astype(np.float32)
    stats = (ctypes.c_long * 3)()
    for lam0, lam1 in ((29.8, 29.8), (1.0, 20.0), (0.01, 1000.0)):
        a = model.seq_naive_reg(_p(x), n, n0, lam0, lam1)
        b = model.seq_model_reg(_p(x), n, n0, lam0, lam1, stats)
        assert _bits(a) == _bits(b)


def test_exact_ties_zeros_and_binade_crossings(model):
    rng = np.random.default_rng(4)
    # exact ties: running value 2^24 (u = 2), terms of exactly 1.0 = u/2 -> round-half-even decides
    t = np.concatenate([[2.0 ** 24], np.ones(5000), [3.0], np.ones(5000)]).astype(np.float32)
    _check_sum(model, t)
    # odd mantissa + ties, negative running value
    t = np.concatenate([[-(2.0 ** 24 + 2)], -np.ones(3000), np.ones(3000) * 3, -np.ones(3000)]).astype(np.float32)
    _check_sum(model, t)
    # all zeros, leading zeros, signed zeros
    _check_sum(model, np.zeros(10000))
    _check_sum(model, np.concatenate([np.zeros(5000), -np.zeros(10), rng.standard_normal(9000)]))
    # slow upward crossing of many binades with tiny terms
    _check_sum(model, np.full(300_000, 1e-3))
    # repeated crossings of a power of two in both directions
    t = np.concatenate([[1024.0], np.tile(np.concatenate([np.full(300, 0.37), np.full(300, -0.37)]), 40)])
    _check_sum(model, t)
    # sign changes of the running value
    _check_sum(model, np.tile(np.concatenate([np.full(200, 1.3), np.full(400, -1.3), np.full(200, 1.3)]), 30))
    # huge and tiny magnitudes mixed, overflow to inf, nan
    t = rng.standard_normal(50_000) * np.exp(rng.uniform(-60, 60, 50_000))
    _check_sum(model, t)
    _check_sum(model, np.concatenate([np.full(10, 3e38), rng.standard_normal(5000)]))
    _check_sum(model, np.concatenate([rng.standard_normal(5000), [np.nan], rng.standard_normal(5000)]))
    # subnormal running values
    _check_sum(model, np.full(5000, 1e-42))


def test_random_magnitude_profiles(model):
    rng = np.random.default_rng(5)
    for trial in range(30):
        n = int(rng.integers(1, 60_000))
        scale = np.exp(rng.uniform(-8, 8))
        drift = rng.uniform(-1, 1)
        t = (rng.standard_normal(n) + drift) * scale * np.exp(rng.uniform(-3, 3, n) * rng.integers(0, 2))
        if trial % 3 == 0:                       # quantised terms provoke exact ties
            t = np.round(t * 4) / 4
        _check_sum(model, t)


# ---- the producer / consumer pipeline's decisions (seqsum_pipe_kernel), sequentially on the CPU ----
@pytest.fixture(scope="module")
def pipe(model):
    for name, args in [("seq_pipe_model_sum", [_f32p, ctypes.c_long, _i64p]),
                       ("seq_pipe_model_dot", [_f32p, _f32p, ctypes.c_long, _i64p])]:
        f = getattr(model, name)
        f.argtypes = args
        f.restype = ctypes.c_float
    return model


def _check_pipe(pipe, t, y=None):
    t = np.ascontiguousarray(t, dtype=np.float32)
    stats = (ctypes.c_long * 4)()
    if y is None:
        a = pipe.seq_naive_sum(_p(t), t.size)
        b = pipe.seq_pipe_model_sum(_p(t), t.size, stats)
    else:
        y = np.ascontiguousarray(y, dtype=np.float32)
        a = pipe.seq_naive_dot(_p(t), _p(y), t.size)
        b = pipe.seq_pipe_model_dot(_p(t), _p(y), t.size, stats)
    assert _bits(a) == _bits(b) or (np.isnan(a) and np.isnan(b)), (a, b, list(stats))
    return dict(zip(("records", "row_walks", "exact", "plain_scans"), stats))


def test_pipeline_model_gradient_like_and_cancelling(pipe):
    """The pipeline's own sequence of decisions -- chunk records for the three exponents around a stale
    prediction, steps over 32 records with the plain integer scan where it applies, the row walk, the
    consumer CTA's exact path with per-row predictions -- equals the naive loop, and every path is taken."""
    rng = np.random.default_rng(11)
    n = 700_001
    g = (rng.standard_normal(n) * 0.02).astype(np.float32)
    g[:3150] = (rng.standard_normal(3150) * 40).astype(np.float32)
    st = _check_pipe(pipe, g, g)
    assert st["records"] > 0.9 * (n // 4096) and st["exact"] >= 1 and st["plain_scans"] > 0
    d = (-g * rng.uniform(0.5, 1.5, n)).astype(np.float32)
    _check_pipe(pipe, g, d)
    s = (rng.standard_normal(n) * 1e-3).astype(np.float32)
    _check_pipe(pipe, s, d)                              # an inner product like those of the two-loop recursion
    st = _check_pipe(pipe, rng.standard_normal(n), rng.standard_normal(n))     # random signs: the exponent wanders
    assert st["row_walks"] > 10 and st["records"] > 10


def test_pipeline_model_adversarial(pipe):
    rng = np.random.default_rng(12)
    n = 300_017
    hover = np.tile(np.concatenate([np.full(300, 0.37), np.full(300, -0.37)]), n // 600)
    st = _check_pipe(pipe, np.concatenate([[1024.0], hover]))
    assert st["row_walks"] + st["exact"] > 1             # the running value hovers at 2^10
    _check_pipe(pipe, np.concatenate([[2048.0 - 30.0], hover + 1e-4, -hover - 3e-3]))
    z = np.concatenate([np.zeros(3 * 4096 + 17), rng.standard_normal(100_000), [0.0]])
    _check_pipe(pipe, np.concatenate([z, -z[::-1], np.zeros(5000), rng.standard_normal(90_000)]))
    _check_pipe(pipe, np.concatenate([np.ones(2 ** 17), -np.ones(2 ** 17), np.full(100_000, 2.0 ** -30)]))
    _check_pipe(pipe, rng.standard_normal(n) * np.exp(rng.uniform(-40, 40, n)))
    wide = rng.standard_normal(n) * 0.01
    wide[rng.integers(0, n, 100)] *= 1e9
    _check_pipe(pipe, wide)
    _check_pipe(pipe, np.round(rng.standard_normal(n) * 64) / 4)             # ties at every step: no plain scan
    _check_pipe(pipe, np.concatenate([[2.0 ** 24], np.tile([1.0, 1.0, 3.0, -1.0], n // 4)]))
    stag = np.full(n, 1e-4)
    stag[0] = 4096.0
    stag[rng.integers(1, n, 300)] = 0.5
    _check_pipe(pipe, stag)
    for bad in (np.inf, -np.inf, np.nan):
        v = rng.standard_normal(n) * 0.1
        v[n // 2 + 77] = bad
        _check_pipe(pipe, v)
    for m in (1, 4095, 4096, 4097, 70_001, 262_144, 262_145):              # ragged lengths
        _check_pipe(pipe, rng.standard_normal(m))
