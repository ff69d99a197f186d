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
