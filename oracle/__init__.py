"""CPU oracle for the DCA hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker.
Nothing under ``pydca_b200/`` imports it (tests/test_boundary.py enforces that).

Two checkers live here:

* ``port``  -- ``dca_oracle.c`` (our C restatement of plmdca_numerics.cpp) and
  ``mf_oracle.py`` (numpy restatement of meanfield_dca/msa_numerics.py and of
  the FN/APC loops of plmdca.py / meanfield_dca.py);
* ``ref``   -- the UNMODIFIED reference C++ compiled from /root/reference into
  ``oracle/_ref/libpydca_ref.so`` (see Makefile), reached through ``ref_shim.cpp``.

Parity status: pinned (tests/test_oracle_golden.py, tests/golden/).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "_build", "libdca_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libpydca_ref.so")

_u8p = ctypes.POINTER(ctypes.c_uint8)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)


def build(ref=True):
    """Compile the checkers (building the checker is not using it)."""
    targets = ["port", "seqmodel"]
    if ref and os.path.isdir("/root/reference/pydca/plmdca"):
        targets.append("ref")
    subprocess.run(["make", "-s", "-C", _HERE] + targets, check=True)


def _ptr(a, t):
    return a.ctypes.data_as(t)


class _Port:
    def __init__(self):
        if not os.path.exists(PORT_SO):
            build(ref=False)
        L = ctypes.CDLL(PORT_SO)
        L.orc_read_fasta.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(_u8p)]
        L.orc_read_fasta.restype = ctypes.c_long
        L.orc_free.argtypes = [ctypes.c_void_p]
        L.orc_reweight_counts.argtypes = [_u8p, ctypes.c_long, ctypes.c_int, ctypes.c_double, ctypes.c_int, _u32p]
        L.orc_reweight_counts_rows.argtypes = [_u8p, ctypes.c_long, ctypes.c_int, ctypes.c_double, ctypes.c_int,
                                               ctypes.POINTER(ctypes.c_long), ctypes.c_long, _u32p]
        L.orc_plm_weights.argtypes = [_u32p, ctypes.c_long, _f32p]
        L.orc_plm_meff.argtypes = [_f32p, ctypes.c_long]
        L.orc_plm_meff.restype = ctypes.c_float
        L.orc_plm_init.argtypes = [_u8p, _f32p, ctypes.c_long, ctypes.c_int, ctypes.c_int, _f32p]
        L.orc_plm_gradient_range.argtypes = [_u8p, _f32p, ctypes.c_long, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_float, ctypes.c_float, _f32p, _f32p, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_int]
        L.orc_plm_gradient_range.restype = ctypes.c_float
        L.orc_plm_gradient_f64.argtypes = [_u8p, _f32p, ctypes.c_long, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_double, ctypes.c_double, _f32p, _f64p, ctypes.c_int]
        L.orc_plm_gradient_f64.restype = ctypes.c_double
        L.orc_num_threads.restype = ctypes.c_int
        L.orc_set_num_threads.argtypes = [ctypes.c_int]
        self.L = L

    # -- FASTA -----------------------------------------------------------------
    def read_fasta(self, path, biomolecule, seqs_len):
        """Integer MSA (uint8, states 0..q-1, gap=q-1), unique rows, file order."""
        bio = 1 if biomolecule.strip().upper() == "PROTEIN" else 2
        out = _u8p()
        n = self.L.orc_read_fasta(os.fsencode(path), bio, int(seqs_len), ctypes.byref(out))
        if n < 0:
            raise IOError("orc_read_fasta(%s) failed with %d" % (path, n))
        arr = np.ctypeslib.as_array(out, shape=(n, seqs_len)).copy()
        self.L.orc_free(out)
        return arr

    # -- reweighting -------------------------------------------------------------
    def reweight_counts(self, msa, seqid, f64=False):
        msa = np.ascontiguousarray(msa, dtype=np.uint8)
        n, l = msa.shape
        c = np.zeros(n, dtype=np.uint32)
        self.L.orc_reweight_counts(_ptr(msa, _u8p), n, l, float(seqid), int(bool(f64)), _ptr(c, _u32p))
        return c

    def reweight_counts_rows(self, msa, seqid, rows, f64=False):
        """c_i for the given rows only (each against all N rows): spot checks at full size."""
        msa = np.ascontiguousarray(msa, dtype=np.uint8)
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        n, l = msa.shape
        c = np.zeros(rows.size, dtype=np.uint32)
        self.L.orc_reweight_counts_rows(_ptr(msa, _u8p), n, l, float(seqid), int(bool(f64)),
                                        rows.ctypes.data_as(ctypes.POINTER(ctypes.c_long)), rows.size, _ptr(c, _u32p))
        return c

    def plm_weights(self, counts):
        counts = np.ascontiguousarray(counts, dtype=np.uint32)
        w = np.zeros(counts.size, dtype=np.float32)
        self.L.orc_plm_weights(_ptr(counts, _u32p), counts.size, _ptr(w, _f32p))
        return w

    def plm_meff(self, w):
        w = np.ascontiguousarray(w, dtype=np.float32)
        return float(self.L.orc_plm_meff(_ptr(w, _f32p), w.size))

    # -- plmDCA -------------------------------------------------------------------
    def plm_init(self, msa, w, q):
        msa = np.ascontiguousarray(msa, dtype=np.uint8)
        w = np.ascontiguousarray(w, dtype=np.float32)
        n, l = msa.shape
        x = np.zeros(l * q + l * (l - 1) // 2 * q * q, dtype=np.float32)
        self.L.orc_plm_init(_ptr(msa, _u8p), _ptr(w, _f32p), n, l, q, _ptr(x, _f32p))
        return x

    def plm_gradient(self, msa, w, q, lambda_h, lambda_J, x, carry=True, sites=None):
        msa = np.ascontiguousarray(msa, dtype=np.uint8)
        w = np.ascontiguousarray(w, dtype=np.float32)
        x = np.ascontiguousarray(x, dtype=np.float32)
        n, l = msa.shape
        lo, hi = (0, l) if sites is None else sites
        g = np.zeros_like(x)
        fx = self.L.orc_plm_gradient_range(_ptr(msa, _u8p), _ptr(w, _f32p), n, l, q,
                                           float(lambda_h), float(lambda_J), _ptr(x, _f32p),
                                           _ptr(g, _f32p), int(bool(carry)), int(lo), int(hi))
        return float(fx), g

    def plm_gradient_f64(self, msa, w, q, lambda_h, lambda_J, x, carry=True):
        msa = np.ascontiguousarray(msa, dtype=np.uint8)
        w = np.ascontiguousarray(w, dtype=np.float32)
        x = np.ascontiguousarray(x, dtype=np.float32)
        n, l = msa.shape
        g = np.zeros(x.size, dtype=np.float64)
        fx = self.L.orc_plm_gradient_f64(_ptr(msa, _u8p), _ptr(w, _f32p), n, l, q,
                                         float(lambda_h), float(lambda_J), _ptr(x, _f32p),
                                         _ptr(g, _f64p), int(bool(carry)))
        return float(fx), g

    def num_threads(self):
        return int(self.L.orc_num_threads())

    def set_num_threads(self, n):
        self.L.orc_set_num_threads(int(n))


class _RefPlm:
    """The reference's own PlmDCA object (plmdca.h:16) behind ref_shim.cpp."""

    def __init__(self, lib, msa_file, biomolecule, seqs_len, q, seqid, lambda_h, lambda_J, num_threads):
        self._lib = lib
        bio = 1 if biomolecule.strip().upper() == "PROTEIN" else 2
        self._path = os.fsencode(msa_file)  # the C++ object keeps the pointer
        self._h = lib.ref_plm_create(self._path, bio, seqs_len, q, seqid, lambda_h, lambda_J, num_threads)
        if not self._h:
            raise IOError("reference PlmDCA constructor failed for %s" % msa_file)
        self.L, self.q = seqs_len, q
        self.N = int(lib.ref_plm_num_seqs(self._h))
        self.P = seqs_len * q + seqs_len * (seqs_len - 1) // 2 * q * q

    def close(self):
        if self._h:
            self._lib.ref_plm_destroy(self._h)
            self._h = None

    __del__ = close

    def seqs(self):
        out = np.zeros((self.N, self.L), dtype=np.uint8)
        self._lib.ref_plm_get_seqs(self._h, _ptr(out, _u8p))
        return out

    def weights(self):
        out = np.zeros(self.N, dtype=np.float32)
        self._lib.ref_plm_get_weights(self._h, _ptr(out, _f32p))
        return out

    def meff(self):
        return float(self._lib.ref_plm_meff(self._h))

    def set_threads(self, n):
        self._lib.ref_plm_set_threads(self._h, int(n))

    def init(self):
        x = np.zeros(self.P, dtype=np.float32)
        self._lib.ref_plm_init(self._h, _ptr(x, _f32p))
        return x

    def gradient(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        g = np.zeros_like(x)
        fx = self._lib.ref_plm_gradient(self._h, _ptr(x, _f32p), _ptr(g, _f32p))
        return float(fx), g

    def reweight(self):
        out = np.zeros(self.N, dtype=np.float32)
        self._lib.ref_plm_reweight(self._h, _ptr(out, _f32p))
        return out

    def fit(self, max_iterations):
        """Reference lbfgs() on the reference gradient (plmdcaBackend.cpp:61-81
        parameters); returns (x, status, evaluations, fx)."""
        x = np.zeros(self.P, dtype=np.float32)
        evals = ctypes.c_int(0)
        fx = ctypes.c_float(0)
        st = self._lib.ref_plm_fit(self._h, _ptr(x, _f32p), int(max_iterations), ctypes.byref(evals),
                                   ctypes.byref(fx))
        return x, int(st), int(evals.value), float(fx.value)

    def fit_trace(self, max_iterations, max_trace=512):
        """fit() with the reference's progress callback recorded: returns (x, status,
        evaluations, fx, trace[iterations, 6] = fx, xnorm, gnorm, step, ls, evaluations so far)."""
        x = np.zeros(self.P, dtype=np.float32)
        evals = ctypes.c_int(0)
        iters = ctypes.c_int(0)
        fx = ctypes.c_float(0)
        trace = np.zeros((max_trace, 6), dtype=np.float32)
        st = self._lib.ref_plm_fit_trace(self._h, _ptr(x, _f32p), int(max_iterations), ctypes.byref(evals),
                                         ctypes.byref(fx), _ptr(trace, _f32p), int(max_trace), ctypes.byref(iters))
        return x, int(st), int(evals.value), float(fx.value), trace[:min(max_trace, int(iters.value))].copy()


class _Ref:
    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(
                REF_SO + " is missing: run `make -C oracle ref` where /root/reference exists")
        L = ctypes.CDLL(REF_SO)
        L.ref_plm_create.argtypes = [ctypes.c_char_p, ctypes.c_uint, ctypes.c_uint, ctypes.c_uint,
                                     ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_uint]
        L.ref_plm_create.restype = ctypes.c_void_p
        for name in ("ref_plm_destroy", "ref_plm_get_seqs", "ref_plm_get_weights", "ref_plm_init",
                     "ref_plm_set_threads", "ref_plm_reweight"):
            getattr(L, name).restype = None
        L.ref_plm_destroy.argtypes = [ctypes.c_void_p]
        L.ref_plm_num_seqs.argtypes = [ctypes.c_void_p]
        L.ref_plm_num_seqs.restype = ctypes.c_uint
        L.ref_plm_get_seqs.argtypes = [ctypes.c_void_p, _u8p]
        L.ref_plm_get_weights.argtypes = [ctypes.c_void_p, _f32p]
        L.ref_plm_meff.argtypes = [ctypes.c_void_p]
        L.ref_plm_meff.restype = ctypes.c_float
        L.ref_plm_set_threads.argtypes = [ctypes.c_void_p, ctypes.c_uint]
        L.ref_plm_init.argtypes = [ctypes.c_void_p, _f32p]
        L.ref_plm_gradient.argtypes = [ctypes.c_void_p, _f32p, _f32p]
        L.ref_plm_gradient.restype = ctypes.c_float
        L.ref_plm_reweight.argtypes = [ctypes.c_void_p, _f32p]
        L.ref_plm_fit.argtypes = [ctypes.c_void_p, _f32p, ctypes.c_uint, ctypes.POINTER(ctypes.c_int),
                                  ctypes.POINTER(ctypes.c_float)]
        L.ref_plm_fit.restype = ctypes.c_int
        L.ref_plm_fit_trace.argtypes = [ctypes.c_void_p, _f32p, ctypes.c_uint, ctypes.POINTER(ctypes.c_int),
                                        ctypes.POINTER(ctypes.c_float), _f32p, ctypes.c_int,
                                        ctypes.POINTER(ctypes.c_int)]
        L.ref_plm_fit_trace.restype = ctypes.c_int
        # the reference's public C ABI (plmdcaBackend.cpp:151, :204)
        L.plmdcaBackend.argtypes = (ctypes.c_ushort, ctypes.c_ushort, ctypes.c_char_p, ctypes.c_uint,
                                    ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_uint,
                                    ctypes.c_uint, ctypes.c_bool)
        L.plmdcaBackend.restype = ctypes.c_void_p
        L.freeFieldsAndCouplings.argtypes = [ctypes.c_void_p]
        L.freeFieldsAndCouplings.restype = None
        self.L = L

    def plm(self, msa_file, biomolecule, seqs_len, q, seqid, lambda_h, lambda_J, num_threads=1):
        return _RefPlm(self.L, msa_file, biomolecule, seqs_len, q, seqid, lambda_h, lambda_J, num_threads)

    def plmdca_backend(self, msa_file, biomolecule, seqs_len, q, seqid, lambda_h, lambda_J,
                       max_iterations, num_threads=1, verbose=False):
        """Reference end-to-end fit through its own C ABI; returns x (float32[P])."""
        bio = 1 if biomolecule.strip().upper() == "PROTEIN" else 2
        P = seqs_len * q + seqs_len * (seqs_len - 1) // 2 * q * q
        ptr = self.L.plmdcaBackend(bio, q, os.fsencode(msa_file), seqs_len, seqid, lambda_h, lambda_J,
                                   max_iterations, num_threads, verbose)
        if not ptr:
            raise RuntimeError("reference plmdcaBackend returned NULL")
        x = np.ctypeslib.as_array(ctypes.cast(ptr, _f32p), shape=(P,)).copy()
        self.L.freeFieldsAndCouplings(ptr)
        return x


_port = None
_ref = None


def port():
    global _port
    if _port is None:
        _port = _Port()
    return _port


def ref():
    global _ref
    if _ref is None:
        _ref = _Ref()
    return _ref


def have_ref():
    return os.path.exists(REF_SO)
