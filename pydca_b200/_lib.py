"""ctypes binding of pydca_b200/_plmdcaBackend_b200.so (include/dca_b200.h).

The library is looked up exactly like the reference looks up its own backend
(``glob('<pkgdir>/_plmdcaBackend*')``, pydca/plmdca/plmdca.py:33-38).  There is
no CPU fallback: if the shared object is missing, or no CUDA device is visible,
every compute call raises ``DcaBackendError``.
"""
import ctypes
import glob
import os
import weakref

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))


class DcaBackendError(RuntimeError):
    """Raised when the CUDA backend is missing or a backend call fails."""


_u8p = ctypes.POINTER(ctypes.c_uint8)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)

#: int (*)(void *ctx, float *g_dev, long n, double *scalars_dev, int n_scalars)
ALLREDUCE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long,
                                ctypes.c_void_p, ctypes.c_int)


class FitResult(ctypes.Structure):
    _fields_ = [("status", ctypes.c_int), ("iterations", ctypes.c_int), ("evaluations", ctypes.c_int),
                ("fx", ctypes.c_float), ("device_ms", ctypes.c_double), ("timed_ms", ctypes.c_double),
                ("timed_evaluations", ctypes.c_int)]


def library_path():
    paths = sorted(glob.glob(os.path.join(_PKG, "_plmdcaBackend*")))
    if not paths:
        raise DcaBackendError(
            "pydca_b200 CUDA backend not built: no _plmdcaBackend*.so in %s "
            "(run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C pydca_b200/csrc`)" % _PKG)
    return paths[0]


_lib = None


def lib():
    """The loaded shared object with argtypes set (loaded once)."""
    global _lib
    if _lib is not None:
        return _lib
    L = ctypes.CDLL(library_path())
    L.dca_last_error.restype = ctypes.c_char_p
    L.dca_device_count.restype = ctypes.c_int
    L.dca_set_device.argtypes = [ctypes.c_int]
    L.dca_kernel_launches.restype = ctypes.c_ulonglong
    # the reference's own ABI, same argtypes/restype as pydca/plmdca/plmdca.py:81-89
    L.plmdcaBackend.argtypes = (ctypes.c_ushort, ctypes.c_ushort, ctypes.c_char_p, ctypes.c_uint,
                                ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_uint,
                                ctypes.c_uint, ctypes.c_bool)
    L.plmdcaBackend.restype = ctypes.c_void_p
    L.freeFieldsAndCouplings.argtypes = [ctypes.c_void_p]
    L.freeFieldsAndCouplings.restype = None
    L.dca_read_fasta.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(_u8p)]
    L.dca_read_fasta.restype = ctypes.c_long
    L.dca_free_host.argtypes = [ctypes.c_void_p]
    L.dca_free_host.restype = None
    L.dca_weights.argtypes = [_u8p, ctypes.c_long, ctypes.c_int, ctypes.c_double, ctypes.c_int,
                              ctypes.c_long, ctypes.c_long, _u32p]
    L.dca_weights_part.argtypes = [_u8p, ctypes.c_long, ctypes.c_int, ctypes.c_double, ctypes.c_int,
                                   ctypes.c_int, ctypes.c_int, _u32p, _f32p]
    L.dca_plm_create.argtypes = [_u8p, _f32p, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                 ctypes.c_float, ctypes.c_int, ctypes.c_int]
    L.dca_plm_create.restype = ctypes.c_void_p
    L.dca_plm_create_sharded.argtypes = [_u8p, _f32p, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                         ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    L.dca_plm_create_sharded.restype = ctypes.c_void_p
    L.dca_plm_create_asym.argtypes = L.dca_plm_create.argtypes
    L.dca_plm_create_asym.restype = ctypes.c_void_p
    L.dca_plm_asym_export.argtypes = [ctypes.c_void_p, _f32p, _f32p, ctypes.c_int, _f32p]
    L.dca_plm_asym_pack.argtypes = [_f32p, ctypes.c_int, _f32p, ctypes.c_int, ctypes.c_int, _f32p]
    L.dca_plm_fit_asymmetric.argtypes = [_u8p, _f32p, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                         ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p,
                                         ctypes.POINTER(FitResult)]
    L.dca_set_estimator.argtypes = [ctypes.c_int]
    L.dca_plm_local_params.argtypes = [ctypes.c_void_p]
    L.dca_plm_local_params.restype = ctypes.c_long
    L.dca_plm_local_layout.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_long)]
    L.dca_shard_layout.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_long)]
    L.dca_plm_arena.argtypes = [ctypes.c_void_p]
    L.dca_plm_arena.restype = ctypes.c_void_p
    L.dca_plm_x_device.argtypes = [ctypes.c_void_p]
    L.dca_plm_x_device.restype = ctypes.c_void_p
    L.dca_plm_ipc_handle.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
    L.dca_plm_connect_ipc.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]
    L.dca_plm_connect_ptrs.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]
    L.dca_plm_disconnect.argtypes = [ctypes.c_void_p]
    L.dca_plm_destroy.argtypes = [ctypes.c_void_p]
    L.dca_plm_destroy.restype = None
    L.dca_plm_num_params.argtypes = [ctypes.c_void_p]
    L.dca_plm_num_params.restype = ctypes.c_long
    L.dca_plm_set_carry.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.dca_plm_set_fidelity.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.dca_set_arithmetic.argtypes = [ctypes.c_int]
    L.dca_seq_sum_stats.argtypes = [_f32p, _f32p, ctypes.c_long, ctypes.c_long, ctypes.c_float, ctypes.c_float, _f32p,
                                    _u32p, _f32p]
    L.dca_set_gpus.argtypes = [ctypes.c_int]
    L.dca_seq_sum.argtypes = [_f32p, _f32p, ctypes.c_long, ctypes.c_long, ctypes.c_float, ctypes.c_float, _f32p]
    L.dca_plm_init.argtypes = [ctypes.c_void_p, _f32p]
    L.dca_plm_eval.argtypes = [ctypes.c_void_p, _f32p, _f32p, _f64p]
    L.dca_plm_eval_device.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, _f64p]
    L.dca_plm_fit.argtypes = [ctypes.c_void_p, _f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                              ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(FitResult)]
    L.dca_plm_fit_from_init.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(FitResult)]
    L.dca_plm_eval_timed.argtypes = [ctypes.c_void_p, _f32p, _f32p]
    L.dca_lbfgs_vector_bench.argtypes = [ctypes.c_long, ctypes.c_int, ctypes.c_int, _f32p]
    L.dca_plm_cuda_stream.argtypes = [ctypes.c_void_p]
    L.dca_plm_cuda_stream.restype = ctypes.c_void_p
    L.dca_plm_describe.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_long)]
    L.dca_plm_layout_stats.argtypes = [ctypes.c_void_p, _f64p]
    L.dca_last_fit.argtypes = [ctypes.POINTER(FitResult)]
    L.dca_plm_fn_apc.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, _f64p, _f64p]
    L.dca_mf_freqs.argtypes = [_u8p, _f64p, ctypes.c_long, ctypes.c_int, ctypes.c_int, _f64p, _f64p]
    L.dca_mf_corr_mat.argtypes = [_u8p, _f64p, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                  _f64p, _f64p]
    L.dca_mf_couplings.argtypes = [_f64p, ctypes.c_int, _f64p]
    L.dca_mf_fn_apc.argtypes = [_f64p, ctypes.c_int, ctypes.c_int, _f64p, _f64p]
    L.dca_mf_scores.argtypes = [_u8p, _f64p, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                _f64p, _f64p, _f64p, _f64p, _f64p, _f32p]
    L.dca_mf_corr_from_freqs.argtypes = [_f64p, _f64p, ctypes.c_int, ctypes.c_int, _f64p]
    L.dca_fp64_mma_peak.argtypes = [ctypes.c_int, _f64p]
    L.dca_sw_scores.argtypes = [_u8p, ctypes.c_long, ctypes.c_int, ctypes.c_int, _u8p, ctypes.c_int,
                                ctypes.POINTER(ctypes.c_int)]
    L.dca_mf_direct_info.argtypes = [_f64p, _f64p, ctypes.c_int, ctypes.c_int, _f64p, _f64p]
    L.dca_plm_direct_info.argtypes = [_f32p, _f64p, ctypes.c_int, ctypes.c_int, _f64p, _f64p]
    _lib = L
    return L


def last_error():
    return lib().dca_last_error().decode("utf-8", "replace")


ESTIMATORS = {None: -1, "joint": 0, "asymmetric": 1}


def set_estimator(mode):
    """Estimator of plmdcaBackend(): 'joint' (the reference's), 'asymmetric' (independent per-site
    fits, symmetrised at the end) or None = the DCA_B200_ESTIMATOR environment variable."""
    if mode not in ESTIMATORS:
        raise ValueError("estimator must be 'joint', 'asymmetric' or None, got %r" % (mode,))
    check(lib().dca_set_estimator(ESTIMATORS[mode]), "dca_set_estimator")


def asym_pack(W_full, h_full, seqs_len, num_site_states):
    """Directed couplings of all sites -> the reference's parameter vector (symmetrised)."""
    W_full = np.ascontiguousarray(W_full, dtype=np.float32)
    h_full = np.ascontiguousarray(h_full, dtype=np.float32).ravel()
    l, q = int(seqs_len), int(num_site_states)
    x = np.empty(l * q + l * (l - 1) // 2 * q * q, dtype=np.float32)
    check(lib().dca_plm_asym_pack(ptr(W_full, _f32p), int(W_full.shape[1]), ptr(h_full, _f32p), l, q, ptr(x, _f32p)),
          "dca_plm_asym_pack")
    return x


def fit_asymmetric(msa, weights, num_site_states, lambda_h, lambda_J, max_iterations, gpus=0, verbose=False):
    """Asymmetric plmDCA on `gpus` GPUs (0 = all visible / DCA_B200_GPUS) -> (x in the reference layout, FitResult)."""
    msa = np.ascontiguousarray(msa, dtype=np.uint8)
    weights = np.ascontiguousarray(weights, dtype=np.float32)
    n, l = msa.shape
    q = int(num_site_states)
    x = np.empty(l * q + l * (l - 1) // 2 * q * q, dtype=np.float32)
    res = FitResult()
    check(lib().dca_plm_fit_asymmetric(ptr(msa, _u8p), ptr(weights, _f32p), n, l, q, float(lambda_h), float(lambda_J),
                                       int(max_iterations), int(gpus), int(bool(verbose)), ptr(x, _f32p),
                                       ctypes.byref(res)), "dca_plm_fit_asymmetric")
    return x, res


def check(rc, what):
    if rc != 0:
        raise DcaBackendError("%s failed: %s" % (what, last_error()))


def ptr(a, t):
    return a.ctypes.data_as(t)


def biomolecule_int(biomolecule):
    return 1 if biomolecule.strip().upper() == "PROTEIN" else 2


class _HostBuffer:
    """Owner of a host buffer handed out by the library (dca_free_host when the last user is gone)."""

    def __init__(self, p):
        self._p = p

    def __del__(self):
        if self._p:
            try:
                lib().dca_free_host(self._p)
            except Exception:      # interpreter shutdown
                pass
            self._p = None


# ---- thin array-level wrappers -----------------------------------------------
def read_fasta(path, biomolecule, seqs_len):
    """uint8[N, L] unique sequences (states 0..q-1) in first-occurrence order."""
    out = _u8p()
    n = lib().dca_read_fasta(os.fsencode(path), biomolecule_int(biomolecule), int(seqs_len), ctypes.byref(out))
    if n <= 0:
        raise DcaBackendError("reading %s failed: %s" % (path, last_error()))
    # the library's buffer itself (no copy of the alignment), released with the last view of it
    size = int(n) * int(seqs_len)
    buf = (ctypes.c_uint8 * size).from_address(ctypes.addressof(out.contents))
    buf._owner = _HostBuffer(ctypes.cast(out, ctypes.c_void_p))
    return np.frombuffer(buf, dtype=np.uint8).reshape(int(n), int(seqs_len))


def sequence_counts(msa, seqid, compare_f64=False, rows=None):
    """Neighbour counts c_i (uint32[N]); rows=(lo,hi) computes a shard only."""
    msa = np.ascontiguousarray(msa, dtype=np.uint8)
    n, l = msa.shape
    lo, hi = (0, n) if rows is None else rows
    counts = np.zeros(n, dtype=np.uint32)
    check(lib().dca_weights(ptr(msa, _u8p), n, l, float(seqid), int(bool(compare_f64)), lo, hi,
                            ptr(counts, _u32p)), "dca_weights")
    return counts


def sequence_counts_part(msa, seqid, part, nparts, compare_f64=False, want_ms=False):
    """Part `part` of `nparts` of the symmetric all-pairs comparison: partial neighbour counts of
    ALL rows (uint32[N]); summing the parts gives sequence_counts(msa, seqid)."""
    msa = np.ascontiguousarray(msa, dtype=np.uint8)
    n, l = msa.shape
    counts = np.zeros(n, dtype=np.uint32)
    ms = ctypes.c_float()
    check(lib().dca_weights_part(ptr(msa, _u8p), n, l, float(seqid), int(bool(compare_f64)), int(part),
                                 int(nparts), ptr(counts, _u32p), ctypes.byref(ms)), "dca_weights_part")
    return (counts, ms.value) if want_ms else counts


def shard_layout(seqs_len, num_site_states, site_lo, site_hi):
    """(field offset, field count, coupling offset, coupling count) of the parameter slice owned
    by the sites [site_lo, site_hi) inside the reference's vector (host arithmetic only)."""
    out = (ctypes.c_long * 4)()
    check(lib().dca_shard_layout(int(seqs_len), int(num_site_states), int(site_lo), int(site_hi), out),
          "dca_shard_layout")
    return tuple(int(v) for v in out)


class PlmObjective:
    """Device-resident plmDCA objective (the reference's C++ PlmDCA object).

    ``sites=(lo, hi)``: data term of those sites only over the whole (replicated) vector.
    ``shard=(rank, world, bounds)``: owner-sharded multi-GPU objective (include/dca_b200.h);
    init / eval / fit then work on the LOCAL slice of ``Ploc`` floats and are collective."""

    def __init__(self, msa, weights, num_site_states, lambda_h, lambda_J, sites=None, shard=None, asymmetric=False):
        msa = np.ascontiguousarray(msa, dtype=np.uint8)
        weights = np.ascontiguousarray(weights, dtype=np.float32)
        n, l = msa.shape
        self.N, self.L, self.q = n, l, int(num_site_states)
        self.shard = None
        self.asymmetric = bool(asymmetric)
        if asymmetric:
            # per-site estimator on the sites [lo, hi): x = [h (ld) | W (L*q rows x ld)], include/dca_b200.h
            if shard is not None:
                raise ValueError("the asymmetric estimator needs no owner-sharding: give every GPU its own sites=(lo, hi)")
            lo, hi = (0, l) if sites is None else sites
            self.sites = (int(lo), int(hi))
            self._h = lib().dca_plm_create_asym(ptr(msa, _u8p), ptr(weights, _f32p), n, l, self.q,
                                                float(lambda_h), float(lambda_J), int(lo), int(hi))
        elif shard is not None:
            rank, world, bounds = shard
            b = (ctypes.c_int * (int(world) + 1))(*[int(v) for v in bounds])
            self.shard = (int(rank), int(world), tuple(int(v) for v in bounds))
            self._h = lib().dca_plm_create_sharded(ptr(msa, _u8p), ptr(weights, _f32p), n, l, self.q,
                                                   float(lambda_h), float(lambda_J), int(rank), int(world), b)
        else:
            lo, hi = (0, l) if sites is None else sites
            self._h = lib().dca_plm_create(ptr(msa, _u8p), ptr(weights, _f32p), n, l, self.q,
                                           float(lambda_h), float(lambda_J), int(lo), int(hi))
        if not self._h:
            raise DcaBackendError("dca_plm_create failed: %s" % last_error())
        self.P = int(lib().dca_plm_num_params(self._h))
        self.Ploc = int(lib().dca_plm_local_params(self._h))

    # ---- asymmetric estimator ------------------------------------------------------------
    def asym_shape(self):
        """(ncol, ld): valid and padded column counts of the local matrix."""
        ncol = (self.sites[1] - self.sites[0]) * self.q
        ld = self.Ploc // (self.L * self.q + 1)
        return ncol, ld

    def asym_split(self, x):
        """x -> (h [sites, q], W [L*q, sites*q]) views without the pad columns."""
        ncol, ld = self.asym_shape()
        return x[:ncol].reshape(-1, self.q), x[ld:].reshape(self.L * self.q, ld)[:, :ncol]

    def asym_join(self, h, W):
        ncol, ld = self.asym_shape()
        x = np.zeros(self.Ploc, dtype=np.float32)
        x[:ncol] = np.asarray(h, dtype=np.float32).ravel()
        x[ld:].reshape(self.L * self.q, ld)[:, :ncol] = W
        return x

    def asym_export(self, W_full, h_full, x=None):
        """Write my columns of x (default: the device-resident result of the last fit) into the
        all-sites host arrays W_full [L*q, L*q] and h_full [L*q] (float32, C order)."""
        assert W_full.dtype == np.float32 and W_full.flags.c_contiguous and h_full.dtype == np.float32
        xp = None if x is None else ptr(np.ascontiguousarray(x, dtype=np.float32), _f32p)
        check(lib().dca_plm_asym_export(self._h, xp, ptr(W_full, _f32p), int(W_full.shape[1]), ptr(h_full, _f32p)),
              "dca_plm_asym_export")

    # ---- owner-sharded wiring -----------------------------------------------------------
    def local_layout(self):
        out = (ctypes.c_long * 4)()
        check(lib().dca_plm_local_layout(self._h, out), "dca_plm_local_layout")
        return tuple(int(v) for v in out)

    def local_slice(self, x_full):
        """The part of a full parameter vector this rank owns."""
        fo, fc, co, cc = self.local_layout()
        return np.concatenate([x_full[fo:fo + fc], x_full[co:co + cc]]).astype(np.float32)

    def arena(self):
        return int(lib().dca_plm_arena(self._h) or 0)

    def x_device(self):
        return int(lib().dca_plm_x_device(self._h) or 0)

    def ipc_handle(self):
        buf = ctypes.create_string_buffer(64)
        check(lib().dca_plm_ipc_handle(self._h, buf), "dca_plm_ipc_handle")
        return buf.raw

    def connect_ipc(self, handles):
        blob = b"".join(handles)
        check(lib().dca_plm_connect_ipc(self._h, blob, len(handles)), "dca_plm_connect_ipc")

    def connect_ptrs(self, arena_ptrs):
        arr = (ctypes.c_void_p * len(arena_ptrs))(*[int(v) for v in arena_ptrs])
        check(lib().dca_plm_connect_ptrs(self._h, arr), "dca_plm_connect_ptrs")

    def disconnect(self):
        if getattr(self, "_h", None):
            check(lib().dca_plm_disconnect(self._h), "dca_plm_disconnect")

    def close(self):
        if getattr(self, "_h", None):
            lib().dca_plm_destroy(self._h)
            self._h = None

    __del__ = close

    def set_carry(self, carry=True, chunk=0, warm=-1):
        check(lib().dca_plm_set_carry(self._h, int(bool(carry)), int(chunk), int(warm)), "dca_plm_set_carry")

    def set_arithmetic(self, mode):
        """'reference': the reference's sequential float32 sums, bit for bit (its L-BFGS
        trajectory); 'accurate': float64 fixed-order reductions (faster, FN differs ~1e-4)."""
        modes = {"reference": 1, "accurate": 0}
        if mode not in modes:
            raise ValueError("arithmetic mode must be 'reference' or 'accurate', got %r" % (mode,))
        check(lib().dca_plm_set_fidelity(self._h, modes[mode]), "dca_plm_set_fidelity")

    def init(self):
        x = np.zeros(self.Ploc, dtype=np.float32)
        check(lib().dca_plm_init(self._h, ptr(x, _f32p)), "dca_plm_init")
        return x

    def eval(self, x, out=None):
        x = np.ascontiguousarray(x, dtype=np.float32)
        g = np.empty(self.Ploc, dtype=np.float32) if out is None else out
        fx = ctypes.c_double()
        check(lib().dca_plm_eval(self._h, ptr(x, _f32p), ptr(g, _f32p), ctypes.byref(fx)), "dca_plm_eval")
        return fx.value, g

    def eval_device(self, x_ptr, g_ptr, dir_ptr=None, read_scalars=True):
        sc = np.zeros(4, dtype=np.float64)
        check(lib().dca_plm_eval_device(self._h, x_ptr, g_ptr, dir_ptr,
                                        ptr(sc, _f64p) if read_scalars else None), "dca_plm_eval_device")
        return sc

    KERNELS = ("unpack", "logits", "softmax_scan", "scatter", "export", "assemble")

    def eval_timed(self, x):
        """Per-kernel CUDA-event milliseconds of one evaluation (dict by kernel)."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        ms = np.zeros(6, dtype=np.float32)
        check(lib().dca_plm_eval_timed(self._h, ptr(x, _f32p), ptr(ms, _f32p)), "dca_plm_eval_timed")
        return dict(zip(self.KERNELS, ms.tolist()))

    def layout_stats(self):
        """{'sorted', 'group_sums', 'uniform4', 'logits_columns'}: the row layout the gather kernels run on."""
        out = np.zeros(4, dtype=np.float64)
        check(lib().dca_plm_layout_stats(self._h, ptr(out, _f64p)), "dca_plm_layout_stats")
        return {"sorted": bool(out[0]), "group_sums": bool(out[1]), "uniform4": float(out[2]), "logits_columns": int(out[3])}

    def cuda_stream(self):
        """Raw cudaStream_t (int) the handle's kernels run on."""
        return int(lib().dca_plm_cuda_stream(self._h) or 0)

    def describe(self):
        out = (ctypes.c_long * 10)()
        check(lib().dca_plm_describe(self._h, out), "dca_plm_describe")
        keys = ("N", "L", "q", "ld", "P", "nsplit", "nchunk", "col_tiles", "Npad", "nscan")
        return dict(zip(keys, [int(v) for v in out]))

    def fit(self, x0, max_iterations, verbose=False, warmup_iterations=0, allreduce=None):
        x = np.array(x0, dtype=np.float32, copy=True)
        res = FitResult()
        cb = ctypes.cast(allreduce, ctypes.c_void_p) if allreduce is not None else None
        check(lib().dca_plm_fit(self._h, ptr(x, _f32p), int(max_iterations), int(bool(verbose)),
                                int(warmup_iterations), cb, None, ctypes.byref(res)), "dca_plm_fit")
        return x, res


def _fit_from_init(self, max_iterations, verbose=False):
    """Initial point built on the device, L-BFGS, result left in the handle's device buffer
    (x_device()): no host copy of the parameters in either direction."""
    res = FitResult()
    check(lib().dca_plm_fit_from_init(self._h, int(max_iterations), int(bool(verbose)), ctypes.byref(res)),
          "dca_plm_fit_from_init")
    return res


PlmObjective.fit_from_init = _fit_from_init


def plm_fn_apc(x, seqs_len, num_site_states):
    x = np.ascontiguousarray(x, dtype=np.float32)
    npairs = seqs_len * (seqs_len - 1) // 2
    fn = np.zeros(npairs, dtype=np.float64)
    apc = np.zeros(npairs, dtype=np.float64)
    check(lib().dca_plm_fn_apc(ptr(x, _f32p), int(seqs_len), int(num_site_states), ptr(fn, _f64p),
                               ptr(apc, _f64p)), "dca_plm_fn_apc")
    return fn, apc


ARITHMETIC_MODES = {"accurate": 0, "reference": 1, None: -1}


def set_backend_arithmetic(mode):
    """Arithmetic of plmdcaBackend() for this process: 'accurate', 'reference' or None (follow
    the DCA_B200_ARITHMETIC environment variable; default accurate)."""
    if mode not in ARITHMETIC_MODES:
        raise ValueError("arithmetic must be 'accurate', 'reference' or None, got %r" % (mode,))
    check(lib().dca_set_arithmetic(ARITHMETIC_MODES[mode]), "dca_set_arithmetic")


def set_backend_gpus(n):
    """GPUs plmdcaBackend() drives in this process (0 = DCA_B200_GPUS, default all visible)."""
    check(lib().dca_set_gpus(int(n)), "dca_set_gpus")


def seq_dot(x, y):
    """Sequential float32 dot product (arithmetic_ansi.h:114-121) evaluated exactly on the GPU."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.ascontiguousarray(y, dtype=np.float32)
    out = ctypes.c_float()
    check(lib().dca_seq_sum(ptr(x, _f32p), ptr(y, _f32p), x.size, 0, 0.0, 0.0, ctypes.byref(out)), "dca_seq_sum")
    return np.float32(out.value)


def seq_dot_stats(x, y):
    """(value, kernel ms of one warm launch, pipeline counters) of the sequential float32 dot."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.ascontiguousarray(y, dtype=np.float32)
    out, ms = ctypes.c_float(), ctypes.c_float()
    st = np.zeros(8, dtype=np.uint32)
    check(lib().dca_seq_sum_stats(ptr(x, _f32p), ptr(y, _f32p), x.size, 0, 0.0, 0.0, ctypes.byref(out),
                                  ptr(st, _u32p), ctypes.byref(ms)), "dca_seq_sum_stats")
    keys = ("from_records", "exact", "timeouts", "stale", "range", "kcycles", "kcycles_exact", "row_walks")
    return np.float32(out.value), ms.value, dict(zip(keys, st.tolist()))


def seq_reg(x, n0, lambda0, lambda1):
    """Sequential float32 L2 accumulation (plmdca_numerics.cpp:463-486) evaluated exactly on the GPU."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = ctypes.c_float()
    check(lib().dca_seq_sum(ptr(x, _f32p), None, x.size, int(n0), float(lambda0), float(lambda1),
                            ctypes.byref(out)), "dca_seq_sum")
    return np.float32(out.value)


def lbfgs_vector_bench(n, stored_pairs=5, reps=20):
    """CUDA-event ms of {history+inner products, direction, step} passes on n floats."""
    ms = np.zeros(3, dtype=np.float32)
    check(lib().dca_lbfgs_vector_bench(int(n), int(stored_pairs), int(reps), ptr(ms, _f32p)), "dca_lbfgs_vector_bench")
    return dict(zip(("update_gram", "direction", "step"), ms.tolist()))


def sw_scores(msa, num_site_states, ref_codes):
    """Local alignment score of the reference (codes 0..q-2) against every row of the zero-based
    alignment, gaps removed (int32[N])."""
    msa = np.ascontiguousarray(msa, dtype=np.uint8)
    ref = np.ascontiguousarray(ref_codes, dtype=np.uint8)
    n, l = msa.shape
    out = np.zeros(n, dtype=np.int32)
    check(lib().dca_sw_scores(ptr(msa, _u8p), n, l, int(num_site_states), ptr(ref, _u8p), ref.size,
                              out.ctypes.data_as(ctypes.POINTER(ctypes.c_int))), "dca_sw_scores")
    return out


def fp64_mma_peak(reps=5):
    """Measured FP64 tensor-pipe peak (TFLOP/s): a register-only DMMA stream on every SM."""
    out = ctypes.c_double()
    check(lib().dca_fp64_mma_peak(int(reps), ctypes.byref(out)), "dca_fp64_mma_peak")
    return out.value


def last_fit():
    """L-BFGS statistics of this thread's most recent plmdcaBackend() call."""
    res = FitResult()
    check(lib().dca_last_fit(ctypes.byref(res)), "dca_last_fit")
    return res


def plmdca_backend(msa_file, biomolecule, seqs_len, num_site_states, seqid, lambda_h, lambda_J,
                   max_iterations, num_threads=1, verbose=False, copy=False):
    """The reference's C entry point, called the way pydca/plmdca/plmdca.py:216-230 calls it.

    The reference copies the returned buffer element by element and then calls
    freeFieldsAndCouplings (:225-230).  Here the float32[P] result is a zero-copy view of the
    library's host buffer; freeFieldsAndCouplings runs when the array (and every view of it) is
    garbage collected.  ``copy=True`` gives the reference's copy-then-free behaviour."""
    p = seqs_len * num_site_states + seqs_len * (seqs_len - 1) // 2 * num_site_states ** 2
    L = lib()
    h = L.plmdcaBackend(biomolecule_int(biomolecule), num_site_states, os.fsencode(msa_file), seqs_len,
                        seqid, lambda_h, lambda_J, max_iterations, num_threads, verbose)
    if not h:
        raise DcaBackendError("plmdcaBackend failed: %s" % last_error())
    x = np.ctypeslib.as_array(ctypes.cast(h, _f32p), shape=(p,))
    if copy:
        x = x.copy()
        L.freeFieldsAndCouplings(h)
        return x
    weakref.finalize(x, L.freeFieldsAndCouplings, h)
    return x
