"""ctypes binding of include/ms_b200.h (one Solver = one ms_ctx = one rank = one GPU)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MS_VEC = {"d": 0, "b": 1, "y": 2, "x": 3, "r": 4}
dp = C.POINTER(C.c_double)


class MSError(RuntimeError):
    """CUDA / NCCL / argument error from the library (negative return code)."""


class NumericalFailure(ArithmeticError):
    """Return code 1: a pivot fell below EPS*norm (the reference's `return 1`)."""


class Stats(C.Structure):
    _fields_ = [
        ("factor_ms", C.c_double), ("update_ms", C.c_double), ("forward_ms", C.c_double),
        ("backward_ms", C.c_double), ("init_ms", C.c_double), ("bnorm_ms", C.c_double),
        ("residual_ms", C.c_double), ("update_flops", C.c_double),
        ("kernel_launches", C.c_longlong), ("update_launches", C.c_longlong),
        ("h2d_bytes", C.c_longlong), ("d2h_bytes", C.c_longlong), ("used_dmma", C.c_int),
        ("forward_fused", C.c_int),
        ("update_sum_ms", C.c_double),
        ("block_internal", C.c_int),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def lib_path():
    return os.path.join(_HERE, "lib", "libms_b200.so")


def load():
    """Load the CUDA library; fail loudly if it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise MSError(f"{path} not built: run `make lib` (or __graft_entry__.build()); "
                      "there is no CPU fallback")
    L = C.CDLL(path, mode=C.RTLD_GLOBAL)
    sz, vp, ci, cd = C.c_size_t, C.c_void_p, C.c_int, C.c_double
    sig = {
        "ms_last_error": (C.c_char_p, []),
        "ms_device_count": (ci, [C.POINTER(ci)]),
        "ms_get_unique_id": (ci, [vp]),
        "ms_create": (ci, [C.POINTER(vp), sz, sz, ci, ci, ci, vp]),
        "ms_destroy": (ci, [vp]),
        "ms_set_option": (ci, [vp, C.c_char_p, C.c_long]),
        "ms_get_stats": (ci, [vp, C.POINTER(Stats)]),
        "ms_sync": (ci, [vp]),
        "ms_init_formula": (ci, [vp, ci]),
        "ms_upload_rows": (ci, [vp, vp, sz, sz, sz]),
        "ms_upload_rows_async": (ci, [vp, vp, sz, sz, sz]),
        "ms_download_rows": (ci, [vp, vp, sz, sz, sz]),
        "ms_host_alloc": (ci, [C.POINTER(vp), sz]),
        "ms_host_free": (ci, [vp]),
        "ms_upload_packed": (ci, [vp, vp]),
        "ms_download_packed": (ci, [vp, vp]),
        "ms_get_row": (ci, [vp, sz, dp]),
        "ms_get_col": (ci, [vp, sz, dp]),
        "ms_get_block": (ci, [vp, sz, sz, sz, sz, dp]),
        "ms_b_and_norm": (ci, [vp, C.POINTER(cd)]),
        "ms_factor": (ci, [vp, cd, C.POINTER(cd)]),
        "ms_solve_forward": (ci, [vp, cd]),
        "ms_solve_backward": (ci, [vp, cd]),
        "ms_residual": (ci, [vp, C.POINTER(cd), C.POINTER(cd)]),
        "ms_set_vec": (ci, [vp, ci, dp, sz]),
        "ms_get_vec": (ci, [vp, ci, dp, sz]),
        "ms_block_diag_factor": (ci, [dp, dp, ci, cd]),
        "ms_block_upper_inverse": (ci, [dp, dp, ci, cd]),
        "ms_block_panel": (ci, [dp, dp, dp, ci, ci]),
        "ms_block_trsm_panel": (ci, [dp, dp, dp, ci, ci]),
        "ms_block_update": (ci, [dp, dp, dp, dp, ci, ci, ci]),
        "ms_block_update_diag": (ci, [dp, dp, dp, ci, ci]),
        "ms_norm2": (ci, [dp, sz, C.POINTER(cd)]),
        "ms_error_vs_planted": (ci, [dp, sz, C.POINTER(cd)]),
        "ms_probe_dmma_peak": (ci, [ci, C.POINTER(cd)]),
        "ms_probe_hbm_copy": (ci, [ci, sz, C.POINTER(cd)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _LIB = L
    return L


EXPORTED = [
    "ms_last_error", "ms_device_count", "ms_get_unique_id", "ms_create", "ms_destroy",
    "ms_set_option", "ms_get_stats", "ms_sync", "ms_init_formula", "ms_upload_rows", "ms_upload_rows_async",
    "ms_download_rows", "ms_host_alloc", "ms_host_free", "ms_upload_packed", "ms_download_packed", "ms_get_row", "ms_get_col", "ms_get_block", "ms_b_and_norm",
    "ms_factor", "ms_solve_forward", "ms_solve_backward", "ms_residual", "ms_set_vec",
    "ms_get_vec", "ms_block_diag_factor", "ms_block_upper_inverse", "ms_block_panel",
    "ms_block_trsm_panel", "ms_block_update", "ms_block_update_diag", "ms_norm2", "ms_error_vs_planted",
    "ms_probe_dmma_peak", "ms_probe_hbm_copy",
]


def _check(rc, allow_numerical=False):
    if rc == 0:
        return 0
    if rc == 1 and allow_numerical:
        return 1
    if rc == 1:
        raise NumericalFailure("pivot below EPS*norm")
    raise MSError(f"ms_b200 error {rc}: {load().ms_last_error().decode()}")


def _p(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(dp)


def device_count():
    n = C.c_int(0)
    _check(load().ms_device_count(C.byref(n)))
    return n.value


def get_unique_id():
    buf = C.create_string_buffer(128)
    _check(load().ms_get_unique_id(buf))
    return buf.raw


class Solver:
    """Device-resident problem: A (n x n, block size m), vectors d b y x r."""

    def __init__(self, n, m, device=0, rank=0, nranks=1, unique_id=None):
        self.L = load()
        self.n, self.m, self.rank, self.nranks = n, m, rank, nranks
        self.h = C.c_void_p()
        uid = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        _check(self.L.ms_create(C.byref(self.h), n, m, device, rank, nranks, uid))

    def close(self):
        if self.h:
            self.L.ms_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_option(self, key, value):
        _check(self.L.ms_set_option(self.h, key.encode(), int(value)))

    def stats(self):
        s = Stats()
        _check(self.L.ms_get_stats(self.h, C.byref(s)))
        return s.as_dict()

    def sync(self):
        _check(self.L.ms_sync(self.h))

    # ---- init
    def init_formula(self, kind=0):
        _check(self.L.ms_init_formula(self.h, kind))

    def upload_rows(self, rows, row0=0):
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        _check(self.L.ms_upload_rows(self.h, rows.ctypes.data, row0, rows.shape[0], rows.shape[1]))

    def upload_rows_ptr(self, ptr, row0, nrows, ld_host):
        _check(self.L.ms_upload_rows(self.h, ptr, row0, nrows, ld_host))

    def download_rows_ptr(self, ptr, row0, nrows, ld_host):
        _check(self.L.ms_download_rows(self.h, ptr, row0, nrows, ld_host))

    def download_rows(self, row0, nrows):
        out = np.zeros((nrows, self.n))
        _check(self.L.ms_download_rows(self.h, out.ctypes.data, row0, nrows, self.n))
        return out

    def upload_packed(self, body):
        ptr = body if isinstance(body, int) else body.ctypes.data
        _check(self.L.ms_upload_packed(self.h, ptr))

    def download_packed(self, body):
        ptr = body if isinstance(body, int) else body.ctypes.data
        _check(self.L.ms_download_packed(self.h, ptr))

    def get_row(self, i):
        buf = np.zeros(self.n)
        _check(self.L.ms_get_row(self.h, i, _p(buf)))
        return buf

    def get_block(self, row0, nrows, col0, ncols):
        buf = np.zeros((nrows, ncols))
        _check(self.L.ms_get_block(self.h, row0, nrows, col0, ncols, _p(buf)))
        return buf

    def get_col(self, j):
        buf = np.zeros(self.n)
        _check(self.L.ms_get_col(self.h, j, _p(buf)))
        return buf

    def get_upper(self):
        """Dense (n, n) array holding the stored upper triangle (small n only)."""
        out = np.zeros((self.n, self.n))
        for i in range(self.n):
            out[i] = self.get_row(i)
        return out

    # ---- pipeline
    def b_and_norm(self):
        norm = C.c_double(0)
        _check(self.L.ms_b_and_norm(self.h, C.byref(norm)))
        return norm.value

    def factor(self, norm):
        """Returns (failed, elapsed_seconds)."""
        el = C.c_double(0)
        rc = _check(self.L.ms_factor(self.h, norm, C.byref(el)), allow_numerical=True)
        return bool(rc), el.value

    def solve_forward(self, norm):
        return bool(_check(self.L.ms_solve_forward(self.h, norm), allow_numerical=True))

    def solve_backward(self, norm):
        return bool(_check(self.L.ms_solve_backward(self.h, norm), allow_numerical=True))

    def residual(self):
        r, e = C.c_double(0), C.c_double(0)
        _check(self.L.ms_residual(self.h, C.byref(r), C.byref(e)))
        return r.value, e.value

    def get_vec(self, which):
        buf = np.zeros(self.n)
        _check(self.L.ms_get_vec(self.h, MS_VEC[which], _p(buf), self.n))
        return buf

    def set_vec(self, which, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        _check(self.L.ms_set_vec(self.h, MS_VEC[which], _p(v), len(v)))


# ---- stateless block kernels (host buffers in the reference's layouts)
def _pack(u):
    return np.ascontiguousarray(u[np.triu_indices(u.shape[0])], dtype=np.float64)


def _unpack(p, w):
    u = np.zeros((w, w))
    u[np.triu_indices(w)] = p
    return u


def block_diag_factor(blk, norma):
    w = blk.shape[0]
    a, d = _pack(blk), np.zeros(w)
    rc = _check(load().ms_block_diag_factor(_p(a), _p(d), w, norma), allow_numerical=True)
    return _unpack(a, w), d, rc


def block_upper_inverse(blk, norma):
    w = blk.shape[0]
    a, inv = _pack(blk), np.zeros(w * (w + 1) // 2)
    rc = _check(load().ms_block_upper_inverse(_p(a), _p(inv), w, norma), allow_numerical=True)
    return _unpack(inv, w), rc


def block_panel(inv, A, D):
    A = np.ascontiguousarray(A, dtype=np.float64).copy()
    h, w = A.shape
    _check(load().ms_block_panel(_p(_pack(inv)), _p(A), _p(np.ascontiguousarray(D, dtype=np.float64)), h, w))
    return A


def block_trsm_panel(R, A, D):
    A = np.ascontiguousarray(A, dtype=np.float64).copy()
    h, w = A.shape
    _check(load().ms_block_trsm_panel(_p(_pack(R)), _p(A), _p(np.ascontiguousarray(D, dtype=np.float64)), h, w))
    return A


def block_update(A, R1, D, R2):
    A = np.ascontiguousarray(A, dtype=np.float64).copy()
    R1 = np.ascontiguousarray(R1, dtype=np.float64)
    R2 = np.ascontiguousarray(R2, dtype=np.float64)
    k, h = R1.shape
    w = R2.shape[1]
    _check(load().ms_block_update(_p(A), _p(R1), _p(np.ascontiguousarray(D, dtype=np.float64)), _p(R2), k, h, w))
    return A


def block_update_diag(Aup, R1, D):
    R1 = np.ascontiguousarray(R1, dtype=np.float64)
    k, w = R1.shape
    a = _pack(Aup)
    _check(load().ms_block_update_diag(_p(a), _p(R1), _p(np.ascontiguousarray(D, dtype=np.float64)), k, w))
    return _unpack(a, w)


def norm2(v):
    out = C.c_double(0)
    v = np.ascontiguousarray(v, dtype=np.float64)
    _check(load().ms_norm2(_p(v), len(v), C.byref(out)))
    return out.value


def error_vs_planted(x):
    out = C.c_double(0)
    x = np.ascontiguousarray(x, dtype=np.float64)
    _check(load().ms_error_vs_planted(_p(x), len(x), C.byref(out)))
    return out.value


class PinnedBuffer:
    """Page-locked host array of float64 (cudaHostAlloc through the C ABI)."""

    def __init__(self, count):
        self.count = count
        self.ptr = C.c_void_p()
        _check(load().ms_host_alloc(C.byref(self.ptr), count * 8))
        self.array = np.ctypeslib.as_array(C.cast(self.ptr, dp), shape=(count,))

    def free(self):
        if self.ptr:
            self.array = None
            load().ms_host_free(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def probe_dmma_peak(device=0):
    out = C.c_double(0)
    _check(load().ms_probe_dmma_peak(device, C.byref(out)))
    return out.value


def probe_hbm_copy(device=0, nbytes=1 << 30):
    out = C.c_double(0)
    _check(load().ms_probe_hbm_copy(device, nbytes, C.byref(out)))
    return out.value
