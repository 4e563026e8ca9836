"""ctypes access to the CPU checkers (TEST INFRASTRUCTURE ONLY).

* `Oracle`  — oracle/_build/liboracle.so, our C restatement (oracle/ms_oracle.c).
* `RefBlocks` / `ref_dump()` — oracle/_ref/, the unmodified reference compiled from
  /root/reference by oracle/Makefile (only present where `make -C oracle ref` ran or
  where gpurun shipped the prebuilt files).

Nothing in matrix_solver_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "liboracle.so")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
REF_BLOCKS_SO = os.path.join(REF_DIR, "libref_blocks.so")
REF_DUMP = os.path.join(REF_DIR, "ref_dump")
REF_CLI = os.path.join(REF_DIR, "test_mat_ref")
REF_CLI_O2 = os.path.join(REF_DIR, "test_mat_ref_O2")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
FIXTURE_DIR = os.path.join(GOLDEN_DIR, "inputs")

dp = C.POINTER(C.c_double)


def _p(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(dp)


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "oracle"], check=True)


def have_ref():
    return all(os.path.exists(p) for p in (REF_BLOCKS_SO, REF_DUMP, REF_CLI_O2))


class Oracle:
    """Dense-storage restatement; arrays are numpy float64, A/R are (n, n) upper."""

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        L = C.CDLL(ORACLE_SO)
        sz = C.c_size_t
        L.ora_alloc_size.restype = sz
        L.ora_alloc_size.argtypes = [sz, sz, C.c_int, C.c_int]
        L.ora_packed_index.restype = sz
        L.ora_packed_index.argtypes = [sz, sz, sz]
        L.ora_packed_offset.restype = sz
        L.ora_packed_offset.argtypes = [sz, sz, sz, sz]
        L.ora_init_formula.argtypes = [dp, sz]
        L.ora_init_file.restype = C.c_int
        L.ora_init_file.argtypes = [dp, sz, C.c_char_p]
        L.ora_b_norm.argtypes = [dp, sz, dp, dp]
        L.ora_diag_factor.restype = C.c_int
        L.ora_diag_factor.argtypes = [dp, sz, dp, C.c_int, C.c_double]
        L.ora_upper_inverse.restype = C.c_int
        L.ora_upper_inverse.argtypes = [dp, sz, dp, C.c_int, C.c_double]
        L.ora_update.argtypes = [dp, sz, dp, sz, dp, dp, sz, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ora_panel.argtypes = [dp, dp, sz, dp, C.c_int, C.c_int]
        L.ora_factor.restype = C.c_int
        L.ora_factor.argtypes = [dp, dp, sz, sz, C.c_double]
        L.ora_forward.restype = C.c_int
        L.ora_forward.argtypes = [dp, dp, dp, sz, C.c_double]
        L.ora_backward.restype = C.c_int
        L.ora_backward.argtypes = [dp, dp, dp, dp, sz, C.c_double]
        L.ora_residual.argtypes = [dp, dp, dp, dp, sz]
        L.ora_norm2.restype = C.c_double
        L.ora_norm2.argtypes = [dp, sz]
        L.ora_error.restype = C.c_double
        L.ora_error.argtypes = [dp, sz]
        self.L = L

    # ---- layout
    def alloc_size(self, n, m, rank=0, nranks=1):
        return self.L.ora_alloc_size(n, m, rank, nranks)

    def packed_offset(self, i, j, n, m):
        return self.L.ora_packed_offset(i, j, n, m)

    # ---- init
    def init_formula(self, n):
        a = np.zeros((n, n))
        self.L.ora_init_formula(_p(a), n)
        return a

    def init_file(self, n, path):
        a = np.zeros((n, n))
        st = self.L.ora_init_file(_p(a), n, os.fsencode(path))
        return a, st

    def b_norm(self, a):
        n = a.shape[0]
        b = np.zeros(n)
        norm = C.c_double(0)
        self.L.ora_b_norm(_p(a), n, _p(b), C.byref(norm))
        return b, norm.value

    # ---- blocks (operate on copies, return results)
    def diag_factor(self, blk, norma):
        a = np.ascontiguousarray(blk, dtype=np.float64).copy()
        w = a.shape[0]
        d = np.zeros(w)
        fail = self.L.ora_diag_factor(_p(a), w, _p(d), w, norma)
        return a, d, fail

    def upper_inverse(self, blk, norma):
        a = np.ascontiguousarray(blk, dtype=np.float64)
        w = a.shape[0]
        inv = np.zeros((w, w))
        fail = self.L.ora_upper_inverse(_p(a), w, _p(inv), w, norma)
        return inv, fail

    def update(self, A, R1, D, R2, upper_only=False):
        A = np.ascontiguousarray(A, dtype=np.float64).copy()
        R1 = np.ascontiguousarray(R1, dtype=np.float64)
        R2 = np.ascontiguousarray(R2, dtype=np.float64)
        D = np.ascontiguousarray(D, dtype=np.float64)
        k, h = R1.shape
        w = R2.shape[1]
        assert A.shape == (h, w) and R2.shape[0] == k
        self.L.ora_update(_p(A), w, _p(R1), h, _p(D), _p(R2), w, k, h, w, int(upper_only))
        return A

    def panel(self, inv, A, D):
        A = np.ascontiguousarray(A, dtype=np.float64).copy()
        inv = np.ascontiguousarray(inv, dtype=np.float64)
        D = np.ascontiguousarray(D, dtype=np.float64)
        h, w = A.shape
        self.L.ora_panel(_p(inv), _p(A), w, _p(D), h, w)
        return A

    # ---- pipeline
    def factor(self, a, m, norma):
        r = a.copy()
        n = r.shape[0]
        d = np.zeros(n)
        fail = self.L.ora_factor(_p(r), _p(d), n, m, norma)
        return r, d, fail

    def forward(self, r, b, norma):
        n = r.shape[0]
        y = np.zeros(n)
        fail = self.L.ora_forward(_p(r), _p(b), _p(y), n, norma)
        return y, fail

    def backward(self, r, y, d, norma):
        n = r.shape[0]
        x = np.zeros(n)
        fail = self.L.ora_backward(_p(r), _p(x), _p(y), _p(d), n, norma)
        return x, fail

    def residual(self, a, b, x):
        n = a.shape[0]
        r = np.zeros(n)
        self.L.ora_residual(_p(a), _p(b), _p(x), _p(r), n)
        return r

    def norm2(self, v):
        return self.L.ora_norm2(_p(np.ascontiguousarray(v)), len(v))

    def error(self, x):
        return self.L.ora_error(_p(np.ascontiguousarray(x)), len(x))

    def solve_pipeline(self, n, m, path=None):
        """The whole test_mat flow (test/test_mat.cpp:89-165) on the oracle."""
        out = {"n": n, "m": m}
        if path:
            a, st = self.init_file(n, path)
            out["init_ok"] = st == 1
            if st != 1:
                return out
        else:
            a = self.init_formula(n)
            out["init_ok"] = True
        out["A"] = a
        b, norm = self.b_norm(a)
        out["b"], out["norm"] = b, norm
        r, d, fail = self.factor(a, m, norm)
        out["factor_fail"] = bool(fail)
        out["d"] = d
        if fail:
            return out
        out["R"] = r
        y, f2 = self.forward(r, b, norm)
        out["y"], out["forward_fail"] = y, bool(f2)
        if f2:
            return out
        x, f3 = self.backward(r, y, d, norm)
        out["x"], out["backward_fail"] = x, bool(f3)
        if f3:
            return out
        res = self.residual(a, b, x)
        out["r"] = res
        out["residual"] = self.norm2(res)
        out["error"] = self.error(x)
        return out


class RefBlocks:
    """The reference's own block kernels (packed-upper layouts as in the reference)."""

    def __init__(self):
        L = C.CDLL(REF_BLOCKS_SO)
        L.ref_diag_factor.restype = C.c_int
        L.ref_diag_factor.argtypes = [dp, dp, C.c_int, C.c_double]
        L.ref_upper_inverse.restype = C.c_int
        L.ref_upper_inverse.argtypes = [dp, dp, C.c_int, C.c_double]
        L.ref_panel.argtypes = [dp, dp, dp, C.c_int, C.c_int]
        L.ref_update.argtypes = [dp, dp, dp, dp, C.c_int, C.c_int]
        L.ref_update_diag.argtypes = [dp, dp, dp, C.c_int, C.c_int]
        L.ref_norm2.restype = C.c_double
        L.ref_norm2.argtypes = [dp, C.c_int]
        L.ref_error.restype = C.c_double
        L.ref_error.argtypes = [dp, C.c_int]
        L.ref_packed_index.restype = C.c_size_t
        L.ref_packed_index.argtypes = [C.c_size_t] * 3
        self.L = L

    @staticmethod
    def pack(u):
        w = u.shape[0]
        return np.ascontiguousarray(u[np.triu_indices(w)])

    @staticmethod
    def unpack(p, w):
        u = np.zeros((w, w))
        u[np.triu_indices(w)] = p
        return u

    def diag_factor(self, blk, norma):
        w = blk.shape[0]
        a = self.pack(blk)
        d = np.zeros(w)
        fail = self.L.ref_diag_factor(_p(a), _p(d), w, norma)
        return self.unpack(a, w), d, fail

    def upper_inverse(self, blk, norma):
        w = blk.shape[0]
        a = self.pack(blk)
        b = np.zeros_like(a)
        fail = self.L.ref_upper_inverse(_p(a), _p(b), w, norma)
        return self.unpack(b, w), fail

    def panel(self, inv, A, D):
        A = np.ascontiguousarray(A).copy()
        h, w = A.shape
        self.L.ref_panel(_p(self.pack(inv)), _p(A), _p(np.ascontiguousarray(D)), h, w)
        return A

    def update(self, A, R1, D, R2):
        A = np.ascontiguousarray(A).copy()
        k, h = R1.shape
        assert k == h, "reference update kernels take square R1"
        w = R2.shape[1]
        self.L.ref_update(_p(A), _p(np.ascontiguousarray(R1)), _p(np.ascontiguousarray(D)),
                          _p(np.ascontiguousarray(R2)), h, w)
        return A

    def update_diag(self, A, R1, D):
        """A is (w, w) upper; R1 is (k, w) with the reference's n := k, l := w."""
        k, w = R1.shape
        a = self.pack(np.ascontiguousarray(A))
        self.L.ref_update_diag(_p(a), _p(np.ascontiguousarray(R1)), _p(np.ascontiguousarray(D)), k, w)
        return self.unpack(a, w)


def parse_dump(raw, n):
    tri = n * (n + 1) // 2
    head = raw[:10]
    o = 10
    out = {
        "n": int(head[0]), "m": int(head[1]), "alloc": int(head[2]), "init_ok": bool(head[3]),
        "norm": head[4], "factor_fail": bool(head[5]), "forward_fail": bool(head[6]),
        "backward_fail": bool(head[7]), "residual": head[8], "error": head[9],
    }
    for k in ("b", "d", "y", "x", "r"):
        out[k] = raw[o:o + n].copy()
        o += n
    iu = np.triu_indices(n)
    for k in ("A", "R"):
        mtx = np.zeros((n, n))
        mtx[iu] = raw[o:o + tri]
        out[k] = mtx
        o += tri
    assert o == raw.size
    return out


def ref_dump(n, m, path=None, np_ranks=1):
    """Run the reference pipeline (oracle/_ref/ref_dump) and return all intermediates."""
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "dump.bin")
        cmd = [REF_DUMP, str(n), str(m)] + ([path] if path else []) + [out]
        env = dict(os.environ, MINIMPI_NP=str(np_ranks))
        subprocess.run(cmd, check=True, env=env, stdout=subprocess.DEVNULL)
        raw = np.fromfile(out, dtype=np.float64)
    return parse_dump(raw, n)


def compare_cli_output(got, want, case=""):
    """The drop-in CLI against the reference's stdout: same tokens in the same order, text equal, numbers
    compared as numbers (residual / er: same order of magnitude), `elapsed` ignored."""
    import re
    num = re.compile(r"-?\d+\.\d+(?:e[+-]\d+)?")
    shape = re.compile(r"[ -]?\d+\.\d+(?:e[+-]\d+)?")    # a sign occupies the pad blank
    gl, wl = got.strip().split("\n"), want.strip().split("\n")
    assert len(gl) == len(wl), (case, got, want)
    for a, b in zip(gl, wl):
        a2, b2 = a.split(", elapsed")[0], b.split(", elapsed")[0]
        # same tokens in the same order: text must match, numbers are compared as numbers
        ta, tb = a2.split(), b2.split()
        assert len(ta) == len(tb), (case, a, b)
        for x, y in zip(ta, tb):
            if num.fullmatch(x.rstrip(",")) and num.fullmatch(y.rstrip(",")):
                continue
            assert x == y, (case, a, b)
        if "-" not in num.sub("", a2) + num.sub("", b2) and not re.search(r"\d-|\s-\d", a2 + b2):
            # no minus signs anywhere: the column layout must be identical character by character
            assert shape.sub("#", a2) == shape.sub("#", b2), (case, a, b)
        fa, fb = [float(v) for v in num.findall(a2)], [float(v) for v in num.findall(b2)]
        if "residual" in a2:
            # n, m are integers (not matched); residual and er: same order of magnitude
            assert fa[0] <= max(4 * fb[0], 1e-13) and fa[1] <= max(4 * fb[1], 1e-12), (case, a, b)
        else:
            assert np.allclose(fa, fb, rtol=2e-2, atol=1e-5), (case, a, b)
