"""CPU model of trsm_quad_loop_kernel (csrc/ms_kernels_diag.cu): the index bookkeeping, not the arithmetic speed.

The kernel gives a column to a quad of lanes (lane t keeps rows k = t mod 4, 16 values in a register window
a[0..15]) and runs the 64 steps of a forward substitution as 16 iterations of ONE loop body.  For that the window
always starts at the current row group — after a group its finished value leaves and the window shifts by one —
and the shared-memory image of R is stored pre-shifted the same way:

    Rq[(i * 4 + t) * 18 + kk] = R(i, k),  k = 4 (i / 4 + kk) + t,  if i < k < h (k <= 63), else 0.

The zeros do the masking: a finished value (k <= i) and the padding rows (k >= h) see only 0 * x_i.  This file
replays exactly that bookkeeping (the staging loop's thread -> element map included) and checks the result
against X = D R^-T A, for full and ragged pieces.
"""
import numpy as np
import pytest

SUB, STRIDE, THREADS = 64, 18, 256


def stage_r_image(R, h):
    """The kernel's staging loop: thread tid, pass u -> element e = tid + 256 u of the 64 x 64 image."""
    img = np.full(SUB * 4 * STRIDE, np.nan)        # NaN = never written; the padding columns 16, 17 stay unread
    reads = []
    for u in range(16):
        for tid in range(THREADS):
            e = tid + THREADS * u
            i, rem = e >> 6, e & 63
            k = 4 * (i >> 2) + rem
            ok = k < SUB and i < k < h
            if ok:
                reads.append((i, k))
            img[(i * 4 + (rem & 3)) * STRIDE + (rem >> 2)] = R[i, k] if ok else 0.0
    return img, reads


def quad_solve(R, A, d, h):
    img, _ = stage_r_image(R, h)
    rinv = np.array([1.0 / R[i, i] if i < h else 1.0 for i in range(SUB)])
    ds = np.array([d[i] if i < h else 1.0 for i in range(SUB)])
    out = np.zeros_like(A)
    for c in range(A.shape[1]):
        a = np.zeros((4, 16))                       # a[t] = the register window of lane t of the quad
        for t in range(4):
            for kk in range(16):
                k = 4 * kk + t
                a[t, kk] = A[k, c] if k < h else 0.0
        for q in range(16):
            for o in range(4):
                i = 4 * q + o
                xi = a[o, 0] * rinv[i]              # lane o's product is the one the quad shuffle broadcasts
                a[o, 0] = xi
                for t in range(4):
                    row = img[(i * 4 + t) * STRIDE:(i * 4 + t) * STRIDE + 16]
                    assert not np.any(np.isnan(row))
                    a[t, :] = a[t, :] - row * xi
            for t in range(4):
                k = 4 * q + t
                if k < h:
                    out[k, c] = a[t, 0] * ds[k]
                a[t, :15] = a[t, 1:]
                a[t, 15] = 0.0
    return out


@pytest.mark.parametrize("h", [64, 63, 50, 33, 17, 4, 1])
def test_quad_solve_is_the_triangular_solve(h):
    rng = np.random.default_rng(h)
    R = np.zeros((SUB, SUB))
    R[:h, :h] = np.triu(rng.standard_normal((h, h)) * 0.1) + 2.0 * np.eye(h)
    d = np.where(rng.random(h) < 0.3, -1.0, 1.0)
    A = rng.standard_normal((h, 7))
    X = quad_solve(R, A, d, h)
    ref = np.diag(d) @ np.linalg.solve(R[:h, :h].T, A)
    assert np.max(np.abs(X - ref)) <= 1e-13 * max(1.0, np.max(np.abs(ref)))


def test_image_holds_every_strict_upper_entry_once_and_zeros_elsewhere():
    rng = np.random.default_rng(0)
    for h in (64, 40):
        R = np.triu(rng.standard_normal((SUB, SUB))) + 3.0 * np.eye(SUB)
        img, reads = stage_r_image(R, h)
        assert sorted(reads) == [(i, k) for i in range(h) for k in range(i + 1, h)]   # each entry read exactly once
        for i in range(SUB):
            for t in range(4):
                for kk in range(16):
                    k = 4 * (i // 4 + kk) + t
                    want = R[i, k] if (k < SUB and i < k < h) else 0.0
                    assert img[(i * 4 + t) * STRIDE + kk] == want


def test_bank_layout_of_the_quads_16_byte_loads():
    """The four lanes of a quad read 16 bytes each at (i * 4 + t) * 18 + kk: with the stride of 18 doubles the four
    addresses of one request fall into four different groups of 4 banks (a stride of 16 would put them on one)."""
    for i in (0, 5, 63):
        for kk in (0, 2, 14):
            banks = {(((i * 4 + t) * STRIDE + kk) * 2) % 32 // 4 for t in range(4)}
            assert len(banks) == 4
            clash = {(((i * 4 + t) * 16 + kk) * 2) % 32 // 4 for t in range(4)}
            assert len(clash) == 1
