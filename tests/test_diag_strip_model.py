"""CPU model of the schedule of diag_strip_kernel (csrc/ms_kernels_diag.cu).

The kernel factors a 64 x 64 diagonal piece in strips of DIAG_NB = 4 pivots.  The pivot warp factors strip b out
of registers while the other seven warps apply strip b-1's finished rows — first to the rows of strip b ("S1",
the only thing the pivot warp waits for), then to every row below it ("S2").  The claim the exact variant rests
on: each entry still receives its products (r_ki * r_kj) * d_k in ascending k, one subtraction each, so the
result is BIT-IDENTICAL to the reference's cholesky_decomp_U (src/solver.cpp:106-160), whatever the strip
width.  This file replays the three phases in numpy (IEEE double, no FMA) in the kernel's order and compares
with the oracle bit for bit; the fast variant (rsqrt + multiply, FMAs on the pivot chain) is modelled too and
must agree to a few ulp.
"""
import numpy as np
import pytest

import oracle_lib as ol

NB = 4


def strip_factor_model(blk, thresh, fast=False, nb=NB):
    a = np.triu(np.array(blk, dtype=np.float64))
    h = a.shape[0]
    d = np.zeros(h)
    rdiag = np.zeros(h)
    failed = False

    def apply_prev(rows_lo, rows_hi, i0):
        # strip [i0 - nb, i0) (final) into rows [rows_lo, rows_hi), columns j >= row, pivots ascending
        for k in range(rows_lo, min(rows_hi, h)):
            for i in range(i0 - nb, i0):
                a[k, k:] = a[k, k:] - (a[i, k] * a[i, k:]) * d[i]

    nstrips = (h + nb - 1) // nb
    for b in range(nstrips):
        i0 = nb * b
        if b > 0:
            apply_prev(i0, i0 + nb, i0)                      # S1
        # pivot warp: the strip out of registers
        for t in range(nb):
            row = i0 + t
            if row >= h:
                break
            s = a[row, row]
            di = 1.0 if s > 0 else -1.0
            if fast:
                y = 1.0 / np.sqrt(abs(s)) if s != 0 else np.inf
                rii = abs(s) * y if s != 0 else 0.0
                a[row, row + 1:] = a[row, row + 1:] * (y * di)
            else:
                rii = np.sqrt(abs(s))
                a[row, row + 1:] = a[row, row + 1:] / (rii * di)
            d[row], rdiag[row] = di, rii
            if abs(rii) < thresh:
                failed = True
            for u in range(t + 1, nb):
                k = i0 + u
                if k >= h:
                    break
                if fast:
                    a[k, k:] = a[k, k:] + (-(a[row, k] * di)) * a[row, k:]      # one FMA in the kernel
                else:
                    a[k, k:] = a[k, k:] - (a[row, k] * a[row, k:]) * di
        if b > 0:
            apply_prev(i0 + nb, h, i0)                       # S2 (runs beside the pivot warp; disjoint rows)
    # the last strip's rows never feed anything below them
    for i in range(h):
        a[i, i] = rdiag[i]
    return a, d, failed


@pytest.fixture(scope="module")
def ora():
    return ol.Oracle()


def _matrices(h, seed):
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((h, h))
    spd = g @ g.T / h + np.eye(h)
    sign = np.where(np.arange(h) % 5 == 3, -1.0, 1.0)
    L = np.tril(rng.standard_normal((h, h)) * 0.1, -1) + np.eye(h)
    indef = L @ np.diag(sign * (1 + rng.random(h))) @ L.T
    minmat = np.fromfunction(lambda i, j: h - np.maximum(i, j), (h, h))
    return [spd, indef, minmat]


@pytest.mark.parametrize("h", [64, 61, 33, 8, 5])
@pytest.mark.parametrize("nb", [4, 2, 1])
def test_strip_schedule_is_bit_identical_to_the_reference_order(ora, h, nb):
    for mat in _matrices(h, 100 + h):
        norma = float(np.max(np.sum(np.abs(mat), axis=1)))
        ref, d_ref, fail_ref = ora.diag_factor(np.triu(mat), norma)
        got, d, fail = strip_factor_model(mat, 1e-14 * norma, fast=False, nb=nb)
        assert not fail_ref and not fail
        assert np.array_equal(d, d_ref)
        assert np.array_equal(np.triu(got), np.triu(ref))


def test_fast_pivots_agree_to_a_few_ulp(ora):
    for mat in _matrices(64, 7):
        norma = float(np.max(np.sum(np.abs(mat), axis=1)))
        ref, d_ref, _ = ora.diag_factor(np.triu(mat), norma)
        got, d, fail = strip_factor_model(mat, 1e-14 * norma, fast=True)
        assert not fail and np.array_equal(d, d_ref)
        scale = np.max(np.abs(ref))
        assert np.max(np.abs(np.triu(got) - np.triu(ref))) <= 64 * 2.0 ** -52 * scale * np.linalg.cond(mat) ** 0.5


def test_zero_pivot_fails_in_both_variants():
    mat = np.array([[1.0, 1.0, 0.0], [1.0, 1.0, 0.0], [0.0, 0.0, 1.0]])   # second pivot: 1 - 1 = 0
    for fast in (False, True):
        with np.errstate(all="ignore"):            # rows behind the zero pivot fill with inf / NaN, as on the device
            _, _, fail = strip_factor_model(mat, 1e-14 * 2.0, fast=fast)
        assert fail
