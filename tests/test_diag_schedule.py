"""CPU model of the blocked schedule of diag_kernel_blocked (csrc/ms_kernels_diag.cu, the experimental
MS_DIAG_VARIANT=2): strips of 8 pivots factored by one warp out of registers, the strip after it
updated by that same warp, everything further down by the other warps one block later.  The model
executes the kernel's operations in the kernel's order with IEEE doubles (numpy scalars: every
multiply, subtract, divide and sqrt is rounded separately, like the _rn intrinsics) and must reproduce
the oracle's cholesky_decomp_U restatement (hence the reference, src/solver.cpp:106-160) bit for bit.
"""
import numpy as np
import pytest

import oracle_lib as ol

NB = 8
f64 = np.float64


def blocked_factor(a_in, thresh):
    h = a_in.shape[0]
    a = np.triu(a_in).astype(f64).copy()
    rdiag, dd = np.zeros(h), np.zeros(h)
    fail = False
    nblk = (h + NB - 1) // NB

    def apply_block(i0p, rows):
        # pivots i0p .. i0p+7 (final rows of `a`) applied, in ascending order, to the given rows
        for k in rows:
            for j in range(k, h):
                c = a[k, j]
                for t in range(NB):
                    i = i0p + t
                    if i >= h:
                        break
                    c = f64(c - f64(f64(a[i, k] * a[i, j]) * dd[i]))
                a[k, j] = c

    for b in range(nblk):
        i0 = NB * b
        # "warps 1-7": the block before this one goes into every row below the current strip
        if b > 0:
            apply_block(i0 - NB, range(i0 + NB, h))
        # "warp 0": the strip, first brought up to date with the previous block, then factored in registers
        strip = range(i0, min(i0 + NB, h))
        if b > 0:
            apply_block(i0 - NB, strip)
        for row in strip:
            s = a[row, row]
            di = f64(1.0) if s > 0 else f64(-1.0)
            rii = np.sqrt(np.abs(s))
            den = f64(rii * di)
            rdiag[row], dd[row] = rii, di
            fail |= bool(abs(rii) < thresh)
            for j in range(row + 1, h):
                a[row, j] = f64(a[row, j] / den)
            for k in range(row + 1, min(i0 + NB, h)):          # the rest of the strip
                rk = a[row, k]
                for j in range(k, h):
                    a[k, j] = f64(a[k, j] - f64(f64(rk * a[row, j]) * di))
    r = np.triu(a, 1) + np.diag(rdiag)
    return r, dd, fail


@pytest.mark.parametrize("h", [1, 2, 3, 7, 8, 9, 16, 17, 31, 40, 63, 64])
def test_blocked_schedule_is_bit_identical_to_the_reference_order(h):
    ora = ol.Oracle()
    rng = np.random.default_rng(100 + h)
    with np.errstate(all="ignore"):
        for trial in range(3):
            g = rng.standard_normal((h, h))
            spd = g @ g.T + h * np.eye(h)
            if trial == 1 and h > 2:
                spd[h // 2:, h // 2:] *= -1.0          # indefinite: d = -1 entries
            if trial == 2:
                spd = spd * 1e-3 + np.outer(np.arange(h), np.arange(h)) * 1e-2
            blk = np.triu(spd)
            norma = np.abs(spd).sum(1).max()
            r0, d0, f0 = ora.diag_factor(blk, norma)
            r1, d1, f1 = blocked_factor(blk, 1e-14 * norma)
            assert f0 == f1
            assert np.array_equal(d0, d1)
            assert np.array_equal(np.triu(r0), r1, equal_nan=True)
