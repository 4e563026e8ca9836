"""CPU model of the order in which factor_block_row (csrc/ms_api.cu) touches the rows of one block row.

Two things about that order are design decisions of this library and not the reference's:

* the INTERNAL blocking — on one GPU a caller's m = 64 / 128 is walked in block rows of 256 —, and
* the LEFT-LOOKING order of the 64-row pieces inside a block row (a piece receives the pieces above it in
  one update right before it is factored).

Neither may change what the reference computes: R and d (beyond rounding) and, above all, the VERDICT of the
per-block-row pivot pre-test (src/solver.cpp:117), which the reference takes on the m diagonal entries of a
block row after the block rows above it have been applied and before its own elimination starts — so which
rows form a block row is visible in the verdict.  The model below replays the C++ schedule step by step in
numpy (same piece table, same (rows, K, k-offset) of every in-block update, same points at which the pre-test
reads the diagonal) and is compared with the oracle run at the caller's m.
"""
import numpy as np
import pytest

import oracle_lib as ol

SUB = 64


def model_factor(a_upper, m_api, m_int, left, norma):
    """Returns (R, d, failed).  Works on a full symmetric copy; only j >= i is meaningful at the end."""
    n = a_upper.shape[0]
    A = np.triu(a_upper) + np.triu(a_upper, 1).T
    d = np.zeros(n)
    thresh = 1e-14 * norma
    assert m_int % m_api == 0 and (m_int == m_api or m_api % SUB == 0)

    def apply(rows, ks):
        # rows (global indices) -= P[ks]^T D[ks] P[ks], P = factored rows `ks` of R (the panel buffer)
        if len(ks) == 0 or len(rows) == 0:
            return
        P = A[ks, :]
        A[np.ix_(rows, np.arange(n))] -= (P[:, rows] * d[ks, None]).T @ P

    for r0 in range(0, n, m_int):
        r1 = min(r0 + m_int, n)
        for p0 in range(r0, r1, SUB):
            p1 = min(p0 + SUB, r1)
            k0 = p0 - r0
            sub_first = k0 % m_api == 0
            sub_rows = min(m_api, r1 - p0)
            if left and k0 > 0:
                ksub = k0 // m_api * m_api
                if sub_first:
                    apply(np.arange(p0, p0 + sub_rows), np.arange(r0, p0))
                elif k0 > ksub:
                    apply(np.arange(p0, p1), np.arange(r0 + ksub, p0))
            if sub_first:
                # the reference's pre-test of one of ITS block rows
                if np.any(np.abs(np.diag(A)[p0:p0 + sub_rows]) < thresh):
                    return A, d, True
            # diag piece + TRSM, row by row (what K1 + K3 compute)
            for i in range(p0, p1):
                piv = A[i, i]
                if abs(piv) < thresh:
                    return A, d, True
                d[i] = 1.0 if piv > 0 else -1.0
                rii = np.sqrt(abs(piv))
                A[i, i:] = A[i, i:] / rii * d[i]
                A[i, i] = rii
                for k in range(i + 1, p1):
                    A[k, k:] -= A[i, k] * d[i] * A[i, k:]
            if not left:
                apply(np.arange(p1, r1), np.arange(p0, p1))
        apply(np.arange(r1, n), np.arange(r0, r1))
    return np.triu(A), d, False


def pretest_matrix(n=256, row=69):
    """A[row, row] = 0 in the input and -1 once row 0 has been eliminated (d_row = -1)."""
    rng = np.random.default_rng(21)
    R = np.eye(n) + np.triu(rng.standard_normal((n, n)) * 0.01, 1)
    R[:row, row] = 0.0
    R[0, row] = 1.0
    d = np.ones(n)
    d[row] = -1.0
    A = R.T @ np.diag(d) @ R
    A[row, row] = 0.0
    return np.triu(A), R, d


@pytest.fixture(scope="module")
def ora():
    return ol.Oracle()


@pytest.mark.parametrize("left", [False, True])
@pytest.mark.parametrize("m_api,m_int", [(64, 64), (64, 256), (128, 128), (128, 256), (256, 256), (48, 48)])
def test_verdict_and_factor_follow_the_callers_block_rows(ora, m_api, m_int, left):
    a, R, d = pretest_matrix()
    _, norm = ora.b_norm(a)
    r_ref, d_ref, f_ref = ora.factor(a, m_api, norm)
    r, dd, f = model_factor(a, m_api, m_int, left, norm)
    # row 69 lies in the caller's first block row for m = 128 / 256 (pre-test sees the 0) and in the second one
    # for m = 64 / 48 (pre-test sees -1)
    assert bool(f_ref) == (m_api in (128, 256))
    assert f == bool(f_ref)
    if not f:
        assert np.array_equal(dd, d_ref) and dd[69] == -1.0
        assert np.max(np.abs(np.triu(r) - np.triu(r_ref))) <= 1e-12


@pytest.mark.parametrize("left", [False, True])
@pytest.mark.parametrize("n,m_api,m_int", [(300, 64, 256), (520, 128, 256), (200, 64, 64), (333, 128, 128)])
def test_indefinite_factor_independent_of_the_internal_order(ora, n, m_api, m_int, left):
    rng = np.random.default_rng(n)
    sign = np.where(np.arange(n) % 5 == 3, -1.0, 1.0)
    L = np.tril(rng.standard_normal((n, n)) * 0.1, -1) + np.eye(n)
    mat = L @ np.diag(sign * (1 + rng.random(n))) @ L.T
    a = np.triu(mat)
    _, norm = ora.b_norm(a)
    r_ref, d_ref, f_ref = ora.factor(a, m_api, norm)
    r, dd, f = model_factor(a, m_api, m_int, left, norm)
    assert not f_ref and not f
    assert np.array_equal(dd, d_ref) and np.array_equal(dd, sign)
    assert np.max(np.abs(np.triu(r) - np.triu(r_ref))) <= 1e-10 * np.max(np.abs(r_ref))


def test_a_pretest_taken_per_internal_block_row_would_be_wrong(ora):
    """What the sub-block bookkeeping is for: testing all 256 rows of the internal block row up front
    (= the reference run with m = 256) fails on a matrix the reference accepts with m = 64."""
    a, _, _ = pretest_matrix()
    _, norm = ora.b_norm(a)
    assert not ora.factor(a, 64, norm)[2]
    assert model_factor(a, 256, 256, True, norm)[2]         # pre-test over 256 rows: sees the 0
    assert not model_factor(a, 64, 256, True, norm)[2]      # per caller block row: sees -1
