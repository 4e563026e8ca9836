"""Pins oracle/ms_oracle.c (the CPU restatement) to the reference.

Two anchors:
  * the committed goldens under tests/golden/ (generated from the unmodified reference by
    tests/golden/make_golden.py) — always available;
  * the reference itself, live (oracle/_ref built from /root/reference) — when present.
The restatement follows the reference's summation order, so with both compiled at plain
-O2 the agreement demanded here is BIT-EXACT (np.array_equal), not a tolerance.
"""
import json
import os

import numpy as np
import pytest

import oracle_lib as ol

ORA = ol.Oracle()
DUMPS = np.load(os.path.join(ol.GOLDEN_DIR, "ref_dumps.npz"))
CASES = sorted({k.split("/")[0] for k in DUMPS.files})
with open(os.path.join(ol.GOLDEN_DIR, "ref_summary.json")) as f:
    SUMMARY = json.load(f)

needs_ref = pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (no /root/reference)")


def golden(case):
    return {k.split("/", 1)[1]: DUMPS[k] for k in DUMPS.files if k.startswith(case + "/")}


def fixture_path(case):
    stem = case.split("_m")[0]
    return None if stem.startswith("formula") else os.path.join(ol.FIXTURE_DIR, stem + ".txt")


def run_case(case):
    g = golden(case)
    n, m = int(g["n"]), int(g["m"])
    return g, ORA.solve_pipeline(n, m, fixture_path(case))


@pytest.mark.parametrize("case", CASES)
def test_pipeline_matches_reference_golden_bitwise(case):
    g, o = run_case(case)
    assert o["init_ok"] == bool(g["init_ok"])
    assert np.array_equal(o["A"], g["A"])
    assert o["norm"] == float(g["norm"])
    assert np.array_equal(o["b"], g["b"])
    assert o["factor_fail"] == bool(g["factor_fail"])
    if o["factor_fail"]:
        return
    assert np.array_equal(np.triu(o["R"]), g["R"])
    assert np.array_equal(o["d"], g["d"])
    assert np.array_equal(o["y"], g["y"])
    assert np.array_equal(o["x"], g["x"])
    assert np.array_equal(o["r"], g["r"])
    assert o["residual"] == float(g["residual"])
    assert o["error"] == float(g["error"])


def test_fixture_outcomes_are_the_published_ones():
    """SURVEY.md Appendix B.1 / BASELINE.md: the reference's own fixture results."""
    want = {"a_m3": (2.512148e-15, 1.125169e-13), "a20_m3": (4.169761e-35, 1.858181e-13),
            "e_m3": (4.965068e-16, 3.886940e-16), "f_m3": (1.335963e-14, 3.682652e-15)}
    for case, (res, err) in want.items():
        _, o = run_case(case)
        assert o["residual"] == pytest.approx(res, rel=1e-6)
        assert o["error"] == pytest.approx(err, rel=1e-6)
    for case in ("b_m3", "c_m3", "d_m3"):
        _, o = run_case(case)
        assert o["factor_fail"], case
    _, o = run_case("f_m3")
    assert o["d"].tolist() == [1.0, 1.0, -1.0, 1.0]
    # m > n, m == n, m == 1 all reduce to the same solution on a.txt
    for case in ("a_m1", "a_m4", "a_m7"):
        _, o = run_case(case)
        assert o["residual"] == pytest.approx(2.512148e-15, rel=1e-6)


@pytest.mark.parametrize("case", sorted(SUMMARY))
def test_large_formula_matches_reference_summary(case):
    s = SUMMARY[case]
    if s["n"] > 2048:
        pytest.skip("kept for the GPU parity test; too slow for the CPU suite")
    o = ORA.solve_pipeline(s["n"], s["m"])
    assert o["norm"] == s["norm"] == s["n"] * (s["n"] + 1) / 2
    assert o["residual"] == s["residual"] and o["error"] == s["error"]
    assert o["x"][:16].tolist() == s["x_head"]
    idx = np.array(s["sample_idx"])
    assert np.array_equal(o["R"][np.ix_(idx, idx)] * (idx[:, None] <= idx[None, :]),
                          np.array(s["R_samples"]))
    assert np.all(o["d"] == 1.0) == s["d_all_one"]


def test_formula_analytic_factor():
    """closed form of R for A_ij = n - max(i,j)."""
    n, m = 300, 32
    o = ORA.solve_pipeline(n, m)
    i = np.arange(n)[:, None].astype(float)
    j = np.arange(n)[None, :].astype(float)
    with np.errstate(divide="ignore", invalid="ignore"):
        want = np.where(i == 0, (n - j) / np.sqrt(n), (n - j) / np.sqrt((n - i) * (n - i + 1)))
    want = np.triu(want)
    assert np.max(np.abs(np.triu(o["R"]) - want)) < 1e-9
    assert o["norm"] == n * (n + 1) / 2


def test_oracle_against_numpy_independently():
    """Independent cross-check: for SPD input R^T R = A is the Cholesky factor
    numpy computes; for an indefinite matrix with non-zero leading minors R^T D R = A; the two
    triangular sweeps solve A x = b."""
    rng = np.random.default_rng(42)
    for n, m in [(7, 3), (40, 16), (65, 64), (130, 48)]:
        g = rng.standard_normal((n, n))
        spd = g @ g.T + n * np.eye(n)
        a = np.triu(spd)
        b, norm = ORA.b_norm(a)
        assert np.allclose(b, spd[:, ::2].sum(1)) and np.isclose(norm, np.abs(spd).sum(1).max())
        r, d, fail = ORA.factor(a, m, norm)
        assert not fail and np.all(d == 1.0)
        assert np.allclose(np.triu(r), np.linalg.cholesky(spd).T, rtol=1e-10, atol=1e-10)
        y, f2 = ORA.forward(r, b, norm)
        x, f3 = ORA.backward(r, y, d, norm)
        assert not f2 and not f3
        assert np.allclose(x, np.linalg.solve(spd, b), rtol=1e-9, atol=1e-9)
        # indefinite: flip the sign of a trailing principal block of the SPD matrix's LDL factors
        L = np.tril(rng.standard_normal((n, n)) * 0.1, -1) + np.eye(n)
        sign = np.where(np.arange(n) % 4 == 2, -1.0, 1.0)
        ind = L @ np.diag(sign * (1.0 + rng.random(n))) @ L.T
        ai = np.triu(ind)
        bi, normi = ORA.b_norm(ai)
        ri, di, faili = ORA.factor(ai, m, normi)
        assert not faili and np.array_equal(di, sign)
        Ru = np.triu(ri)
        assert np.allclose(Ru.T @ np.diag(di) @ Ru, ind, rtol=1e-10, atol=1e-10)
        yi, _ = ORA.forward(ri, bi, normi)
        xi, _ = ORA.backward(ri, yi, di, normi)
        assert np.allclose(xi, np.linalg.solve(ind, bi), rtol=1e-8, atol=1e-8)


def test_layout_restatement():
    """compute_alloc_size / get_bl / get_elU (src/init.cpp:4-58, solver.h:4-9, defines.h:12-18)."""
    for n, m in [(4, 3), (6, 3), (10, 7), (30, 9), (4, 7), (64, 16), (5, 1)]:
        assert ORA.alloc_size(n, m) == n * (n + 1) // 2
        for p in (2, 3, 5):
            assert sum(ORA.alloc_size(n, m, r, p) for r in range(p)) == n * (n + 1) // 2
        seen = set()
        for i in range(n):
            for j in range(i, n):
                seen.add(ORA.packed_offset(i, j, n, m))
        assert seen == set(range(n * (n + 1) // 2))
    for case in CASES:
        g = golden(case)
        assert ORA.alloc_size(int(g["n"]), int(g["m"])) == int(g["alloc"])


@needs_ref
@pytest.mark.needs_ref
def test_block_kernels_match_reference_live_bitwise():
    ref = ol.RefBlocks()
    rng = np.random.default_rng(7)
    for w in (1, 2, 3, 5, 8, 17, 32, 48):
        g = rng.standard_normal((w, w))
        spd = g @ g.T + w * np.eye(w)
        spd[w // 2:, w // 2:] *= -1 if w > 2 else 1  # make it indefinite: exercises d = -1
        blk = np.triu(spd)
        norma = np.abs(spd).sum(1).max()
        r0, d0, f0 = ref.diag_factor(blk, norma)
        r1, d1, f1 = ORA.diag_factor(blk, norma)
        assert f0 == f1 and np.array_equal(d0, d1) and np.array_equal(np.triu(r1), r0)
        i0, g0 = ref.upper_inverse(r0, norma)
        i1, g1 = ORA.upper_inverse(r0, norma)
        assert g0 == g1 and np.array_equal(i0, np.triu(i1))
        D = np.where(rng.random(w) < 0.3, -1.0, 1.0)
        for l in (w, max(1, w - 1), w + 3, 1):
            A = rng.standard_normal((w, l))
            R1 = rng.standard_normal((w, w))
            R2 = rng.standard_normal((w, l))
            assert np.array_equal(ref.update(A, R1, D, R2), ORA.update(A, R1, D, R2))
            assert np.array_equal(ref.panel(i0, A, D), ORA.panel(i0, A, D))
        # diagonal-block update: square (R1 is w x w) and tail (R1 is k x l) forms
        Aup = np.triu(rng.standard_normal((w, w)))
        R1 = rng.standard_normal((w, w))
        assert np.array_equal(ref.update_diag(Aup, R1, D), np.triu(ORA.update(Aup, R1, D, R1, True)))
        k = w + 2
        Dk = np.where(rng.random(k) < 0.3, -1.0, 1.0)
        R1 = rng.standard_normal((k, w))
        assert np.array_equal(ref.update_diag(Aup, R1, Dk), np.triu(ORA.update(Aup, R1, Dk, R1, True)))


@needs_ref
@pytest.mark.needs_ref
@pytest.mark.parametrize("n,m", [(7, 2), (33, 8), (50, 50), (61, 20), (129, 32), (257, 64)])
def test_pipeline_matches_reference_live_bitwise(n, m):
    g = ol.ref_dump(n, m)
    o = ORA.solve_pipeline(n, m)
    for k in ("b", "d", "y", "x", "r"):
        assert np.array_equal(o[k], g[k]), k
    assert np.array_equal(np.triu(o["R"]), g["R"])
    assert (o["norm"], o["residual"], o["error"]) == (g["norm"], g["residual"], g["error"])


@needs_ref
@pytest.mark.needs_ref
def test_reference_values_do_not_depend_on_rank_count():
    """one rank is a sufficient value oracle (also exercises the minimpi shim)."""
    one = ol.ref_dump(30, 9, np_ranks=1)
    for p in (2, 3, 5):
        many = ol.ref_dump(30, 9, np_ranks=p)
        for k in ("b", "d", "y", "x", "r", "R"):
            assert np.array_equal(one[k], many[k]), (p, k)


class _OracleBackedSolver:
    """What selfcheck.check_parity needs from a Solver, served from an oracle run."""

    def __init__(self, w):
        self.w = w

    def get_row(self, i):
        return np.triu(self.w["R"])[i].copy()

    def get_vec(self, which):
        return self.w[which].copy()


@pytest.mark.parametrize("n,m", [(200, 64), (777, 128), (1024, 256)])
def test_the_bench_self_check_accepts_the_oracle_and_rejects_a_damaged_factor(n, m):
    """bench.py's `parity` block (matrix_solver_b200/selfcheck.py) is the only factor check at the sizes no CPU run
    reaches, so the checker itself is checked here: fed with the oracle's R, d, x for the formula matrix it must
    pass, and it must fail when one entry of one checked row, one d or the norm is off."""
    from matrix_solver_b200 import selfcheck
    o = ol.Oracle()
    w = o.solve_pipeline(n, m)
    fake = _OracleBackedSolver(w)
    for world in (1, 2, 8):
        rows = selfcheck.parity_rows(n, m, world)
        assert rows == sorted(set(rows)) and rows[0] == 0 and rows[-1] == n - 1 and len(rows) >= 48
        for r in (m - 1, m, 63, 64):
            assert r >= n or r in rows          # both sides of block-row and piece boundaries
    res = selfcheck.check_parity(fake, n, m, 1, w["norm"], False)
    assert res["ok"] and res["norm_exact"] and res["d_all_plus_one"] and res["r_max_rel"] < 1e-9
    assert not selfcheck.check_parity(fake, n, m, 1, w["norm"], True)["ok"]            # a reported failure
    assert not selfcheck.check_parity(fake, n, m, 1, w["norm"] + 1.0, False)["ok"]     # wrong norm
    row = selfcheck.parity_rows(n, m, 1)[7]
    bad = dict(w)
    bad["R"] = w["R"].copy()
    bad["R"][row, row] *= 1.0 + 1e-5            # (the bound is on max|R - R_ref| / max|R_ref| per row)
    out = selfcheck.check_parity(_OracleBackedSolver(bad), n, m, 1, w["norm"], False)
    assert not out["ok"] and out["r_worst_row"] == row
    bad = dict(w)
    bad["d"] = w["d"].copy()
    bad["d"][n // 2] = -1.0
    assert not selfcheck.check_parity(_OracleBackedSolver(bad), n, m, 1, w["norm"], False)["ok"]
    bad = dict(w)
    bad["x"] = w["x"].copy()
    bad["x"][3] += 0.5
    assert not selfcheck.check_parity(_OracleBackedSolver(bad), n, m, 1, w["norm"], False)["ok"]
