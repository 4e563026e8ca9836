"""Self-check of a finished factorisation of the formula matrix A_ij = n - max(i,j) against its
ANALYTIC factor — the known answer that exists at every n, including the sizes
no CPU oracle can reach (BASELINE configs[3], [4]).  Used by bench.py (the `parity` block, every N) and
tools/multi_gpu_check.py; no oracle and no reference code involved, only closed forms:
    ||A||_inf = n(n+1)/2,  d = 1,  R_0j = (n-j)/sqrt(n),  R_ij = (n-j)/sqrt((n-i)(n-i+1)),  x* = (1,0,1,0,...).
"""
import numpy as np

EPS = 2.0 ** -52


def parity_rows(n, m, world):
    """Rows of R to compare with the analytic factor: both sides of block-row boundaries, of the
    boundaries between ranks' turns in the cyclic map, of 64-row piece boundaries, the whole range
    evenly, and the last block row."""
    nb = (n + m - 1) // m
    rows = {0, 1, 63, 64, 65, m - 1, m, m + 1, n - 1, n - 2, (nb - 1) * m, (nb - 1) * m + 1, (nb - 1) * m - 1}
    for I in (1, 2, world - 1, world, world + 1, nb // 2, nb // 2 + 1, nb - world, nb - 2):
        for off in (-1, 0, 1, 64):
            rows.add(I * m + off)
    rows.update(int(t) for t in [k * (n - 1) / 47.0 for k in range(48)])
    return sorted(r for r in rows if 0 <= r < n)


def analytic_row(n, i):
    """Row i (columns i..n-1) of the factor of A_ij = n - max(i,j): R_0j = (n-j)/sqrt(n),
    R_ij = (n-j)/sqrt((n-i)(n-i+1))."""
    j = np.arange(i, n, dtype=np.float64)
    den = np.sqrt(float(n)) if i == 0 else np.sqrt(float(n - i) * float(n - i + 1))
    return (n - j) / den


def check_parity(solver, n, m, world, norm, any_fail):
    """Runs on every rank (ms_get_row is collective); the matrix still holds the last factor."""
    cond = 1.62 * n * n
    tol_stated = 10.0 * n * EPS * cond              # north star: c * n * eps * cond(A), c = 10
    tol_r = min(tol_stated, 1e-6)                   # the stated bound is vacuous at n = 65536 (~1)
    worst, worst_row = 0.0, -1
    rows = parity_rows(n, m, world)
    for i in rows:
        got = solver.get_row(i)[i:]
        want = analytic_row(n, i)
        err = float(np.max(np.abs(got - want)) / np.max(np.abs(want)))
        if not np.isfinite(err):
            err = float("inf")
        if err > worst:
            worst, worst_row = err, i
    d = solver.get_vec("d")
    x = solver.get_vec("x")
    xs = (np.arange(n) % 2 == 0).astype(np.float64)
    x_err = float(np.linalg.norm(x - xs))
    x_rel = x_err / float(np.linalg.norm(xs))
    out = {
        "norm_exact": bool(norm == n * (n + 1) / 2.0),
        "d_all_plus_one": bool(np.all(d == 1.0)),
        "r_rows_checked": len(rows), "r_max_rel": worst, "r_worst_row": worst_row,
        "r_tol": tol_r, "tol_stated_10_n_eps_cond": tol_stated,
        "x_err_vs_planted": x_err, "x_rel": x_rel, "x_tol_rel": min(tol_stated, 1e-4),
        "numerical_failure": bool(any_fail),
    }
    out["ok"] = bool(out["norm_exact"] and out["d_all_plus_one"] and worst <= tol_r
                     and x_rel <= out["x_tol_rel"] and not any_fail)
    return out
