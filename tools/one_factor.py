#!/usr/bin/env python
"""One generate + factor (+ optional solves) — the short command ncu captures are taken on."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import matrix_solver_b200 as ms  # noqa: E402

n, m = int(sys.argv[1]), int(sys.argv[2])
solve = len(sys.argv) > 3 and sys.argv[3] == "solve"
with ms.Solver(n, m) as s:
    s.set_option("lookahead", int(os.environ.get("MS_LOOKAHEAD", "1")))
    s.init_formula(0)
    norm = s.b_and_norm()
    fail, el = s.factor(norm)
    if solve:
        s.solve_forward(norm)
        s.solve_backward(norm)
    print(f"n={n} m={m} fail={fail} factor_s={el:.4f} tflops={n**3/3/el/1e12:.2f}")
