#!/usr/bin/env python
"""Multi-GPU parity check, launched by torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 \
        --master-port 29511 tools/multi_gpu_check.py
Every rank runs the device pipeline on its block rows; rank 0 compares with the oracle / analytic
answers, prints one JSON line per case and the process exits non-zero if any case is out of tolerance
(tests/test_gpu_multi.py runs this under torchrun when >= 2 devices are visible).
Sizes: MS_MULTI_QUICK=1 -> a third of the small cases + (8192, 256); MS_MULTI_BIG=1 adds (65536, 256)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import matrix_solver_b200 as ms  # noqa: E402
from matrix_solver_b200 import dist, selfcheck  # noqa: E402
import oracle_lib as ol  # noqa: E402


def rel(a, b):
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den else 1.0))


def main():
    rank, local, world = dist.env_rank()
    torch.cuda.set_device(local)
    dist.init("nccl")
    cases = [(4, 3), (30, 9), (100, 48), (130, 64), (200, 64), (512, 64), (1000, 64), (1024, 128), (2048, 256),
             (777, 192), (1500, 100)]
    big = [(8192, 256), (16384, 128), (32768, 256)]
    if os.environ.get("MS_MULTI_QUICK"):
        cases, big = cases[::3], [(8192, 256)]
    if os.environ.get("MS_MULTI_BIG"):
        big.append((65536, 256))
    ora = ol.Oracle()
    bad = []
    for n, m in cases + big:
        raw = ms.get_unique_id() if rank == 0 else b""
        uid = dist.broadcast_bytes(raw, 128)
        t0 = time.time()
        out = {"n": n, "m": m, "P": world}
        with ms.Solver(n, m, device=local, rank=rank, nranks=world, unique_id=uid) as s:
            s.init_formula(0)
            norm = s.b_and_norm()
            b = s.get_vec("b")
            small = n <= 2048
            A = s.get_upper() if small else None
            fail, el = s.factor(norm)
            R = s.get_upper() if small else None
            d = s.get_vec("d")
            f2 = s.solve_forward(norm)
            f3 = s.solve_backward(norm)
            y, x = s.get_vec("y"), s.get_vec("x")
            par = selfcheck.check_parity(s, n, m, world, norm, fail or f2 or f3) if n >= 1024 else None
            s.init_formula(0)
            res, err = s.residual()
            st = s.stats()
            out.update(norm_ok=norm == n * (n + 1) / 2, fail=[fail, f2, f3], factor_ms=st["factor_ms"],
                       tflops=n ** 3 / 3 / (st["factor_ms"] * 1e-3) / 1e12, fwd_ms=st["forward_ms"],
                       bwd_ms=st["backward_ms"], residual=res, error=err, d_ok=bool(np.all(d == 1.0)))
            if small:
                o = ora.solve_pipeline(n, m)
                out.update(A_eq=bool(np.array_equal(np.triu(A), np.triu(o["A"]))), b_eq=bool(np.array_equal(b, o["b"])),
                           R_rel=rel(np.triu(R), np.triu(o["R"])), y_rel=rel(y, o["y"]), x_rel=rel(x, o["x"]),
                           ora_res=o["residual"], ora_err=o["error"])
            # verdicts: the north star's tolerance c*n*eps*cond (c = 10, cond ~ 1.62 n^2), bit equality for
            # the integer data, the residual no worse than the oracle's / 1e-13 relative where no oracle runs
            tol = 10.0 * n * 2.0 ** -52 * 1.62 * n * n
            ok = out["norm_ok"] and not any(out["fail"]) and out["d_ok"]
            if small:
                ok = ok and out["A_eq"] and out["b_eq"] and max(out["R_rel"], out["y_rel"], out["x_rel"]) <= tol
                ok = ok and res <= max(2 * o["residual"], 8 * n * 2.0 ** -52 * float(np.linalg.norm(b)))
            else:
                ok = ok and res / float(np.linalg.norm(b)) <= 1e-13
            if par is not None:
                out.update(r_max_rel=par["r_max_rel"], r_rows=par["r_rows_checked"], x_rel_planted=par["x_rel"])
                ok = ok and par["ok"]
            out["forward_fused"] = st["forward_fused"]
            out["ok"] = bool(ok)
        out["wall"] = time.time() - t0
        dist.barrier()
        if rank == 0:
            print(json.dumps(out), flush=True)
            if not out["ok"]:
                bad.append((n, m))
    nbad = int(dist.max_over_ranks(len(bad)))
    dist.shutdown()
    if nbad:
        print("multi_gpu_check: out of tolerance:", bad, file=sys.stderr)
        sys.exit(1)


if __name__ == "__main__":
    main()
