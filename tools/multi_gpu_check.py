#!/usr/bin/env python
"""Multi-GPU parity check, launched by torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 \
        --master-port 29511 tools/multi_gpu_check.py
Every rank runs the device pipeline on its block rows; rank 0 compares with the oracle / analytic
answers and prints one JSON line per case."""
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
from matrix_solver_b200 import dist  # noqa: E402
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
        out["wall"] = time.time() - t0
        dist.barrier()
        if rank == 0:
            print(json.dumps(out), flush=True)
    dist.shutdown()


if __name__ == "__main__":
    main()
