#!/usr/bin/env python
"""Per-step timeline of one factorisation (MS_TRACE=1).  Single GPU: python tools/trace_factor.py n m
Multi GPU: torchrun ... tools/trace_factor.py n m   (rank 0 and the last rank print)."""
import ctypes
import os
import sys

os.environ["MS_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import matrix_solver_b200 as ms  # noqa: E402
from matrix_solver_b200 import dist  # noqa: E402

n, m = int(sys.argv[1]), int(sys.argv[2])
rank, local, world = dist.env_rank()
torch.cuda.set_device(local)
dist.init("nccl" if world > 1 else None)
uid = None
if world > 1:
    uid = dist.broadcast_bytes(ms.get_unique_id() if rank == 0 else b"", 128)
L = ms.load()
L.ms_debug_trace.restype = ctypes.c_int
fp = ctypes.POINTER(ctypes.c_float)
L.ms_debug_trace.argtypes = [ctypes.c_void_p, fp, fp, fp, ctypes.c_size_t]
with ms.Solver(n, m, device=local, rank=rank, nranks=world, unique_id=uid) as s:
    for rep in range(2):
        s.init_formula(0)
        norm = s.b_and_norm()
        fail, el = s.factor(norm)
    nb = (n + m - 1) // m
    a, b, c = (np.zeros(nb, dtype=np.float32) for _ in range(3))
    # (a single GPU may walk block rows of 256 for m = 64 / 128: the trace has one entry per INTERNAL block row)
    nb = L.ms_debug_trace(s.h, a.ctypes.data_as(fp), b.ctypes.data_as(fp), c.ctypes.data_as(fp), nb)
    if rank in (0, world - 1):
        print(f"rank {rank}/{world} n={n} m={m} factor={el*1e3:.1f} ms")
        print(" step  panel_ready  main_done  strip_ready   d_panel  d_main   wait(main_done - next panel)")
        for I in range(0, nb, max(1, nb // 32)):
            dp = a[I] - a[I - 1] if I else a[I]
            dm = b[I] - b[I - 1] if I and b[I] >= 0 else b[I]
            print(f"{I:5d} {a[I]:11.3f} {b[I]:10.3f} {c[I]:11.3f} {dp:9.3f} {dm:8.3f}")
        # how often does the main stream wait for a panel?  (panel I ready after main finished I-1)
        waits = [(I, a[I] - b[I - 1]) for I in range(1, nb - 1) if a[I] > b[I - 1]]
        print("steps where main waited for the panel:", len(waits), "total wait ms", sum(w for _, w in waits))
dist.shutdown()
