#!/usr/bin/env python
"""Device time of one diag-kernel launch on a 64 x 64 block, per variant (debug hook ms_debug_time_diag)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import matrix_solver_b200 as ms  # noqa: E402

L = ms.load()
L.ms_debug_time_diag.restype = ctypes.c_int
L.ms_debug_time_diag.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double),
                                  ctypes.POINTER(ctypes.c_longlong)]
for h in (16, 32, 64):
    for variant, name in ((0, "diag_kernel<5,4>"), (1, "diag_strip_kernel"), (2, "strip, pivot warp solo"), (3, "strip, fast pivots")):
        us = ctypes.c_double(0)
        cyc = (ctypes.c_longlong * 40)()
        rc = L.ms_debug_time_diag(variant, h, 20, ctypes.byref(us), cyc)
        print(f"h={h:3d} {name:20s} rc={rc} {us.value:8.2f} us per launch (events around the launch, idle GPU: clocks "
              f"may be low); cycles load/strips/writeback/pivot-wait/pivot-factor = {list(cyc)[:5]}")
