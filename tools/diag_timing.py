import ctypes, os, sys, subprocess
import numpy as np
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import matrix_solver_b200 as ms
# the block API leaves the phase clocks in the private ctx flags; read them via a tiny debug hook
L = ms.load()
L.ms_debug_blk_flags.restype = ctypes.c_int
L.ms_debug_blk_flags.argtypes = [ctypes.POINTER(ctypes.c_int)]
rng=np.random.default_rng(0)
for w in (16, 32, 64, 96):
    g=rng.standard_normal((w,w)); spd=g@g.T+w*np.eye(w)
    for rep in range(3):
        ms.block_diag_factor(np.triu(spd), np.abs(spd).sum(1).max())
    buf=(ctypes.c_int*8)()
    L.ms_debug_blk_flags(buf)
    print(w, "cycles load/factor/inverse/writeback:", list(buf)[4:8], "pivot-warp busy / update-warp busy:", list(buf)[2:4])
