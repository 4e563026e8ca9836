"""matrix_solver_b200 — B200-native blocked R^T D R (Cholesky) factor + solve.

Python here is plumbing only: a ctypes binding of the C ABI (include/ms_b200.h) used by the
tests and by bench.py.  The product is the sm_100a library under csrc/ and the C++ host
mirror of the reference's entry points under host/.  There is no CPU fallback: importing
works anywhere, but every compute call raises unless the CUDA library is built and a B200
is visible.
"""
from .capi import (  # noqa: F401
    MSError,
    NumericalFailure,
    PinnedBuffer,
    Solver,
    block_diag_factor,
    block_panel,
    block_trsm_panel,
    block_update,
    block_update_diag,
    block_upper_inverse,
    device_count,
    error_vs_planted,
    get_unique_id,
    lib_path,
    load,
    norm2,
    probe_dmma_peak,
    probe_hbm_copy,
)
