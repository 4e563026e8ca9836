"""CPU-side checks: the C-ABI library loads and exports everything include/ms_b200.h declares,
fails loudly without a GPU (no CPU fallback), and the host-side layout logic matches the
reference's allocation rule.  No compute calls here."""
import ctypes
import os
import re
import subprocess

import pytest

import matrix_solver_b200 as ms
from matrix_solver_b200 import capi, layout
import oracle_lib as ol

ROOT = ol.ROOT
HEADER = os.path.join(ROOT, "include", "ms_b200.h")


def _ensure_lib():
    if not os.path.exists(ms.lib_path()):
        subprocess.run(["make", "-s", "lib"], cwd=ROOT, check=True)


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ms_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    _ensure_lib()
    lib = ctypes.CDLL(ms.lib_path())
    names = declared_symbols()
    assert len(names) >= 28
    for name in names:
        assert hasattr(lib, name), f"{name} declared in ms_b200.h but not exported"
    assert sorted(capi.EXPORTED) == names, "python binding and header disagree"


def test_no_cpu_fallback():
    _ensure_lib()
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert ms.device_count() == 0
    with pytest.raises(ms.MSError, match="no CPU fallback"):
        ms.Solver(64, 64)
    import numpy as np
    with pytest.raises(ms.MSError):
        ms.block_diag_factor(np.eye(4), 1.0)
    with pytest.raises(ms.MSError):
        ms.norm2(np.ones(4))


def test_product_code_never_touches_the_oracle():
    """The oracle is a checker: nothing under matrix_solver_b200/ may import, link or exec it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "matrix_solver_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"oracle_lib|liboracle|ms_oracle|oracle/_ref|ref_dump", txt):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_cli_usage_and_loud_failure(tmp_path):
    exe = os.path.join(ROOT, "matrix_solver_b200", "bin", "test_mat")
    if not os.path.exists(exe):
        subprocess.run(["make", "-s", "cli"], cwd=ROOT, check=True)
    # usage line, exit 0 (test/test_mat.cpp:40-44)
    for args in ([], ["0", "3"], ["4", "0"], ["4"], ["1", "2", "3", "4"]):
        p = subprocess.run([exe] + args, capture_output=True, text=True, cwd=tmp_path)
        assert p.returncode == 0
        assert p.stdout == f"Usage: {exe} n m <filename>\n"
    import torch
    if not torch.cuda.is_available():
        p = subprocess.run([exe, "4", "3"], capture_output=True, text=True, cwd=tmp_path)
        assert p.returncode != 0 and "no CPU fallback" in p.stderr
        assert p.stdout.startswith("n = 4, m = 3, file = \n")


@pytest.mark.parametrize("n,m", [(4, 3), (6, 3), (10, 7), (30, 9), (4, 7), (64, 16), (65536, 256), (1000, 48)])
def test_layout_matches_reference_allocation_rule(n, m):
    ora = ol.Oracle()
    assert layout.ref_alloc_size(n, m) == n * (n + 1) // 2 == ora.alloc_size(n, m)
    for p in (1, 2, 3, 8):
        assert sum(layout.rows_local(n, m, r, p) for r in range(p)) == n
        for r in range(p):
            assert layout.ref_alloc_size(n, m, r, p) == ora.alloc_size(n, m, r, p)
        # local rows are a bijection onto each rank's storage
        seen = {}
        step = max(1, n // 257)
        for g in range(0, n, step):
            r = layout.owner(g // m, p)
            lr = layout.local_row(n, m, g, p)
            assert lr < layout.rows_local(n, m, r, p)
            assert seen.setdefault((r, lr), g) == g


def test_update_flops_close_to_n_cubed_over_three():
    n, m = 8192, 256
    assert abs(layout.update_flops(n, m) / layout.factor_flops(n) - 1) < 0.06
    assert sum(layout.update_flops(n, m, r, 4) for r in range(4)) == pytest.approx(layout.update_flops(n, m))


def test_rank_runtime_forks_and_barriers(tmp_path):
    """The CLI's replacement for `mpirun -np P`: MS_NGPU=P forks P ranks before CUDA is touched
    (host/ms_rt.cpp).  Rank numbering, barriers and the join are exercised here without a GPU."""
    _ensure_lib()
    exe = tmp_path / "rt_ranks_test"
    host = os.path.join(ROOT, "matrix_solver_b200", "host")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + host, "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "rt_ranks_test.cpp"), os.path.join(host, "ms_rt.cpp"),
                    "-L" + os.path.dirname(ms.lib_path()), "-lms_b200",
                    "-Wl,-rpath," + os.path.dirname(ms.lib_path()), "-o", str(exe)], check=True)
    for p in (1, 3, 4):
        out = subprocess.run([str(exe)], env=dict(os.environ, MS_NGPU=str(p)), capture_output=True, text=True,
                             timeout=120)
        assert out.returncode == 0, out.stderr
        lines = out.stdout.strip().split("\n")
        assert lines[-1] == f"rank 0 done, size {p}"
        rounds = [ln for ln in lines if ln.startswith("round")]
        assert rounds == [f"round {r} by rank {r % p} of {p}" for r in range(50)]


def test_bench_reference_arm_contract(tmp_path):
    """`bench.py --impl reference`: the unmodified reference on the host cores (oracle/_ref) — or the
    oracle port when _ref was not shipped — reporting the same metric/unit as our arm, plus the
    cpu_baseline and e2e objects the contract asks for.  Small sample so that it runs in seconds."""
    import json
    import sys
    env = dict(os.environ, MS_BENCH_REF_N="1024")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().split("\n")[-1])
    assert line["impl"] == "reference"
    if "unavailable" in line:
        pytest.skip(line["unavailable"])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["metric"] == "cholesky_factor_gflops_n3_over_3" and line["unit"] == "GFLOP/s"
    assert line["dtype"] == "f64" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    # other ranks of a torchrun launch do nothing and exit 0
    quiet = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference"],
                           capture_output=True, text=True, env=dict(env, RANK="1", WORLD_SIZE="2"), cwd=ROOT)
    assert quiet.returncode == 0 and quiet.stdout.strip() == ""


def test_bench_own_arm_refuses_to_run_without_a_gpu():
    import json
    import sys
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert out.returncode != 0
    assert "no CPU fallback" in json.loads(out.stdout.strip().split("\n")[-1])["error"]
