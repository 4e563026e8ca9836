"""CPU-side checks: the C-ABI library loads and exports everything include/ms_b200.h declares,
fails loudly without a GPU (no CPU fallback), and the host-side layout logic matches the
reference's allocation rule.  No compute calls here."""
import ctypes
import os
import re
import subprocess

import numpy as np

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


def test_ms_stats_layout_matches_the_python_binding(tmp_path):
    """ms_stats crosses the ABI by value into a caller-owned struct: the ctypes mirror must have the header's
    fields at the header's offsets (a field appended on one side only would corrupt the caller's stack)."""
    fields = [name for name, _ in capi.Stats._fields_]
    src = tmp_path / "stats_layout.c"
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void) {',
             '  printf("sizeof %zu\\n", sizeof(ms_stats));']
    lines += [f'  printf("{f} %zu %zu\\n", offsetof(ms_stats, {f}), sizeof(((ms_stats *)0)->{f}));' for f in fields]
    lines += ['  return 0;', '}']
    src.write_text("\n".join(lines) + "\n")
    exe = tmp_path / "stats_layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    assert int(out[0].split()[1]) == ctypes.sizeof(capi.Stats)
    for line, (name, ctype) in zip(out[1:], capi.Stats._fields_):
        f, off, size = line.split()
        assert f == name and int(off) == getattr(capi.Stats, name).offset and int(size) == ctypes.sizeof(ctype), line
    # and every field of the header is mirrored: the struct body names exactly these members
    body = re.search(r"typedef struct ms_stats\s*\{(.*?)\}\s*ms_stats;", re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S), re.S).group(1)
    declared = re.findall(r"\b([a-z_0-9]+)\s*[;,]", body)
    assert declared == fields


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
    # a rank that dies must take the whole job down (exit 3) instead of leaving its peers in a barrier:
    # once a child (SIGCHLD -> the parent stops everybody) and once rank 0 (children die with the parent)
    for victim in (2, 0):
        out = subprocess.run([str(exe), "die", str(victim)], env=dict(os.environ, MS_NGPU="4"), capture_output=True,
                             text=True, timeout=60)
        assert out.returncode == 3, (victim, out.returncode, out.stderr)
        assert "simulated failure" in out.stderr


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


def _matrix_file_driver(tmp_path):
    exe = tmp_path / "matrix_file_test"
    host = os.path.join(ROOT, "matrix_solver_b200", "host")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + host, os.path.join(ROOT, "tests", "matrix_file_test.cpp"),
                    os.path.join(host, "ms_matrix_file.cpp"), "-o", str(exe)], check=True)

    def run(path, n, mode, chunk=3):
        out = subprocess.run([str(exe), str(path), str(n), mode, str(chunk)], capture_output=True, text=True,
                             timeout=120, check=True)
        return out.stdout
    return run


def test_matrix_file_reader_agrees_with_fscanf(tmp_path):
    """host/ms_matrix_file.cpp replaces the reference's per-entry fscanf("%lf") (src/init.cpp:100-112)
    with a buffered strtod parser: same values bit for bit, same acceptance of odd spacing, same
    verdict on short or malformed files."""
    import glob
    run = _matrix_file_driver(tmp_path)
    # the reference's own fixtures
    for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "inputs", "*.txt"))):
        n = len(open(path).readline().split())
        assert run(path, n, "fast") == run(path, n, "scanf"), path
        assert run(path, n, "fast", chunk=1000).split("\n")[0] == str(n)
    rng = np.random.default_rng(3)
    n = 37
    a = rng.standard_normal((n, n)) * 10.0 ** rng.integers(-30, 30, (n, n))
    fmts = ["%.17g", "%.6f", "%e", "%+.3E", "%g"]
    text = tmp_path / "ragged.txt"
    with open(text, "w") as f:
        for i in range(n):
            for j in range(n):
                f.write(fmts[(i + j) % len(fmts)] % a[i, j])
                f.write([" ", "\t", "\n", "   ", " \r\n"][(3 * i + j) % 5])   # line breaks anywhere
    for chunk in (1, 5, 64):
        assert run(text, n, "fast", chunk) == run(text, n, "scanf")
    # values round-trip exactly when written with 17 significant digits
    exact = tmp_path / "exact.txt"
    np.savetxt(exact, a, fmt="%.17g")
    got = run(exact, n, "fast").split()
    assert int(got[0]) == n
    assert np.array_equal(np.array([int(h, 16) for h in got[1:]], dtype=np.uint64).view(np.float64).reshape(n, n), a)
    # short file: the row that hit EOF does not count; garbage stops the parse like "%lf" does
    body = open(exact).read().split("\n")
    short = tmp_path / "short.txt"
    short.write_text("\n".join(body[:20]) + "\n" + " ".join(body[20].split()[:5]))
    assert run(short, n, "fast") == run(short, n, "scanf")
    assert run(short, n, "fast").split("\n")[0] == "20"
    bad = tmp_path / "bad.txt"
    bad.write_text("\n".join(body[:4]) + "\n1.0 abc 2.0\n" + "\n".join(body[5:]))
    assert run(bad, n, "fast") == run(bad, n, "scanf")
    assert run(bad, n, "fast").split("\n")[0] == "4"
    assert run(tmp_path / "missing.txt", n, "fast") == "cant open\n"
    # special tokens strtod and "%lf" both take
    special = tmp_path / "special.txt"
    special.write_text("inf -inf nan 0x1p-3\n+1.5 1e400 -1e400 1e-400\n4.9e-324 -0x1.8P1 .5 5.\n"
                       "1E5 -0 2.2250738585072011e-308 0X10\n")
    assert run(special, 4, "fast") == run(special, 4, "scanf")
    assert run(special, 4, "fast").split("\n")[0] == "4"


def test_matrix_file_reader_tokens_across_buffer_refills(tmp_path):
    """A file larger than the reader's 8 MB buffer: numbers that straddle a refill boundary must come
    out whole."""
    run = _matrix_file_driver(tmp_path)
    n = 800
    a = np.random.default_rng(5).standard_normal((n, n))
    big = tmp_path / "big.txt"
    np.savetxt(big, a, fmt="%.17g")
    assert os.path.getsize(big) > 12 << 20
    got = run(big, n, "fast", chunk=100).split()
    assert int(got[0]) == n
    assert np.array_equal(np.array([int(h, 16) for h in got[1:]], dtype=np.uint64).view(np.float64).reshape(n, n), a)


def test_matrix_file_reader_raw_f64(tmp_path):
    """`name.f64`: n*n little-endian doubles, row-major (the large-input format)."""
    run = _matrix_file_driver(tmp_path)
    n = 50
    a = np.random.default_rng(6).standard_normal((n, n))
    raw = tmp_path / "m.f64"
    a.astype("<f8").tofile(raw)
    got = run(raw, n, "fast", chunk=7).split()
    assert int(got[0]) == n
    assert np.array_equal(np.array([int(h, 16) for h in got[1:]], dtype=np.uint64).view(np.float64).reshape(n, n), a)
    a[:33].astype("<f8").tofile(raw)
    with open(raw, "ab") as f:
        f.write(b"\0" * 12)
    assert run(raw, n, "fast", chunk=7).split("\n")[0] == "33"


def test_matrix_file_reader_skips_rows_of_other_ranks(tmp_path):
    """init_mat_file lets a rank skip the chunks it owns nothing of: a seek for the raw format, a parse
    without delivery for text.  Alternating read/skip must deliver exactly the right rows, and a short
    file must be noticed while skipping too."""
    run = _matrix_file_driver(tmp_path)
    n = 23
    a = np.random.default_rng(8).standard_normal((n, n))
    text, raw = tmp_path / "m.txt", tmp_path / "m.f64"
    np.savetxt(text, a, fmt="%.17g")
    a.astype("<f8").tofile(raw)
    for path in (text, raw):
        for chunk in (1, 4, 7):
            out = run(path, n, "alternate", chunk).split("\n")
            pos, row, take = 0, 0, True
            while row < n:
                want = min(chunk, n - row)
                kind, first, got = out[pos].split()
                assert (kind, int(first), int(got)) == ("read" if take else "skip", row, want)
                pos += 1
                if take:
                    vals = np.array([int(h, 16) for h in out[pos:pos + want * n]], dtype=np.uint64).view(np.float64)
                    assert np.array_equal(vals.reshape(want, n), a[row:row + want])
                    pos += want * n
                row += want
                take = not take
    # short files: 10 complete rows and a bit
    np.savetxt(text, a[:10], fmt="%.17g")
    with open(text, "a") as f:
        f.write("1.0 2.0\n")
    a[:10].astype("<f8").tofile(raw)
    with open(raw, "ab") as f:
        f.write(b"\0" * 20)
    for path in (text, raw):
        out = run(path, n, "alternate", 4).split("\n")
        heads = [ln for ln in out if ln.startswith(("read", "skip"))]
        assert heads == ["read 0 4", "skip 4 4", "read 8 2"]
        out = run(path, n, "alternate", 3).split("\n")
        heads = [ln for ln in out if ln.startswith(("read", "skip"))]
        assert heads == ["read 0 3", "skip 3 3", "read 6 3", "skip 9 1"]
