"""Multi-GPU parity (BASELINE configs[3], [4] plumbing): runs only where >= 2 devices are visible.

  * tools/multi_gpu_check.py under torchrun on 2 ranks — every exchange of the factorisation and the
    solves against the oracle (small n) and the analytic factor (n >= 1024), once per panel transport
    (peer-memory copy engines, and ncclBroadcast with MS_P2P=0);
  * the drop-in CLI with MS_NGPU=2 on a reference fixture and a formula case, against the committed
    reference stdout.
"""
import os
import subprocess
import sys

import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu


def visible_devices():
    import matrix_solver_b200 as m
    return m.device_count()


def torchrun(nproc, script, env_extra, port):
    env = dict(os.environ, **env_extra)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), script]
    return subprocess.run(cmd, cwd=ol.ROOT, env=env, capture_output=True, text=True, timeout=900)


@pytest.mark.parametrize("transport", ["p2p", "nccl"])
def test_two_rank_pipeline_against_oracle_and_analytic_factor(transport):
    if visible_devices() < 2:
        pytest.skip("needs >= 2 GPUs")
    env = {"MS_MULTI_QUICK": "1"}
    if transport == "nccl":
        env["MS_P2P"] = "0"
    p = torchrun(2, os.path.join("tools", "multi_gpu_check.py"), env, 29531 if transport == "p2p" else 29532)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert p.stdout.count('"ok": true') >= 4 and '"ok": false' not in p.stdout


@pytest.mark.parametrize("args,golden", [(("4", "3", "a.txt"), "a_m3.txt"), (("100", "48"), "formula_n100_m48.txt")])
def test_cli_on_two_ranks_prints_the_reference_output(args, golden):
    if visible_devices() < 2:
        pytest.skip("needs >= 2 GPUs")
    exe = os.path.join(ol.ROOT, "matrix_solver_b200", "bin", "test_mat")
    argv = list(args)
    if len(argv) == 3:
        argv[2] = os.path.relpath(os.path.join(ol.FIXTURE_DIR, argv[2]), ol.ROOT)  # the path is printed
    # NCCL_DEBUG=VERSION (set on some boxes) makes NCCL print its banner on stdout: not the CLI's text
    env = dict(os.environ, MS_NGPU="2", NCCL_DEBUG="WARN")
    p = subprocess.run([exe] + argv, cwd=ol.ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    want = open(os.path.join(ol.GOLDEN_DIR, "cli", golden)).read()
    got = "\n".join(ln for ln in p.stdout.split("\n") if not ln.startswith("NCCL version"))
    ol.compare_cli_output(got, want, golden)
