"""World-size-2 `gloo` test of the multi-process plumbing bench.py uses at N > 1: rendezvous on
127.0.0.1, hand-off of the 128-byte NCCL unique id, max-over-ranks timing, and the block-row
cyclic partition (every row owned exactly once)."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    from matrix_solver_b200 import dist, layout
    rank, local, world = dist.init("gloo")
    assert world == 2
    uid = bytes(range(128)) if rank == 0 else b""
    got = dist.broadcast_bytes(uid, 128)
    assert got == bytes(range(128)), got[:8]
    assert dist.max_over_ranks(1.5 + rank) == 2.5
    n, m = 1000, 64
    mine = layout.rows_local(n, m, rank, world)
    assert dist.sum_over_ranks(mine) == n
    flops = dist.sum_over_ranks(layout.update_flops(n, m, rank, world))
    assert abs(flops - layout.update_flops(n, m)) < 1e-3 * flops
    dist.barrier()
    dist.shutdown()
    print("rank", rank, "ok")
""") % ROOT


def test_two_rank_plumbing(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for r, p in enumerate(procs):
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out
        assert f"rank {r} ok" in out
