#!/usr/bin/env python
"""bench.py — factor GFLOP/s (n^3/3) + solve time of the blocked R^T D R path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config c2|c3|c4|c5] [--n N --m M]

--config picks one of BASELINE.json's formula configurations: c2 = n 4096 / m 64, c3 = 16384 / 128,
c4 = 32768 / 256, c5 = 65536 / 256 (default: the configuration the headline metric is quoted on).

N > 1 is launched by torchrun (one rank per GPU; RANK / LOCAL_RANK / WORLD_SIZE from the env).
Prints ONE JSON line on rank 0.

Workload (BASELINE.json configs[4], the configuration the headline metric is quoted on; it fits
one 180 GB B200): formula-generated SPD A_ij = n - max(i,j), n = 65536, block size m = 256, fixed
as N grows ("strong" scaling).  One step = one pass of the hot path with the inputs resident in
HBM: device generator -> b and ||A||_inf -> factor -> forward sweep -> backward sweep.
  value        = (n^3/3) / (mean device time of ms_factor, CUDA events on the library's stream,
                 max over ranks)                                          [GFLOP/s]
  ms_per_step  = wall time of the whole step between device syncs (max over ranks)
  e2e          = (n^3/3) / (host-buffer time-to-solution): the step's matrix comes from page-locked
                 HOST memory through ms_upload_rows, then b/norm, factor, both sweeps and the D2H
                 read of x and d — all inside the timed region
  roofline     = the trailing-update tile kernel (K4+K5): algorithmic flops of its launches / the
                 device time in which at least one of them was running (union of the per-launch
                 CUDA-event intervals, so never more than factor_ms), against the fp64 tensor-pipe
                 peak measured in this run (register-only DMMA.8x8x4 issue loop; a cuBLAS Dgemm
                 8192^3 figure is printed beside it as a library yardstick)
  parity       = checked OUTSIDE the timed region on the last factorisation, at every N: ||A||_inf ==
                 n(n+1)/2, d == 1, >= 64 rows of R (block-row / rank boundaries, last block row)
                 against the analytic factor R_ij = (n-j)/sqrt((n-i)(n-i+1)), x against the planted
                 solution, the residual.  A failed check makes the process exit non-zero.
  cpu_baseline = the UNMODIFIED reference (oracle/_ref/ref_time: cholesky_decomp, then compute_y +
                 compute_x, wall clock) on the box's host cores through the minimpi shim; at the
                 same n for c2 / c3, on a bounded sample (n = 16384, same generator and m) above
`--impl reference` prints the same line for the reference's CPU path alone.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cholesky_factor_gflops_n3_over_3"
UNIT = "GFLOP/s"


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.proc = None
        self.lines = []
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- reference arm
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference_sample(n, m, cores):
    """One run of the unmodified reference (oracle/_ref/ref_time: the reference's own functions in the
    order of test/test_mat.cpp) on `cores` minimpi ranks; factor and solve wall clock."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_time")
    if not os.path.exists(exe):
        return None, "oracle/_ref/ref_time not built (needs /root/reference at build time)"
    env = dict(os.environ, MINIMPI_NP=str(cores))
    cwd = os.path.join(ROOT, "gpurun_out")
    os.makedirs(cwd, exist_ok=True)
    p = subprocess.run([exe, str(n), str(m)], capture_output=True, text=True, env=env, cwd=cwd)
    try:
        d = json.loads(p.stdout.strip().splitlines()[-1])
    except (ValueError, IndexError):
        d = None
    if p.returncode != 0 or not d or d.get("failed"):
        return None, f"reference run failed rc={p.returncode}: {p.stdout[-200:]} {p.stderr[-200:]}"
    return {"elapsed": d["factor_s"], "solve": d["solve_s"], "residual": d["residual"], "error": d["error"]}, None


def run_oracle_port_sample(n, m):
    """Fallback when oracle/_ref was not shipped: time the oracle's C restatement of the same
    algorithm (oracle/ms_oracle.c, one host thread).  bench.py is one of the places allowed to
    execute oracle/ — as the reported CPU baseline only."""
    import ctypes
    import numpy as np
    so = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True)
    L = ctypes.CDLL(so)
    dp = ctypes.POINTER(ctypes.c_double)
    L.ora_init_formula.argtypes = [dp, ctypes.c_size_t]
    L.ora_b_norm.argtypes = [dp, ctypes.c_size_t, dp, dp]
    L.ora_factor.restype = ctypes.c_int
    L.ora_factor.argtypes = [dp, dp, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_double]
    a = np.zeros((n, n))
    d = np.zeros(n)
    b = np.zeros(n)
    norm = ctypes.c_double(0)
    L.ora_init_formula(a.ctypes.data_as(dp), n)
    L.ora_b_norm(a.ctypes.data_as(dp), n, b.ctypes.data_as(dp), ctypes.byref(norm))
    t0 = time.perf_counter()
    fail = L.ora_factor(a.ctypes.data_as(dp), d.ctypes.data_as(dp), n, m, norm.value)
    el = time.perf_counter() - t0
    if fail:
        return None, "oracle port reported a numerical failure"
    return {"elapsed": el, "solve": None, "residual": None, "error": None}, None


def reference_sample_size(n_full, repeated):
    """Bounded sample: same generator and m; the same n up to 16384 (about 20 s of CPU work per run on
    16 cores), n = 16384 above that.  The reference ARM repeats the run steps + warmup times, so its
    bound is n = 8192 (about 5 s per run) unless MS_BENCH_REF_N says otherwise."""
    cap = int(os.environ.get("MS_BENCH_REF_N", "8192" if repeated else "16384"))
    return min(n_full, cap)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0  # the other ranks of a torchrun launch have nothing to do
    cores = host_cores()
    ns = reference_sample_size(args.n, repeated=True)
    kind = "reference"
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_time")):
        kind, cores, ns = "port", 1, min(ns, 2048)
    times, solves, last = [], [], None
    for it in range(args.warmup + args.steps):
        if kind == "reference":
            r, err = run_reference_sample(ns, args.m, cores)
        else:
            r, err = run_oracle_port_sample(ns, args.m)
        if err:
            print(json.dumps({"impl": "reference", "unavailable": err}))
            return 0
        last = r
        if it >= args.warmup:
            times.append(r["elapsed"])
            solves.append(r["solve"])
    t = sum(times) / len(times)
    solve_ms = sum(solves) / len(solves) * 1e3 if all(v is not None for v in solves) else None
    gf = ns ** 3 / 3.0 / t / 1e9
    if kind == "reference":
        sample = (f"formula A_ij=n-max(i,j), n={ns}, m={args.m}: unmodified reference, cholesky_decomp wall clock "
                  f"(what its own `elapsed` covers), {cores} minimpi ranks on the box's host cores")
    else:
        sample = (f"formula A_ij=n-max(i,j), n={ns}, m={args.m}: oracle/ms_oracle.c (C restatement of the "
                  f"reference, one host thread) because oracle/_ref was not shipped")
    line = {
        "impl": "reference", "metric": METRIC, "value": gf, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"formula SPD n={args.n} m={args.m}; reference timed on the bounded sample n={ns}",
                   "n": args.n, "m": args.m, "sample_n": ns},
        "cpu_baseline": {"value": gf, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "solve_ms": solve_ms},
        "e2e": {"value": gf, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "factor_ms": t * 1e3, "solve_ms": solve_ms, "same_n_as_workload": ns == args.n,
        "residual_abs": last["residual"], "error_abs": last["error"],
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------- our arm
def dgemm_yardstick():
    """cuBLAS Dgemm 8192^3 through torch.matmul: burst (best of 5) and 4 s back to back.  A library
    reference for "how good is the DMMA kernel", not a denominator."""
    import torch
    try:
        a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
        b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
        torch.matmul(a, b)
        torch.cuda.synchronize()
        flops = 2.0 * 8192 ** 3
        best = 0.0
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            e1.synchronize()
            best = max(best, flops / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(4, int(4.0 / (flops / (best * 1e12))))
        e0.record()
        for _ in range(reps):
            torch.matmul(a, b)
        e1.record()
        e1.synchronize()
        sustained = reps * flops / (e0.elapsed_time(e1) * 1e-3) / 1e12
        del a, b
        torch.cuda.empty_cache()
        return {"burst_tflops": best, "sustained_tflops": sustained}
    except Exception as exc:  # noqa: BLE001 - a yardstick must never break the bench
        return {"error": str(exc)[:120]}


def our_arm(args):
    import numpy as np
    import torch
    import matrix_solver_b200 as ms
    from matrix_solver_b200 import dist, layout, selfcheck

    rank, local, world = dist.env_rank()
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            print(json.dumps({"error": f"--gpus {args.gpus} needs torchrun --nproc-per-node {args.gpus}"}))
            return 2
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: bench.py measures the B200 path, there is no CPU fallback"}))
        return 2
    torch.cuda.set_device(local)
    dist.init("nccl" if world > 1 else None)
    n, m = args.n, args.m

    uid = None
    if world > 1:
        raw = ms.get_unique_id() if rank == 0 else b""
        uid = dist.broadcast_bytes(raw, 128)
    solver = ms.Solver(n, m, device=local, rank=rank, nranks=world, unique_id=uid)
    solver.set_option("profile_update", 1)

    peak_dmma = ms.probe_dmma_peak(local) if rank == 0 else 0.0
    dgemm = dgemm_yardstick() if (rank == 0 and not args.no_dgemm) else None
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass

    def step():
        solver.init_formula(0)
        norm = solver.b_and_norm()
        failed, _ = solver.factor(norm)
        f2 = solver.solve_forward(norm)
        f3 = solver.solve_backward(norm)
        return norm, failed or f2 or f3

    def sync_all():
        solver.sync()
        torch.cuda.synchronize()
        dist.barrier()

    for _ in range(args.warmup):
        step()
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    st0 = solver.stats()
    acc = {k: 0.0 for k in ("factor_ms", "forward_ms", "backward_ms", "init_ms", "bnorm_ms", "update_ms",
                            "update_sum_ms", "update_flops")}
    any_fail = False
    norm = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        norm, failed = step()
        any_fail |= failed
        st = solver.stats()
        for k in acc:
            acc[k] += st[k]
    sync_all()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        # the probe before the timed steps is sometimes taken while the GPU still sits at a lower clock level
        # (29.7 = 0.8 x 37.1 TFLOP/s seen on fresh boxes): probe again on the warm GPU and keep the larger one
        peak_dmma = max(peak_dmma, ms.probe_dmma_peak(local))
    st1 = solver.stats()
    forward_fused = bool(st1["forward_fused"])

    # ---- parity of the last step (outside the timed region; collective: every rank takes part)
    parity = selfcheck.check_parity(solver, n, m, world, norm, any_fail)
    # residual with the regenerated matrix (the factor overwrote A, test/test_mat.cpp:149-160)
    solver.init_formula(0)
    resid, err = solver.residual()
    bnorm2 = float(np.linalg.norm(solver.get_vec("b")))
    parity["residual_rel"] = resid / bnorm2 if bnorm2 else None
    parity["residual_tol_rel"] = 1e-13
    parity["ok"] = bool(parity["ok"] and bnorm2 > 0 and resid / bnorm2 <= 1e-13)
    parity_ok_all = dist.max_over_ranks(0.0 if parity["ok"] else 1.0) == 0.0

    K = args.steps
    factor_ms = dist.max_over_ranks(acc["factor_ms"] / K)
    wall_ms = dist.max_over_ranks(wall / K * 1e3)
    fwd_ms = dist.max_over_ranks(acc["forward_ms"] / K)
    bwd_ms = dist.max_over_ranks(acc["backward_ms"] / K)
    init_ms = dist.max_over_ranks(acc["init_ms"] / K)
    bnorm_ms = dist.max_over_ranks(acc["bnorm_ms"] / K)
    upd_ms_rank = acc["update_ms"] / K          # union of the launch intervals on this rank
    upd_sum_ms_rank = acc["update_sum_ms"] / K
    upd_flops_rank = acc["update_flops"] / K
    upd_tflops_rank = upd_flops_rank / upd_ms_rank / 1e9 if upd_ms_rank > 0 else 0.0
    upd_tflops_min = -dist.max_over_ranks(-upd_tflops_rank)
    upd_ms_max = dist.max_over_ranks(upd_ms_rank)
    launches = dist.sum_over_ranks(st1["kernel_launches"] - st0["kernel_launches"])
    value = layout.factor_flops(n) / (factor_ms * 1e-3) / 1e9

    # ---- e2e: host buffers in, host results out (rows of A from page-locked memory)
    e2e = None
    if not args.no_e2e:
        try:
            rows_mine = layout.rows_local(n, m, rank, world)
            # every rank keeps the rows it owns, full length, contiguous in block-row order
            pin = ms.PinnedBuffer(max(rows_mine, 1) * n)
            blocks = [I for I in range(rank, layout.n_blocks(n, m), world)]
            solver.init_formula(0)
            off = 0
            for I in blocks:
                h = layout.block_rows(n, m, I)
                # write block row I so that global column j sits at offset j of each row
                solver.download_rows_ptr(pin.ptr.value + off * n * 8, I * m, h, n)
                off += h
            sync_all()
            e2e_times = []
            h2d0, d2h0 = solver.stats()["h2d_bytes"], solver.stats()["d2h_bytes"]
            for it in range(1 + args.e2e_steps):
                sync_all()
                t1 = time.perf_counter()
                off = 0
                for I in blocks:
                    h = layout.block_rows(n, m, I)
                    solver.upload_rows_ptr(pin.ptr.value + off * n * 8, I * m, h, n)
                    off += h
                nrm = solver.b_and_norm()
                solver.factor(nrm)
                solver.solve_forward(nrm)
                solver.solve_backward(nrm)
                x = solver.get_vec("x")
                d = solver.get_vec("d")
                sync_all()
                if it > 0:
                    e2e_times.append(time.perf_counter() - t1)
                else:
                    h2d0, d2h0 = solver.stats()["h2d_bytes"], solver.stats()["d2h_bytes"]
            t_e2e = dist.max_over_ranks(sum(e2e_times) / len(e2e_times))
            stx = solver.stats()
            e2e = {"value": layout.factor_flops(n) / t_e2e / 1e9, "unit": UNIT,
                   "h2d_bytes_per_step": dist.sum_over_ranks((stx["h2d_bytes"] - h2d0) / len(e2e_times)),
                   "d2h_bytes_per_step": dist.sum_over_ranks((stx["d2h_bytes"] - d2h0) / len(e2e_times)),
                   "seconds": t_e2e, "steps": len(e2e_times),
                   "x_err_vs_planted": float(np.linalg.norm(x - (np.arange(n) % 2 == 0))),
                   "what": "pinned host rows -> ms_upload_rows -> b/norm -> factor (+ fused forward) -> backward -> D2H x, d"}
            pin.free()

        except (ms.MSError, MemoryError) as exc:  # e.g. not enough page-lockable host memory
            e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                   "error": str(exc)[:200]}

    solver.close()
    if rank != 0:
        dist.shutdown()
        return 0 if parity_ok_all else 3

    # ---- cpu baseline (rank 0, N = 1 only): the reference on the host cores, bounded sample
    cpu = None
    if world == 1 and not args.no_cpu:
        cores = host_cores()
        ns = reference_sample_size(n, repeated=False)
        r, errmsg = run_reference_sample(ns, m, cores)
        if r:
            cpu = {"value": ns ** 3 / 3.0 / r["elapsed"] / 1e9, "unit": UNIT, "cores": cores, "kind": "reference",
                   "solve_ms": r["solve"] * 1e3, "factor_ms": r["elapsed"] * 1e3, "same_n_as_workload": ns == n,
                   "sample": f"formula n={ns}, m={m}: unmodified reference (oracle/_ref/ref_time), {cores} minimpi "
                             f"ranks: cholesky_decomp {r['elapsed']:.2f} s, compute_y + compute_x {r['solve']:.3f} s; "
                             f"residual {r['residual']:.3e}"}
        else:
            ns = min(ns, 2048)
            r, err2 = run_oracle_port_sample(ns, m)
            if r:
                cpu = {"value": ns ** 3 / 3.0 / r["elapsed"] / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
                       "sample": f"formula n={ns}, m={m}: oracle/ms_oracle.c, one host thread, {r['elapsed']:.2f} s "
                                 f"({errmsg})"}
            else:
                cpu = {"value": None, "unit": UNIT, "cores": cores, "kind": "reference", "sample": errmsg}

    # DRAM traffic of the dominant kernel: one `ncu --set full` capture of ONE launch (the step-0 bulk
    # update of this n, m on one GPU), regenerated by tools/ncu_traffic.sh; quoted only for the
    # configuration it was captured on
    traffic, traffic_note = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "update_kernel_traffic.json")) as f:
            tj = json.load(f)
        if world == 1 and tj.get("n") == n and tj.get("m") == m:
            traffic = tj.get("dram_bytes_per_launch")
            traffic_note = {k: tj.get(k) for k in ("launch", "algorithmic_bytes_per_launch", "duration_ms_under_ncu",
                                                   "dmma_pipe_active_pct", "source")}
    except (OSError, ValueError):
        pass
    peak_src = ("measured in this run: register-only DMMA.8x8x4 issue loop on all SMs (ms_probe_dmma_peak) = "
                f"{peak_dmma:.2f} TFLOP/s; MEASURED_PEAKS.json has no fp64 entry; nominal 37 TFLOP/s")
    if dgemm and "burst_tflops" in dgemm:
        peak_src += (f"; library yardstick (not the denominator): cuBLAS Dgemm 8192^3 "
                     f"{dgemm['burst_tflops']:.2f} burst / {dgemm['sustained_tflops']:.2f} sustained TFLOP/s")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": wall_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"formula-generated SPD A_ij=n-max(i,j), n={n}, block m={m}, factor+solve "
                               f"(BASELINE configs[{CONFIG_INDEX.get((n, m), '-')}]); block-row cyclic over {world} GPU(s)",
                   "n": n, "m": m,
                   # rows per block row the kernels walked: m, or the internal 256 a single GPU uses for m = 64 / 128
                   # (same R and d)
                   "block_internal": int(st1["block_internal"]),
                   "parallelism": f"block-row-cyclic x{world}, look-ahead 2, panel over NVLink peer memory "
                                  f"(copy engines; ncclBroadcast when MS_P2P=0)" if world > 1 else "single GPU, look-ahead 2",
                   "forward_solve": "fused into the factorisation (b as augmented column)" if forward_fused
                                    else "separate sweep",
                   "l2": "inputs larger than L2 (matrix %.1f GB per GPU)" % (8.0 * n * n / world / 1e9)},
        "factor_ms": factor_ms, "forward_ms": fwd_ms, "backward_ms": bwd_ms, "solve_ms": fwd_ms + bwd_ms,
        "init_ms": init_ms, "bnorm_ms": bnorm_ms,
        "time_to_solution_ms": init_ms + bnorm_ms + factor_ms + fwd_ms + bwd_ms,
        "residual_abs": resid, "residual_rel": resid / bnorm2 if bnorm2 else None, "error_abs": err,
        "numerical_failure": bool(any_fail), "parity": parity,
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "tile_dmma_kernel<UPDATE> (fp64 DMMA.8x8x4 trailing update)",
                     "achieved": upd_tflops_min, "peak": peak_dmma, "unit": "TFLOP/s",
                     "frac": upd_tflops_min / peak_dmma if peak_dmma else None, "traffic": traffic,
                     "traffic_capture": traffic_note,
                     "peak_source": peak_src, "dgemm_yardstick": dgemm,
                     "algorithmic_flops_per_step": float(upd_flops_rank),
                     "kernel_ms_per_step": upd_ms_max,
                     "kernel_ms_per_step_is": "union of the update launches' CUDA-event intervals (max over ranks); "
                                              "<= factor_ms by construction",
                     "kernel_ms_sum_of_launches": upd_sum_ms_rank,
                     "whole_factor_frac_of_peak": value / 1e3 / (peak_dmma * world) if peak_dmma else None,
                     "hbm_peak_gbs": peaks.get("hbm_gbs"),
                     "backward_achieved_gbs": layout.solve_bytes(n) / (bwd_ms * 1e-3) / 1e9 if bwd_ms else None,
                     "backward_frac_of_hbm": (layout.solve_bytes(n) / (bwd_ms * 1e-3) / 1e9 / (peaks["hbm_gbs"] * world)
                                              if bwd_ms and peaks.get("hbm_gbs") else None)},
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    dist.shutdown()
    if not parity_ok_all:
        print("bench.py: PARITY CHECK FAILED: " + json.dumps(parity), file=sys.stderr)
        return 3
    return 0


CONFIGS = {"c2": (4096, 64), "c3": (16384, 128), "c4": (32768, 256), "c5": (65536, 256)}
CONFIG_INDEX = {(4096, 64): 1, (16384, 128): 2, (32768, 256): 3, (65536, 256): 4}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=os.environ.get("MS_BENCH_CONFIG", "c5"), choices=sorted(CONFIGS))
    ap.add_argument("--n", type=int, default=int(os.environ.get("MS_BENCH_N", "0")))
    ap.add_argument("--m", type=int, default=int(os.environ.get("MS_BENCH_M", "0")))
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-dgemm", action="store_true")
    args = ap.parse_args()
    cn, cm = CONFIGS[args.config]
    args.n = args.n or cn
    args.m = args.m or cm
    if args.impl == "reference":
        return reference_arm(args)
    return our_arm(args)


if __name__ == "__main__":
    sys.exit(main())
