#!/usr/bin/env python
"""Development diagnostic run on a B200 box: every stage in its own subprocess with a timeout
(a wedged mbarrier must not eat the whole gpurun call).  Writes gpurun_out/gpu_check.log.
Not part of the test-suite; tests/ holds the real parity tests.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def rel(a, b):
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den else 1.0))


def stage_probe():
    import matrix_solver_b200 as ms
    print("devices", ms.device_count())
    print("dmma_peak_tflops", ms.probe_dmma_peak(0))
    print("hbm_copy_gbs", ms.probe_hbm_copy(0, 1 << 30))


def stage_blocks():
    import matrix_solver_b200 as ms
    import oracle_lib as ol
    ora = ol.Oracle()
    rng = np.random.default_rng(3)
    for w in (1, 3, 8, 33, 64, 100):
        g = rng.standard_normal((w, w))
        spd = g @ g.T + w * np.eye(w)
        if w > 2:
            spd[w // 2:, w // 2:] *= -1
        blk = np.triu(spd)
        norma = np.abs(spd).sum(1).max()
        r0, d0, f0 = ora.diag_factor(blk, norma)
        r1, d1, f1 = ms.block_diag_factor(blk, norma)
        print(f"diag w={w}: fail {f0}/{f1} d_equal {np.array_equal(d0, d1)} bitexact "
              f"{np.array_equal(np.triu(r0), r1)} rel {rel(r1, np.triu(r0)):.2e}")
        i0, g0 = ora.upper_inverse(np.triu(r0), norma)
        i1, g1 = ms.block_upper_inverse(np.triu(r0), norma)
        print(f"  inverse: fail {g0}/{g1} bitexact {np.array_equal(np.triu(i0), i1)} rel {rel(i1, np.triu(i0)):.2e}")
        D = np.where(rng.random(w) < 0.3, -1.0, 1.0)
        for l in (w, w + 5, 1):
            A = rng.standard_normal((w, l))
            R1 = rng.standard_normal((w, w))
            R2 = rng.standard_normal((w, l))
            print(f"  update l={l}: rel {rel(ms.block_update(A, R1, D, R2), ora.update(A, R1, D, R2)):.2e}"
                  f"  panel rel {rel(ms.block_panel(np.triu(i0), A, D), ora.panel(np.triu(i0), A, D)):.2e}")
        Aup = np.triu(rng.standard_normal((w, w)))
        R1 = rng.standard_normal((w + 2, w))
        Dk = np.where(rng.random(w + 2) < 0.3, -1.0, 1.0)
        print(f"  update_diag: rel {rel(ms.block_update_diag(Aup, R1, Dk), np.triu(ora.update(Aup, R1, Dk, R1, True))):.2e}")


def run_pipeline(n, m, path=None, generic=False, lookahead=True, compare=True, dense=None):
    import matrix_solver_b200 as ms
    import oracle_lib as ol
    out = {"n": n, "m": m, "generic": generic, "lookahead": lookahead}
    t0 = time.time()
    with ms.Solver(n, m) as s:
        s.set_option("force_generic", generic)
        s.set_option("lookahead", lookahead)
        s.set_option("profile_update", int(os.environ.get("MS_PROFILE_UPDATE", "0")))
        if path or dense is not None:
            a = dense
            if a is None:
                ora = ol.Oracle()
                a, st = ora.init_file(n, path)
            s.upload_rows(a)
        else:
            s.init_formula(0)
        norm = s.b_and_norm()
        b = s.get_vec("b")
        A0 = s.get_upper() if n <= 2048 and compare else None
        fail, el = s.factor(norm)
        out.update(norm=norm, factor_fail=fail, factor_s=el, gflops=n ** 3 / 3 / el / 1e9 if el else 0)
        if not fail:
            R = s.get_upper() if n <= 2048 and compare else None
            d = s.get_vec("d")
            f2 = s.solve_forward(norm)
            f3 = s.solve_backward(norm)
            y, x = s.get_vec("y"), s.get_vec("x")
            if path or dense is not None:
                s.upload_rows(a)
            else:
                s.init_formula(0)
            res, err = s.residual()
            out.update(forward_fail=f2, backward_fail=f3, residual=res, error=err,
                       d_neg=int(np.sum(d < 0)), x_head=[float(v) for v in x[:4]])
        st = s.stats()
        out["stats"] = {k: st[k] for k in ("factor_ms", "forward_ms", "backward_ms", "init_ms", "bnorm_ms",
                                           "residual_ms", "kernel_launches", "used_dmma", "update_ms", "update_flops",
                                           "update_launches")}
        if st["update_ms"] > 0:
            out["update_tflops"] = st["update_flops"] / st["update_ms"] / 1e9
        if compare and n <= 2048:
            ora = ol.Oracle()
            o = ora.solve_pipeline(n, m, path) if dense is None else None
            if o is not None:
                out["A_equal"] = bool(np.array_equal(np.triu(A0), np.triu(o["A"])))
                out["b_rel"] = rel(b, o["b"])
                out["norm_equal"] = norm == o["norm"]
                out["ora_fail"] = o.get("factor_fail")
                if not fail and not o.get("factor_fail"):
                    out["R_rel"] = rel(np.triu(R), np.triu(o["R"]))
                    out["d_equal"] = bool(np.array_equal(d, o["d"]))
                    out["y_rel"] = rel(y, o["y"])
                    out["x_rel"] = rel(x, o["x"])
                    out["ora_residual"] = o["residual"]
                    out["ora_error"] = o["error"]
    out["wall_s"] = time.time() - t0
    print(json.dumps(out))
    return out


def stage_small_generic():
    import oracle_lib as ol
    fx = {"a": 4, "a20": 4, "b": 4, "c": 6, "d": 6, "e": 6, "f": 4}
    for k, n in fx.items():
        run_pipeline(n, 3, os.path.join(ol.FIXTURE_DIR, k + ".txt"))
    for n, m in [(3, 3), (10, 7), (30, 9), (4, 7), (100, 48), (130, 50), (200, 100), (257, 96)]:
        run_pipeline(n, m)


def stage_fast_small():
    for n, m in [(64, 64), (128, 64), (256, 64), (200, 64), (512, 128), (1000, 64), (1024, 256), (777, 192)]:
        run_pipeline(n, m)
        run_pipeline(n, m, lookahead=False)


def stage_fast_vs_generic():
    for n, m in [(512, 64), (2048, 128), (2048, 256)]:
        run_pipeline(n, m, generic=True)
        run_pipeline(n, m)


def stage_mid():
    for n, m in [(4096, 64), (8192, 128), (8192, 256)]:
        run_pipeline(n, m, compare=False)


def stage_big():
    for n, m in [(16384, 128), (16384, 256), (32768, 256)]:
        run_pipeline(n, m, compare=False)
        run_pipeline(n, m, compare=False, lookahead=False)


def stage_huge():
    run_pipeline(65536, 256, compare=False)


STAGES = {"probe": stage_probe, "blocks": stage_blocks, "small_generic": stage_small_generic,
          "fast_small": stage_fast_small, "fast_vs_generic": stage_fast_vs_generic, "mid": stage_mid,
          "big": stage_big, "huge": stage_huge}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage")
    ap.add_argument("--stages", default="probe,blocks,small_generic,fast_small,fast_vs_generic,mid")
    ap.add_argument("--timeout", type=int, default=240)
    a = ap.parse_args()
    if a.stage:
        STAGES[a.stage]()
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "gpu_check.log"), "w")
    for st in a.stages.split(","):
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, __file__, "--stage", st], capture_output=True, text=True,
                               timeout=a.timeout, cwd=ROOT)
            txt = f"=== {st} rc={p.returncode} {time.time() - t0:.1f}s\n{p.stdout}\n{p.stderr[-3000:]}\n"
        except subprocess.TimeoutExpired as e:
            txt = f"=== {st} TIMEOUT after {a.timeout}s\n{(e.stdout or b'').decode() if isinstance(e.stdout, bytes) else e.stdout}\n"
        log.write(txt)
        log.flush()
        print(txt)


if __name__ == "__main__":
    main()
