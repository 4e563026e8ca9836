#!/usr/bin/env python
"""Regenerates tests/golden/ from the UNMODIFIED reference (run in the build container).

Needs /root/reference and `make -C oracle ref`.  Produces:
  inputs/*.txt          the reference's own tiny fixture matrices (test_data/*.txt), which
                        are BASELINE config 1 and the negative cases of
                        test_data/test13_matr_mpi.sh — copied because /root/reference does
                        not exist on the GPU box
  ref_dumps.npz         full-precision intermediates (A, b, norm, R, d, y, x, r, flags) of
                        the reference pipeline for small cases, from oracle/_ref/ref_dump
  ref_summary.json      residual / error / samples for cases too large to store whole
  cli/*.txt             stdout of the reference CLI (oracle/_ref/test_mat_ref_O2)
The GPU tests and the oracle tests compare against these files only.
"""
import json
import os
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402

REF_DATA = "/root/reference/test_data"
FIXTURES = {"a": 4, "a20": 4, "b": 4, "c": 6, "d": 6, "e": 6, "f": 4}

# (name, n, m, fixture or None)
SMALL = [(f"{k}_m3", n, 3, k) for k, n in FIXTURES.items()]
SMALL += [("a_m1", 4, 1, "a"), ("a_m4", 4, 4, "a"), ("a_m7", 4, 7, "a"), ("f_m2", 4, 2, "f"),
          ("e_m2", 6, 2, "e"), ("e_m4", 6, 4, "e")]
SMALL += [(f"formula_n{n}_m{m}", n, m, None) for n, m in
          [(3, 3), (10, 7), (11, 3), (30, 15), (30, 9), (64, 16), (64, 64), (96, 32), (100, 48),
           (130, 64), (200, 64)]]
LARGE = [(1000, 48), (1024, 64), (2048, 64), (2048, 128), (1536, 256), (4096, 64)]


def main():
    assert ol.have_ref(), "run `make -C oracle ref` first"
    inp = os.path.join(HERE, "inputs")
    cli = os.path.join(HERE, "cli")
    os.makedirs(inp, exist_ok=True)
    os.makedirs(cli, exist_ok=True)
    for k in FIXTURES:
        shutil.copyfile(os.path.join(REF_DATA, k + ".txt"), os.path.join(inp, k + ".txt"))

    store = {}
    for name, n, m, fx in SMALL:
        path = os.path.join(inp, fx + ".txt") if fx else None
        d = ol.ref_dump(n, m, path)
        for key, val in d.items():
            store[f"{name}/{key}"] = np.asarray(val)
    np.savez_compressed(os.path.join(HERE, "ref_dumps.npz"), **store)

    summary = {}
    for n, m in LARGE:
        d = ol.ref_dump(n, m)
        idx = np.linspace(0, n - 1, 17).astype(int)
        summary[f"formula_n{n}_m{m}"] = {
            "n": n, "m": m, "norm": d["norm"], "residual": d["residual"], "error": d["error"],
            "b_norm2": float(np.linalg.norm(d["b"])),
            "d_all_one": bool(np.all(d["d"] == 1.0)),
            "x_head": d["x"][:16].tolist(), "y_head": d["y"][:16].tolist(),
            "sample_idx": idx.tolist(),
            "R_samples": d["R"][np.ix_(idx, idx)].tolist(),
            "R_diag_sum": float(np.sum(np.diag(d["R"]))),
            "R_fro": float(np.linalg.norm(d["R"])),
        }
    with open(os.path.join(HERE, "ref_summary.json"), "w") as f:
        json.dump(summary, f, indent=1)

    # CLI stdout goldens (the elapsed field is timing noise; tests mask it).
    for name, n, m, fx in SMALL:
        args = [ol.REF_CLI_O2, str(n), str(m)]
        if fx:
            args.append(os.path.join("tests/golden/inputs", fx + ".txt"))
        out = subprocess.run(args, cwd=os.path.dirname(os.path.dirname(HERE)), check=True,
                             capture_output=True, text=True).stdout
        with open(os.path.join(cli, name + ".txt"), "w") as f:
            f.write(out)
    lg = os.path.join(os.path.dirname(os.path.dirname(HERE)), "log.txt")
    if os.path.exists(lg):
        os.remove(lg)
    print("golden files written:", len(store), "arrays,", len(summary), "summaries")


if __name__ == "__main__":
    main()
