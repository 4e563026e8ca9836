#!/bin/bash
# single-GPU gpurun call D: latency probe, GPU suite (one-pass b/norm + residual, file staging), quick benches
set -u
mkdir -p gpurun_out
matrix_solver_b200/bin/lat_probe | tee gpurun_out/r02_latency_probe.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02d_pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02d_pytest_gpu.log
tail -8 gpurun_out/r02d_pytest_gpu.log
B="--steps 3 --warmup 2 --no-e2e --no-cpu --no-dgemm"
timeout 300 python bench.py --config c5 $B > gpurun_out/r02d_c5.json 2> gpurun_out/r02d_c5.err; echo "c5 rc=$?"
timeout 300 python bench.py --config c2 $B > gpurun_out/r02d_c2.json 2> gpurun_out/r02d_c2.err; echo "c2 rc=$?"
MS_NO_LOOKAHEAD=1 timeout 300 python bench.py --config c2 $B > gpurun_out/r02d_c2_nola.json 2> /dev/null; echo "c2 nola rc=$?"
python - <<'PY'
import json
for cfg in ("c5", "c2", "c2_nola"):
    try:
        d = json.load(open(f"gpurun_out/r02d_{cfg}.json")); r = d["roofline"]
        print(cfg, "factor", round(d["factor_ms"], 2), "TF", round(d["value"] / 1e3, 2), "bwd", round(d["backward_ms"], 3),
              "bnorm", round(d["bnorm_ms"], 3), "init", round(d["init_ms"], 3), "frac", round(r["frac"], 3), "parity", d["parity"]["ok"])
    except Exception as e:
        print(cfg, "no line", e)
PY
