#!/bin/bash
# single-GPU gpurun call C: suite, diag kernel A/B, c5/c2/c3 bench (banded tile walk, prefetching backward sweep)
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02c_pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02c_pytest_gpu.log
tail -8 gpurun_out/r02c_pytest_gpu.log
python tools/diag_timing.py > gpurun_out/r02c_diag_timing.log 2>&1; tail -5 gpurun_out/r02c_diag_timing.log
B="--steps 3 --warmup 2 --no-e2e --no-cpu --no-dgemm"
for cfg in c5 c2 c3; do
  timeout 300 python bench.py --config $cfg $B > gpurun_out/r02c_${cfg}_strip.json 2> gpurun_out/r02c_${cfg}_strip.err; echo "$cfg strip rc=$?"
  MS_DIAG_STRIP=0 timeout 300 python bench.py --config $cfg $B > gpurun_out/r02c_${cfg}_old.json 2> /dev/null; echo "$cfg old-diag rc=$?"
done
python - <<'PY'
import json
for cfg in ("c5", "c2", "c3"):
    for v in ("strip", "old"):
        try:
            d = json.load(open(f"gpurun_out/r02c_{cfg}_{v}.json")); r = d["roofline"]
            print(cfg, v, "factor", round(d["factor_ms"], 2), "TF", round(d["value"] / 1e3, 2), "bwd", round(d["backward_ms"], 3),
                  "upd union", round(r["kernel_ms_per_step"], 1), "frac", round(r["frac"], 3), "parity", d["parity"]["ok"],
                  d["parity"]["r_max_rel"])
        except Exception as e:
            print(cfg, v, "no line", e)
PY
