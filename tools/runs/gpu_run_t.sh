#!/bin/bash
# single-GPU call T (the last GPU minute of the round): bench lines of the two small configurations on HEAD
set -u
mkdir -p gpurun_out
timeout 25 python bench.py --config c3 --steps 5 --warmup 3 --e2e-steps 5 --no-cpu --no-dgemm 2> /dev/null | tail -1 > gpurun_out/r02t_c3.json
timeout 20 python bench.py --config c2 --steps 5 --warmup 3 --e2e-steps 5 --no-cpu --no-dgemm 2> /dev/null | tail -1 > gpurun_out/r02t_c2.json
python - <<'PY'
import json
for cfg in ("c2", "c3"):
    try:
        d = json.loads(open(f"gpurun_out/r02t_{cfg}.json").read().strip().splitlines()[-1]); r = d["roofline"]
        print(cfg, "factor", round(d["factor_ms"], 3), "TF", round(d["value"] / 1e3, 2), "blk", d["config"].get("block_internal"), "bwd", round(d["backward_ms"], 3),
              "upd", round(r["kernel_ms_per_step"], 2), "frac", round(r["frac"], 3), "tts", round(d["time_to_solution_ms"], 2), "parity", d["parity"]["ok"],
              "e2e", round((d.get("e2e") or {}).get("value") or 0, 1), "launches", d["gpu_launches"])
    except Exception as e:
        print(cfg, "no line", e)
PY
