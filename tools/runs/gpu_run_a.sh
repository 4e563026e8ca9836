#!/bin/bash
# one single-GPU gpurun call: the whole GPU suite, then the bench lines of BASELINE configs c5, c2, c3
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest_gpu.log
tail -15 gpurun_out/r02_pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --e2e-steps 2 > gpurun_out/r02_bench_c5.json 2> gpurun_out/r02_bench_c5.err
echo "bench c5 rc=$?"; tail -c 3000 gpurun_out/r02_bench_c5.json
timeout 600 python bench.py --config c2 --steps 10 --warmup 3 > gpurun_out/r02_bench_c2.json 2> gpurun_out/r02_bench_c2.err
echo "bench c2 rc=$?"
timeout 900 python bench.py --config c3 --steps 5 --warmup 3 > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err
echo "bench c3 rc=$?"
python - <<'PY'
import json
for c in ("c5", "c2", "c3"):
    try:
        d = json.load(open(f"gpurun_out/r02_bench_{c}.json"))
        print(c, {k: d.get(k) for k in ("value", "factor_ms", "forward_ms", "backward_ms", "ms_per_step")},
              "parity", d["parity"]["ok"], d["parity"]["r_max_rel"], "roof", d["roofline"]["frac"],
              d["roofline"]["kernel_ms_per_step"], "cpu", (d.get("cpu_baseline") or {}).get("value"),
              "e2e", (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(c, "no line", e)
PY
