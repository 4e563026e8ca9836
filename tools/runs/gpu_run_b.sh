#!/bin/bash
# single-GPU gpurun call B: GPU suite with the new backward sweep, layout A/B, ncu launch list + full captures
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02b_pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02b_pytest_gpu.log
tail -8 gpurun_out/r02b_pytest_gpu.log
B="--steps 3 --warmup 2 --no-e2e --no-cpu --no-dgemm"
timeout 300 python bench.py $B > gpurun_out/r02b_c5_default.json 2> gpurun_out/r02b_c5_default.err; echo "default rc=$?"
MS_FUSE_FORWARD=0 timeout 300 python bench.py $B > gpurun_out/r02b_c5_nofuse.json 2> gpurun_out/r02b_c5_nofuse.err; echo "nofuse rc=$?"
timeout 300 python bench.py $B > gpurun_out/r02b_c5_default2.json 2> /dev/null; echo "default2 rc=$?"
python - <<'PY'
import json
for c in ("default", "nofuse", "default2"):
    try:
        d = json.load(open(f"gpurun_out/r02b_c5_{c}.json")); r = d["roofline"]
        print(c, "factor", round(d["factor_ms"], 1), "fwd", round(d["forward_ms"], 3), "bwd", round(d["backward_ms"], 3),
              "upd union", round(r["kernel_ms_per_step"], 1), "peak", round(r["peak"], 2), "frac", round(r["frac"], 3),
              "parity", d["parity"]["ok"], "x_rel", d["parity"]["x_rel"])
    except Exception as e:
        print(c, "no line", e)
PY
# launch list of the bench command (one step, serialised under ncu: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02_launches_c5.csv \
  python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --no-dgemm > gpurun_out/r02_launches_c5.out 2>&1
echo "ncu launch list rc=$?"
# one bulk update launch (step 0, look-ahead off so that the launch order is fixed: 3 strip updates, U1a, U1b, U2a, U2b)
MS_LOOKAHEAD=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_dmma --launch-skip 6 --launch-count 1 \
  -o gpurun_out/r02_update_full -f python tools/one_factor.py 65536 256 > gpurun_out/r02_update_full.out 2>&1
echo "ncu update rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bsweep --launch-count 1 \
  -o gpurun_out/r02_bsweep_full -f python tools/one_factor.py 65536 256 solve > gpurun_out/r02_bsweep_full.out 2>&1
echo "ncu bsweep rc=$?"
ls -la gpurun_out/*.ncu-rep
