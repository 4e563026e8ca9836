#!/bin/bash
# single-GPU call L: final validation of HEAD — GPU suite, smoke, bench lines of every BASELINE configuration,
# launch list of the bench command, ncu captures of the dominant kernels
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02l_pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02l_pytest_gpu.log
tail -5 gpurun_out/r02l_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02l_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02l_smoke.log
timeout 900 python bench.py > gpurun_out/r02l_c5.json 2> gpurun_out/r02l_c5.err; echo "c5 rc=$?"
for cfg in c2 c3 c4; do
  timeout 900 python bench.py --config $cfg --steps 5 --warmup 3 --no-dgemm > gpurun_out/r02l_$cfg.json 2> gpurun_out/r02l_$cfg.err; echo "$cfg rc=$?"
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02l_reference_arm.json 2> /dev/null; echo "ref arm rc=$?"
python - <<'PY'
import json
for cfg in ("c5", "c2", "c3", "c4"):
    try:
        d = json.loads(open(f"gpurun_out/r02l_{cfg}.json").read().strip().splitlines()[-1]); r = d["roofline"]
        print(cfg, "factor", round(d["factor_ms"], 2), "TF", round(d["value"] / 1e3, 2), "bwd", round(d["backward_ms"], 3),
              "bnorm", round(d["bnorm_ms"], 3), "tts", round(d["time_to_solution_ms"], 2), "frac", round(r["frac"], 3), "peak", round(r["peak"], 2),
              "parity", d["parity"]["ok"], "e2e", (d.get("e2e") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(cfg, "no line", e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02_launches_c5.csv \
  python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --no-dgemm > gpurun_out/r02_launches_c5.out 2>&1
echo "ncu launch list rc=$?"
MS_LOOKAHEAD=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_dmma --launch-skip 6 --launch-count 1 \
  -o gpurun_out/r02_update_full -f python tools/one_factor.py 65536 256 > gpurun_out/r02_update_full.out 2>&1
echo "ncu update rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bsweep_kernel|sym_onepass" --launch-count 2 \
  -o gpurun_out/r02_bsweep_sym_full -f python tools/one_factor.py 65536 256 solve > gpurun_out/r02_bsweep_sym_full.out 2>&1
echo "ncu bsweep+sym rc=$?"
