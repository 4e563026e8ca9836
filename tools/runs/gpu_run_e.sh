#!/bin/bash
# single-GPU gpurun call E: GPU suite, the bench lines of c5 (full default line) / c2 / c3 / c4, ncu captures
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02e_pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02e_pytest_gpu.log
tail -8 gpurun_out/r02e_pytest_gpu.log
python tools/diag_timing.py > gpurun_out/r02_diag_timing.txt 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02e_c5.json 2> gpurun_out/r02e_c5.err; echo "c5 rc=$?"
MS_DIAG_STRIP=0 timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu --no-dgemm > gpurun_out/r02e_c5_olddiag.json 2> /dev/null; echo "c5 old diag rc=$?"
for cfg in c2 c3 c4; do
  timeout 900 python bench.py --config $cfg --steps 5 --warmup 3 --no-dgemm > gpurun_out/r02e_$cfg.json 2> gpurun_out/r02e_$cfg.err; echo "$cfg rc=$?"
done
MS_DIAG_STRIP=0 timeout 300 python bench.py --config c3 --steps 5 --warmup 3 --no-e2e --no-cpu --no-dgemm > gpurun_out/r02e_c3_olddiag.json 2> /dev/null; echo "c3 old diag rc=$?"
python - <<'PY'
import json
for cfg in ("c5", "c5_olddiag", "c2", "c3", "c3_olddiag", "c4"):
    try:
        d = json.load(open(f"gpurun_out/r02e_{cfg}.json")); r = d["roofline"]
        print(cfg, "factor", round(d["factor_ms"], 2), "TF", round(d["value"] / 1e3, 2), "bwd", round(d["backward_ms"], 3),
              "bnorm", round(d["bnorm_ms"], 3), "tts", round(d["time_to_solution_ms"], 2), "frac", round(r["frac"], 3),
              "parity", d["parity"]["ok"], "e2e", (d.get("e2e") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(cfg, "no line", e)
PY
MS_LOOKAHEAD=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_dmma --launch-skip 6 --launch-count 1 \
  -o gpurun_out/r02_update_full -f python tools/one_factor.py 65536 256 > gpurun_out/r02_update_full.out 2>&1
echo "ncu update rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bsweep|sym_onepass" --launch-count 2 \
  -o gpurun_out/r02_bsweep_sym_full -f python tools/one_factor.py 65536 256 solve > gpurun_out/r02_bsweep_sym_full.out 2>&1
echo "ncu bsweep+sym rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"diag_strip|trsm_panel" --launch-skip 40 --launch-count 2 \
  -o gpurun_out/r02_panel_full -f python tools/one_factor.py 16384 256 > gpurun_out/r02_panel_full.out 2>&1
echo "ncu panel rc=$?"
