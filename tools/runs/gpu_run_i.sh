#!/bin/bash
# single-GPU call I: TMA-reduction epilogue: bit-identity tests and A/B at c5 / c3 / c2
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tile_kernel_rank3 or baseline_sizes or block_update_on_the_dmma" > gpurun_out/r02i_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02i_pytest.log
B="--no-cpu --no-dgemm --no-e2e"
for red in 0 1; do
  MS_TILE_RED=$red timeout 600 python bench.py --steps 3 --warmup 2 $B > gpurun_out/r02i_c5_red$red.json 2> /dev/null; echo "c5 red$red rc=$?"
  MS_TILE_RED=$red timeout 600 python bench.py --config c3 --steps 5 --warmup 3 $B > gpurun_out/r02i_c3_red$red.json 2> /dev/null; echo "c3 red$red rc=$?"
  MS_TILE_RED=$red timeout 600 python bench.py --config c2 --steps 10 --warmup 3 $B > gpurun_out/r02i_c2_red$red.json 2> /dev/null; echo "c2 red$red rc=$?"
done
python - <<'PY'
import json
for cfg in ("c5_red0", "c5_red1", "c3_red0", "c3_red1", "c2_red0", "c2_red1"):
    try:
        d = json.loads(open(f"gpurun_out/r02i_{cfg}.json").read().strip().splitlines()[-1]); r = d["roofline"]
        print(cfg, "factor", round(d["factor_ms"], 2), "TF", round(d["value"] / 1e3, 2), "upd", round(r["kernel_ms_per_step"], 1), "frac", round(r["frac"], 3),
              "peak", round(r["peak"], 2), "parity", d["parity"]["ok"], d["parity"]["r_max_rel"])
    except Exception as e:
        print(cfg, "no line", e)
PY
