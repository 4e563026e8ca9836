#!/bin/bash
# One gpurun call that decides whether the experimental kernels (tile variant 2, blocked
# diag kernel) can become the default:
#   1. their opt-in bit-identity tests,
#   2. the whole GPU suite with both variants switched on through the environment,
#   3. A/B timing of the four combinations at the headline size and at the small, chain-bound size.
# Usage:  gpurun --timeout 1500 -- 'bash tools/validate_experimental.sh'
# Results: gpurun_out/exp_*.log, gpurun_out/ab_*.json (one bench line each).
set -u
mkdir -p gpurun_out
MS_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k experimental \
  > gpurun_out/exp_optin_tests.log 2>&1
echo "opt-in tests rc=$?" | tee -a gpurun_out/exp_optin_tests.log
tail -3 gpurun_out/exp_optin_tests.log
# tile variant 2 once more with its 2-D TMA loads (the default above is the rank-3 tensor map)
MS_TILE_TMA3=0 MS_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu \
  -k experimental_tile > gpurun_out/exp_optin_tests_tma2d.log 2>&1
echo "opt-in tile tests, 2-D TMA rc=$?" | tee -a gpurun_out/exp_optin_tests_tma2d.log
tail -3 gpurun_out/exp_optin_tests_tma2d.log
MS_TILE_VARIANT=2 MS_DIAG_VARIANT=2 timeout 900 python -m pytest tests -x -q -m gpu \
  > gpurun_out/exp_full_suite.log 2>&1
echo "full suite with variants rc=$?" | tee -a gpurun_out/exp_full_suite.log
tail -3 gpurun_out/exp_full_suite.log
for combo in "1 1 1" "2 1 1" "2 1 0" "1 2 1" "2 2 1"; do       # tile variant, diag variant, rank-3 TMA
  for size in "65536 256" "4096 64"; do
    set -- $combo $size
    out=gpurun_out/ab_tile$1_diag$2_tma3$3_n$4
    MS_TILE_VARIANT=$1 MS_DIAG_VARIANT=$2 MS_TILE_TMA3=$3 timeout 300 python bench.py --n $4 --m $5 --steps 3 --warmup 2 \
      --no-e2e --no-cpu 2> $out.err | tail -1 > $out.json
    python - "$1" "$2" "$4" $out.json "$3" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[4]))
    r = d.get("roofline", {})
    print(f"tile={sys.argv[1]} diag={sys.argv[2]} tma3={sys.argv[5]} n={sys.argv[3]}: factor {d.get('factor_ms', float('nan')):.2f} ms, "
          f"{d['value'] / 1e3:.2f} TFLOP/s, update kernel frac {r.get('frac')}")
except Exception as e:
    print("no bench line for", sys.argv[1:4], e)
PY
  done
done
