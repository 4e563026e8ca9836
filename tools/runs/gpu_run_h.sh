#!/bin/bash
# 4-GPU gpurun call H: scatter + all-gather panel transport (parity at P = 4, 3, 2; bench at N = 4), lean prologue at N = 1
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B="--no-cpu --no-dgemm --no-e2e"
timeout 900 $TR --nproc-per-node 4 --master-port 29571 tools/multi_gpu_check.py > gpurun_out/r02h_multi4.log 2>&1
echo "multi4 rc=$?"; grep -c '"ok": true' gpurun_out/r02h_multi4.log; grep '"ok": false' gpurun_out/r02h_multi4.log | cut -c1-300 | head -3; tail -2 gpurun_out/r02h_multi4.log | cut -c1-300
MS_MULTI_QUICK=1 timeout 600 $TR --nproc-per-node 3 --master-port 29572 tools/multi_gpu_check.py > gpurun_out/r02h_multi3.log 2>&1
echo "multi3 rc=$?"; grep -c '"ok": true' gpurun_out/r02h_multi3.log
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r02h_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/r02h_pytest_multi.log
timeout 600 $TR --nproc-per-node 4 --master-port 29573 bench.py --gpus 4 --steps 3 --warmup 2 $B > gpurun_out/r02h_c5_n4.json 2> gpurun_out/r02h_c5_n4.err; echo "n4 rc=$?"
timeout 600 $TR --nproc-per-node 4 --master-port 29574 bench.py --gpus 4 --config c4 --steps 5 --warmup 3 $B > gpurun_out/r02h_c4_n4.json 2> /dev/null; echo "c4 n4 rc=$?"
timeout 600 python bench.py --steps 3 --warmup 2 $B > gpurun_out/r02h_c5_n1.json 2> /dev/null; echo "n1 rc=$?"
timeout 600 python bench.py --config c3 --steps 5 --warmup 3 $B > gpurun_out/r02h_c3_n1.json 2> /dev/null; echo "c3 rc=$?"
timeout 600 python bench.py --config c2 --steps 10 --warmup 3 $B > gpurun_out/r02h_c2_n1.json 2> /dev/null; echo "c2 rc=$?"
python - <<'PY'
import json
for cfg in ("c5_n4", "c4_n4", "c5_n1", "c3_n1", "c2_n1"):
    try:
        d = json.loads(open(f"gpurun_out/r02h_{cfg}.json").read().strip().splitlines()[-1]); r = d["roofline"]
        print(cfg, "factor", round(d["factor_ms"], 2), "TF", round(d["value"] / 1e3, 2), "fwd", round(d["forward_ms"], 3), "bwd", round(d["backward_ms"], 3),
              "bnorm", round(d["bnorm_ms"], 3), "tts", round(d["time_to_solution_ms"], 2), "upd", round(r["kernel_ms_per_step"], 1), "frac", round(r["frac"], 3),
              "parity", d["parity"]["ok"], d["parity"]["r_max_rel"])
    except Exception as e:
        print(cfg, "no line", e)
PY
