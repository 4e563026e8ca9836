#!/bin/bash
# 2-GPU gpurun call F: whole GPU suite (multi-GPU tests included), 2-rank parity check, bench at N = 1 and 2
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r02f_pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02f_pytest_gpu.log
tail -8 gpurun_out/r02f_pytest_gpu.log
timeout 900 $TR --nproc-per-node 2 --master-port 29541 tools/multi_gpu_check.py > gpurun_out/r02f_multi2.log 2>&1
echo "multi_gpu_check rc=$?"; grep -c '"ok": true' gpurun_out/r02f_multi2.log; grep '"ok": false' gpurun_out/r02f_multi2.log | head -3
tail -4 gpurun_out/r02f_multi2.log | cut -c1-400
B="--steps 3 --warmup 2 --no-cpu --no-dgemm"
timeout 600 python bench.py $B --no-e2e > gpurun_out/r02f_c5_n1.json 2> gpurun_out/r02f_c5_n1.err; echo "n1 rc=$?"
timeout 600 $TR --nproc-per-node 2 --master-port 29542 bench.py --gpus 2 $B --e2e-steps 2 > gpurun_out/r02f_c5_n2.json 2> gpurun_out/r02f_c5_n2.err; echo "n2 rc=$?"
for T in 2 4; do
  MS_TILES_PER_CTA=$T timeout 600 $TR --nproc-per-node 2 --master-port 2955$T bench.py --gpus 2 $B --no-e2e > gpurun_out/r02f_c5_n2_tpc$T.json 2> /dev/null; echo "tpc$T rc=$?"
done
MS_P2P=0 timeout 600 $TR --nproc-per-node 2 --master-port 29546 bench.py --gpus 2 $B --no-e2e > gpurun_out/r02f_c5_n2_nccl.json 2> /dev/null; echo "nccl rc=$?"
python - <<'PY'
import json
for cfg in ("c5_n1", "c5_n2", "c5_n2_tpc2", "c5_n2_tpc4", "c5_n2_nccl"):
    try:
        d = json.loads(open(f"gpurun_out/r02f_{cfg}.json").read().strip().splitlines()[-1]); r = d["roofline"]
        print(cfg, "factor", round(d["factor_ms"], 2), "TF", round(d["value"] / 1e3, 2), "fwd", round(d["forward_ms"], 3), "bwd", round(d["backward_ms"], 3),
              "bnorm", round(d["bnorm_ms"], 3), "tts", round(d["time_to_solution_ms"], 2), "frac", round(r["frac"], 3),
              "parity", d["parity"]["ok"], d["parity"]["r_max_rel"], "e2e", (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(cfg, "no line", e)
PY
