#!/bin/bash
# 8-GPU gpurun call J: headline at N = 8 / 4 / 2 / 1 on ONE box with the scatter + all-gather panel transport, trace
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B="--no-cpu --no-dgemm"
timeout 600 $TR --nproc-per-node 8 --master-port 29581 bench.py --gpus 8 --steps 5 --warmup 3 --e2e-steps 3 $B > gpurun_out/r02j_c5_n8.json 2> gpurun_out/r02j_c5_n8.err; echo "n8 rc=$?"
timeout 600 $TR --nproc-per-node 8 --master-port 29582 tools/trace_factor.py 65536 256 > gpurun_out/r02j_trace_n8.log 2>&1; echo "trace rc=$?"
timeout 600 $TR --nproc-per-node 8 --master-port 29583 bench.py --gpus 8 --config c4 --steps 5 --warmup 3 --e2e-steps 3 $B > gpurun_out/r02j_c4_n8.json 2> /dev/null; echo "c4 n8 rc=$?"
timeout 600 $TR --nproc-per-node 4 --master-port 29584 bench.py --gpus 4 --steps 3 --warmup 3 --e2e-steps 2 $B > gpurun_out/r02j_c5_n4.json 2> /dev/null; echo "n4 rc=$?"
timeout 600 $TR --nproc-per-node 2 --master-port 29585 bench.py --gpus 2 --steps 3 --warmup 3 --e2e-steps 2 $B > gpurun_out/r02j_c5_n2.json 2> /dev/null; echo "n2 rc=$?"
timeout 600 python bench.py --steps 3 --warmup 3 --e2e-steps 2 $B > gpurun_out/r02j_c5_n1.json 2> /dev/null; echo "n1 rc=$?"
MS_MULTI_QUICK=1 timeout 600 $TR --nproc-per-node 8 --master-port 29586 tools/multi_gpu_check.py > gpurun_out/r02j_multi8.log 2>&1; echo "multi8 rc=$?"
python - <<'PY'
import json
for cfg in ("c5_n8", "c4_n8", "c5_n4", "c5_n2", "c5_n1"):
    try:
        d = json.loads(open(f"gpurun_out/r02j_{cfg}.json").read().strip().splitlines()[-1]); r = d["roofline"]
        print(cfg, "factor", round(d["factor_ms"], 2), "TF", round(d["value"] / 1e3, 2), "fwd", round(d["forward_ms"], 3), "bwd", round(d["backward_ms"], 3),
              "bnorm", round(d["bnorm_ms"], 3), "tts", round(d["time_to_solution_ms"], 2), "upd", round(r["kernel_ms_per_step"], 1), "frac", round(r["frac"], 3),
              "peak", round(r["peak"], 2), "parity", d["parity"]["ok"], d["parity"]["r_max_rel"], "e2e", (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(cfg, "no line", e)
PY
tail -3 gpurun_out/r02j_trace_n8.log
grep -c '"ok": true' gpurun_out/r02j_multi8.log
