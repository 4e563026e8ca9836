#!/bin/bash
# 8-GPU gpurun call G: headline bench at N = 8 (and 4, 1 on the same box), transport A/B, per-step trace, c4 at N = 8
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B="--no-cpu --no-dgemm"
timeout 600 $TR --nproc-per-node 8 --master-port 29561 bench.py --gpus 8 --steps 5 --warmup 3 --e2e-steps 2 $B > gpurun_out/r02g_c5_n8.json 2> gpurun_out/r02g_c5_n8.err; echo "n8 rc=$?"
MS_P2P=0 timeout 600 $TR --nproc-per-node 8 --master-port 29562 bench.py --gpus 8 --steps 3 --warmup 2 --no-e2e $B > gpurun_out/r02g_c5_n8_nccl.json 2> /dev/null; echo "n8 nccl rc=$?"
MS_DIAG_STRIP=0 timeout 600 $TR --nproc-per-node 8 --master-port 29563 bench.py --gpus 8 --steps 3 --warmup 2 --no-e2e $B > gpurun_out/r02g_c5_n8_olddiag.json 2> /dev/null; echo "n8 olddiag rc=$?"
timeout 600 $TR --nproc-per-node 8 --master-port 29564 bench.py --gpus 8 --config c4 --steps 5 --warmup 3 --e2e-steps 2 $B > gpurun_out/r02g_c4_n8.json 2> /dev/null; echo "c4 n8 rc=$?"
timeout 600 $TR --nproc-per-node 4 --master-port 29565 bench.py --gpus 4 --steps 3 --warmup 2 --e2e-steps 2 $B > gpurun_out/r02g_c5_n4.json 2> /dev/null; echo "n4 rc=$?"
timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e $B > gpurun_out/r02g_c5_n1.json 2> /dev/null; echo "n1 rc=$?"
MS_TILES_PER_CTA=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e $B > gpurun_out/r02g_c5_n1_tpc1.json 2> /dev/null; echo "n1 tpc1 rc=$?"
timeout 600 $TR --nproc-per-node 8 --master-port 29566 tools/trace_factor.py 65536 256 > gpurun_out/r02g_trace_n8.log 2>&1; echo "trace rc=$?"
MS_MULTI_QUICK=1 timeout 600 $TR --nproc-per-node 8 --master-port 29567 tools/multi_gpu_check.py > gpurun_out/r02g_multi8.log 2>&1; echo "multi8 rc=$?"
python - <<'PY'
import json
for cfg in ("c5_n8", "c5_n8_nccl", "c5_n8_olddiag", "c4_n8", "c5_n4", "c5_n1", "c5_n1_tpc1"):
    try:
        d = json.loads(open(f"gpurun_out/r02g_{cfg}.json").read().strip().splitlines()[-1]); r = d["roofline"]
        print(cfg, "factor", round(d["factor_ms"], 2), "TF", round(d["value"] / 1e3, 2), "fwd", round(d["forward_ms"], 3), "bwd", round(d["backward_ms"], 3),
              "bnorm", round(d["bnorm_ms"], 3), "tts", round(d["time_to_solution_ms"], 2), "upd", round(r["kernel_ms_per_step"], 1), "frac", round(r["frac"], 3),
              "parity", d["parity"]["ok"], d["parity"]["r_max_rel"], "e2e", (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(cfg, "no line", e)
PY
tail -5 gpurun_out/r02g_trace_n8.log
grep -c '"ok": true' gpurun_out/r02g_multi8.log
