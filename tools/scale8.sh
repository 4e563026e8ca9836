#!/bin/bash
# 8-GPU factor time for a few MS_NCCL_CTAS settings (bench.py, no e2e)
for C in 0 2 4 8; do
  MS_NCCL_CTAS=$C python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2952$C bench.py --gpus 8 --steps 2 --warmup 2 --no-e2e 2>/dev/null | grep '^{' | python -c "
import sys, json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('MS_NCCL_CTAS=$C', f\"factor={d['factor_ms']:.1f}ms value={d['value']/1e3:.1f}TF fwd={d['forward_ms']:.1f} bwd={d['backward_ms']:.1f} upd_ms={r['kernel_ms_per_step']:.1f}\")
"
done
