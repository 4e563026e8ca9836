#!/bin/bash
# 2-GPU call Q: every GPU test (incl. the 2-rank ones), smoke and the headline at N = 2 on HEAD
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02q_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02q_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 2 --master-port 29591 bench.py --gpus 2 --steps 3 --warmup 3 --e2e-steps 2 --no-cpu --no-dgemm 2> gpurun_out/r02q_c5_n2.err | tail -1 > gpurun_out/r02q_c5_n2.json
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02q_c5_n2.json').read().strip().splitlines()[-1]); r = d['roofline']
print('c5 N', d['n_gpus'], 'factor', round(d['factor_ms'], 2), 'TF', round(d['value'] / 1e3, 2), 'fwd', round(d['forward_ms'], 3), 'bwd', round(d['backward_ms'], 3),
      'upd', round(r['kernel_ms_per_step'], 1), 'frac', round(r['frac'], 3), 'tts', round(d['time_to_solution_ms'], 2), 'parity', d['parity']['ok'], d['parity'].get('r_max_rel'),
      'e2e', (d.get('e2e') or {}).get('value'))
PY
