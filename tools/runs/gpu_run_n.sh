#!/bin/bash
# single-GPU call N: last check of HEAD
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02n_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02n_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 3 --warmup 3 --e2e-steps 2 --no-cpu 2> /dev/null | tail -1 > gpurun_out/r02n_c5.json
python -c "import json; d=json.load(open('gpurun_out/r02n_c5.json')); r=d['roofline']; print('c5 factor', round(d['factor_ms'],2), 'TF', round(d['value']/1e3,2), 'bnorm', round(d['bnorm_ms'],3), 'bwd', round(d['backward_ms'],3), 'frac', round(r['frac'],3), 'parity', d['parity']['ok'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])"
