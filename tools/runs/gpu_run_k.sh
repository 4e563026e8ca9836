#!/bin/bash
# single-GPU call K: early operand request in the tile kernel: parity subset + A/B at c5 / c3 / c2
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tile_kernel_rank3 or baseline_sizes or block_update_on_the_dmma or families or golden" > gpurun_out/r02k_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02k_pytest.log
B="--no-cpu --no-dgemm --no-e2e"
for e in 0 1 0 1; do
  MS_TILE_EARLY=$e timeout 600 python bench.py --steps 3 --warmup 2 $B 2> /dev/null | tail -1 > gpurun_out/r02k_c5_e$e.json; echo "c5 early$e rc=$?"
  python -c "import json; d=json.load(open('gpurun_out/r02k_c5_e$e.json')); print('c5 early=$e factor', round(d['factor_ms'],2), 'frac', round(d['roofline']['frac'],3), 'parity', d['parity']['ok'])"
done
for e in 0 1; do
  MS_TILE_EARLY=$e timeout 600 python bench.py --config c3 --steps 5 --warmup 3 $B 2> /dev/null | tail -1 > gpurun_out/r02k_c3_e$e.json
  MS_TILE_EARLY=$e timeout 600 python bench.py --config c2 --steps 10 --warmup 3 $B 2> /dev/null | tail -1 > gpurun_out/r02k_c2_e$e.json
  python -c "
import json
for c in ('c3','c2'):
    d=json.load(open('gpurun_out/r02k_%s_e$e.json' % c)); print(c, 'early=$e factor', round(d['factor_ms'],2), 'TF', round(d['value']/1e3,2), 'frac', round(d['roofline']['frac'],3), 'parity', d['parity']['ok'])"
done
