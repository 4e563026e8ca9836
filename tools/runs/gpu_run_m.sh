#!/bin/bash
# single-GPU call M: rolled TRSM panel kernel (parity + timing at the chain-bound sizes), BNORM batch size A/B
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "trsm or golden or formula_sweep or families or baseline_sizes or fused or pivots" > gpurun_out/r02m_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02m_pytest.log
B="--no-cpu --no-dgemm --no-e2e"
for rb in 2 4; do
  MS_SYM_RB=$rb timeout 600 python bench.py --steps 3 --warmup 2 $B 2> /dev/null | tail -1 > gpurun_out/r02m_c5_rb$rb.json
  python -c "import json; d=json.load(open('gpurun_out/r02m_c5_rb$rb.json')); print('c5 rb=$rb factor', round(d['factor_ms'],2), 'bnorm', round(d['bnorm_ms'],3), 'bwd', round(d['backward_ms'],3), 'parity', d['parity']['ok'])"
done
for c in c2 c3 c4; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 $B 2> /dev/null | tail -1 > gpurun_out/r02m_$c.json
  python -c "import json; d=json.load(open('gpurun_out/r02m_$c.json')); print('$c factor', round(d['factor_ms'],2), 'TF', round(d['value']/1e3,2), 'frac', round(d['roofline']['frac'],3), 'parity', d['parity']['ok'], d['parity']['r_max_rel'])"
done
timeout 300 ncu --set full --clock-control none -k regex:"trsm_panel" --launch-skip 40 --launch-count 1 -o gpurun_out/r02_trsm_full -f python tools/one_factor.py 16384 256 > gpurun_out/r02_trsm_full.out 2>&1; echo "ncu trsm rc=$?"
