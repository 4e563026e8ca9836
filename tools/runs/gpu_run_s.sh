#!/bin/bash
# single-GPU call S: quad-loop TRSM with all global loads batched, against the one-thread-per-column kernel
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "trsm or sweep or golden or internal or families or dense" > gpurun_out/r02s_pytest_subset.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02s_pytest_subset.log
show() {
python - "$@" <<'PY'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d['roofline']
        print(f.split('/')[-1], 'factor', round(d['factor_ms'], 3), 'TF', round(d['value'] / 1e3, 2), 'upd union', round(r.get('kernel_ms_per_step') or 0, 2),
              'frac', round(r['frac'], 3), 'tts', round(d['time_to_solution_ms'], 2), 'parity', d['parity']['ok'], d['parity'].get('r_max_rel'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
}
B="--no-cpu --no-dgemm --no-e2e"
for cfg in c2 c3; do
  timeout 300 python bench.py --config $cfg --steps 5 --warmup 3 $B 2> /dev/null | tail -1 > gpurun_out/r02s_${cfg}_quad.json
  MS_TRSM_QUAD=0 timeout 300 python bench.py --config $cfg --steps 5 --warmup 3 $B 2> /dev/null | tail -1 > gpurun_out/r02s_${cfg}_percol.json
  show gpurun_out/r02s_${cfg}_quad.json gpurun_out/r02s_${cfg}_percol.json
done
timeout 300 python bench.py --steps 3 --warmup 2 $B 2> /dev/null | tail -1 > gpurun_out/r02s_c5_quad.json
MS_TRSM_QUAD=0 timeout 300 python bench.py --steps 3 --warmup 2 $B 2> /dev/null | tail -1 > gpurun_out/r02s_c5_percol.json
show gpurun_out/r02s_c5_quad.json gpurun_out/r02s_c5_percol.json
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02s_ncu_c2_launches.csv \
  python tools/one_factor.py 4096 64 > gpurun_out/r02s_ncu_c2.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r02s_ncu_c2_launches.csv', errors='replace')) if len(r) > 10]
hdr = next(r for r in rows if 'Kernel Name' in r)
ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows:
    if r is hdr or r[0] == 'ID':
        continue
    try:
        v = float(r[iv].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r[iu], 1e-3)
    except ValueError:
        continue
    k = r[ik].split('(')[0][:60]
    a = agg.setdefault(k, [0, 0.0, 0.0])
    a[0] += 1; a[1] += v; a[2] = max(a[2], v)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:4]:
    print(f'{k:62s} n={a[0]:5d} sum={a[1]:9.1f} us mean={a[1]/a[0]:7.2f} max={a[2]:7.2f}')
PY
