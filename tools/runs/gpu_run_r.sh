#!/bin/bash
# single-GPU call R: the quad TRSM as a loop (instruction-cache resident) against the unrolled one;
# left-looking rule — all single-GPU tests, the TRSM A/B on the chain-bound configuration, every configuration once
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02r_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02r_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
show() {
python - "$@" <<'PY'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d['roofline']
        print(f.split('/')[-1], 'N', d['n_gpus'], 'factor', round(d['factor_ms'], 3), 'TF', round(d['value'] / 1e3, 2), 'blk', d['config'].get('block_internal'),
              'fwd', round(d['forward_ms'], 3), 'bwd', round(d['backward_ms'], 3), 'bnorm', round(d['bnorm_ms'], 3), 'upd union', round(r.get('kernel_ms_per_step') or 0, 2),
              'frac', round(r['frac'], 3), 'tts', round(d['time_to_solution_ms'], 2), 'parity', d['parity']['ok'], d['parity'].get('r_max_rel'),
              'e2e', (d.get('e2e') or {}).get('value'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
}
B="--no-cpu --no-dgemm"
for cfg in c2 c3; do
  timeout 300 python bench.py --config $cfg --steps 5 --warmup 3 --e2e-steps 5 $B 2> gpurun_out/r02r_${cfg}.err | tail -1 > gpurun_out/r02r_${cfg}.json
  MS_TRSM_QUAD=1 timeout 300 python bench.py --config $cfg --steps 5 --warmup 3 --no-e2e $B 2> /dev/null | tail -1 > gpurun_out/r02r_${cfg}_trsm_unrolled.json
  show gpurun_out/r02r_${cfg}.json gpurun_out/r02r_${cfg}_trsm_unrolled.json
done
timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e $B 2> gpurun_out/r02r_c5_n1.err | tail -1 > gpurun_out/r02r_c5_n1.json
show gpurun_out/r02r_c5_n1.json
# isolated duration of every launch of one small factorisation (TRSM and diag on the chain)
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02r_ncu_c2_launches.csv \
  python tools/one_factor.py 4096 64 > gpurun_out/r02r_ncu_c2.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r02r_ncu_c2_launches.csv', errors='replace')) if len(r) > 10]
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
tot = sum(a[1] for a in agg.values())
print('c2 one_factor launches: total us', round(tot, 1))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{k:62s} n={a[0]:5d} sum={a[1]:9.1f} us mean={a[1]/a[0]:7.2f} max={a[2]:7.2f}')
PY
