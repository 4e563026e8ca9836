#!/bin/bash
# quick perf sweep on the GPU box: env knobs x sizes; prints one line per run
for T in 1 4 16 32; do
  echo "== MS_TILES_PER_CTA=$T"
  MS_TILES_PER_CTA=$T MS_PROFILE_UPDATE=1 python tools/gpu_check.py --stage big 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); s=d['stats']
    print(f\"n={d['n']} m={d['m']} la={d['lookahead']} factor={s['factor_ms']:.1f}ms {d['gflops']/1e3:.2f}TF upd={s['update_ms']:.1f}ms upd_tflops={d.get('update_tflops',0):.2f}\")
"
done
