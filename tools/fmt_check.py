import sys, json
for l in sys.stdin:
    if not l.startswith('{'):
        continue
    d = json.loads(l); s = d["stats"]
    print(f"n={d['n']} m={d['m']} la={d['lookahead']} factor={s['factor_ms']:.1f}ms {d['gflops']/1e3:.2f}TF "
          f"upd={s['update_ms']:.1f}ms upd_tflops={d.get('update_tflops',0):.2f} fwd={s['forward_ms']:.1f} "
          f"bwd={s['backward_ms']:.1f} res={d.get('residual',-1):.2e} err={d.get('error',-1):.2e}")
