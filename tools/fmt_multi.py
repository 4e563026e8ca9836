import sys, json
for l in sys.stdin:
    if not l.startswith('{'):
        continue
    d = json.loads(l)
    keep = ("n", "m", "P", "fail", "tflops", "factor_ms", "R_rel", "x_rel", "residual", "A_eq", "b_eq", "d_ok")
    print({k: (round(v, 3) if isinstance(v, float) and abs(v) > 1e-3 else v) for k, v in d.items() if k in keep})
