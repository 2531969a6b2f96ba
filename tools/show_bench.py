#!/usr/bin/env python3
"""Print the interesting fields of a bench.py JSON line (file argument)."""
import json
import sys

for path in sys.argv[1:]:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(f"== {path}: N={d['n_gpus']} value={d['value']:.2f} {d['unit']} ms/step={d['ms_per_step']:.3f} scaling={d['scaling']} launches={d.get('gpu_launches')}")
    print(f"   e2e={d['e2e']['value']:.2f} h2d={d['e2e']['h2d_bytes_per_step']} d2h={d['e2e']['d2h_bytes_per_step']} clocks={d.get('clocks')}")
    if r:
        print(f"   single={d['config']['latency_ms_single_matvec']:.3f} ms err={d['config']['max_abs_err_vs_float64']:.2e}")
        print(f"   dominant={r['class']} frac={r['frac']:.3f} share={r['share_of_step']:.3f} traffic={r.get('traffic')}")
        for k, v in sorted(r["kernels"].items(), key=lambda kv: -kv[1]["ms_per_matvec"]):
            print(f"     {k:14s} {v['ms_per_matvec']:.3f} ms share {v['share_of_step']:.3f} launches {v['launches_per_matvec']:.0f} frac {v['frac']:.3f}")
        print(f"   matvec frac={r['matvec']['frac']:.3f} integer frac={r['integer']['frac']:.3f}")
    for k in ("cpu_baseline", "token", "tuned_split"):
        if d.get(k):
            print(f"   {k}: " + json.dumps({a: b for a, b in d[k].items() if a not in ('note', 'metric', 'sample')}))
