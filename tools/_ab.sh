for kb in 0 44 56 74; do
  echo "== ks smem pad $kb KB"
  SPEAR_KS_SMEM_KB=$kb python bench.py --steps 5 --warmup 3 --no-cpu-baseline --tuned-weight 8 > gpurun_out/pad_$kb.json 2> gpurun_out/pad_$kb.err
  tail -3 gpurun_out/pad_$kb.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/pad_$kb.json').read().strip().splitlines()[-1])
r=d['roofline']
print(d['value'], d['e2e']['value'], d['config']['latency_ms_single_matvec'], r['avg_launch_ms'], r['frac'])
t=d.get('tuned_split'); print(t['split'], t['value'], t['latency_ms_single_matvec'])
"
done
