#!/bin/bash
# bench.py value / roofline per kernel variant: r02_bench_ab.sh "cfg1;cfg2" [extra bench args]
mkdir -p gpurun_out; rm -f gpurun_out/bench_ab.log
IFS=';' read -ra CFGS <<< "$1"
for cfg in "${CFGS[@]}"; do
  echo "== $cfg $2" >> gpurun_out/bench_ab.log
  env $cfg timeout 400 python bench.py --steps 40 --warmup 5 --secondary 0 --cpu-sample 1 $2 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print(json.dumps({'value':round(d['value']), 'ms_per_step':round(d['ms_per_step'],3), 'e2e':round(d['e2e']['value']), 'roof_frac':round(r['frac'],4), 'ms_per_launch':round(r['ms_per_launch'],4), 'launches':d['gpu_launches'], 'lat':d.get('latency')}))" >> gpurun_out/bench_ab.log
done
cat gpurun_out/bench_ab.log
