#!/bin/bash
# bench A/B on the number of batches the 64 sequences of a GPU are split over
mkdir -p gpurun_out
for nb in 2 4 1; do
  echo "== bench --batches $nb"
  timeout 500 python bench.py --steps 80 --warmup 5 --secondary 0 --cpu-sample 1 --batches $nb 2> gpurun_out/bench_ab.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print(json.dumps({'value':round(d['value']), 'ms_per_step':round(d['ms_per_step'],3), 'e2e':round(d['e2e']['value']), 'roof_frac':round(r['frac'],4), 'ms_per_launch':round(r['ms_per_launch'],4), 'stages':r['stage_ms_per_step'], 'launches':d['gpu_launches']}))"
done | tee gpurun_out/bench_batches.log
