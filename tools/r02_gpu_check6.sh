#!/bin/bash
# bounds-in-assembly check (parity / batch / sequence tests), then bench A/B on the grid of the registration kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py tests/test_gpu_sequence.py -x -q -m gpu > gpurun_out/gputests_voxel.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_voxel.log
tail -4 gpurun_out/gputests_voxel.log
timeout 300 python tools/batch_cycle_step.py 32 100 6 2>&1 | tail -n 1
for g in "" "LLB_ITER_GRID=600"; do
  echo "== bench $g"
  env $g timeout 500 python bench.py --steps 100 --warmup 5 --secondary 0 --cpu-sample 1 2> gpurun_out/bench_ab.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print(json.dumps({'value':round(d['value']), 'ms_per_step':round(d['ms_per_step'],3), 'e2e':round(d['e2e']['value']), 'roof_frac':round(r['frac'],4), 'ms_per_launch':round(r['ms_per_launch'],4), 'stages':r['stage_ms_per_step'], 'launches':d['gpu_launches']}))"
done | tee gpurun_out/bench_ab.log
