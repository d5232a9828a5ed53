#!/bin/bash
# voxel-path check: parity tests, step time for three radix tile sizes, launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py -x -q -m gpu > gpurun_out/gputests_voxel.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_voxel.log
tail -4 gpurun_out/gputests_voxel.log
for k in 4096 8192 16384; do echo "LLB_RADIX_KEYS=$k"; LLB_RADIX_KEYS=$k timeout 300 python tools/batch_cycle_step.py 32 100 6 2>&1 | tail -n 1; done | tee gpurun_out/cycle_radix_keys.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 200 --csv --log-file gpurun_out/launches_cycle4.csv python tools/batch_cycle_step.py 32 100 3 > gpurun_out/ncu_cycle.log 2>&1
LLB_RADIX_KEYS=8192 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 200 --csv --log-file gpurun_out/launches_cycle4_8192.csv python tools/batch_cycle_step.py 32 100 3 > gpurun_out/ncu_cycle.log 2>&1
