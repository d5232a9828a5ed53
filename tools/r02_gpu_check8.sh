#!/bin/bash
# keys+histogram fusion check (parity / batch tests), step time for the assembly CTA size, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py -x -q -m gpu > gpurun_out/gputests_voxel.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_voxel.log
tail -4 gpurun_out/gputests_voxel.log
for k in 256 1024 2048; do echo "LLB_ASM_PTS_PER_CTA=$k"; LLB_ASM_PTS_PER_CTA=$k timeout 300 python tools/batch_cycle_step.py 32 100 6 2>&1 | tail -n 1; done | tee gpurun_out/cycle_asm.log
LLB_ASM_PTS_PER_CTA=1024 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 200 --csv --log-file gpurun_out/launches_cycle6.csv python tools/batch_cycle_step.py 32 100 3 > gpurun_out/ncu_cycle.log 2>&1
tail -n 1 gpurun_out/ncu_cycle.log
