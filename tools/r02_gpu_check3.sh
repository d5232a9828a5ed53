#!/bin/bash
# voxel-path change check: voxel / batch / key-frame parity tests, then the launch list of mapping-cycle steps
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/gputests_voxel.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_voxel.log
tail -15 gpurun_out/gputests_voxel.log
timeout 300 python tools/batch_cycle_step.py 32 100 6 > gpurun_out/cycle_plain.log 2>&1; tail -n 2 gpurun_out/cycle_plain.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 200 --csv --log-file gpurun_out/launches_cycle2.csv python tools/batch_cycle_step.py 32 100 3 > gpurun_out/ncu_cycle.log 2>&1
tail -n 3 gpurun_out/ncu_cycle.log
