#!/bin/bash
# centroid unroll check + ncu full capture of the first radix scatter pass of a mapping-cycle step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py -x -q -m gpu > gpurun_out/gputests_voxel.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_voxel.log
tail -4 gpurun_out/gputests_voxel.log
timeout 300 python tools/batch_cycle_step.py 32 100 6 > gpurun_out/cycle_plain.log 2>&1; tail -n 1 gpurun_out/cycle_plain.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 200 --csv --log-file gpurun_out/launches_cycle3.csv python tools/batch_cycle_step.py 32 100 3 > gpurun_out/ncu_cycle.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:radix_scatter -s 4 -c 1 -o gpurun_out/radix_scatter_full -f python tools/batch_cycle_step.py 32 100 3 > gpurun_out/ncu_full_rs.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:voxel_centroid -s 1 -c 1 -o gpurun_out/centroid_full -f python tools/batch_cycle_step.py 32 100 3 > gpurun_out/ncu_full_c.log 2>&1
tail -n 2 gpurun_out/ncu_full_rs.log
