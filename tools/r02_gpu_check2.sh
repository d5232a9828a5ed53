#!/bin/bash
# new-row tests (loop closure, sharded map, adapter), launch list of mapping-cycle steps (profiling starts at cuProfilerStart, after the
# set-up), full capture of the registration kernel of a real step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_loop.py tests/test_gpu_adapter.py tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/gputests_new.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_new.log
tail -30 gpurun_out/gputests_new.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 200 --csv --log-file gpurun_out/launches_cycle.csv python tools/batch_cycle_step.py 32 100 3 > gpurun_out/ncu_cycle.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:batch_lm_kernel -s 1 -c 1 -o gpurun_out/batch_lm_full -f python tools/batch_cycle_step.py 32 100 3 > gpurun_out/ncu_full.log 2>&1
tail -n 3 gpurun_out/ncu_cycle.log; tail -n 3 gpurun_out/ncu_full.log
