#!/bin/bash
# s2m kernel change (interleaved queries for sharded maps) + bounds-in-assembly: parity / sharded / sequence tests, cycle launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_sequence.py -x -q -m gpu > gpurun_out/gputests_s2m.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_s2m.log
tail -4 gpurun_out/gputests_s2m.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 200 --csv --log-file gpurun_out/launches_cycle5.csv python tools/batch_cycle_step.py 32 100 3 > gpurun_out/ncu_cycle.log 2>&1
tail -n 1 gpurun_out/ncu_cycle.log
