#!/bin/bash
# sharded-map virtual-rank test, launch list of mapping-cycle steps (set-up launches skipped), full capture of the registration kernel of a real step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/gputests_sharded.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_sharded.log
tail -30 gpurun_out/gputests_sharded.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 604 -c 120 --csv --log-file gpurun_out/launches_cycle.csv python tools/batch_cycle_step.py 32 100 3 > gpurun_out/ncu_cycle.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:batch_lm_kernel -s 101 -c 1 -o gpurun_out/batch_lm_full -f python tools/batch_cycle_step.py 32 100 3 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_cycle.log gpurun_out/ncu_full.log
