#!/bin/bash
# round-2 state check on the GPU box: tests, bench line, launch lists (bench + one mapping-cycle step), full capture of the registration kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests.log
tail -5 gpurun_out/gputests.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_cycle.csv python tools/batch_cycle_step.py 32 100 2 > gpurun_out/ncu_cycle.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --secondary 0 --cpu-sample 1 > gpurun_out/ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:batch_lm_kernel -s 2 -c 1 -o gpurun_out/batch_lm_full -f python tools/batch_cycle_step.py 32 100 2 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
