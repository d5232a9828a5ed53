#!/bin/bash
# round-2 state check on the GPU box: every GPU test, smoke, the bench line, launch lists (timed region of the bench, one
# mapping-cycle step) and one full capture of the registration kernel.  Profilers attach at cuProfilerStart, i.e. after the
# set-up (hundreds of launches) of the two programs.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/gputests_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_final.log
tail -4 gpurun_out/gputests_final.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_final.log 2>&1; tail -n 2 gpurun_out/smoke_final.log
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
LLB_BENCH_CUPROF=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches_bench_final.csv python bench.py --steps 4 --warmup 3 --secondary 0 --cpu-sample 1 > gpurun_out/ncu_bench.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 200 --csv --log-file gpurun_out/launches_cycle_final.csv python tools/batch_cycle_step.py 32 100 3 > gpurun_out/ncu_cycle.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:batch_lm_kernel -s 1 -c 1 -o gpurun_out/batch_lm_full_final -f python tools/batch_cycle_step.py 32 100 3 > gpurun_out/ncu_full.log 2>&1
tail -n 2 gpurun_out/ncu_bench.log gpurun_out/ncu_full.log 2>/dev/null | cut -c1-300
