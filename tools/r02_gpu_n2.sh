#!/bin/bash
# 2-GPU check: sharded registration test (query-sharded + map-sharded, fused exchange) and the bench line at N=2 with the `sharded` object
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/gputests_n2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests_n2.log
tail -15 gpurun_out/gputests_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 50 --warmup 5 --secondary 0 --cpu-sample 1 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
