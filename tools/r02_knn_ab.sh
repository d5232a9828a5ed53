#!/bin/bash
# A/B of the iteration kernels of the batched engine (run on the GPU box): parity tests, throughput per variant, ncu capture
# usage: r02_knn_ab.sh workload "cfg1;cfg2;..." [ncu env]
set -x
mkdir -p gpurun_out
rm -f gpurun_out/ab_bench.log
python -m pytest tests/test_gpu_batch.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/ab_tests.log
cat gpurun_out/ab_tests.log
W=${1:-vlp16_100k}
IFS=';' read -ra CFGS <<< "${2:-LLB_BATCH_FUSED=0;LLB_BATCH_FUSED=1}"
for cfg in "${CFGS[@]}"; do
  echo "== $cfg" >> gpurun_out/ab_bench.log
  env $cfg python tools/batch_bench.py $W ${AB_B:-32} 20 2>&1 | tail -2 >> gpurun_out/ab_bench.log
done
cat gpurun_out/ab_bench.log
if [ -n "$3" ]; then
env $3 ncu --set full --clock-control none --import-source on -k regex:batch_lm_kernel -s 1 -c 2 -o gpurun_out/prof_knnfit -f python tools/batch_step.py $W 32 3 > gpurun_out/ncu_knnfit.log 2>&1
tail -3 gpurun_out/ncu_knnfit.log
fi
