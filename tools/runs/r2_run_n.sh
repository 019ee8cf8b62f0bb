#!/bin/bash
cd "$(dirname "$0")/../.."
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py tests/test_gpu_label_assoc.py -q > gpurun_out/r2_pytest_n.log 2>&1; tail -5 gpurun_out/r2_pytest_n.log | cut -c1-300
for shape in "96 96" "128 96" "64 64" "32 32"; do
  set -- $shape
  timeout 300 python tools/conv_bench.py --cin $1 --cout $2 --which wgrad --gather cpasync --wgrows 0,64 2>&1 | grep -v "^rows" >> gpurun_out/r2_conv_bench_n.log
done
cat gpurun_out/r2_conv_bench_n.log
timeout 300 python tools/profile_step.py --dump gpurun_out/r2_launches_n.json > gpurun_out/r2_profile_step_n.log 2>&1; head -4 gpurun_out/r2_profile_step_n.log; tail -8 gpurun_out/r2_profile_step_n.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n.log 2>&1; tail -c 600 gpurun_out/r2_bench_n.log
