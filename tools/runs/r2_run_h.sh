#!/bin/bash
cd "$(dirname "$0")/../.."
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -q > gpurun_out/r2_pytest_h.log 2>&1; tail -5 gpurun_out/r2_pytest_h.log | cut -c1-300
for shape in "96 96" "128 96" "64 64"; do
  set -- $shape
  timeout 300 python tools/conv_bench.py --cin $1 --cout $2 --which fwd,wgrad --gather cpasync,tma --wgrows 0,64 --wggroup 0,2 --l2pf 0,16,40 >> gpurun_out/r2_conv_bench_h.log 2>&1
done
cat gpurun_out/r2_conv_bench_h.log
timeout 900 python -m pytest tests/test_gpu_net.py -q > gpurun_out/r2_pytest_h_net.log 2>&1; tail -8 gpurun_out/r2_pytest_h_net.log | cut -c1-300; grep -n "median gradient cosine\|worst outside" gpurun_out/r2_pytest_h_net.log | cut -c1-600
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_h.log 2>&1; tail -c 1200 gpurun_out/r2_bench_h.log
