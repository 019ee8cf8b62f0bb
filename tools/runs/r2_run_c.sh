#!/bin/bash
cd "$(dirname "$0")/../.."
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -q > gpurun_out/r2_pytest_c.log 2>&1; tail -6 gpurun_out/r2_pytest_c.log
for shape in "96 96" "128 96" "256 256" "64 64"; do
  set -- $shape
  timeout 300 python tools/conv_bench.py --cin $1 --cout $2 --which fwd --gather cpasync,cpasync2,tma --issuer lean,general >> gpurun_out/r2_conv_bench_c.log 2>&1
done
timeout 300 python tools/conv_bench.py --cin 96 --cout 96 --which wgrad --gather cpasync,tma --wgrows 0,64 >> gpurun_out/r2_conv_bench_c.log 2>&1
cat gpurun_out/r2_conv_bench_c.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c.log 2>&1; tail -c 2500 gpurun_out/r2_bench_c.log
