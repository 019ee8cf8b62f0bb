#!/bin/bash
cd "$(dirname "$0")/../.."
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -q > gpurun_out/r2_pytest_g.log 2>&1; tail -5 gpurun_out/r2_pytest_g.log | cut -c1-300
for shape in "96 96" "128 96" "256 256" "64 64"; do
  set -- $shape
  timeout 300 python tools/conv_bench.py --cin $1 --cout $2 --which fwd,wgrad --gather cpasync >> gpurun_out/r2_conv_bench_g.log 2>&1
done
cat gpurun_out/r2_conv_bench_g.log
timeout 300 python tools/profile_step.py --dump gpurun_out/r2_launches_g.json > gpurun_out/r2_profile_step_g.log 2>&1; head -8 gpurun_out/r2_profile_step_g.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_g.log 2>&1; tail -c 1200 gpurun_out/r2_bench_g.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_fwd -s 1 -c 1 -f -o gpurun_out/r2_ncu_fwd_96_g python tools/conv_bench.py --cin 96 --cout 96 --which fwd --gather cpasync --iters 1 > gpurun_out/r2_ncu_fwd_g.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad -s 1 -c 1 -f -o gpurun_out/r2_ncu_wgrad_96_g python tools/conv_bench.py --cin 96 --cout 96 --which wgrad --gather cpasync --iters 1 > gpurun_out/r2_ncu_wgrad_g.log 2>&1
