#!/bin/bash
cd "$(dirname "$0")/../.."
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_d.log 2>&1; tail -12 gpurun_out/r2_pytest_d.log; grep -n "train-mode loss terms\|worst gradient\|head cosines" gpurun_out/r2_pytest_d.log | cut -c1-600
for shape in "96 96" "128 96" "256 256" "64 64" "32 32"; do
  set -- $shape
  timeout 300 python tools/conv_bench.py --cin $1 --cout $2 --which wgrad --gather cpasync --wgrows 0,64 >> gpurun_out/r2_conv_bench_d.log 2>&1
done
cat gpurun_out/r2_conv_bench_d.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_d.log 2>&1; tail -c 1800 gpurun_out/r2_bench_d.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-wgrad-overlap > gpurun_out/r2_bench_d_noverlap.log 2>&1; tail -c 600 gpurun_out/r2_bench_d_noverlap.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-trunk-executor > gpurun_out/r2_bench_d_modules.log 2>&1; tail -c 600 gpurun_out/r2_bench_d_modules.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad -s 1 -c 1 -f -o gpurun_out/r2_ncu_wgrad_96_d python tools/conv_bench.py --cin 96 --cout 96 --which wgrad --gather cpasync --iters 1 > gpurun_out/r2_ncu_wgrad_d.log 2>&1
