#!/bin/bash
cd "$(dirname "$0")/../.."
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_f.log 2>&1; tail -12 gpurun_out/r2_pytest_f.log | cut -c1-300; grep -n "worst gradient\|head cosines\|worst train" gpurun_out/r2_pytest_f.log | cut -c1-900
for shape in "96 96" "128 96" "256 256" "64 64"; do
  set -- $shape
  timeout 300 python tools/conv_bench.py --cin $1 --cout $2 --which wgrad --gather cpasync >> gpurun_out/r2_conv_bench_f.log 2>&1
done
cat gpurun_out/r2_conv_bench_f.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_f.log 2>&1; tail -c 1500 gpurun_out/r2_bench_f.log
timeout 400 python bench.py --workload eval --steps 10 --warmup 3 > gpurun_out/r2_bench_f_eval.log 2>&1; tail -c 2500 gpurun_out/r2_bench_f_eval.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad -s 1 -c 1 -f -o gpurun_out/r2_ncu_wgrad_96_lpt python tools/conv_bench.py --cin 96 --cout 96 --which wgrad --gather cpasync --iters 1 > gpurun_out/r2_ncu_wgrad_f.log 2>&1
