#!/bin/bash
cd "$(dirname "$0")/../.."
timeout 300 python tools/split_debug.py > gpurun_out/r2_split_debug.log 2>&1; cat gpurun_out/r2_split_debug.log
F="tests/test_gpu_kernels.py tests/test_gpu_net.py"
timeout 900 python -m pytest $F -q > gpurun_out/r2_pytest_e.log 2>&1; tail -12 gpurun_out/r2_pytest_e.log; grep -n "worst gradient\|head cosines\|worst train" gpurun_out/r2_pytest_e.log | cut -c1-900
for shape in "96 96" "128 96" "256 256" "64 64"; do
  set -- $shape
  timeout 300 python tools/conv_bench.py --cin $1 --cout $2 --which wgrad --gather cpasync >> gpurun_out/r2_conv_bench_e.log 2>&1
done
cat gpurun_out/r2_conv_bench_e.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_e.log 2>&1; tail -c 1500 gpurun_out/r2_bench_e.log
timeout 400 python bench.py --workload eval --steps 10 --warmup 3 > gpurun_out/r2_bench_e_eval.log 2>&1; tail -c 2500 gpurun_out/r2_bench_e_eval.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_fwd -s 1 -c 1 -f -o gpurun_out/r2_ncu_fwd_96_lean python tools/conv_bench.py --cin 96 --cout 96 --which fwd --gather cpasync --iters 1 > gpurun_out/r2_ncu_fwd_e.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad -s 1 -c 1 -f -o gpurun_out/r2_ncu_wgrad_96_lpt python tools/conv_bench.py --cin 96 --cout 96 --which wgrad --gather cpasync --iters 1 > gpurun_out/r2_ncu_wgrad_e.log 2>&1
