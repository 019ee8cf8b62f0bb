#!/bin/bash
# first GPU pass of round 2: parity tests, conv micro-benchmarks in both gather modes, whole-step bench in both modes
cd "$(dirname "$0")/../.."
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_a.log 2>&1; tail -15 gpurun_out/r2_pytest_a.log
for shape in "96 96" "128 96" "256 256" "64 64"; do
  set -- $shape
  timeout 300 python tools/conv_bench.py --cin $1 --cout $2 --which fwd,wgrad >> gpurun_out/r2_conv_bench_a.log 2>&1
done
cat gpurun_out/r2_conv_bench_a.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_a_cpasync.log 2>&1; tail -c 3000 gpurun_out/r2_bench_a_cpasync.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --gather-mode tma > gpurun_out/r2_bench_a_tma.log 2>&1; tail -c 1500 gpurun_out/r2_bench_a_tma.log
