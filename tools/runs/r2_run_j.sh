#!/bin/bash
cd "$(dirname "$0")/../.."
for lib in "" build/libb2m_leanaddr.so build/libb2m_t2.so; do
  echo "=== lib: ${lib:-default}" >> gpurun_out/r2_conv_bench_j.log
  for shape in "96 96" "128 96" "64 64"; do
    set -- $shape
    B2M_BENCH_LIB=$lib timeout 300 python tools/conv_bench.py --cin $1 --cout $2 --which fwd --gather cpasync 2>&1 | grep -v "^rows" >> gpurun_out/r2_conv_bench_j.log
  done
  echo "--- 1 scene x 0.6 (L2-size), 2 scenes (L1-size)" >> gpurun_out/r2_conv_bench_j.log
  B2M_BENCH_LIB=$lib timeout 300 python tools/conv_bench.py --scenes 1 --scale 0.6 --cin 128 --cout 128 --which fwd --gather cpasync --iters 30 2>&1 >> gpurun_out/r2_conv_bench_j.log
  B2M_BENCH_LIB=$lib timeout 300 python tools/conv_bench.py --scenes 2 --cin 96 --cout 96 --which fwd --gather cpasync --iters 30 2>&1 >> gpurun_out/r2_conv_bench_j.log
done
echo "=== wgrad dY ring depth (TMA)" >> gpurun_out/r2_conv_bench_j.log
timeout 300 python tools/conv_bench.py --cin 96 --cout 96 --which wgrad --gather tma --wgrows 64 --bslots 0,2,3,6 2>&1 | grep -v "^rows" >> gpurun_out/r2_conv_bench_j.log
cat gpurun_out/r2_conv_bench_j.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_j.log 2>&1; tail -6 gpurun_out/r2_pytest_j.log | cut -c1-300
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_j.log 2>&1; tail -c 1200 gpurun_out/r2_bench_j.log
