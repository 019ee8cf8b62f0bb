#!/bin/bash
cd "$(dirname "$0")/../.."
: > gpurun_out/r2_conv_bench_af.log
for cfg in "1 0.45 64 64" "1 0.45 128 128" "1 0.45 256 256" "1 0.22 256 256" "2 0.84 96 96" "2 0.84 128 96" "2 0.84 64 64"; do
  set -- $cfg
  timeout 300 python tools/conv_bench.py --scenes $1 --scale $2 --cin $3 --cout $4 --which wgrad --gather tma,cpasync_all 2>&1 | grep -v "^rows" >> gpurun_out/r2_conv_bench_af.log
done
cat gpurun_out/r2_conv_bench_af.log
