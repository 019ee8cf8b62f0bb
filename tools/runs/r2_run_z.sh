#!/bin/bash
cd "$(dirname "$0")/../.."
: > gpurun_out/r2_conv_bench_z.log
for cfg in "8 0.84 96" "2 0.84 96" "2 0.84 128" "1 0.45 128" "1 0.45 256"; do
  set -- $cfg
  timeout 300 python tools/conv_bench.py --scenes $1 --scale $2 --cin $3 --cout $3 --which dgrad --gather cpasync 2>&1 | grep -v "^rows" >> gpurun_out/r2_conv_bench_z.log
done
cat gpurun_out/r2_conv_bench_z.log
