#!/bin/bash
cd "$(dirname "$0")/../.."
timeout 300 python tools/conv_bench.py --scenes 8 --cin 96 --cout 96 --which fwd,dgrad --gather cpasync 2>&1 | grep -v "^rows" > gpurun_out/r2_conv_bench_ab.log
timeout 300 python tools/conv_bench.py --scenes 2 --cin 96 --cout 96 --which fwd,dgrad --gather cpasync 2>&1 | grep -v "^rows" >> gpurun_out/r2_conv_bench_ab.log
cat gpurun_out/r2_conv_bench_ab.log
