#!/bin/bash
cd "$(dirname "$0")/../.."
for lib in "" build/libb2m_wgremap.so; do
  echo "=== lib: ${lib:-default}" >> gpurun_out/r2_conv_bench_q.log
  for shape in "96 96" "128 96"; do
    set -- $shape
    B2M_BENCH_LIB=$lib timeout 300 python tools/conv_bench.py --cin $1 --cout $2 --which wgrad --gather tma,cpasync_all --wggroup 0,2 2>&1 | grep -v "^rows" >> gpurun_out/r2_conv_bench_q.log
  done
done
cat gpurun_out/r2_conv_bench_q.log
B2M_LIB=build/libb2m_wgremap.so B2M_GATHER_MODE=cpasync_all timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -q -k "wgrad or conv" > gpurun_out/r2_pytest_q.log 2>&1; tail -3 gpurun_out/r2_pytest_q.log | cut -c1-200
