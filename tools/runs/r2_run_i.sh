#!/bin/bash
cd "$(dirname "$0")/../.."
for lib in "" build/libb2m_nopf.so build/libb2m_g.so; do
  echo "=== lib: ${lib:-default (rolled loops + prefetch code)}" >> gpurun_out/r2_conv_bench_i.log
  for shape in "96 96" "128 96"; do
    set -- $shape
    B2M_BENCH_LIB=$lib timeout 300 python tools/conv_bench.py --cin $1 --cout $2 --which fwd,wgrad --gather cpasync,tma 2>&1 | grep -v "^rows" >> gpurun_out/r2_conv_bench_i.log
  done
done
echo "=== repeat default" >> gpurun_out/r2_conv_bench_i.log
timeout 300 python tools/conv_bench.py --cin 96 --cout 96 --which fwd,wgrad --gather cpasync,tma 2>&1 | grep -v "^rows" >> gpurun_out/r2_conv_bench_i.log
cat gpurun_out/r2_conv_bench_i.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x > gpurun_out/r2_pytest_i.log 2>&1; tail -3 gpurun_out/r2_pytest_i.log | cut -c1-200
timeout 300 python -m pytest tests/test_gpu_net.py -q -k well_conditioned > gpurun_out/r2_pytest_i_net.log 2>&1; tail -3 gpurun_out/r2_pytest_i_net.log | cut -c1-300; grep -n "median gradient cosine\|^E  " gpurun_out/r2_pytest_i_net.log | cut -c1-400 | head
