#!/bin/bash
cd "$(dirname "$0")/../.."
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x > gpurun_out/r2_pytest_l.log 2>&1; tail -3 gpurun_out/r2_pytest_l.log | cut -c1-200
timeout 300 python tools/conv_bench.py --cin 96 --cout 96 --which wgrad --gather tma,cpasync_all 2>&1 | grep -v "^rows" | tee gpurun_out/r2_conv_bench_l.log
timeout 300 python tools/profile_step.py --dump gpurun_out/r2_launches_l.json > gpurun_out/r2_profile_step_l.log 2>&1; head -5 gpurun_out/r2_profile_step_l.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_l.log 2>&1; tail -c 800 gpurun_out/r2_bench_l.log
timeout 300 python -m pytest tests/test_gpu_net.py -q -k well_conditioned > gpurun_out/r2_pytest_l_net.log 2>&1; tail -3 gpurun_out/r2_pytest_l_net.log | cut -c1-300
