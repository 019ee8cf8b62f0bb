#!/bin/bash
cd "$(dirname "$0")/../.."
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "nms or NMS or decode" > gpurun_out/r2_pytest_u.log 2>&1; tail -3 gpurun_out/r2_pytest_u.log | cut -c1-300
timeout 300 python tools/decode_bench.py > gpurun_out/r2_decode_bench_u.log 2>&1; tail -12 gpurun_out/r2_decode_bench_u.log
timeout 300 python tools/profile_step.py --cprofile 5 > gpurun_out/r2_host_profile_u.log 2>&1; grep -n "host returned" gpurun_out/r2_host_profile_u.log | head -3
