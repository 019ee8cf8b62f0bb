#!/bin/bash
cd "$(dirname "$0")/../.."
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_m.log 2>&1; tail -6 gpurun_out/r2_pytest_m.log | cut -c1-300
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_m.log 2>&1; tail -c 700 gpurun_out/r2_bench_m.log
