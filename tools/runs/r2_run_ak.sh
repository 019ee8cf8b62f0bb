#!/bin/bash
cd "$(dirname "$0")/../.."
timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -q -x > gpurun_out/r2_pytest_ak.log 2>&1; tail -2 gpurun_out/r2_pytest_ak.log | cut -c1-300
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_ak.log 2>&1; grep '^{' gpurun_out/r2_bench_ak.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('value %.1f  %.2f ms/step e2e %.1f (%.2f ms)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']), d['config']['timing'])" || tail -20 gpurun_out/r2_bench_ak.log
