#!/bin/bash
cd "$(dirname "$0")/../.."
timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -q -x > gpurun_out/r2_pytest_x.log 2>&1; tail -4 gpurun_out/r2_pytest_x.log | cut -c1-400
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_x.log 2>&1; grep '^{' gpurun_out/r2_bench_x.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('value %.1f  %.2f ms/step e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']), d['config']['timing']); print('roofline', d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], 'launches', d['gpu_launches'])" || tail -20 gpurun_out/r2_bench_x.log
timeout 300 python tools/profile_step.py --cprofile 5 > gpurun_out/r2_profile_step_x.log 2>&1; grep "host returned" gpurun_out/r2_profile_step_x.log | head -3; tail -8 gpurun_out/r2_profile_step_x.log
