#!/bin/bash
cd "$(dirname "$0")/../.."
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_r.log 2>&1; tail -4 gpurun_out/r2_pytest_r.log | cut -c1-300
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_r.log 2>&1; grep '^{' gpurun_out/r2_bench_r.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('value %.1f  %.2f ms/step e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']), d['config']['timing']); print({k: round(v['ms_per_step'],2) for k,v in list(d['kernels'].items())[:9]})"
timeout 300 python tools/profile_step.py > gpurun_out/r2_profile_step_r.log 2>&1; tail -8 gpurun_out/r2_profile_step_r.log
