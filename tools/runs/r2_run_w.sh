#!/bin/bash
cd "$(dirname "$0")/../.."
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_w.log 2>&1; tail -6 gpurun_out/r2_pytest_w.log | cut -c1-400
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_w.log 2>&1; grep '^{' gpurun_out/r2_bench_w.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('value %.1f  %.2f ms/step e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']), d['config']['timing']); print('roofline', d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], 'launches', d['gpu_launches']); print({k: round(v['ms_per_step'],2) for k,v in list(d['kernels'].items())[:12]})" || tail -20 gpurun_out/r2_bench_w.log
timeout 300 python tools/profile_step.py > gpurun_out/r2_profile_step_w.log 2>&1; grep "host returned" gpurun_out/r2_profile_step_w.log | head -3; tail -8 gpurun_out/r2_profile_step_w.log
