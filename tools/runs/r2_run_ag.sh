#!/bin/bash
cd "$(dirname "$0")/../.."
timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -q -x > gpurun_out/r2_pytest_ag.log 2>&1; tail -2 gpurun_out/r2_pytest_ag.log | cut -c1-300
for flags in "" "--flush-every 4" "--flush-every 24"; do
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $flags > gpurun_out/r2_bench_ag.log 2>&1; grep '^{' gpurun_out/r2_bench_ag.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('[$flags] value %.1f  %.2f ms/step e2e %.1f (%.2f ms)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']), d['config']['timing'])" || tail -20 gpurun_out/r2_bench_ag.log
done
timeout 300 python tools/profile_step.py > gpurun_out/r2_profile_step_ag.log 2>&1; grep "host returned" gpurun_out/r2_profile_step_ag.log | head -2; tail -7 gpurun_out/r2_profile_step_ag.log
