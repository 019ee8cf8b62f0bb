#!/bin/bash
# last sanity pass of the round on the committed tree: smoke(), the whole GPU test suite, the default bench line, the reference arm
cd "$(dirname "$0")/../.."
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2f_smoke.log 2>&1; tail -2 gpurun_out/r2f_smoke.log | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest_sanity.log 2>&1; tail -2 gpurun_out/r2f_pytest_sanity.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/r2f_bench_default.log 2>&1; grep '^{' gpurun_out/r2f_bench_default.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('default bench: value %.1f  %.2f ms/step e2e %.1f frac %.4f steps %d warmup %d cpu %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['steps'], d['warmup'], d['cpu_baseline']['value']), d['roofline']['traffic_source'])" || tail -20 gpurun_out/r2f_bench_default.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_reference_arm2.log 2>&1; tail -c 200 gpurun_out/r2f_bench_reference_arm2.log
