#!/bin/bash
cd "$(dirname "$0")/../.."
for flags in "--hp-stream" "--hp-stream --prefetch"; do
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $flags > gpurun_out/r2_bench_ad.log 2>&1; grep '^{' gpurun_out/r2_bench_ad.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('[$flags] value %.1f  %.2f ms/step e2e %.1f (%.2f ms)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']), d['config']['timing'])" || tail -20 gpurun_out/r2_bench_ad.log
done
