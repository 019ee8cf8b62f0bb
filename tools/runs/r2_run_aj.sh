#!/bin/bash
cd "$(dirname "$0")/../.."
for t in 1 2 4; do
timeout 400 python bench.py --workload eval --steps 10 --warmup 3 --decode-threads $t > gpurun_out/r2_bench_aj_eval.log 2>&1; grep '^{' gpurun_out/r2_bench_aj_eval.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('threads $t eval: value %.1f  %.2f ms/step e2e %.1f stages %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['stage_ms']))" || tail -20 gpurun_out/r2_bench_aj_eval.log
done
