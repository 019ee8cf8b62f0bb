#!/bin/bash
cd "$(dirname "$0")/../.."
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_net.py -m gpu -q -x -k "detection2mask or decode or nms or variant or reference_selection" > gpurun_out/r2_pytest_ai.log 2>&1; tail -2 gpurun_out/r2_pytest_ai.log | cut -c1-300
timeout 400 python bench.py --workload eval --steps 10 --warmup 3 > gpurun_out/r2_bench_ai_eval.log 2>&1; grep '^{' gpurun_out/r2_bench_ai_eval.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('eval: value %.1f  %.2f ms/step e2e %.1f stages %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['stage_ms']), d['config']['timing'])" || tail -20 gpurun_out/r2_bench_ai_eval.log
