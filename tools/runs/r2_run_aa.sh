#!/bin/bash
cd "$(dirname "$0")/../.."
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_net.py -m gpu -q -x > gpurun_out/r2_pytest_aa.log 2>&1; tail -3 gpurun_out/r2_pytest_aa.log | cut -c1-400
timeout 300 python tools/conv_bench.py --scenes 8 --cin 96 --cout 96 --which dgrad --gather cpasync 2>&1 | grep -v "^rows" > gpurun_out/r2_conv_bench_aa.log
timeout 300 python tools/conv_bench.py --scenes 1 --scale 0.45 --cin 128 --cout 128 --which dgrad --gather cpasync 2>&1 | grep -v "^rows" >> gpurun_out/r2_conv_bench_aa.log
cat gpurun_out/r2_conv_bench_aa.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_aa.log 2>&1; grep '^{' gpurun_out/r2_bench_aa.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('value %.1f  %.2f ms/step e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']), d['config']['timing']); print('roofline', d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], 'launches', d['gpu_launches']); print({k: round(v['ms_per_step'],2) for k,v in list(d['kernels'].items())[:8]})" || tail -20 gpurun_out/r2_bench_aa.log
