#!/bin/bash
cd "$(dirname "$0")/../.."
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "conv or dgrad" > gpurun_out/r2_pytest_ae.log 2>&1; tail -3 gpurun_out/r2_pytest_ae.log | cut -c1-400
timeout 300 python tools/conv_bench.py --scenes 2 --cin 32 --cout 32 --which fwd --gather tma,cpasync 2>&1 | grep -v "^rows" > gpurun_out/r2_conv_bench_ae.log
timeout 300 python tools/conv_bench.py --scenes 8 --cin 32 --cout 32 --which fwd --gather tma,cpasync 2>&1 | grep -v "^rows" >> gpurun_out/r2_conv_bench_ae.log
timeout 300 python tools/conv_bench.py --scenes 2 --cin 32 --cout 64 --which fwd --gather tma,cpasync 2>&1 | grep -v "^rows" >> gpurun_out/r2_conv_bench_ae.log
cat gpurun_out/r2_conv_bench_ae.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_ae.log 2>&1; grep '^{' gpurun_out/r2_bench_ae.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('value %.1f  %.2f ms/step e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']), d['config']['timing']); print('roofline', d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], 'launches', d['gpu_launches']); print({k: round(v['ms_per_step'],2) for k,v in list(d['kernels'].items())[:8]})" || tail -20 gpurun_out/r2_bench_ae.log
