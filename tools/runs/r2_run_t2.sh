#!/bin/bash
cd "$(dirname "$0")/../.."
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_2gpu_balanced.log 2>&1; grep '^{' gpurun_out/r2_bench_2gpu_balanced.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('2 gpus value %.1f scenes/s  %.2f ms/step  e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']), d['config']['timing'], d['config']['scene_sampling'])" || tail -20 gpurun_out/r2_bench_2gpu_balanced.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.log 2>&1; tail -c 900 gpurun_out/r2_bench_reference_arm.log
