#!/bin/bash
# 8-GPU box: evaluation at N = 8 (no communication) and the training step at N = 4, final build
cd "$(dirname "$0")/../.."
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 --steps 10 --warmup 3 "${@:4}" > gpurun_out/$3 2>&1; grep '^{' gpurun_out/$3 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$3', 'value %.1f scenes/s  %.2f ms/step  e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']), d['config']['timing'])" || tail -5 gpurun_out/$3; }
run 8 29523 r2f_bench_8gpu_eval.log --workload eval
run 4 29522 r2f_bench_4gpu_overlap.log
