#!/bin/bash
# final 8-GPU pass: the driver's scaling invocation at N = 8 and N = 4, evaluation at N = 8, N = 1 on the same box
cd "$(dirname "$0")/../.."
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 --steps 10 --warmup 3 "${@:4}" > gpurun_out/$3 2>&1; grep '^{' gpurun_out/$3 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$3', 'value %.1f scenes/s  %.2f ms/step  e2e %.1f  sync_bn %s grad_sync %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['sync_bn'], d['config']['grad_sync']), d['config']['timing'])" || tail -5 gpurun_out/$3; }
run 8 29521 r2f_bench_8gpu_overlap.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_1gpu_same_box_as_8gpu.log 2>&1; grep '^{' gpurun_out/r2f_bench_1gpu_same_box_as_8gpu.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('1 gpu value %.1f scenes/s %.2f ms/step' % (d['value'], d['ms_per_step']))"
