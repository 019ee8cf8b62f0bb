#!/bin/bash
# final 2-GPU pass: the whole GPU test suite (incl. the 2-GPU distributed tests), training bench overlap / SyncBatchNorm
cd "$(dirname "$0")/../.."
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_m2.log 2>&1; tail -4 gpurun_out/r2_pytest_m2.log | cut -c1-300
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 10 --warmup 3 "${@:3}" > gpurun_out/$2 2>&1; grep '^{' gpurun_out/$2 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$2', 'value %.1f scenes/s  %.2f ms/step  e2e %.1f  sync_bn %s grad_sync %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['sync_bn'], d['config']['grad_sync']), d['config']['timing'])" || tail -5 gpurun_out/$2; }
run 29511 r2f_bench_2gpu_overlap.log --grad-sync overlap
run 29513 r2f_bench_2gpu_overlap_syncbn.log --grad-sync overlap --sync-bn
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_1gpu_same_box_as_2gpu.log 2>&1; grep '^{' gpurun_out/r2f_bench_1gpu_same_box_as_2gpu.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('1 gpu value %.1f scenes/s %.2f ms/step' % (d['value'], d['ms_per_step']))"
