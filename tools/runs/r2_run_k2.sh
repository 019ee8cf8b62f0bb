#!/bin/bash
# 2-GPU pass: SyncBatchNorm test, training bench with the gradient-sync modes and with SyncBatchNorm
cd "$(dirname "$0")/../.."
timeout 600 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_label_assoc.py -q > gpurun_out/r2_pytest_k2.log 2>&1; tail -3 gpurun_out/r2_pytest_k2.log | cut -c1-300
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 10 --warmup 3 "${@:3}" > gpurun_out/$2 2>&1; grep '^{' gpurun_out/$2 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$2', 'value %.1f scenes/s  %.2f ms/step  e2e %.1f  sync_bn %s grad_sync %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['sync_bn'], d['config']['grad_sync']))" || tail -5 gpurun_out/$2; }
run 29511 r2_bench_2gpu_overlap.log --grad-sync overlap
run 29512 r2_bench_2gpu_flat.log --grad-sync flat
run 29513 r2_bench_2gpu_overlap_syncbn.log --grad-sync overlap --sync-bn
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_1gpu_k.log 2>&1; grep '^{' gpurun_out/r2_bench_1gpu_k.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('1 gpu value %.1f scenes/s %.2f ms/step' % (d['value'], d['ms_per_step']))"
