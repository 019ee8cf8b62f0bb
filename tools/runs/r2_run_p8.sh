#!/bin/bash
# 8-GPU pass for the NVLink peer exchange: the exchange test on all 8 GPUs, then the SyncBatchNorm training bench
cd "$(dirname "$0")/../.."
timeout 600 python -m pytest tests/test_gpu_distributed.py -q -x -k peer > gpurun_out/r2_pytest_p8.log 2>&1; tail -4 gpurun_out/r2_pytest_p8.log | cut -c1-300
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 10 --warmup 3 "${@:3}" > gpurun_out/$2 2>&1; grep '^{' gpurun_out/$2 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$2', 'value %.1f scenes/s  %.2f ms/step  e2e %.1f  sync_bn %s grad_sync %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['sync_bn'], d['config']['grad_sync']), d['config']['timing'])" || tail -15 gpurun_out/$2; }
run 29513 r2f_bench_8gpu_overlap_syncbn_peer.log --grad-sync overlap --sync-bn
