#!/bin/bash
# validation + evidence pass: tests, bench lines (train / prefetch / eval / reference arm), ncu launch list of a step,
# ncu --set full captures of the forward, wgrad, map, hash, pooling and NMS kernels
cd "$(dirname "$0")/../.."
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_o.log 2>&1; tail -4 gpurun_out/r2_pytest_o.log | cut -c1-300
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_o.log 2>&1; tail -c 400 gpurun_out/r2_bench_o.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --prefetch > gpurun_out/r2_bench_o_prefetch.log 2>&1; grep '^{' gpurun_out/r2_bench_o_prefetch.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('prefetch: value %.1f  %.2f ms/step e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
timeout 400 python bench.py --workload eval --steps 10 --warmup 3 > gpurun_out/r2_bench_o_eval.log 2>&1; grep '^{' gpurun_out/r2_bench_o_eval.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('eval: value %.1f  %.2f ms/step e2e %.1f stages %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['stage_ms']))"
timeout 300 python tools/decode_bench.py > gpurun_out/r2_decode_bench.log 2>&1; cat gpurun_out/r2_decode_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_ncu_step_launches.csv python tools/profile_step.py > gpurun_out/r2_ncu_step_run.log 2>&1; wc -l gpurun_out/r2_ncu_step_launches.csv
cap() { timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -f -o gpurun_out/$3 ${@:4} > gpurun_out/$3.log 2>&1; }
cap conv_fwd 1 r2_ncu_fwd_k27_96x96_L0_final python tools/conv_bench.py --cin 96 --cout 96 --which fwd --gather cpasync --iters 1
cap conv_wgrad 1 r2_ncu_wgrad_k27_96x96_L0_final python tools/conv_bench.py --cin 96 --cout 96 --which wgrad --gather cpasync --iters 1
cap kmap_from_coarse_rows 2 r2_ncu_kmap_from_coarse python tools/profile_step.py --warm 0
cap hash_insert 2 r2_ncu_hash_insert python tools/profile_step.py --warm 0
cap segsum 2 r2_ncu_segsum python tools/profile_step.py --warm 0
cap nms_scan 1 r2_ncu_nms_scan python tools/decode_bench.py --iters 2
cap mask_inter 1 r2_ncu_mask_inter python tools/decode_bench.py --iters 2
ls -la gpurun_out/r2_ncu_*final*.ncu-rep gpurun_out/r2_ncu_kmap* gpurun_out/r2_ncu_hash* gpurun_out/r2_ncu_seg* gpurun_out/r2_ncu_nms* gpurun_out/r2_ncu_mask* 2>&1 | cut -c1-150
