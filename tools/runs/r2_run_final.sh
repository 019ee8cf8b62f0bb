#!/bin/bash
# final validation + evidence pass of round 2: tests, bench lines (train / eval / reference arm), decode bench, per-launch
# event table, ncu launch list of one step, ncu --set full captures of the forward, wgrad, finalize and NMS-scan kernels
cd "$(dirname "$0")/../.."
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; tail -3 gpurun_out/r2f_pytest.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench_train_1gpu.log 2>&1; grep '^{' gpurun_out/r2f_bench_train_1gpu.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('train: value %.1f  %.2f ms/step e2e %.1f frac %.4f kernel_ms %.2f cpu %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], d['cpu_baseline']['value']), d['config']['timing'])" || tail -20 gpurun_out/r2f_bench_train_1gpu.log
timeout 400 python bench.py --workload eval --steps 10 --warmup 3 > gpurun_out/r2f_bench_eval_1gpu.log 2>&1; grep '^{' gpurun_out/r2f_bench_eval_1gpu.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('eval: value %.1f  %.2f ms/step e2e %.1f stages %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['stage_ms']))" || tail -20 gpurun_out/r2f_bench_eval_1gpu.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_reference_arm.log 2>&1; tail -c 300 gpurun_out/r2f_bench_reference_arm.log
timeout 300 python tools/decode_bench.py > gpurun_out/r2f_decode_bench.log 2>&1; cat gpurun_out/r2f_decode_bench.log
timeout 300 python tools/profile_step.py --dump gpurun_out/r2f_launches.json > gpurun_out/r2f_profile_step.log 2>&1; grep "host returned" gpurun_out/r2f_profile_step.log | head -2; tail -7 gpurun_out/r2f_profile_step.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2f_ncu_step_launches.csv python tools/profile_step.py > gpurun_out/r2f_ncu_step_run.log 2>&1; wc -l gpurun_out/r2f_ncu_step_launches.csv
cap() { timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -f -o gpurun_out/$3 ${@:4} > gpurun_out/$3.log 2>&1; }
cap conv_fwd 1 r2f_ncu_fwd_k27_96x96_L0 python tools/conv_bench.py --cin 96 --cout 96 --which fwd --gather cpasync --iters 1
cap conv_wgrad 1 r2f_ncu_wgrad_k27_96x96_L0 python tools/conv_bench.py --cin 96 --cout 96 --which wgrad --gather cpasync --iters 1
cap conv_fwd 1 r2f_ncu_fwd_k27_32x32_L1 python tools/conv_bench.py --scenes 2 --cin 32 --cout 32 --which fwd --gather cpasync --iters 1
cap nms_scan 1 r2f_ncu_nms_scan python tools/decode_bench.py --iters 2
ls -la gpurun_out/r2f_ncu_*.ncu-rep 2>&1 | cut -c1-150
