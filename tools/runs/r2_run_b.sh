#!/bin/bash
cd "$(dirname "$0")/../.."
F="tests/test_gpu_kernels.py::test_conv_forward_offset_split_equals_unsplit tests/test_gpu_net.py::test_eval_forward_vs_reference_golden_and_oracle tests/test_gpu_net.py::test_variant_configs_vs_reference_golden_and_oracle tests/test_gpu_net.py::test_train_mode_whole_network_well_conditioned"
B2M_GATHER_MODE=tma timeout 600 python -m pytest $F -q > gpurun_out/r2_pytest_b_tma.log 2>&1; tail -12 gpurun_out/r2_pytest_b_tma.log
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -q > gpurun_out/r2_pytest_b.log 2>&1; tail -12 gpurun_out/r2_pytest_b.log
for shape in "96 96" "128 96" "256 256" "64 64" "32 32"; do
  set -- $shape
  timeout 300 python tools/conv_bench.py --cin $1 --cout $2 --which wgrad --gather cpasync >> gpurun_out/r2_conv_bench_b.log 2>&1
done
cat gpurun_out/r2_conv_bench_b.log
timeout 300 python tools/profile_step.py --dump gpurun_out/r2_launches_b.json > gpurun_out/r2_profile_step_b.log 2>&1; head -60 gpurun_out/r2_profile_step_b.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_fwd -s 1 -c 1 -f -o gpurun_out/r2_ncu_fwd_96_cpasync python tools/conv_bench.py --cin 96 --cout 96 --which fwd --gather cpasync --iters 1 > gpurun_out/r2_ncu_fwd_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad -s 1 -c 1 -f -o gpurun_out/r2_ncu_wgrad_96_cpasync python tools/conv_bench.py --cin 96 --cout 96 --which wgrad --gather cpasync --iters 1 > gpurun_out/r2_ncu_wgrad_b.log 2>&1
ls -la gpurun_out/*.ncu-rep
