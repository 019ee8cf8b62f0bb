#!/bin/bash
# ncu --set full captures of the two kernels added last: the column copy of the launch lists and the offset-split finalize
# with the fused BatchNorm-backward reduction (skip counts chosen to land on full-resolution / deep-level backward launches)
cd "$(dirname "$0")/../.."
cap() { timeout 400 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -f -o gpurun_out/$3 ${@:4} > gpurun_out/$3.log 2>&1; }
cap copy_columns 60 r2f_ncu_copy_columns python tools/profile_step.py --warm 1
cap conv_finalize 200 r2f_ncu_conv_finalize_bnr python tools/profile_step.py --warm 1
ls -la gpurun_out/r2f_ncu_copy_columns.ncu-rep gpurun_out/r2f_ncu_conv_finalize_bnr.ncu-rep 2>&1 | cut -c1-120
