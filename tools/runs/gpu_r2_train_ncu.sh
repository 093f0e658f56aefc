#!/bin/bash
# ncu --set full captures of the training step's top kernels (one launch each, taken inside tools/profile_train.py)
mkdir -p gpurun_out
cap() {  # name regex skip
  timeout 600 ncu --set full --clock-control none -k regex:$2 -s $3 -c 1 -o gpurun_out/r2_train_$1 -f python tools/profile_train.py > gpurun_out/r2_ncu_train_$1.log 2>&1
  grep -c "==PROF== Report" gpurun_out/r2_ncu_train_$1.log
}
cap bn_apply_act bn_apply_act 20
cap bn_stats bn_stats 20
cap dcn_bwd_data dcn_bwd_data_kernel 1
cap dcn_bwd_weight dcn_bwd_weight_kernel 1
cap conv_wgrad_tf32 conv_wgrad_tf32_kernel 20
ls gpurun_out/*.ncu-rep | tail -8
