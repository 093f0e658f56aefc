#!/bin/bash
mkdir -p gpurun_out
BN_BY_SHAPE=1 TOPN=70 timeout 600 python tools/profile_train.py > gpurun_out/r2_profile_train_bn_by_shape.txt 2>&1
grep -E "bn_apply_act|bn_stats" gpurun_out/r2_profile_train_bn_by_shape.txt | head -40
