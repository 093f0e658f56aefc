#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/profile_train.py 32 tf32 tf32 > gpurun_out/r2_q_profile_train.txt 2>&1
head -45 gpurun_out/r2_q_profile_train.txt
