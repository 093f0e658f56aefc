#!/bin/bash
# rows per gather warp: 1 (16 warps) vs 2 (8 warps) vs 4 (4 warps)
cp fami_pose_b200/libfami_b200.so /tmp/lib1.so
echo "RPW 1"; BLOCKED=1 timeout 200 python tools/time_dcn.py 2>&1 | grep sigma
for r in 2 4; do
cp fami_pose_b200/libfami_b200_rpw$r.so fami_pose_b200/libfami_b200.so
echo "RPW $r"; BLOCKED=1 timeout 200 python tools/time_dcn.py 2>&1 | grep sigma
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "dcn or offset_conv" 2>&1 | tail -2
done
cp /tmp/lib1.so fami_pose_b200/libfami_b200.so
