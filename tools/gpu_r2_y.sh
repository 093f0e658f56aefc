#!/bin/bash
# stage-count scaling of the DCN kernel with the epilogue ablated (bit 1): skeleton (31), no-far (9), full-minus-epilogue (1)
cp fami_pose_b200/libfami_b200.so /tmp/lib3.so
for n in 2 3 4 5; do
  if [ $n -eq 3 ]; then cp /tmp/lib3.so fami_pose_b200/libfami_b200.so; else cp fami_pose_b200/libfami_b200_a$n.so fami_pose_b200/libfami_b200.so; fi
  for ab in 31 9 1; do
    echo -n "stages $n ablate $ab: "; FAMI_DCN_ABLATE=$ab BLOCKED=1 timeout 100 python tools/time_dcn.py 2>&1 | grep sigma | cut -d: -f2 | cut -d'>' -f1 | tr '\n' '|'; echo
  done
done
cp /tmp/lib3.so fami_pose_b200/libfami_b200.so
