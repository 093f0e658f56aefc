import os
os.environ["FAMI_PROBES"] = "1"   # fami_debug_* live in libfami_b200_probes.so (csrc/build.py --probes)
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fami_pose_b200 import _lib, ops
out = torch.zeros(2, dtype=torch.int64, device="cuda")
for variant in (3, 203, 4, 204):
    for N in (48, 192):
        for iters in (810,):
            _lib.call("fami_debug_umma_rate", ops._ptr(out), N, iters, variant, ops._stream())
            torch.cuda.synchronize()
            tot, iss = out.tolist()
            print("variant %d N=%3d iters=%4d: %7.1f clk/MMA total, %7.1f clk/MMA issue  (ideal %.0f)" % (variant, N, iters, tot / iters, iss / iters, 128 * N / 256))
