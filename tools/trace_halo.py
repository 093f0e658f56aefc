"""Per-role timeline of CTA 0 of a halo conv (FAMI_HALO_TRACE=1)."""
import os
os.environ["FAMI_PROBES"] = "1"   # fami_debug_* live in libfami_b200_probes.so (csrc/build.py --probes)
import os, sys, ctypes
os.environ.setdefault("FAMI_HALO_TRACE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fami_pose_b200 as fp
from fami_pose_b200 import ops, _lib
fp.set_precision("fp16")
C, H, W = {"c48": (48, 96, 72), "c96": (96, 48, 36), "c192": (192, 24, 18)}[sys.argv[1] if len(sys.argv) > 1 else "c48"]
N = 160
x = ops.empty_nhwc(N, C, H, W, torch.float16, "cuda").normal_()
conv = torch.nn.Conv2d(C, C, 3, 1, 1, bias=False).cuda()
bn = torch.nn.BatchNorm2d(C).cuda().eval()
out = ops.empty_nhwc(N, C, H, W, torch.float16, "cuda")
for _ in range(3):
    ops.conv_bn_act(x, conv, bn, relu=True, residual=x, out=out)
buf = np.zeros(8192, dtype=np.uint64)
_lib.call("fami_debug_read_trace", buf.ctypes.data_as(ctypes.c_void_p), 8192)
t = buf[:4 * 64 * 8].reshape(4, 64, 8).astype(np.int64)
t0 = t[1, 0, 3]
rel = lambda v: (v - t0) / 1000.0
print("tile |  MMA: start  tempty_ok  fullA_ok  issued |  EPI: wait_start tfull_ok done |  PROD: emptyA_ok   (us)")
for i in range(12):
    print("%4d | %9.2f %9.2f %9.2f %9.2f | %9.2f %9.2f %9.2f | %9.2f" % (
        i, rel(t[1, i, 3]), rel(t[1, i, 0]), rel(t[1, i, 1]), rel(t[1, i, 2]), rel(t[2, i, 2]), rel(t[2, i, 0]), rel(t[2, i, 1]), rel(t[0, i, 0])))
print("per-tap issue timestamps (us after fullA_ok) for tiles 1..3:")
for i in (1, 2, 3):
    print(i, " ".join("%6.2f" % ((t[3, i, k] - t[1, i, 1]) / 1000.0) for k in range(8)))
dclk = t[1, 11, 5] - t[1, 1, 5]
dns = t[1, 11, 2] - t[1, 1, 2]
print("SM clock during kernel: %.0f MHz (clock64 delta %d over %d ns)" % (1000.0 * dclk / dns, dclk, dns))
