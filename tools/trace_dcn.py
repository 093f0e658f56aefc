"""Timeline of CTA 0 of the tensor-core DCN kernel (FAMI_DCN_TRACE=1)."""
import os
os.environ.setdefault("FAMI_DCN_WP", "0")   # the per-tap timeline is the tcgen05 kernel's (csrc/dcn_tc.cu)
os.environ["FAMI_PROBES"] = "1"   # fami_debug_* live in libfami_b200_probes.so (csrc/build.py --probes)
import os, sys, ctypes
os.environ.setdefault("FAMI_DCN_TRACE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fami_pose_b200 as fp
from fami_pose_b200 import ops, _lib
fp.set_precision("fp16")
B, C, G, H, W = 32, 48, 12, 96, 72
XP = int(os.environ.get("XPITCH", C))
x = ops.empty_nhwc(B, XP, H, W, torch.float16, "cuda").normal_()[:, :C]
om = ops.empty_nhwc(B, 27 * G, H, W, torch.float32, "cuda").normal_() * float(os.environ.get("SIGMA", "2"))
blk = ops.om_to_blocked(om, G)
dcn = fp.DeformConv2d(C, C, 3, padding=3, dilation=3).cuda()
out = ops.empty_nhwc(B, C, H, W, torch.float16, "cuda")
for _ in range(3):
    dcn(x, None, None, out=out, blocked_om=blk, groups=G)
buf = np.zeros(4096, dtype=np.uint64)
_lib.call("fami_debug_read_trace", buf.ctypes.data_as(ctypes.c_void_p), -4096)
t = buf[:256].reshape(16, 16).astype(np.int64)
t0 = t[0, 0]
print("tile | start   window ready | kernel row 0 / 1 / 2 gathered (us after window ready)   [CTA 0, gather warp 0]")
for i in range(10):
    print("%3d | %7.2f %7.2f | %s" % (i, (t[i, 0] - t0) / 1e3, (t[i, 1] - t0) / 1e3,
          "  ".join("%5.2f" % ((t[i, 2 + k] - t[i, 1]) / 1e3) for k in range(3))))
w = buf[2048:2048 + 512].reshape(2, 32, 8).astype(np.int64)
for k in range(2):
    base = w[k, :16, 0].min()
    print("tap %d of tile 3, every gather warp (us after the first warp started the tap): start | loads+blends done | A stage free | stores done | fenced+arrived" % (3 + k))
    for wp in range(16):
        print("  warp %2d: %s" % (wp, "  ".join("%6.2f" % ((w[k, wp, e] - base) / 1e3) for e in range(5))))
    print("  issuer : all arrivals seen %6.2f   MMAs + commit issued %6.2f" % ((w[k, 16, 0] - base) / 1e3, (w[k, 16, 1] - base) / 1e3))
