"""Time one conv shape: python tools/time_conv_shape.py N Cin Cout H W k stride [res]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fami_pose_b200 as fp
from fami_pose_b200 import ops
fp.set_precision("fp16")
N, Cin, Cout, H, W, k, s = map(int, sys.argv[1:8])
res = len(sys.argv) > 8 and sys.argv[8] == "res"
x = ops.empty_nhwc(N, Cin, H, W, torch.float16, "cuda").normal_()
conv = torch.nn.Conv2d(Cin, Cout, k, s, k // 2, bias=False).cuda()
bn = torch.nn.BatchNorm2d(Cout).cuda().eval()
Ho, Wo = (H + 2 * (k // 2) - k) // s + 1, (W + 2 * (k // 2) - k) // s + 1
out = ops.empty_nhwc(N, Cout, Ho, Wo, torch.float16, "cuda")
r = ops.empty_nhwc(N, Cout, Ho, Wo, torch.float16, "cuda").normal_() if res else None
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
with torch.no_grad():
    for i in range(12):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.conv_bn_act(x, conv, bn, relu=True, residual=r, out=out); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
ts = sorted(ts[3:])
fl = 2.0 * N * Ho * Wo * k * k * Cin * Cout
print("conv %s: median %.1f us  %.0f TFLOP/s" % (sys.argv[1:], ts[len(ts) // 2], fl / ts[len(ts) // 2] / 1e6))
