"""Time the tensor-core DCN kernel alone (bench shape) for a few offset magnitudes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fami_pose_b200 as fp
from fami_pose_b200 import ops
fp.set_precision("fp16")
B, C, G, H, W = 32, 48, 12, 96, 72
XP = int(os.environ.get("XPITCH", C))
x = ops.empty_nhwc(B, XP, H, W, torch.float16, "cuda").normal_()[:, :C]
dcn = fp.DeformConv2d(C, C, 3, padding=3, dilation=3).cuda()
out = ops.empty_nhwc(B, C, H, W, torch.float16, "cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for sigma in (0.5, 2.0, 4.0):
    om = ops.empty_nhwc(B, 27 * G, H, W, torch.float32, "cuda").normal_() * sigma
    blk = ops.om_to_blocked(om, G, layout=ops.dcn_blocked_layout(C, C, G)) if os.environ.get("BLOCKED") else None
    run = (lambda: dcn(x, None, None, out=out, blocked_om=blk, groups=G)) if blk is not None else (lambda: dcn(x, None, None, out=out, fused_om=om))
    if blk is not None:
        ref = ops.empty_nhwc(B, C, H, W, torch.float16, "cuda")
        dcn(x, None, None, out=ref, fused_om=om); run()
        print("blocked vs tap-major max diff", float((ref.float() - out.float()).abs().max()))
    for _ in range(3):
        run()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    alg = B * H * W * (27 * G * 4 + C * 2 * 2)
    print("sigma %.1f: median %.1f us  min %.1f us  -> %.0f GB/s algorithmic" % (sigma, ts[len(ts) // 2], ts[0], alg / ts[len(ts) // 2] / 1e3))
