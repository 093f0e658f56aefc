"""Micro-driver for ncu: runs the dominant conv classes (HRNet stage-4 3x3 s1) and the DCN launch a
few times.  usage: python tools/prof_conv.py [bf16|fp32] [which]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fami_pose_b200 as fp
from fami_pose_b200 import ops

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
which = sys.argv[2] if len(sys.argv) > 2 else "all"
fp.set_precision(prec)
dev = "cuda"
dt = ops.act_dtype()
N = 160
shapes = {"c48": (48, 96, 72), "c96": (96, 48, 36), "c192": (192, 24, 18), "c384": (384, 12, 9)}
ev = lambda: torch.cuda.Event(enable_timing=True)
_ng = torch.no_grad(); _ng.__enter__()   # inference launches (fp32 / tf32 storage would otherwise take the autograd path)
for name, (C, H, W) in shapes.items():
    if which not in ("all", name):
        continue
    x = ops.empty_nhwc(N, C, H, W, dt, dev).normal_()
    conv = torch.nn.Conv2d(C, C, 3, 1, 1, bias=False).to(dev)
    bn = torch.nn.BatchNorm2d(C).to(dev).eval()
    out = ops.empty_nhwc(N, C, H, W, dt, dev)
    ts = []
    for i in range(6):
        e0, e1 = ev(), ev()
        e0.record()
        ops.conv_bn_act(x, conv, bn, relu=True, residual=x, out=out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts[2:])[len(ts[2:]) // 2] * 1e-3
    fl = 2.0 * N * H * W * 9 * C * C
    by = 3 * N * H * W * C * x.element_size()
    print("%s %s: %.1f us  %.1f TFLOP/s  %.0f GB/s(min traffic)" % (prec, name, t * 1e6, fl / t / 1e12, by / t / 1e9))
if which in ("all", "dcn"):
    B, C, G, H, W = 32, 48, 12, 96, 72
    x = ops.empty_nhwc(B, C, H, W, dt, dev).normal_()
    off = ops.empty_nhwc(B, 18 * G, H, W, torch.float32, dev).normal_() * 2
    msk = ops.empty_nhwc(B, 9 * G, H, W, torch.float32, dev).normal_()
    dcn = fp.DeformConv2d(C, C, 3, padding=3, dilation=3).to(dev)
    out = ops.empty_nhwc(B, C, H, W, dt, dev)
    om = ops.empty_nhwc(B, 27 * G, H, W, torch.float32, dev).normal_() * 2
    fused = dt != torch.float32
    blk = ops.om_to_blocked(om, G) if fused else None
    ts = []
    for i in range(6):
        e0, e1 = ev(), ev()
        e0.record()
        if fused:
            dcn(x, None, None, out=out, blocked_om=blk, groups=G)
        else:
            dcn(x, off, msk, out=out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts[2:])[len(ts[2:]) // 2] * 1e-3
    alg = B * H * W * (2 * C * x.element_size() + 27 * G * 4)
    print("%s dcn: %.1f us %.0f GB/s" % (prec, t * 1e6, alg / t / 1e9))
if which == "bnstats":
    for C in (48, 64, 96, 192, 256, 384, 16, 17):
        x = ops.empty_nhwc(4, C, 24, 18, torch.float32, dev).normal_()
        st = torch.zeros(2 * C, dtype=torch.float64, device=dev)
        try:
            fp._lib.call("fami_bn_stats", ops._ptr(x), 0, C, 4 * 24 * 18, C, ops._ptr(st), ops._stream())
            torch.cuda.synchronize()
            ref = x.float().sum((0, 2, 3)).double()
            print(C, "ok", float((st[:C] - ref).abs().max()))
        except Exception as e:
            print(C, "FAIL", e)
