"""Times the dominant conv shape classes of the model on one arm:  python tools/time_convs.py PREC [out.json] [N]
(L2 flushed between launches, CUDA events, median of 9).  Classes: HRNet stage-4 3x3 s1 (48/96/192/384), layer1,
transition / head many-in few-out, fuse 1x1 + upsample, stride 2, the fused offset|mask producer."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fami_pose_b200 as fp
from fami_pose_b200 import ops

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
out_path = sys.argv[2] if len(sys.argv) > 2 else None
N = int(sys.argv[3]) if len(sys.argv) > 3 else 160
fp.set_precision(prec)
dt = ops.act_dtype()
SHAPES = [  # Cin, Cout, H, W, k, stride, dil, res, up, images
    (48, 48, 96, 72, 3, 1, 1, False, 1, N), (48, 48, 96, 72, 3, 1, 1, True, 1, N),
    (96, 96, 48, 36, 3, 1, 1, False, 1, N), (96, 96, 48, 36, 3, 1, 1, True, 1, N),
    (192, 192, 24, 18, 3, 1, 1, False, 1, N), (192, 192, 24, 18, 3, 1, 1, True, 1, N),
    (384, 384, 12, 9, 3, 1, 1, False, 1, N), (384, 384, 12, 9, 3, 1, 1, True, 1, N),
    (64, 64, 96, 72, 3, 1, 1, False, 1, N), (64, 256, 96, 72, 1, 1, 1, True, 1, N), (256, 64, 96, 72, 1, 1, 1, False, 1, N),
    (256, 48, 96, 72, 3, 1, 1, False, 1, N), (48, 96, 96, 72, 3, 2, 1, True, 1, N), (96, 48, 48, 36, 1, 1, 1, True, 2, N),
    (384, 48, 12, 9, 1, 1, 1, True, 8, N), (64, 64, 192, 144, 3, 2, 1, False, 1, N),
    (48, 324, 96, 72, 3, 1, 3, False, 1, N // 5), (192, 48, 96, 72, 3, 1, 1, False, 1, N // 5), (96, 48, 96, 72, 3, 1, 1, False, 1, N // 5),
]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
rows = []
with torch.no_grad():
    for Cin, Cout, H, W, k, s, d, res, up, n in SHAPES:
        pad = d * (k // 2)
        x = ops.empty_nhwc(n, Cin, H, W, dt, "cuda").normal_()
        conv = torch.nn.Conv2d(Cin, Cout, k, s, pad, d, bias=False).cuda()
        bn = torch.nn.BatchNorm2d(Cout).cuda().eval()
        Ho, Wo = (H + 2 * pad - d * (k - 1) - 1) // s + 1, (W + 2 * pad - d * (k - 1) - 1) // s + 1
        odt = torch.float32 if Cout == 324 else dt
        out = ops.empty_nhwc(n, Cout, Ho * up, Wo * up, odt, "cuda")
        r = ops.empty_nhwc(n, Cout, Ho * up, Wo * up, dt, "cuda").normal_() if res else None
        ts = []
        for i in range(12):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.conv_bn_act(x, conv, bn, relu=True, residual=r, up=up, out=out, out_dtype=odt)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        t = sorted(ts[3:])[4]
        fl = 2.0 * n * Ho * Wo * k * k * Cin * Cout
        es, eo = x.element_size(), out.element_size()
        byt = n * H * W * Cin * es + n * Ho * Wo * up * up * Cout * (eo + (es if res else 0)) + k * k * Cin * Cout * es
        rows.append({"shape": "%d->%d k%d s%d d%d @%dx%d%s%s n=%d" % (Cin, Cout, k, s, d, H, W, " +res" if res else "", " up%d" % up if up > 1 else "", n),
                     "us": round(t, 1), "tflops": round(fl / t / 1e6, 1), "min_gbs": round(byt / t / 1e3, 1)})
        print(rows[-1], flush=True)
        del x, out, r
if out_path:
    json.dump({"precision": prec, "rows": rows}, open(out_path, "w"), indent=1)
