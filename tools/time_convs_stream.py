"""Times the conv shape classes of the model in the fp32-residual-stream mode of the fp16 arm ('fp16s'): residual taken from
an fp32 twin, output written as fp16 + fp32 twin.  python tools/time_convs_stream.py [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fami_pose_b200 as fp
from fami_pose_b200 import ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 160
SHAPES = [  # Cin, Cout, H, W, k, stride, dil, res, up
    (48, 48, 96, 72, 3, 1, 1, True, 1), (96, 96, 48, 36, 3, 1, 1, True, 1), (192, 192, 24, 18, 3, 1, 1, True, 1),
    (384, 384, 12, 9, 3, 1, 1, True, 1), (64, 256, 96, 72, 1, 1, 1, True, 1), (48, 96, 96, 72, 3, 2, 1, True, 1),
    (96, 48, 48, 36, 1, 1, 1, True, 2), (384, 48, 12, 9, 1, 1, 1, True, 8), (256, 48, 96, 72, 3, 1, 1, False, 1),
    (96, 192, 48, 36, 3, 2, 1, True, 1), (192, 96, 24, 18, 1, 1, 1, True, 2),
]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
with torch.no_grad():
    for Cin, Cout, H, W, k, s, d, res, up in SHAPES:
        pad = d * (k // 2)
        conv = torch.nn.Conv2d(Cin, Cout, k, s, pad, d, bias=False).cuda()
        bn = torch.nn.BatchNorm2d(Cout).cuda().eval()
        Ho, Wo = (H + 2 * pad - d * (k - 1) - 1) // s + 1, (W + 2 * pad - d * (k - 1) - 1) // s + 1
        row = []
        for arm in ("fp16", "fp16s"):
            fp.set_precision(arm)
            x = ops.empty_nhwc(N, Cin, H, W, torch.float16, "cuda").normal_()
            r = None
            if res:
                r = ops.empty_nhwc(N, Cout, Ho * up, Wo * up, torch.float16, "cuda").normal_()
                if arm == "fp16s":
                    r._fami_f32 = ops.empty_nhwc(N, Cout, Ho * up, Wo * up, torch.float32, "cuda").normal_()
            ts = []
            for i in range(10):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                y = ops.conv_bn_act(x, conv, bn, relu=True, residual=r, up=up)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
                del y
            row.append(sorted(ts[3:])[3])
        print("%d->%d k%d s%d @%dx%d%s%s: fp16 %.1f us   fp16s %.1f us" % (Cin, Cout, k, s, H, W, " +res" if res else "", " up%d" % up if up > 1 else "", row[0], row[1]), flush=True)
fp.set_precision("fp32")
