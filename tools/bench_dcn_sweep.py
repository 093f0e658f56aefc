"""BASELINE config 5: deformable-align microbench, feature maps 48x36 and 96x72, C = 32..256, G = C/4 (the
reference's 4 channels per offset group) and G = 1; achieved algorithmic HBM GB/s against the roofline.
sigma = 2 px offsets.  usage: python tools/bench_dcn_sweep.py [out.json]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fami_pose_b200 as fp
from fami_pose_b200 import ops

dev = "cuda"
PEAK = 6650.0
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


rows = []
B = 32
for (H, W) in ((48, 36), (96, 72)):
    for C in (32, 48, 64, 128, 256):
        for G in (C // 4, 1):
            dcn = fp.DeformConv2d(C, C, 3, padding=3, dilation=3).to(dev)
            row = {"H": H, "W": W, "C": C, "G": G, "B": B}
            # exact-fp32 arm (torchvision operand layout)
            fp.set_precision("fp32")
            x = ops.empty_nhwc(B, C, H, W, torch.float32, dev).normal_()
            off = ops.empty_nhwc(B, 18 * G, H, W, torch.float32, dev).normal_() * 2
            msk = ops.empty_nhwc(B, 9 * G, H, W, torch.float32, dev).normal_()
            out = ops.empty_nhwc(B, C, H, W, torch.float32, dev)
            t = timeit(lambda: dcn(x, off, msk, out=out))
            alg = 4 * B * H * W * (2 * C + 27 * G) + 4 * (9 * C * C + C)
            row.update({"fp32_us": t, "fp32_GBps": alg / t / 1e3, "fp32_frac": alg / t / 1e3 / PEAK})
            # 16-bit tensor-core arm (fused tap-major offsets), where the kernel takes the shape
            if ops.dcn_fused_supported(C, G, torch.float16):
                fp.set_precision("fp16")
                xh = ops.empty_nhwc(B, C, H, W, torch.float16, dev).normal_()
                om = ops.empty_nhwc(B, 27 * G, H, W, torch.float32, dev).normal_() * 2
                oh = ops.empty_nhwc(B, C, H, W, torch.float16, dev)
                blk = ops.om_to_blocked(om, G, layout=ops.dcn_blocked_layout(C, C, G))      # the layout the model's producer conv writes
                t16 = timeit(lambda: dcn(xh, None, None, out=oh, blocked_om=blk, groups=G))
                del blk
                alg16 = B * H * W * (2 * 2 * C + 4 * 27 * G) + 2 * (9 * C * C) + 4 * C
                row.update({"fp16_us": t16, "fp16_GBps": alg16 / t16 / 1e3, "fp16_frac": alg16 / t16 / 1e3 / PEAK})
            rows.append(row)
            print(row, flush=True)
fp.set_precision("fp32")
if len(sys.argv) > 1:
    json.dump({"gpu": torch.cuda.get_device_name(0), "hbm_peak_GBps": PEAK, "rows": rows}, open(sys.argv[1], "w"), indent=1)
