"""Weight-gradient kernels alone: the 'tf32' arm's mma.sync TF32 kernel against the exact-fp32 FMA kernel on the shapes of the
training step (profiles/r2_profile_train_calls.txt)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fami_pose_b200 as fp
from fami_pose_b200 import ops
SHAPES = [(32, 48, 324, 96, 72, 3, 1, 3, 3), (32, 48, 48, 96, 72, 3, 1, 1, 1), (32, 96, 48, 96, 72, 3, 1, 1, 1), (32, 192, 48, 96, 72, 3, 1, 1, 1),
          (32, 48, 16, 96, 72, 3, 1, 1, 1), (32, 16, 16, 96, 72, 3, 1, 1, 1), (32, 48, 17, 96, 72, 3, 1, 1, 1), (160, 48, 48, 96, 72, 3, 1, 1, 1)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (N, Cin, Cout, H, W, k, s, p, d) in SHAPES:
    x = ops.empty_nhwc(N, Cin, H, W, torch.float32, "cuda").normal_()
    Ho, Wo = (H + 2 * p - d * (k - 1) - 1) // s + 1, (W + 2 * p - d * (k - 1) - 1) // s + 1
    gy = ops.empty_nhwc(N, Cout, Ho, Wo, torch.float32, "cuda").normal_()
    res = {}
    for arm in ("fp32", "tf32"):
        fp.set_precision(arm)
        for _ in range(2):
            gw, _ = ops.conv_wgrad(x, gy, (Cout, Cin, k, k), s, p, d)
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); gw, _ = ops.conv_wgrad(x, gy, (Cout, Cin, k, k), s, p, d); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        res[arm] = (sorted(ts)[2], gw)
    fl = 2.0 * N * Ho * Wo * k * k * Cin * Cout
    err = float((res["tf32"][1] - res["fp32"][1]).abs().max() / res["fp32"][1].abs().max())
    print("wgrad N=%d %d->%d k%d s%d d%d %dx%d: fp32 FMA %.0f us (%.1f TFLOP/s), tf32 mma.sync %.0f us (%.1f TFLOP/s), rel diff %.1e (incl. the memset)"
          % (N, Cin, Cout, k, s, d, H, W, res["fp32"][0], fl / res["fp32"][0] / 1e6, res["tf32"][0], fl / res["tf32"][0] / 1e6, err), flush=True)
fp.set_precision("fp32")
