"""Same-box GPU library baselines: torchvision's CUDA deform_conv2d and cuDNN
convolutions on the same tensors as our kernels.  Measurement tool only -- nothing here is on the
product path.  usage: python tools/bench_vs_libs.py [out.json]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import torchvision
import fami_pose_b200 as fp
from fami_pose_b200 import ops

dev = "cuda"
torch.backends.cudnn.benchmark = True
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=12, warm=4):
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
# ---------------------------------------------------------------- stage-4 3x3 convs (N = 160 images)
N = 160
for name, (C, H, W) in {"c48": (48, 96, 72), "c96": (96, 48, 36), "c192": (192, 24, 18), "c384": (384, 12, 9)}.items():
    fl = 2.0 * N * H * W * 9 * C * C
    conv = torch.nn.Conv2d(C, C, 3, 1, 1, bias=False).to(dev)
    bn = torch.nn.BatchNorm2d(C).to(dev).eval()
    fp.set_precision("fp16")
    x = ops.empty_nhwc(N, C, H, W, torch.float16, dev).normal_()
    out = ops.empty_nhwc(N, C, H, W, torch.float16, dev)
    t_ours = timeit(lambda: ops.conv_bn_act(x, conv, bn, relu=True, residual=x, out=out))
    t_ours_plain = timeit(lambda: ops.conv_bn_act(x, conv, bn, relu=True, out=out))
    # cuDNN: channels_last fp16 conv alone, and conv + BN + add + ReLU as the reference runs them (separate kernels)
    xc = torch.randn(N, C, H, W, device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last)
    wc = conv.weight.detach().half().contiguous(memory_format=torch.channels_last)
    bnh = torch.nn.BatchNorm2d(C).to(dev).half().eval()
    with torch.no_grad():
        t_cudnn = timeit(lambda: F.conv2d(xc, wc, None, 1, 1))
        t_chain = timeit(lambda: F.relu(bnh(F.conv2d(xc, wc, None, 1, 1)) + xc))
        x32 = torch.randn(N, C, H, W, device=dev)
        w32 = conv.weight.detach()
        torch.backends.cudnn.allow_tf32 = False
        t_cudnn32 = timeit(lambda: F.conv2d(x32, w32, None, 1, 1))
    rows.append({"op": "conv3x3 s1 " + name, "shape": [N, C, H, W], "gflop": fl / 1e9,
                 "ours_fp16_conv_bn_res_relu_us": t_ours, "ours_fp16_conv_bn_relu_us": t_ours_plain,
                 "cudnn_fp16_nhwc_conv_only_us": t_cudnn, "cudnn_fp16_conv_bn_add_relu_us": t_chain,
                 "cudnn_fp32_nchw_conv_only_us": t_cudnn32,
                 "ours_tflops": fl / t_ours_plain / 1e6, "cudnn_fp16_tflops": fl / t_cudnn / 1e6})
    print(rows[-1], flush=True)

# ---------------------------------------------------------------- deformable alignment (bench shape)
B, C, G, H, W = 32, 48, 12, 96, 72
dcn = fp.DeformConv2d(C, C, 3, padding=3, dilation=3).to(dev)
x32 = torch.randn(B, C, H, W, device=dev)
off = 2 * torch.randn(B, 18 * G, H, W, device=dev)
msk = torch.randn(B, 9 * G, H, W, device=dev)
with torch.no_grad():
    t_tv32 = timeit(lambda: torchvision.ops.deform_conv2d(x32, off, dcn.weight, dcn.bias, 1, 3, 3, msk))
    t_tv16 = timeit(lambda: torchvision.ops.deform_conv2d(x32.half(), off.half(), dcn.weight.half(), dcn.bias.half(), 1, 3, 3, msk.half()))
    # the reference also runs the offset and mask convs as separate cuDNN launches feeding deform_conv2d
    wo = torch.randn(18 * G, C, 3, 3, device=dev) * 0.05
    wm = torch.randn(9 * G, C, 3, 3, device=dev) * 0.05
    t_tv_chain = timeit(lambda: torchvision.ops.deform_conv2d(x32, F.conv2d(x32, wo, None, 1, 3, 3), dcn.weight, dcn.bias, 1, 3, 3,
                                                              F.conv2d(x32, wm, None, 1, 3, 3)))
fp.set_precision("fp32")
xn = ops.to_nhwc(x32); offn = ops.to_nhwc(off); mskn = ops.to_nhwc(msk)
outn = ops.empty_nhwc(B, C, H, W, torch.float32, dev)
t_ours32 = timeit(lambda: dcn(xn, offn, mskn, out=outn))
fp.set_precision("fp16")
xh = ops.empty_nhwc(B, C, H, W, torch.float16, dev).normal_()
om = ops.empty_nhwc(B, 27 * G, H, W, torch.float32, dev).normal_() * 2
outh = ops.empty_nhwc(B, C, H, W, torch.float16, dev)
blk = ops.om_to_blocked(om, G, layout=ops.dcn_blocked_layout(C, C, G))      # the layout the model's producer conv writes
t_ours16 = timeit(lambda: dcn(xh, None, None, out=outh, blocked_om=blk, groups=G))
alg32 = 4 * B * H * W * (2 * C + 27 * G)
alg16 = B * H * W * (2 * 2 * C + 4 * 27 * G)
rows.append({"op": "deform_conv2d 48->48 k3 d3 G12", "shape": [B, C, H, W],
             "torchvision_cuda_fp32_us": t_tv32, "torchvision_cuda_fp16_us": t_tv16,
             "torchvision_fp32_with_offset_mask_convs_us": t_tv_chain,
             "ours_fp32_simt_us": t_ours32, "ours_fp16_tc_us": t_ours16,
             "ours_fp32_GBps": alg32 / t_ours32 / 1e3, "ours_fp16_GBps": alg16 / t_ours16 / 1e3,
             "torchvision_fp32_GBps": alg32 / t_tv32 / 1e3})
print(rows[-1], flush=True)
if len(sys.argv) > 1:
    with open(sys.argv[1], "w") as f:
        json.dump({"torch": torch.__version__, "torchvision": torchvision.__version__,
                   "cudnn": torch.backends.cudnn.version(), "gpu": torch.cuda.get_device_name(0), "rows": rows}, f, indent=1)
