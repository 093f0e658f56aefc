"""GPU parity of each C-ABI op against the CPU oracle (oracle/fami_oracle.py) and the committed
torchvision/reference goldens.  Tolerances are stated per test; index results are bit-exact."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import fami_oracle as fo  # noqa: E402  (checker only)

DEV = "cuda"


def fp():
    import fami_pose_b200 as m
    m.set_precision("fp32")
    return m


def nhwc(t):
    """NCHW cpu tensor -> channels-last activation on the GPU via the library's own converter."""
    from fami_pose_b200 import ops
    return ops.to_nhwc(t.to(DEV).float().contiguous(), torch.float32)


def back(t):
    from fami_pose_b200 import ops
    return ops.to_nchw(t).cpu()


def test_layout_roundtrip():
    fp()
    g = torch.Generator().manual_seed(1)
    for shape in [(2, 3, 17, 13), (1, 48, 9, 7), (3, 17, 5, 4), (1, 1, 4, 4)]:
        x = torch.randn(*shape, generator=g)
        y = back(nhwc(x))
        assert torch.equal(x, y), shape


CONV_CASES = [
    # Cin, Cout, k, stride, pad, dil, H, W, bias, bn, relu, res, up
    (3, 64, 3, 2, 1, 1, 20, 14, False, True, True, False, 1),     # stem (scalar gather path)
    (64, 64, 3, 2, 1, 1, 19, 13, False, True, True, False, 1),
    (64, 256, 1, 1, 0, 1, 9, 7, False, True, False, True, 1),     # bottleneck conv3 + residual
    (48, 48, 3, 1, 1, 1, 24, 18, False, True, True, True, 1),     # BasicBlock conv2
    (96, 48, 1, 1, 0, 1, 6, 5, False, True, True, True, 2),       # fuse up x2
    (384, 48, 1, 1, 0, 1, 3, 3, False, True, False, True, 8),     # fuse up x8
    (48, 96, 3, 2, 1, 1, 24, 18, False, True, False, True, 1),    # fuse down + running sum
    (48, 324, 3, 1, 3, 3, 12, 9, True, False, False, False, 1),   # offset|mask conv, dilation 3
    (48, 17, 3, 1, 1, 1, 12, 9, True, False, False, False, 1),    # agg_final_layer (Cout=17 scalar store)
    (48, 17, 1, 1, 0, 1, 12, 9, True, False, False, False, 1),    # hrnet.final_layer
    (16, 16, 3, 2, 1, 1, 6, 5, True, True, True, False, 1),       # conv_bn_relu with bias + BN
    (192, 192, 3, 1, 1, 1, 6, 5, False, True, True, True, 1),
    (384, 384, 3, 1, 1, 1, 5, 4, False, True, True, True, 1),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_bn_act_eval(case):
    """fami_conv2d_bn_act_fwd vs torch CPU fp32 conv2d -> batch_norm(eval) -> +res -> relu -> nearest up.
    Tolerance 2e-5 * max|ref| + 1e-5 (fp32, differing summation order only)."""
    m = fp()
    from fami_pose_b200 import ops
    Cin, Cout, k, s, p, d, H, W, bias, bn, relu, res, up = case
    g = torch.Generator().manual_seed(hash(case) % (2 ** 31))
    N = 3
    x = torch.randn(N, Cin, H, W, generator=g)
    conv = torch.nn.Conv2d(Cin, Cout, k, s, p, d, bias=bias)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) / (Cin * k * k) ** 0.5)
        if bias:
            conv.bias.copy_(torch.randn(Cout, generator=g))
    bnm = None
    if bn:
        bnm = torch.nn.BatchNorm2d(Cout).eval()
        with torch.no_grad():
            bnm.weight.copy_(torch.rand(Cout, generator=g) + 0.5)
            bnm.bias.copy_(torch.randn(Cout, generator=g) * 0.2)
            bnm.running_mean.copy_(torch.randn(Cout, generator=g) * 0.2)
            bnm.running_var.copy_(torch.rand(Cout, generator=g) + 0.5)
    with torch.no_grad():
        y = conv(x)
        if bn:
            y = bnm(y)
        if up > 1:
            y = F.interpolate(y, scale_factor=up, mode="nearest")
        r = torch.randn(y.shape, generator=g) if res else None
        if res:
            y = y + r
        if relu:
            y = F.relu(y)
    conv_d = conv.to(DEV)
    bn_d = bnm.to(DEV) if bn else None
    with torch.no_grad():      # the fused inference kernel (with grad enabled the fp32 arm records the unfused graph)
        out = ops.conv_bn_act(nhwc(x), conv_d, bn_d, relu=relu, residual=nhwc(r) if res else None, up=up)
    got = back(out)
    tol = 2e-5 * float(y.abs().max()) + 1e-5
    assert float((got - y).abs().max()) <= tol
    if up == 1:
        # same values through the differentiable (unfused) path
        out_g = ops.conv_bn_act(nhwc(x), conv_d, bn_d, relu=relu, residual=nhwc(r) if res else None, up=up)
        assert out_g.requires_grad
        assert float((back(out_g.detach()) - y).abs().max()) <= tol


def test_conv_bn_train_mode():
    """train-mode BatchNorm (batch statistics + running-stat update) vs torch; tol 1e-4."""
    fp()
    from fami_pose_b200 import ops
    g = torch.Generator().manual_seed(7)
    x = torch.randn(4, 48, 12, 9, generator=g)
    conv = torch.nn.Conv2d(48, 96, 3, 2, 1, bias=False)
    bn = torch.nn.BatchNorm2d(96, momentum=0.1).train()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(96, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(96, generator=g) * 0.1)
    import copy
    conv_d, bn_d = copy.deepcopy(conv).to(DEV), copy.deepcopy(bn).to(DEV)
    with torch.no_grad():
        y = F.relu(bn(conv(x)))
    out = back(ops.conv_bn_act(nhwc(x), conv_d, bn_d, relu=True))
    assert float((out - y).abs().max()) <= 1e-4
    assert float((bn_d.running_mean.cpu() - bn.running_mean).abs().max()) <= 1e-5
    assert float((bn_d.running_var.cpu() - bn.running_var).abs().max()) <= 1e-5
    assert int(bn_d.num_batches_tracked) == 1




@pytest.mark.parametrize("C,dtype", [(48, torch.float32), (17, torch.float32), (48, torch.float16)])
def test_bn_stats_large_mean(C, dtype):
    """per-channel sum / sum-of-squares with |mean| / std = 1e3 (post-ReLU-like features, large conv biases): the variance
    recovered from the statistics must match a float64 computation to 1e-3 relative (E[x^2] - mean^2 from plain float
    partial sums loses it entirely at this ratio)."""
    m = fp()
    from fami_pose_b200 import ops
    g = torch.Generator().manual_seed(11)
    rows = 4 * 24 * 18
    mean = 100.0 if dtype == torch.float16 else 1000.0       # fp16 storage: keep the ulp of the values below the std
    std = 1.0
    x = (torch.randn(rows, C, generator=g) * std + mean).to(dtype).to(DEV).contiguous()
    st = torch.zeros(2 * C, dtype=torch.float64, device=DEV)
    m._lib.call("fami_bn_stats", ops._ptr(x), ops._code(dtype),
                C, rows, C, ops._ptr(st), ops._stream())
    torch.cuda.synchronize()
    xd = x.double()
    ref_mean, ref_var = xd.mean(0), xd.var(0, unbiased=False)
    got_mean = st[:C] / rows
    got_var = st[C:] / rows - got_mean * got_mean
    assert float((got_mean - ref_mean).abs().max()) <= 1e-6 * mean
    assert float(((got_var - ref_var) / ref_var).abs().max()) <= 1e-3


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_cast_nhwc_fast_path_is_round_to_nearest(dtype):
    """ops.cast_nhwc (fami_bn_apply_act with null scale / shift: the plain-cast kernel the tf32 arm's deformable convolutions
    use) equals torch's round-to-nearest conversion bit for bit, for a dense tensor, a channel slice of a wider buffer
    (pitch > C) and a channel count that falls back to the affine kernel (C % 8 != 0)."""
    fp()
    from fami_pose_b200 import ops
    g = torch.Generator().manual_seed(5)
    for C, wide in ((48, None), (48, 96), (20, None)):
        x = torch.randn(3, C, 17, 9, generator=g) * 3
        xd = ops.to_nhwc(x.to(DEV), torch.float32)
        if wide:
            buf = ops.empty_nhwc(3, wide, 17, 9, torch.float32, DEV).zero_()
            buf[:, 24:24 + C].copy_(xd)
            xd = buf[:, 24:24 + C]
        out = ops.cast_nhwc(xd, dtype)
        assert out.dtype == dtype
        assert torch.equal(ops.to_nchw(out).cpu(), x.to(dtype))


def _golden_inputs(name):
    # same generator recipe as tests/golden/make_golden.py::dcn_inputs (kept in sync by test_oracle.py)
    from tests_support import dcn_cases, dcn_inputs
    for case in dcn_cases():
        if case[0] == name:
            return case, dcn_inputs(*case)
    raise KeyError(name)


@pytest.mark.parametrize("name", ["c48g12", "c32g8", "c16g1_oob", "c64g16"])
def test_dcn_fwd_vs_torchvision_golden(name, golden_dir):
    """fami_dcn_fwd vs torchvision CPU deform_conv2d (committed golden, fp64 reference values);
    tolerance 1e-5 abs."""
    fp()
    from fami_pose_b200 import layers
    gold = np.load(os.path.join(golden_dir, "dcn_torchvision.npz"))
    case, (x, off, msk, w, b, go) = _golden_inputs(name)
    _, B, C, Cout, G, H, W, sig = case
    mod = layers.DeformConv2d(C, Cout, 3, padding=3, dilation=3).to(DEV)
    with torch.no_grad():
        mod.weight.copy_(w.to(DEV))
        mod.bias.copy_(b.to(DEV))
    out = back(mod(nhwc(x), nhwc(off), nhwc(msk)))
    ref = torch.from_numpy(gold["%s_f64_out" % name]).float()
    assert float((out - ref).abs().max()) <= 1e-5 * max(1.0, float(ref.abs().max()))


def test_dcn_fwd_vs_oracle_larger():
    """config-2 DCN shape at B=2 vs the numpy oracle; tol 2e-5."""
    fp()
    from fami_pose_b200 import layers
    g = torch.Generator().manual_seed(3)
    B, C, G, H, W = 2, 48, 12, 96, 72
    x = torch.randn(B, C, H, W, generator=g)
    off = 2.0 * torch.randn(B, 18 * G, H, W, generator=g)
    msk = torch.randn(B, 9 * G, H, W, generator=g)
    w = 0.05 * torch.randn(C, C, 3, 3, generator=g)
    b = 0.1 * torch.randn(C, generator=g)
    ref = torch.from_numpy(fo.dcn_fwd(x.numpy(), off.numpy(), msk.numpy(), w.numpy(), b.numpy()))
    mod = layers.DeformConv2d(C, C, 3, padding=3, dilation=3).to(DEV)
    with torch.no_grad():
        mod.weight.copy_(w.to(DEV))
        mod.bias.copy_(b.to(DEV))
    out = back(mod(nhwc(x), nhwc(off), nhwc(msk)))
    assert float((out - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))


def test_dcn_zero_offset_equals_dilated_conv():
    """Property (size independent): zero offsets + unit mask == ordinary dilated convolution."""
    fp()
    from fami_pose_b200 import layers, ops
    g = torch.Generator().manual_seed(5)
    B, C, G, H, W = 4, 48, 12, 96, 72
    x = torch.randn(B, C, H, W, generator=g)
    mod = layers.DeformConv2d(C, C, 3, padding=3, dilation=3).to(DEV)
    conv = torch.nn.Conv2d(C, C, 3, 1, 3, 3).to(DEV)
    with torch.no_grad():
        conv.weight.copy_(mod.weight)
        conv.bias.copy_(mod.bias)
    xd = nhwc(x)
    off = ops.empty_nhwc(B, 18 * G, H, W, torch.float32, DEV).zero_()
    msk = ops.empty_nhwc(B, 9 * G, H, W, torch.float32, DEV).fill_(1.0)
    a = back(mod(xd, off, msk))
    b = back(ops.conv_bn_act(xd, conv, None))
    assert float((a - b).abs().max()) <= 1e-5


def test_dcn_bad_offset_channels_raises():
    fp()
    from fami_pose_b200 import layers, ops
    mod = layers.DeformConv2d(48, 48, 3, padding=3, dilation=3).to(DEV)
    x = ops.empty_nhwc(1, 48, 8, 8, torch.float32, DEV).zero_()
    off = ops.empty_nhwc(1, 20, 8, 8, torch.float32, DEV).zero_()
    msk = ops.empty_nhwc(1, 9, 8, 8, torch.float32, DEV).zero_()
    with pytest.raises(RuntimeError):
        mod(x, off, msk)


def test_warp_translate_vs_kornia_restatement():
    """fami_warp_translate_fwd vs the kornia>=0.6 warp_affine restatement; tol 1e-5 (incl. shifts
    that push the whole image out of view)."""
    fp()
    from fami_pose_b200 import ops
    g = torch.Generator().manual_seed(11)
    B, C, H, W = 5, 48, 24, 18
    src = torch.randn(B, C, H, W, generator=g)
    txy = torch.tensor([[0.0, 0.0], [1.25, -2.5], [-3.75, 0.5], [40.0, 3.0], [0.999, 17.2]])
    M = torch.eye(3)[0:2].view(1, 2, 3).repeat(B, 1, 1)
    M[:, 0, 2], M[:, 1, 2] = txy[:, 0], txy[:, 1]
    ref = fo.warp_affine_kornia(src, M, (H, W))
    out = back(ops.warp_translate(nhwc(src), txy.to(DEV)))
    assert float((out - ref).abs().max()) <= 1e-5
    ref2 = torch.from_numpy(fo.warp_translate(src.numpy(), txy.numpy()))
    assert float((out - ref2).abs().max()) <= 1e-5


def test_sub_copy_linear():
    fp()
    from fami_pose_b200 import ops
    g = torch.Generator().manual_seed(13)
    a = torch.randn(6, 16, 5, 4, generator=g)
    b = torch.randn(2, 16, 5, 4, generator=g)
    out = back(ops.sub_bcast(nhwc(a), nhwc(b), 3))
    assert torch.equal(out, a - b.repeat(3, 1, 1, 1))
    dst = ops.empty_nhwc(2, 32, 5, 4, torch.float32, DEV).zero_()
    ops.copy_into(nhwc(b), dst[:, 16:])
    full = back(dst)
    assert torch.equal(full[:, 16:], b) and float(full[:, :16].abs().max()) == 0.0
    x = torch.randn(8, 144, generator=g)
    w = torch.randn(64, 144, generator=g) / 12
    bias = torch.randn(64, generator=g)
    y = ops.linear(x.to(DEV), w.to(DEV), bias.to(DEV)).cpu()
    assert float((y - F.linear(x, w, bias)).abs().max()) <= 1e-5


def test_joint_mse_vs_oracle():
    """fami_joint_mse_fwd_bwd vs oracle.joint_mse (float64) rel 1e-5, and vs autograd for the gradient."""
    fp()
    from fami_pose_b200 import ops
    g = torch.Generator().manual_seed(17)
    B, J, H, W = 3, 17, 24, 18
    pred = torch.randn(B, J, H, W, generator=g)
    tgt = torch.rand(B, J, H, W, generator=g)
    tw = (torch.rand(B, J, 1, generator=g) < 0.85).float()
    ref = fo.joint_mse(pred.numpy(), tgt.numpy(), tw.numpy())
    loss, grad = ops.joint_mse(pred.to(DEV), tgt.to(DEV), tw.to(DEV), want_grad=True)
    assert abs(float(loss) - ref) <= 1e-5 * abs(ref)
    p = pred.clone().requires_grad_(True)
    l2 = (((p - tgt) * tw.view(B, J, 1, 1)) ** 2).mean(dim=(0, 2, 3)).sum() / J
    l2.backward()
    assert float((grad.permute(0, 3, 1, 2).cpu() - p.grad).abs().max()) <= 1e-8 + 1e-5 * float(p.grad.abs().max())


def test_softmax_pkl_vs_oracle_and_torch():
    """fami_softmax_pkl_fwd vs oracle.softmax_pkl (float64) and torch's own F.kl_div call as the
    reference makes it (Alignment_V15.py:260,275); tol 1e-6 abs + 1e-4 rel."""
    fp()
    from fami_pose_b200 import ops
    g = torch.Generator().manual_seed(19)
    for C in (17, 48):
        B, H, W = 2, 96, 72
        a = torch.randn(B, C, H, W, generator=g) * 0.3
        b = torch.randn(B, C, H, W, generator=g) * 0.3
        ref = fo.softmax_pkl(a.reshape(B * C, -1).numpy(), b.reshape(B * C, -1).numpy(), 0.05)
        tref = F.kl_div(input=F.softmax(a.reshape(B * C, -1) / 0.05, dim=1),
                        target=F.softmax(b.reshape(B * C, -1) / 0.05, dim=1), reduction="mean")
        got = float(ops.softmax_pkl(nhwc(a), nhwc(b), 0.05))
        assert abs(ref - float(tref)) <= 1e-6 + 1e-4 * abs(ref)
        assert abs(got - ref) <= 1e-6 + 1e-4 * abs(ref)


def test_argmax_bit_exact_with_ties():
    """fami_argmax_hw == numpy argmax (first maximum wins), including planted ties."""
    fp()
    import fami_pose_b200 as m
    g = torch.Generator().manual_seed(23)
    B, J, H, W = 4, 17, 96, 72
    hm = torch.randn(B, J, H, W, generator=g)
    hm[0, 0].fill_(0.0)                      # all equal -> index 0
    hm[1, 3, 10, 5] = 9.0
    hm[1, 3, 50, 60] = 9.0                   # tie -> lower flat index
    hm[2, 5] = -hm[2, 5].abs() - 1.0         # all negative -> coords masked to 0
    preds_ref, maxv_ref, idx_ref = fo.get_max_preds(hm.numpy())
    idx = m.argmax_indices(hm.to(DEV)).cpu().numpy()
    assert np.array_equal(idx, idx_ref.astype(np.int32))
    preds, maxv = m.get_max_preds(hm.to(DEV))
    assert np.array_equal(preds.cpu().numpy(), preds_ref)
    assert np.array_equal(maxv.cpu().numpy(), maxv_ref)


def test_cpu_tensors_fail_loudly():
    fp()
    from fami_pose_b200 import ops
    with pytest.raises(RuntimeError):
        ops.to_nhwc(torch.zeros(1, 3, 4, 4))


# ---------------------------------------------------------------------------------------------
# bf16 tensor-core (tcgen05) convolution
# ---------------------------------------------------------------------------------------------
TC_CASES = [
    # Cin, Cout, k, stride, pad, dil, H, W, N, bias, bn, relu, res, up, out_f32
    (48, 48, 3, 1, 1, 1, 24, 18, 3, False, True, True, True, 1, False),
    (64, 64, 1, 1, 0, 1, 16, 8, 1, False, False, False, False, 1, False),   # plain GEMM, M = 128
    (64, 64, 3, 1, 1, 1, 16, 8, 2, False, True, True, False, 1, False),
    (96, 96, 3, 1, 1, 1, 12, 9, 5, False, True, True, True, 1, False),      # 2 chunks, partial last chunk
    (192, 192, 3, 1, 1, 1, 6, 5, 7, False, True, True, True, 1, False),
    (384, 384, 3, 1, 1, 1, 5, 4, 9, False, True, True, True, 1, False),     # 2 N tiles
    (48, 96, 3, 2, 1, 1, 24, 18, 3, False, True, False, True, 1, False),    # stride 2 + running sum
    (64, 64, 3, 2, 1, 1, 19, 13, 2, False, True, True, False, 1, False),    # stride 2, odd size
    (96, 48, 1, 1, 0, 1, 6, 5, 4, False, True, True, True, 2, False),       # fuse up x2
    (384, 48, 1, 1, 0, 1, 3, 3, 5, False, True, False, True, 8, False),     # fuse up x8
    (48, 324, 3, 1, 3, 3, 12, 9, 2, True, False, False, False, 1, True),    # offset|mask conv, fp32 out
    (48, 17, 3, 1, 1, 1, 12, 9, 2, True, False, False, False, 1, True),     # heatmap conv, fp32 out
    (16, 16, 3, 2, 1, 1, 6, 5, 8, True, True, True, False, 1, False),
    (256, 64, 1, 1, 0, 1, 9, 7, 3, False, True, True, False, 1, False),
]


# shapes that take the halo kernel (3x3, stride 1, pad == dil, feature maps large enough)
HALO_CASES = [
    (48, 48, 3, 1, 1, 1, 96, 72, 2, False, True, True, True, 1, False),     # stage-4 branch 0 (weights resident)
    (96, 96, 3, 1, 1, 1, 48, 36, 2, False, True, True, True, 1, False),     # branch 1 (2 chunks, streamed B)
    (192, 192, 3, 1, 1, 1, 24, 18, 3, False, True, True, True, 1, False),   # branch 2
    (64, 64, 3, 1, 1, 1, 96, 72, 1, False, True, True, False, 1, False),    # layer1 bottleneck conv2
    (256, 48, 3, 1, 1, 1, 96, 72, 1, False, True, True, False, 1, False),   # transition1 (4 chunks)
    (192, 48, 3, 1, 1, 1, 96, 72, 1, False, True, True, False, 1, False),   # sup_agg_block conv1
    (48, 324, 3, 1, 3, 3, 96, 72, 1, True, False, False, False, 1, True),   # offset|mask conv: dilation 3, 3 N tiles, fp32 out
    (48, 17, 3, 1, 1, 1, 96, 72, 2, True, False, False, False, 1, True),    # agg_final_layer, fp32 out, Cout 17
    (48, 48, 3, 1, 1, 1, 50, 37, 3, False, True, True, True, 1, False),     # ragged size (partial last row tile)
    (32, 32, 3, 1, 1, 1, 80, 60, 2, False, True, True, True, 1, False),     # W32 variant
]


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("case", TC_CASES + HALO_CASES)
def test_conv_tc_half(case, prec):
    """tcgen05 16-bit conv vs torch CPU fp32 conv on the SAME rounded operands (so only the
    accumulation order and the output rounding differ): tol 1e-2*max|ref| for bf16 outputs
    (2^-8 rounding), 2e-3*max|ref| for fp16 / fp32 outputs."""
    import fami_pose_b200 as m
    from fami_pose_b200 import ops
    m.set_precision(prec)
    hdt = ops.act_dtype()
    try:
        Cin, Cout, k, s, p, d, H, W, N, bias, bn, relu, res, up, out_f32 = case
        g = torch.Generator().manual_seed(abs(hash(case)) % (2 ** 31))
        rb = lambda t: t.to(hdt).float()
        x = rb(torch.randn(N, Cin, H, W, generator=g))
        conv = torch.nn.Conv2d(Cin, Cout, k, s, p, d, bias=bias)
        with torch.no_grad():
            conv.weight.copy_(rb(torch.randn(conv.weight.shape, generator=g) / (Cin * k * k) ** 0.5))
            if bias:
                conv.bias.copy_(torch.randn(Cout, generator=g))
        bnm = None
        if bn:
            bnm = torch.nn.BatchNorm2d(Cout).eval()
            with torch.no_grad():
                bnm.weight.copy_(torch.rand(Cout, generator=g) + 0.5)
                bnm.bias.copy_(torch.randn(Cout, generator=g) * 0.2)
                bnm.running_mean.copy_(torch.randn(Cout, generator=g) * 0.2)
                bnm.running_var.copy_(torch.rand(Cout, generator=g) + 0.5)
        with torch.no_grad():
            y = conv(x)
            if bn:
                y = bnm(y)
            if up > 1:
                y = F.interpolate(y, scale_factor=up, mode="nearest")
            r = rb(torch.randn(y.shape, generator=g)) if res else None
            if res:
                y = y + r
            if relu:
                y = F.relu(y)
        xd = ops.to_nhwc(x.to(DEV), hdt)
        rd = ops.to_nhwc(r.to(DEV), hdt) if res else None
        out = ops.conv_bn_act(xd, conv.to(DEV), bnm.to(DEV) if bn else None, relu=relu, residual=rd, up=up,
                              out_dtype=torch.float32 if out_f32 else None)
        assert out.dtype == (torch.float32 if out_f32 else hdt)
        got = ops.to_nchw(out).cpu()
        tol = (2e-3 if (out_f32 or prec == 'fp16') else 1e-2) * float(y.abs().max()) + 1e-3
        err = float((got - y).abs().max())
        print("tc conv", prec, case, "err", err, "tol", tol)
        assert err <= tol
    finally:
        m.set_precision("fp32")


def _tf32_rna(t):
    """cvt.rna.tf32.f32 on a float32 tensor: round the 23-bit mantissa to 10 bits, ties away from zero."""
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("case", TC_CASES + HALO_CASES)
def test_conv_tc_tf32(case):
    """tcgen05 kind::tf32 conv (fp32 storage, TF32 multiplicands, fp32 accumulate; the 'tf32' arm) vs torch CPU fp32
    conv on operands ALREADY rounded to tf32, so only the accumulation order differs: tol 2e-5*max|ref| + 1e-5.
    Output, residual and the BN/ReLU/upsample epilogue are fp32."""
    import fami_pose_b200 as m
    from fami_pose_b200 import ops
    m.set_precision("tf32")
    try:
        Cin, Cout, k, s, p, d, H, W, N, bias, bn, relu, res, up, out_f32 = case
        g = torch.Generator().manual_seed(abs(hash(case)) % (2 ** 31))
        x = _tf32_rna(torch.randn(N, Cin, H, W, generator=g))
        conv = torch.nn.Conv2d(Cin, Cout, k, s, p, d, bias=bias)
        with torch.no_grad():
            conv.weight.copy_(_tf32_rna(torch.randn(conv.weight.shape, generator=g) / (Cin * k * k) ** 0.5))
            if bias:
                conv.bias.copy_(torch.randn(Cout, generator=g))
        bnm = None
        if bn:
            bnm = torch.nn.BatchNorm2d(Cout).eval()
            with torch.no_grad():
                bnm.weight.copy_(torch.rand(Cout, generator=g) + 0.5)
                bnm.bias.copy_(torch.randn(Cout, generator=g) * 0.2)
                bnm.running_mean.copy_(torch.randn(Cout, generator=g) * 0.2)
                bnm.running_var.copy_(torch.rand(Cout, generator=g) + 0.5)
        with torch.no_grad():
            y = torch.nn.functional.conv2d(x.double(), conv.weight.double(), conv.bias.double() if bias else None, s, p, d).float()
            if bn:
                y = bnm(y)
            if up > 1:
                y = F.interpolate(y, scale_factor=up, mode="nearest")
            r = torch.randn(y.shape, generator=g) if res else None      # the residual stream keeps full fp32
            if res:
                y = y + r
            if relu:
                y = F.relu(y)
        xd = ops.to_nhwc(x.to(DEV), torch.float32)
        assert ops.conv_code(xd, Cin) == m._lib.TF32
        rd = ops.to_nhwc(r.to(DEV), torch.float32) if res else None
        with torch.no_grad():
            out = ops.conv_bn_act(xd, conv.to(DEV), bnm.to(DEV) if bn else None, relu=relu, residual=rd, up=up)
        assert out.dtype == torch.float32
        got = ops.to_nchw(out).cpu()
        tol = 2e-5 * float(y.abs().max()) + 1e-5
        err = float((got - y).abs().max())
        print("tf32 conv", case, "err", err, "tol", tol)
        assert err <= tol
    finally:
        m.set_precision("fp32")


def test_conv_tf32_rounds_activations_to_nearest():
    """The tf32 arm keeps activations as full fp32 in HBM; the multiplicand is rounded to TF32 (to nearest even, measured
    by tools/probe_tma_tf32.py) on its way into shared memory (TFLOAT32 tensor map), not truncated by the tensor core.
    Un-rounded inputs: the result must match the conv of nearest-rounded inputs (2e-5; ties are measure-zero for random
    data, so rna and rne agree) and differ measurably from the conv of truncated inputs."""
    import fami_pose_b200 as m
    from fami_pose_b200 import ops
    m.set_precision("tf32")
    try:
        g = torch.Generator().manual_seed(7)
        x = torch.rand(2, 64, 24, 18, generator=g) + 0.5            # positive: truncation bias is coherent
        conv = torch.nn.Conv2d(64, 64, 3, 1, 1, bias=False)
        with torch.no_grad():
            conv.weight.copy_(_tf32_rna(torch.rand(conv.weight.shape, generator=g) / 576))
        xt = (x.view(torch.int32) & ~0x1FFF).view(torch.float32)
        with torch.no_grad():
            y_rna = F.conv2d(_tf32_rna(x).double(), conv.weight.double(), None, 1, 1).float()
            y_trunc = F.conv2d(xt.double(), conv.weight.double(), None, 1, 1).float()
        with torch.no_grad():
            got = ops.to_nchw(ops.conv_bn_act(ops.to_nhwc(x.to(DEV), torch.float32), conv.to(DEV), None)).cpu()
        e_rna, e_trunc = float((got - y_rna).abs().max()), float((got - y_trunc).abs().max())
        print("tf32 activation rounding: err vs rna %.3e, vs truncation %.3e" % (e_rna, e_trunc))
        assert e_rna <= 2e-5 * float(y_rna.abs().max()) and e_trunc > 5 * e_rna
    finally:
        m.set_precision("fp32")


# ---------------------------------------------------------------------------------------------
# 16-bit tensor-core DCN (fused tap-major offsets)
# ---------------------------------------------------------------------------------------------
def _to_tap_major(off, msk, G):
    from fami_pose_b200 import ops
    cat = torch.cat([off, msk], 1)
    return cat[:, ops.tap_major_perm(G)].contiguous()


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("shape", [(2, 48, 48, 12, 12, 9, 3.0), (1, 32, 32, 8, 10, 7, 2.0), (1, 64, 48, 16, 8, 8, 1.0),
                                   (2, 48, 48, 12, 96, 72, 2.0), (1, 48, 48, 12, 40, 30, 14.0), (1, 48, 17, 12, 33, 21, 2.0),
                                   # C > 64: channel passes of 64 channels, weights streamed per (pass, tap) (config 5: C = 128, 256)
                                   (1, 128, 128, 32, 20, 13, 2.0), (1, 256, 64, 64, 17, 9, 2.0), (2, 128, 48, 32, 16, 8, 7.0),
                                   (1, 256, 256, 64, 16, 8, 2.0)])
def test_dcn_tc_fused_vs_oracle(shape, prec):
    """tensor-core DCN (x/weights/columns in 16 bit, fp32 accumulate, fp32 offsets) vs the numpy oracle run
    on the SAME 16-bit-rounded x and weights; sigma=14 px exercises the out-of-window global path.
    Tolerance: 4e-3*max|ref| (fp16) / 2e-2*max|ref| (bf16) -- column + output rounding only."""
    import fami_pose_b200 as m
    from fami_pose_b200 import layers, ops
    B, C, Cout, G, H, W, sig = shape
    m.set_precision(prec)
    hdt = ops.act_dtype()
    try:
        g = torch.Generator().manual_seed(B * 1000 + C + H)
        x = torch.randn(B, C, H, W, generator=g).to(hdt).float()
        off = sig * torch.randn(B, 18 * G, H, W, generator=g)
        msk = torch.randn(B, 9 * G, H, W, generator=g)
        w = (0.05 * torch.randn(Cout, C, 3, 3, generator=g)).to(hdt).float()
        b = 0.1 * torch.randn(Cout, generator=g)
        ref = torch.from_numpy(fo.dcn_fwd(x.numpy(), off.numpy(), msk.numpy(), w.numpy(), b.numpy()))
        mod = layers.DeformConv2d(C, Cout, 3, padding=3, dilation=3).to(DEV)
        with torch.no_grad():
            mod.weight.copy_(w.to(DEV))
            mod.bias.copy_(b.to(DEV))
        om = ops.to_nhwc(_to_tap_major(off, msk, G).to(DEV), torch.float32)
        out = mod(ops.to_nhwc(x.to(DEV), hdt), None, None, fused_om=om)
        assert out.dtype == hdt
        got = ops.to_nchw(out).cpu()
        tol = (4e-3 if prec == "fp16" else 2e-2) * float(ref.abs().max()) + 1e-3
        err = float((got - ref).abs().max())
        print("dcn tc", prec, shape, "err", err, "tol", tol)
        assert err <= tol
    finally:
        m.set_precision("fp32")


@pytest.mark.parametrize("shape", [(2, 48, 48, 12, 33, 21, 2.0), (1, 32, 32, 8, 16, 8, 1.0), (1, 48, 48, 12, 40, 30, 14.0),
                                   (1, 128, 64, 32, 18, 11, 2.0)])
def test_dcn_tf32_arm_vs_oracle(shape):
    """The 'tf32' arm's deformable convolution: fp32 x in, fp32 out, the gather and the contraction on fp16 multiplicands
    (x cast on the way in: the same 11-bit significand as TF32), fp32 accumulation, fp32 offsets -- vs the numpy oracle on
    the UNROUNDED fp32 operands.  Tolerance 3e-3*max|ref| (three 2^-11 roundings of the sampled column per term)."""
    import fami_pose_b200 as m
    from fami_pose_b200 import layers, ops
    B, C, Cout, G, H, W, sig = shape
    m.set_precision("tf32")
    try:
        g = torch.Generator().manual_seed(B * 1000 + C + H)
        x = torch.randn(B, C, H, W, generator=g)
        off = sig * torch.randn(B, 18 * G, H, W, generator=g)
        msk = torch.randn(B, 9 * G, H, W, generator=g)
        w = 0.05 * torch.randn(Cout, C, 3, 3, generator=g)
        b = 0.1 * torch.randn(Cout, generator=g)
        ref = torch.from_numpy(fo.dcn_fwd(x.numpy(), off.numpy(), msk.numpy(), w.numpy(), b.numpy()))
        mod = layers.DeformConv2d(C, Cout, 3, padding=3, dilation=3).to(DEV)
        with torch.no_grad():
            mod.weight.copy_(w.to(DEV))
            mod.bias.copy_(b.to(DEV))
            om = ops.to_nhwc(_to_tap_major(off, msk, G).to(DEV), torch.float32)
            blk = ops.om_to_blocked(om, G, layout=ops.dcn_blocked_layout(C, Cout, G))
            xd = ops.to_nhwc(x.to(DEV), torch.float32)
            out = mod(xd, None, None, blocked_om=blk, groups=G)
            wide = ops.empty_nhwc(B, 2 * Cout, H, W, torch.float32, DEV)          # output slice of a wider buffer
            mod(xd, None, None, out=wide[:, Cout:], blocked_om=blk, groups=G)
        assert out.dtype == torch.float32
        got = ops.to_nchw(out).cpu()
        assert torch.equal(ops.to_nchw(wide[:, Cout:]).cpu(), got)
        tol = 3e-3 * float(ref.abs().max()) + 1e-3
        err = float((got - ref).abs().max())
        print("dcn tf32 arm", shape, "err", err, "tol", tol)
        assert err <= tol
    finally:
        m.set_precision("fp32")


@pytest.mark.parametrize("shape", [(2, 128, 128, 32, 35, 19), (1, 256, 64, 64, 16, 24), (2, 64, 64, 16, 20, 9), (1, 32, 32, 8, 17, 8),
                                   (2, 48, 48, 12, 35, 19), (1, 48, 48, 12, 100, 13), (2, 32, 32, 8, 49, 24), (1, 48, 17, 12, 20, 9)])
def test_dcn_tc_blocked_equals_tap_major(shape):
    """The blocked offset layout of the shape's deformable kernel (ops.dcn_blocked_layout: 3 for the warp-private kernel,
    C == Cout in {32, 48}; 2 for the tcgen05 kernel) and the tap-major NHWC layout (om_layout 1) give bit-identical outputs,
    incl. C > 64 (channel passes), maps that are not a multiple of the 16x8 layout tile / the 24-row compute tile, and
    several strips per image column."""
    import fami_pose_b200 as m
    from fami_pose_b200 import layers, ops
    B, C, Cout, G, H, W = shape
    m.set_precision("fp16")
    try:
        g = torch.Generator().manual_seed(C + H)
        x = ops.to_nhwc(torch.randn(B, C, H, W, generator=g).to(DEV), torch.float16)
        om = ops.to_nhwc((2.5 * torch.randn(B, 27 * G, H, W, generator=g)).to(DEV), torch.float32)
        mod = layers.DeformConv2d(C, Cout, 3, padding=3, dilation=3).to(DEV)
        with torch.no_grad():
            o1 = mod(x, None, None, fused_om=om)
            o2 = mod(x, None, None, blocked_om=ops.om_to_blocked(om, G, layout=ops.dcn_blocked_layout(C, Cout, G)), groups=G)
        assert torch.isfinite(o1.float()).all() and torch.equal(o1, o2)
    finally:
        m.set_precision("fp32")


@pytest.mark.parametrize("dil", [1, 2, 4])
def test_dcn_tc_other_dilations(dil):
    """The tensor-core kernel at the dilations the ABI admits besides the reference's 3 (pad == dil; window radius dil + 7),
    against the numpy oracle on fp16-rounded operands; offsets up to several pixels beyond the window reach."""
    import fami_pose_b200 as m
    from fami_pose_b200 import layers, ops
    B, C, Cout, G, H, W, sig = 2, 48, 48, 12, 21, 13, 3.0
    m.set_precision("fp16")
    try:
        g = torch.Generator().manual_seed(100 + dil)
        x = torch.randn(B, C, H, W, generator=g).half().float()
        off = sig * torch.randn(B, 18 * G, H, W, generator=g)
        msk = torch.randn(B, 9 * G, H, W, generator=g)
        w = (0.05 * torch.randn(Cout, C, 3, 3, generator=g)).half().float()
        b = 0.1 * torch.randn(Cout, generator=g)
        ref = torch.from_numpy(fo.dcn_fwd(x.numpy(), off.numpy(), msk.numpy(), w.numpy(), b.numpy(), pad=dil, dil=dil))
        mod = layers.DeformConv2d(C, Cout, 3, padding=dil, dilation=dil).to(DEV)
        with torch.no_grad():
            mod.weight.copy_(w.to(DEV))
            mod.bias.copy_(b.to(DEV))
            om = ops.to_nhwc(_to_tap_major(off, msk, G).to(DEV), torch.float32)
            o1 = mod(ops.to_nhwc(x.to(DEV), torch.float16), None, None, fused_om=om)
            o2 = mod(ops.to_nhwc(x.to(DEV), torch.float16), None, None,
                     blocked_om=ops.om_to_blocked(om, G, layout=ops.dcn_blocked_layout(C, Cout, G)), groups=G)
        assert torch.equal(o1, o2)
        got = ops.to_nchw(o1).cpu().float()
        tol = 4e-3 * float(ref.abs().max()) + 1e-3
        assert float((got - ref).abs().max()) <= tol
    finally:
        m.set_precision("fp32")


def test_dcn_tc_zero_offset_equals_dilated_conv_full_size():
    """Property at config-2 size (B=32): zero offsets + unit mask == the tensor-core dilated conv."""
    import fami_pose_b200 as m
    from fami_pose_b200 import layers, ops
    m.set_precision("fp16")
    try:
        g = torch.Generator().manual_seed(5)
        B, C, G, H, W = 32, 48, 12, 96, 72
        xd = ops.empty_nhwc(B, C, H, W, torch.float16, DEV).normal_()
        mod = layers.DeformConv2d(C, C, 3, padding=3, dilation=3).to(DEV)
        conv = torch.nn.Conv2d(C, C, 3, 1, 3, 3).to(DEV)
        with torch.no_grad():
            conv.weight.copy_(mod.weight)
            conv.bias.copy_(mod.bias)
        om = ops.empty_nhwc(B, 27 * G, H, W, torch.float32, DEV).zero_()
        om.view(B, 9, 3, G, H, W)[:, :, 2] = 1.0    # logical channel order [tap][dy|dx|mask][g]
        a = ops.to_nchw(mod(xd, None, None, fused_om=om)).cpu()
        b = ops.to_nchw(ops.conv_bn_act(xd, conv, None)).cpu()
        assert float((a - b).abs().max()) <= 4e-3 * float(b.abs().max()) + 1e-3
    finally:
        m.set_precision("fp32")


# ---------------------------------------------------------------------------------------------
# backward kernels (fp32 arm)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("arm", ["fp32", "tf32"])
@pytest.mark.parametrize("name", ["c48g12", "c32g8", "c16g1_oob", "c64g16"])
def test_dcn_bwd_vs_torchvision_autograd_golden(name, arm, golden_dir):
    """fami_dcn_bwd vs torchvision's autograd (committed fp64 golden: grads wrt input, offset, mask,
    weight, bias, incl. out-of-bounds offsets and G=1); tolerance 2e-4 * max|ref| (fp32, atomics).  On the 'tf32' arm the
    weight gradient's products run on TF32 tensor cores (columns and grad_out rounded to nearest TF32: 2e-3 * max|ref|),
    every other gradient stays exact fp32."""
    m = fp()
    m.set_precision(arm)
    try:
        _dcn_bwd_case(name, golden_dir, 2e-3 if arm == "tf32" else 2e-4)
    finally:
        m.set_precision("fp32")


def _dcn_bwd_case(name, golden_dir, gw_tol):
    from fami_pose_b200 import layers
    gold = np.load(os.path.join(golden_dir, "dcn_torchvision.npz"))
    case, (x, off, msk, w, b, go) = _golden_inputs(name)
    _, B, C, Cout, G, H, W, sig = case
    mod = layers.DeformConv2d(C, Cout, 3, padding=3, dilation=3).to(DEV)
    with torch.no_grad():
        mod.weight.copy_(w.to(DEV))
        mod.bias.copy_(b.to(DEV))
    xd, od, md = nhwc(x).requires_grad_(True), nhwc(off).requires_grad_(True), nhwc(msk).requires_grad_(True)
    out = mod(xd, od, md)
    out.backward(nhwc(go))
    got = {"gx": back(xd.grad), "goff": back(od.grad), "gmask": back(md.grad), "gw": mod.weight.grad.cpu(),
           "gb": mod.bias.grad.cpu()}
    for key, g in got.items():
        ref = torch.from_numpy(gold["%s_f64_%s" % (name, key)]).float()
        err = float((g - ref).abs().max())
        assert err <= (gw_tol if key == "gw" else 2e-4) * max(1.0, float(ref.abs().max())), (key, err)


def test_warp_translate_bwd_vs_torch_autograd():
    """fami_warp_translate_bwd vs torch autograd through the kornia restatement (CPU fp64); tol 1e-4 rel."""
    fp()
    from fami_pose_b200 import kornia_shim
    g = torch.Generator().manual_seed(31)
    B, C, H, W = 3, 16, 12, 9
    src = torch.randn(B, C, H, W, generator=g, dtype=torch.float64)
    txy = torch.tensor([[0.3, -1.7], [2.25, 0.5], [-0.6, 3.4]], dtype=torch.float64)
    go = torch.randn(B, C, H, W, generator=g, dtype=torch.float64)
    s_ref, t_ref = src.clone().requires_grad_(True), txy.clone().requires_grad_(True)
    M = torch.eye(3, dtype=torch.float64)[0:2].view(1, 2, 3).repeat(B, 1, 1)
    M = torch.cat([M[:, :, :2], t_ref.unsqueeze(2)], 2)
    fo.warp_affine_kornia(s_ref, M, (H, W)).backward(go)
    sd = nhwc(src.float()).requires_grad_(True)
    td = txy.float().to(DEV).requires_grad_(True)
    Md = torch.cat([torch.eye(3, device=DEV)[0:2].view(1, 2, 3).repeat(B, 1, 1)[:, :, :2], td.unsqueeze(2)], 2)
    out = kornia_shim.warp_affine(sd, Md, (H, W))
    out.backward(nhwc(go.float()))
    e1 = float((back(sd.grad).double() - s_ref.grad).abs().max())
    e2 = float((td.grad.cpu().double() - t_ref.grad).abs().max())
    assert e1 <= 1e-5 and e2 <= 1e-4 * max(1.0, float(t_ref.grad.abs().max())), (e1, e2)


@pytest.mark.parametrize("layout", [2, 3])
@pytest.mark.parametrize("shape", [(2, 48, 12, 33, 21), (1, 32, 8, 16, 8), (3, 64, 16, 20, 30), (2, 48, 12, 96, 72)])
def test_offset_conv_blocked_layout_and_dcn(shape, layout):
    """The fused offset|mask producer writing a blocked layout (fami_conv_desc.om_groups / om_layout: 2 row-blocked, read by
    the tcgen05 deformable kernel; 3 k-step-blocked, read by the warp-private one) equals the same conv written as an NHWC
    activation and converted on the host (ops.om_to_blocked), bit for bit -- incl. maps that are not a multiple of the 16x8
    DCN tile -- and the deformable kernel gives identical outputs from the blocked and the tap-major layout."""
    m = fp()
    from fami_pose_b200 import ops
    B, C, G, H, W = shape
    if layout == 3 and ops.dcn_blocked_layout(C, C, G) != 3:
        pytest.skip("layout 3 is the warp-private kernel's (C == Cout in {32, 48})")
    m.set_precision("fp16")
    try:
        g = torch.Generator().manual_seed(31 + C)
        x = ops.to_nhwc(torch.randn(B, C, H, W, generator=g).to(DEV), torch.float16)
        conv = torch.nn.Conv2d(C, 27 * G, 3, 1, 3, 3).to(DEV)
        with torch.no_grad():
            conv.weight.copy_(torch.randn(conv.weight.shape, generator=g).to(DEV) * 0.05)
            conv.bias.copy_(torch.randn(27 * G, generator=g).to(DEV))
            nhwc = ops.conv_bn_act(x, conv, None, relu=False, out_dtype=torch.float32)
            blk = ops.conv_offsets_blocked(x, conv, G, layout=layout)
            ref = ops.om_to_blocked(nhwc, G, layout=layout)
            # slots of pixels outside the image are never written by the producer nor read by the consumer
            msk = ops.om_to_blocked(torch.ones_like(nhwc), G, layout=layout) > 0       # the converter zero-fills the slots outside the image
            assert torch.equal(blk[msk], ref[msk])
            dcn = m.DeformConv2d(C, C, 3, padding=3, dilation=3).to(DEV)
            o1 = dcn(x, None, None, fused_om=nhwc)
            o2 = dcn(x, None, None, blocked_om=blk, groups=G)
            if layout == ops.dcn_blocked_layout(C, C, G):
                assert torch.equal(o1, o2)
            else:   # tap-major went to the warp-private kernel, layout 2 to the tcgen05 one: same math, another summation order
                assert float((o1.float() - o2.float()).abs().max()) <= 2e-3 * float(o1.float().abs().max())
    finally:
        m.set_precision("fp32")


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("shape", [(2, 64, 48), (3, 33, 47), (1, 384, 288)])
def test_stem_conv_tensor_core(shape, prec):
    """HRNet stem (3 -> 64, 3x3, stride 2, pad 1, fp32 pixels -> 16-bit activations, BN + ReLU; hrnet.py conv1/bn1) on
    the tensor-core kernel with its in-CTA im2col (csrc/stem_tc.cu) vs torch fp32, incl. odd sizes (image borders,
    ragged last tile).  Tolerance: 16-bit rounding of inputs, weights and outputs (fp16 2e-3, bf16 1.6e-2 relative)."""
    m = fp()
    from fami_pose_b200 import ops
    N, H, W = shape
    g = torch.Generator().manual_seed(77)
    x = torch.randn(N, 3, H, W, generator=g)
    conv = torch.nn.Conv2d(3, 64, 3, 2, 1, bias=False)
    bn = torch.nn.BatchNorm2d(64).eval()
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * 0.2)
        bn.weight.copy_(torch.rand(64, generator=g) + 0.5); bn.bias.copy_(torch.randn(64, generator=g) * 0.2)
        bn.running_mean.copy_(torch.randn(64, generator=g) * 0.2); bn.running_var.copy_(torch.rand(64, generator=g) + 0.5)
        ref = F.relu(bn(conv(x)))
    m.set_precision(prec)
    try:
        dt = ops.act_dtype()
        xn = ops.to_nhwc(x.to(DEV), torch.float32)
        with torch.no_grad():
            y = ops.conv_bn_act(xn, conv.to(DEV), bn.to(DEV), relu=True, out_dtype=dt)
        assert y.dtype == dt
        got = ops.to_nchw(y).float().cpu()
    finally:
        m.set_precision("fp32")
    tol = (2e-3 if prec == "fp16" else 1.6e-2) * float(ref.abs().max()) + 1e-3
    assert float((got - ref).abs().max()) <= tol
