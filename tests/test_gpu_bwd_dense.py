"""Backward of the dense pieces (conv dgrad / wgrad, BatchNorm, MI estimator, Linear) through the C ABI,
checked against torch autograd of the same fp32 ops evaluated in float64 on the CPU (the reference's
training step derives exactly these through autograd: alignment_mi_function_term6_1.py:104-156)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import fami_pose_b200 as fp  # noqa: E402
from fami_pose_b200 import ops  # noqa: E402


def _nhwc(t):
    return ops.to_nhwc(t.float().cuda(), torch.float32)


def _rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


CONV_CASES = [
    # N, Cin, Cout, H, W, k, stride, pad, dil, bias
    (2, 48, 48, 24, 18, 3, 1, 1, 1, False),     # BasicBlock conv
    (2, 48, 324, 12, 9, 3, 1, 3, 3, True),      # fused offset|mask conv (dilated)
    (3, 16, 16, 13, 11, 3, 2, 1, 1, True),      # global-offset stride-2 chain, odd sizes
    (2, 96, 48, 10, 7, 1, 1, 0, 1, False),      # 1x1 downsample branch
    (2, 48, 17, 9, 8, 3, 1, 1, 1, True),        # agg_final_layer
    (1, 192, 48, 6, 5, 3, 1, 1, 1, False),      # Cin > 64 (two ci tiles in wgrad)
    (2, 20, 70, 7, 6, 3, 2, 1, 1, True),        # ragged channel counts
]


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv_dgrad_wgrad_vs_autograd(case):
    N, Cin, Cout, H, W, k, stride, pad, dil, bias = case
    fp.set_precision("fp32")
    g = torch.Generator().manual_seed(100 + Cin + Cout)
    x = torch.randn(N, Cin, H, W, generator=g, dtype=torch.float64, requires_grad=True)
    w = (torch.randn(Cout, Cin, k, k, generator=g, dtype=torch.float64) * 0.1).requires_grad_()
    b = torch.randn(Cout, generator=g, dtype=torch.float64, requires_grad=True) if bias else None
    y = F.conv2d(x, w, b, stride, pad, dil)
    gy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(gy)
    gyn = _nhwc(gy)
    gx = ops.conv_dgrad(gyn, w.detach().float().cuda(), (N, Cin, H, W), stride, pad, dil)
    gw, gb = ops.conv_wgrad(_nhwc(x.detach()), gyn, (Cout, Cin, k, k), stride, pad, dil, want_bias=bias)
    assert _rel(ops.to_nchw(gx).double().cpu(), x.grad) < 2e-5
    assert _rel(gw.double().cpu(), w.grad) < 2e-5
    if bias:
        assert _rel(gb.double().cpu(), b.grad) < 2e-5


def _tf32_rna64(t):
    f = t.float().contiguous()
    i = f.view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32).double()


@pytest.mark.parametrize("case", [c for c in CONV_CASES if c[6] == 1 and c[2] % 4 == 0],
                         ids=lambda c: "x".join(map(str, c)))
def test_conv_dgrad_tf32_vs_autograd(case):
    """Stride-1 dgrad on the 'tf32' arm (the forward tensor-core kernels over the flipped / transposed filter, incl. the
    324-channel offset|mask gradient whose last 8-channel K-step is half zero fill): grad_y and the weights pre-rounded to
    TF32 so that only the accumulation order differs from float64 autograd."""
    N, Cin, Cout, H, W, k, stride, pad, dil, bias = case
    fp.set_precision("tf32")
    try:
        g = torch.Generator().manual_seed(200 + Cin + Cout)
        x = torch.randn(N, Cin, H, W, generator=g, dtype=torch.float64, requires_grad=True)
        w = _tf32_rna64(torch.randn(Cout, Cin, k, k, generator=g, dtype=torch.float64) * 0.1)
        y = F.conv2d(x, w, None, stride, pad, dil)
        gy = _tf32_rna64(torch.randn(y.shape, generator=g, dtype=torch.float64))
        y.backward(gy)
        gyn = _nhwc(gy)
        assert ops.conv_code(gyn, Cout) == fp._lib.TF32
        gx = ops.conv_dgrad(gyn, w.float().cuda(), (N, Cin, H, W), stride, pad, dil)
        assert _rel(ops.to_nchw(gx).double().cpu(), x.grad) < 2e-5
    finally:
        fp.set_precision("fp32")


@pytest.mark.parametrize("case", CONV_CASES + [(2, 48, 324, 40, 30, 3, 1, 3, 3, True)],
                         ids=lambda c: "x".join(map(str, c)))
def test_conv_wgrad_tf32_vs_autograd(case):
    """wgrad on the 'tf32' arm (mma.sync TF32, fp32 accumulation, fp32 atomics over pixel slabs; csrc/bwd_dense.cu
    conv_wgrad_tf32_kernel): x and grad_y pre-rounded to TF32, so that only the accumulation order differs from float64
    autograd -- incl. strides, dilation 3, two ci tiles, ragged channel counts (element-wise staging) and a map with several
    pixel slabs per tile."""
    N, Cin, Cout, H, W, k, stride, pad, dil, bias = case
    fp.set_precision("tf32")
    try:
        g = torch.Generator().manual_seed(300 + Cin + Cout)
        x = _tf32_rna64(torch.randn(N, Cin, H, W, generator=g, dtype=torch.float64))
        w = (torch.randn(Cout, Cin, k, k, generator=g, dtype=torch.float64) * 0.1).requires_grad_()
        b = torch.randn(Cout, generator=g, dtype=torch.float64, requires_grad=True) if bias else None
        y = F.conv2d(x, w, b, stride, pad, dil)
        gy = _tf32_rna64(torch.randn(y.shape, generator=g, dtype=torch.float64))
        y.backward(gy)
        gw, gb = ops.conv_wgrad(_nhwc(x), _nhwc(gy), (Cout, Cin, k, k), stride, pad, dil, want_bias=bias)
        assert _rel(gw.double().cpu(), w.grad) < 2e-5
        if bias:
            assert _rel(gb.double().cpu(), b.grad) < 2e-5
        # unrounded operands: the result moves by TF32's rounding of the operands only (2^-11 relative per factor)
        xu = torch.randn(N, Cin, H, W, generator=g, dtype=torch.float64)
        gyu = torch.randn(y.shape, generator=g, dtype=torch.float64)
        w2 = w.detach().clone().requires_grad_()
        F.conv2d(xu, w2, None, stride, pad, dil).backward(gyu)
        gw2, _ = ops.conv_wgrad(_nhwc(xu), _nhwc(gyu), (Cout, Cin, k, k), stride, pad, dil)
        assert _rel(gw2.double().cpu(), w2.grad) < 2e-3
    finally:
        fp.set_precision("fp32")


def test_conv_dgrad_into_channel_slice():
    """grad_x written into a channel slice of a wider buffer (the concat operands of Alignment_V15.py:143,160)."""
    fp.set_precision("fp32")
    g = torch.Generator().manual_seed(5)
    N, Cin, Cout, H, W = 2, 48, 48, 8, 6
    x = torch.randn(N, Cin, H, W, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(Cout, Cin, 3, 3, generator=g, dtype=torch.float64) * 0.1
    y = F.conv2d(x, w, None, 1, 1, 1)
    gy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(gy)
    wide = ops.empty_nhwc(N, 96, H, W, torch.float32, "cuda").zero_()
    ops.conv_dgrad(_nhwc(gy), w.float().cuda(), (N, Cin, H, W), 1, 1, 1, out=wide[:, 48:])
    assert _rel(ops.to_nchw(wide[:, 48:]).double().cpu(), x.grad) < 2e-5
    assert float(wide[:, :48].abs().max()) == 0.0


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("relu,res", [(True, True), (True, False), (False, False)])
def test_bn_bwd_vs_autograd(training, relu, res):
    fp.set_precision("fp32")
    g = torch.Generator().manual_seed(11)
    N, C, H, W = 3, 48, 9, 7
    x = (torch.randn(N, C, H, W, generator=g, dtype=torch.float64) * 2 + 0.5).requires_grad_()
    r = torch.randn(N, C, H, W, generator=g, dtype=torch.float64, requires_grad=True)
    gamma = (torch.rand(C, generator=g, dtype=torch.float64) + 0.5).requires_grad_()
    beta = torch.randn(C, generator=g, dtype=torch.float64, requires_grad=True)
    rm = torch.randn(C, generator=g, dtype=torch.float64) * 0.1
    rv = torch.rand(C, generator=g, dtype=torch.float64) + 0.5
    eps = 1e-5
    z = F.batch_norm(x, rm.clone(), rv.clone(), gamma, beta, training, 0.1, eps)
    if res:
        z = z + r
    y = F.relu(z) if relu else z
    gy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(gy)
    if training:
        mean = x.detach().mean((0, 2, 3))
        var = x.detach().var((0, 2, 3), unbiased=False)
    else:
        mean, var = rm, rv
    invstd = 1.0 / torch.sqrt(var + eps)
    gx, dgamma, dbeta, gres = ops.bn_bwd(_nhwc(x.detach()), _nhwc(gy), mean.float().cuda(), invstd.float().cuda(),
                                         gamma.detach().float().cuda(), y=_nhwc(y.detach()) if relu else None,
                                         training=training, want_res=res)
    assert _rel(ops.to_nchw(gx).double().cpu(), x.grad) < 5e-5
    assert _rel(dgamma.double().cpu(), gamma.grad) < 5e-5
    assert _rel(dbeta.double().cpu(), beta.grad) < 5e-5
    if res:
        assert _rel(ops.to_nchw(gres).double().cpu(), r.grad) < 1e-6


@pytest.mark.parametrize("C,H,W", [(17, 12, 9), (48, 24, 18)])
def test_softmax_pkl_bwd_vs_autograd(C, H, W):
    """MI estimator gradient incl. the reference's quirk (probabilities fed as log-probs, Alignment_V15.py:250-277)."""
    fp.set_precision("fp32")
    g = torch.Generator().manual_seed(3)
    B, T = 3, 0.05
    a = (torch.randn(B, C, H, W, generator=g, dtype=torch.float64) * 0.1).requires_grad_()
    b = (torch.randn(B, C, H, W, generator=g, dtype=torch.float64) * 0.1).requires_grad_()
    p = F.softmax(a.reshape(B * C, -1) / T, dim=1)
    t = F.softmax(b.reshape(B * C, -1) / T, dim=1)
    loss = F.kl_div(p, t, reduction="mean")
    (loss * 1.7).backward()
    an, bn = _nhwc(a.detach()), _nhwc(b.detach())
    fwd = ops.softmax_pkl(an, bn, T)
    assert abs(float(fwd) - float(loss)) < 1e-5 * max(1.0, abs(float(loss)))
    ga, gb = ops.softmax_pkl_bwd(an, bn, torch.tensor(1.7, device="cuda"), T)
    assert _rel(ops.to_nchw(ga).double().cpu(), a.grad) < 1e-4
    assert _rel(ops.to_nchw(gb).double().cpu(), b.grad) < 1e-4


def test_linear_bwd_vs_autograd():
    fp.set_precision("fp32")
    g = torch.Generator().manual_seed(4)
    M, K, N = 8, 144, 64
    x = torch.randn(M, K, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(N, K, generator=g, dtype=torch.float64, requires_grad=True)
    b = torch.randn(N, generator=g, dtype=torch.float64, requires_grad=True)
    y = F.linear(x, w, b)
    gy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(gy)
    gx, gw, gb = ops.linear_bwd(x.detach().float().cuda(), w.detach().float().cuda(), gy.float().cuda())
    assert _rel(gx.double().cpu(), x.grad) < 1e-5
    assert _rel(gw.double().cpu(), w.grad) < 1e-5
    assert _rel(gb.double().cpu(), b.grad) < 1e-5
