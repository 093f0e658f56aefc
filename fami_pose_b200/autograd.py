"""autograd.Function wrappers (fp32 arm) for the two alignment ops whose backward kernels exist:
modulated deformable convolution and the global translation warp.  Operands are channels-last
activations (see ops.py); gradients come back in the same layout."""
import ctypes

import torch

from . import _lib, ops
from ._lib import DcnDesc, F32, TF32


class DeformConvFunction(torch.autograd.Function):
    """torchvision.ops.deform_conv2d forward/backward (fami_dcn_fwd / fami_dcn_bwd)."""

    @staticmethod
    def forward(ctx, x, offset, mask, weight, bias, owner, pad, dil):
        out = ops.dcn_fwd(x, offset, mask, weight, bias, owner, pad=pad, dil=dil)
        ctx.save_for_backward(x, offset, mask, weight, bias if bias is not None else torch.empty(0, device=x.device))
        ctx.cfg = (owner, pad, dil, bias is not None)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, offset, mask, weight, bias = ctx.saved_tensors
        owner, pad, dil, has_bias = ctx.cfg
        if x.dtype != torch.float32:
            raise NotImplementedError("fami_dcn_bwd is implemented for the fp32 arm")
        B, C, H, W, xp = ops.meta(x)
        Cout, _, kh, kw = weight.shape
        _, OC, _, _, offp = ops.meta(offset)
        _, MC, _, _, mp = ops.meta(mask)
        G = OC // (2 * kh * kw)
        go = ops.to_nhwc(grad_out.float(), torch.float32) if not ops.is_nhwc(grad_out) else grad_out
        gop = ops.meta(go)[4]
        dev = x.device
        gx = ops.empty_nhwc(B, C, H, W, torch.float32, dev)
        goff = ops.empty_nhwc(B, OC, H, W, torch.float32, dev)
        gmask = ops.empty_nhwc(B, MC, H, W, torch.float32, dev)
        cpad = _lib.load().fami_conv_cout_pad(Cout)
        gw = torch.empty(kh * kw * C * cpad, dtype=torch.float32, device=dev)
        gb = torch.empty(Cout, dtype=torch.float32, device=dev)
        w = ops.packed_weight(owner, weight, torch.float32)
        # 'tf32' arm: the weight gradient's products on TF32 tensor cores (everything else of the backward stays exact fp32)
        d = DcnDesc(B, H, W, C, Cout, G, kh, kw, 1, pad, dil, xp, offp, mp, gop, 0, TF32 if ops._PRECISION == "tf32" else F32)
        _lib.call("fami_dcn_bwd", ctypes.byref(d), ops._ptr(x), ops._ptr(offset), ops._ptr(mask), ops._ptr(w),
                  ops._ptr(go), ops._ptr(gx), ops._ptr(goff), ops._ptr(gmask), ops._ptr(gw), ops._ptr(gb), ops._stream())
        # packed [taps][C][CoutPad] -> OIHW
        gw_oihw = gw.view(kh * kw, C, cpad)[:, :, :Cout].permute(2, 1, 0).reshape(Cout, C, kh, kw).contiguous()
        return gx, goff, gmask, gw_oihw, (gb if has_bias else None), None, None, None


class WarpTranslateFunction(torch.autograd.Function):
    """kornia warp_affine for M=[[1,0,tx],[0,1,ty]] forward/backward (fami_warp_translate_fwd/bwd)."""

    @staticmethod
    def forward(ctx, src, txy):
        out = ops.warp_translate(src, txy)
        ctx.save_for_backward(src, txy)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        src, txy = ctx.saved_tensors
        if src.dtype != torch.float32:
            raise NotImplementedError("fami_warp_translate_bwd is implemented for the fp32 arm")
        B, C, H, W, sp = ops.meta(src)
        go = ops.to_nhwc(grad_out.float(), torch.float32) if not ops.is_nhwc(grad_out) else grad_out
        gop = ops.meta(go)[4]
        gs = ops.empty_nhwc(B, C, H, W, torch.float32, src.device)
        gt = torch.empty((B, 2), dtype=torch.float32, device=src.device)
        t = txy.detach().float().contiguous()
        _lib.call("fami_warp_translate_bwd", ops._ptr(src), sp, ops._ptr(t), ops._ptr(go), gop, ops._ptr(gs), C,
                  ops._ptr(gt), B, H, W, C, ops._stream())
        return gs, gt


# ------------------------------------------------------------------------------------------------
# dense pieces of the trainable head (fp32 arm): the training step of
# engine/core/functions/alignment_mi_function_term6_1.py:104-156 differentiates exactly these
# ------------------------------------------------------------------------------------------------

def _nhwc_f32(g):
    if g.dtype == torch.float32 and g.dim() == 4 and ops.is_nhwc(g):
        return g
    return ops.to_nhwc(g.float().contiguous(), torch.float32)


def needs_graph(*ts):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in ts)


class ConvBnActFunction(torch.autograd.Function):
    """conv -> [BatchNorm] -> [+residual] -> [ReLU] (basic_model.py:44-63, basic_layer.py:55-73) with
    fami_conv2d_dgrad / fami_conv2d_wgrad / fami_bn_bwd as the backward."""

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, residual, geom, bn, relu):
        k, stride, pad, dil = geom
        Cout = weight.shape[0]
        dev = x.device
        code = ops.conv_code(x, weight.shape[1])      # F32: exact SIMT kernels; TF32 ('tf32' arm): tcgen05 kind::tf32
        wp = ops.pack_weight(weight, code)
        b = bias.detach().float().contiguous() if bias is not None else None
        raw = mean = invstd = None
        if bn is None:
            y = ops._conv_raw(x, wp, Cout, k, stride, pad, dil, None, b, residual, relu, 1, None, None, torch.float32, code=code)
        else:
            training = bn.training or bn.running_mean is None
            fused_stats = training and code == F32      # the tensor-core kernels take the statistics in a separate pass
            stats = torch.zeros(2 * Cout, dtype=torch.float64, device=dev) if training else None
            raw = ops._conv_raw(x, wp, Cout, k, stride, pad, dil, None, b, None, False, 1, None, stats if fused_stats else None,
                                torch.float32, code=code)
            N, _, Ho, Wo, rawp = ops.meta(raw)
            if training and not fused_stats:
                _lib.call("fami_bn_stats", ops._ptr(raw), F32, rawp, N * Ho * Wo, Cout, ops._ptr(stats), ops._stream())
            scale = torch.empty(Cout, dtype=torch.float32, device=dev)
            shift = torch.empty_like(scale)
            if training:
                mean = torch.empty_like(scale)
                invstd = torch.empty_like(scale)
                track = bn.track_running_stats and bn.running_mean is not None
                mom = bn.momentum if bn.momentum is not None else 0.1
                _lib.call("fami_bn_finalize", ops._ptr(stats), ops._ptr(gamma), ops._ptr(beta),
                          ops._ptr(bn.running_mean if track else None), ops._ptr(bn.running_var if track else None),
                          ops._ptr(scale), ops._ptr(shift), ops._ptr(mean), ops._ptr(invstd), Cout, N * Ho * Wo,
                          float(bn.eps), float(mom), ops._stream())
                if track:
                    torch.autograd.graph.increment_version((bn.running_mean, bn.running_var))   # see ops.conv_bn_act
                    if bn.num_batches_tracked is not None:
                        bn.num_batches_tracked += 1
            else:
                with torch.no_grad():
                    mean = bn.running_mean.float().clone()
                    invstd = torch.rsqrt(bn.running_var.double() + bn.eps).float()
                    g = gamma.detach().float() if gamma is not None else torch.ones_like(mean)
                    bb = beta.detach().float() if beta is not None else torch.zeros_like(mean)
                    scale = (g * invstd).contiguous()
                    shift = (bb - mean * scale).contiguous()
            y = ops.empty_nhwc(N, Cout, Ho, Wo, torch.float32, dev)
            rp = ops.meta(residual)[4] if residual is not None else 0
            _lib.call("fami_bn_apply_act", ops._ptr(raw), F32, rawp, ops._ptr(scale), ops._ptr(shift), ops._ptr(residual),
                      rp, ops._ptr(y), ops.meta(y)[4], F32, N, Ho, Wo, Cout, 1, int(bool(relu)), ops._stream())
            ctx.training = training
        ctx.geom, ctx.relu, ctx.has_bn = geom, bool(relu), bn is not None
        ctx.has_bias, ctx.has_res = bias is not None, residual is not None
        ctx.x_shape = tuple(x.shape)
        ctx.save_for_backward(x, weight, raw, y if relu else None, mean, invstd, gamma)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight, raw, y, mean, invstd, gamma = ctx.saved_tensors
        k, stride, pad, dil = ctx.geom
        need_x, need_w, need_b, need_g, need_beta, need_res = ctx.needs_input_grad[:6]
        g = _nhwc_f32(gy)
        dgamma = dbeta = gres = None
        if ctx.has_bn:
            g, dgamma, dbeta, gres = ops.bn_bwd(raw, g, mean, invstd, gamma, y=y, training=ctx.training,
                                                want_res=ctx.has_res and need_res)
        elif ctx.relu:
            C = g.shape[1]
            zero = torch.zeros(C, dtype=torch.float32, device=g.device)
            one = torch.ones(C, dtype=torch.float32, device=g.device)
            g, _, _, _ = ops.bn_bwd(g, g, zero, one, None, y=y, training=False)
            gres = g if ctx.has_res else None
        else:
            gres = g if ctx.has_res else None
        gw = gb = gx = None
        if need_w or (ctx.has_bias and need_b):
            gw, gb = ops.conv_wgrad(x, g, tuple(weight.shape), stride, pad, dil, want_bias=ctx.has_bias and need_b)
        if need_x:
            gx = ops.conv_dgrad(g, weight, ctx.x_shape, stride, pad, dil)
        return (gx, gw if need_w else None, gb, dgamma if need_g else None, dbeta if need_beta else None,
                gres if need_res else None, None, None, None)


def conv_bn_act(x, conv, bn, relu, residual, weight=None, bias=None):
    """Differentiable ops.conv_bn_act (fp32 arm, no upsample-on-write, freshly allocated output)."""
    if x.dtype != torch.float32:
        raise NotImplementedError("the differentiable path runs on fp32 storage: fami.set_precision('fp32' | 'tf32')")
    weight = conv.weight if weight is None else weight
    bias = conv.bias if bias is None else bias
    geom = (conv.kernel_size[0], conv.stride[0], conv.padding[0], conv.dilation[0])
    gamma = bn.weight if bn is not None else None
    beta = bn.bias if bn is not None else None
    return ConvBnActFunction.apply(x, weight, bias, gamma, beta, residual, geom, bn, relu)


class UpsampleAddReluFunction(torch.autograd.Function):
    """y = [ReLU](residual + nearest_upsample(t, up)): the tail of an HRNet fuse-layer term (hrnet.py:99-112 Interpolate +
    the running sum / final ReLU of :151-172).  Inference fuses this into the 1x1 conv's store; the differentiable path
    runs it as its own launch (fami_bn_apply_act with an identity affine) so that autograd can see it, with
    fami_upsample_add_bwd as the backward (replica sum + ReLU mask + residual gradient in one pass)."""

    @staticmethod
    def forward(ctx, t, residual, up, relu):
        N, C, H, W, tp = ops.meta(t)
        dev = t.device
        one = torch.ones(C, dtype=torch.float32, device=dev)
        zero = torch.zeros(C, dtype=torch.float32, device=dev)
        y = ops.empty_nhwc(N, C, H * up, W * up, torch.float32, dev)
        rp = ops.meta(residual)[4] if residual is not None else 0
        _lib.call("fami_bn_apply_act", ops._ptr(t), F32, tp, ops._ptr(one), ops._ptr(zero), ops._ptr(residual), rp, ops._ptr(y),
                  ops.meta(y)[4], F32, N, H, W, C, up, int(bool(relu)), ops._stream())
        ctx.up, ctx.relu, ctx.has_res = up, bool(relu), residual is not None
        ctx.small = (N, C, H, W)
        ctx.save_for_backward(y if relu else None)
        return y

    @staticmethod
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        N, C, H, W = ctx.small
        g = _nhwc_f32(gy)
        need_res = ctx.has_res and ctx.needs_input_grad[1]
        gt = ops.empty_nhwc(N, C, H, W, torch.float32, g.device)
        gres = ops.empty_nhwc(N, C, H * ctx.up, W * ctx.up, torch.float32, g.device) if need_res else None
        _lib.call("fami_upsample_add_bwd", ops._ptr(g), ops.meta(g)[4], ops._ptr(y), ops.meta(y)[4] if y is not None else 0,
                  ops._ptr(gt), ops.meta(gt)[4], ops._ptr(gres), ops.meta(gres)[4] if gres is not None else 0, N, H, W, C,
                  ctx.up, ops._stream())
        return gt, gres, None, None


class LinearFunction(torch.autograd.Function):
    """nn.Linear of feat_global_offset_layers[7..9] (Alignment_V15.py:69-71)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return ops.linear(x, weight, bias)

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        gx, gw, gb = ops.linear_bwd(x, weight, gy, want_x=ctx.needs_input_grad[0])
        return gx, gw, (gb if ctx.has_bias else None)


class SoftmaxPklFunction(torch.autograd.Function):
    """MI estimator core (Alignment_V15.py:250-277).  The reference detaches the `input` operand, so only
    the target operand b receives a gradient."""

    @staticmethod
    def forward(ctx, a, b, temperature):
        ctx.save_for_backward(a, b)
        ctx.temperature = temperature
        return ops.softmax_pkl(a, b, temperature)

    @staticmethod
    def backward(ctx, gout):
        a, b = ctx.saved_tensors
        _, gb = ops.softmax_pkl_bwd(a, b, gout, ctx.temperature, want_a=False, want_b=True)
        return None, gb, None


class JointMSEFunction(torch.autograd.Function):
    """JointMSELoss (mse_loss.py:21-40): loss and d loss / d pred in one pass (fami_joint_mse_fwd_bwd)."""

    @staticmethod
    def forward(ctx, pred, target, target_weight):
        p = pred if ops.is_nhwc(pred) else ops.to_nhwc(pred.float().contiguous(), torch.float32)
        loss, grad = ops.joint_mse(p, target, target_weight, want_grad=True)
        ctx.save_for_backward(grad)                     # [B,H,W,J]
        return loss

    @staticmethod
    def backward(ctx, gout):
        (grad,) = ctx.saved_tensors
        return (grad * gout).permute(0, 3, 1, 2), None, None


class SubFunction(torch.autograd.Function):
    """sup_bb_feat - kf_bb_feat (Alignment_V15.py:132)."""

    @staticmethod
    def forward(ctx, a, b):
        return ops.sub_bcast(a, b, 1)

    @staticmethod
    def backward(ctx, g):
        return g, -g


def cat_channels(ts):
    """torch.cat(dim=1) of channels-last activations (Alignment_V15.py:139,143,160); autograd slices the
    gradient back.  The inference path writes producers straight into slices instead."""
    out = torch.cat(list(ts), 1)
    return out if ops.is_nhwc(out) else out.contiguous(memory_format=torch.channels_last)
