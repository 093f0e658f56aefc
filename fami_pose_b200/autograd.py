"""autograd.Function wrappers (fp32 arm) for the two alignment ops whose backward kernels exist:
modulated deformable convolution and the global translation warp.  Operands are channels-last
activations (see ops.py); gradients come back in the same layout."""
import ctypes

import torch

from . import _lib, ops
from ._lib import DcnDesc, F32


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
        d = DcnDesc(B, H, W, C, Cout, G, kh, kw, 1, pad, dil, xp, offp, mp, gop, 0, F32)
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
