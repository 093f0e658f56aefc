"""`kornia.geometry.warp_affine` stand-in for the one way the reference calls it
(Alignment_V15.py:133-135): M = [[1,0,tx],[0,1,ty]], dsize = input size, bilinear, zeros."""
from . import ops


def warp_affine(src, M, dsize, mode='bilinear', padding_mode='zeros', align_corners=True):
    if mode != 'bilinear' or padding_mode != 'zeros':
        raise NotImplementedError("only bilinear / zeros, as the reference uses")
    if tuple(dsize) != tuple(src.shape[-2:]):
        raise NotImplementedError("dsize must equal the input size")
    x = ops.to_nhwc(src)
    txy = M[:, :, 2]  # (tx, ty); the linear part must be the identity (not checked: would sync the host)
    import torch
    if torch.is_grad_enabled() and (x.requires_grad or txy.requires_grad):
        from .autograd import WarpTranslateFunction   # differentiable path (fp32 arm): fami_warp_translate_bwd
        return WarpTranslateFunction.apply(x, txy.float().contiguous())
    return ops.warp_translate(x, txy)
