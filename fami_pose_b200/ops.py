"""Functional ops over libfami_b200.so.

Activations travel between ops as torch tensors with LOGICAL shape [N,C,H,W] and channels-last
strides (element (n,c,y,x) at ((n*H+y)*W+x)*pitch + c).  `pitch` may exceed C: a channel slice of a
wider buffer is a valid operand, which is how the reference's torch.cat(dim=1) calls are executed
without copies.  torch is used for memory, streams and parameter bookkeeping only; every FLOP on
activations is issued through the C ABI.
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import BF16, F16, F32, TF32, ConvDesc, DcnDesc

_PRECISION = "fp32"


_DTYPES = {"fp32": torch.float32, "tf32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}


_STREAM_F32 = False


def stream_f32():
    return _STREAM_F32


def set_precision(p, stream_f32=None):
    """'fp32': exact-fp32 SIMT kernels (bit-for-bit fp32 FMA arithmetic).
    'tf32': fp32 STORAGE everywhere (activations, residual stream, offsets, heatmaps) with every convolution on
    tcgen05.mma.kind::tf32 -- multiplicands rounded to TF32 (10-bit mantissa; weights at pack time, activations by
    the TMA load), fp32 accumulation.  This is the arithmetic the reference's fp32 nn.Conv2d gets from cuDNN on a
    GPU (torch.backends.cudnn.allow_tf32 = True by default); it meets the fp32 tier's 1e-3 tolerance.
    'fp16' / 'bf16': tcgen05 tensor-core arm with 16-bit activations (fp32 accumulation, fp32
    offsets/masks/heatmaps); fp16 carries 3 more mantissa bits than bf16 at the same tensor-core rate.
    'fp16s' = 'fp16' with the fp32 residual stream (stream_f32=True): fp16 multiplicands carry the 11-bit significand of
    TF32, every residual / running sum is kept and added in fp32, accumulation is fp32 -- the TF32 error (5.5e-4 / 8.3e-4
    against the reference on config 2) at kind::f16 MMA rates."""
    global _PRECISION, _STREAM_F32
    if p == "fp16s":
        p, stream_f32 = "fp16", True
    if p not in _DTYPES:
        raise ValueError("precision must be one of %s" % sorted(_DTYPES))
    _PRECISION = p
    # fp32 residual stream on the 16-bit arms: every conv output that can later serve as a residual is stored twice --
    # rounded to 16 bit (the tensor-core operand) and unrounded in fp32 (the residual) -- so the trunk's ~100 sequential
    # residual additions accumulate in fp32.  Default: on for bf16 (8-bit significand: 1.1e-2 / 2.1e-2 without it, inside
    # north_star's 1e-2 with it), off for fp16 (meets 1e-2 either way; the extra fp32 traffic costs throughput).
    env = os.environ.get("FAMI_STREAM_F32")
    _STREAM_F32 = (p == "bf16") if stream_f32 is None and env is None else bool(int(env) if stream_f32 is None else stream_f32)
    _STREAM_F32 = _STREAM_F32 and p in ("bf16", "fp16")


def get_precision():
    return _PRECISION


def act_dtype():
    return _DTYPES[_PRECISION]


def _code(dtype):
    if dtype == torch.float32:
        return F32
    if dtype == torch.bfloat16:
        return BF16
    if dtype == torch.float16:
        return F16
    raise TypeError("unsupported activation dtype %s" % dtype)


def conv_code(x, Cin=None):
    """Descriptor dtype of a convolution / deformable convolution over activation x: the storage code of x, or
    TF32 when the 'tf32' arm is active and the shape fits the tensor-core path (Cin multiple of 8, 16-byte pixel
    pitch; the 3-channel stem stays on the exact-fp32 kernel)."""
    if _PRECISION == "tf32" and x.dtype == torch.float32:
        N, C, H, W, p = meta(x)
        if (Cin or C) % 4 == 0 and p % 4 == 0:
            return TF32
    return _code(x.dtype)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("fami_pose_b200 ops need CUDA tensors (no CPU fallback); got device %s" % t.device)


def empty_nhwc(N, C, H, W, dtype, device, pitch=None):
    p = C if pitch is None else pitch
    return torch.empty_strided((N, C, H, W), (H * W * p, 1, W * p, p), dtype=dtype, device=device)


def meta(t):
    """(N, C, H, W, pitch) of a channels-last operand; raises if the strides are not NHWC-with-pitch."""
    if t.dim() != 4:
        raise ValueError("expected a 4-D activation, got shape %s" % (tuple(t.shape),))
    N, C, H, W = t.shape
    s = t.stride()
    if W > 1:
        p = s[3]
    elif H > 1:
        p = s[2]
    elif N > 1:
        p = s[0]
    else:
        p = C
    ok = (C == 1 or s[1] == 1) and (W == 1 or s[3] == p) and (H == 1 or s[2] == W * p) and (N == 1 or s[0] == H * W * p) and p >= C
    if not ok:
        raise ValueError("activation is not channels-last with a uniform pixel pitch: shape %s strides %s"
                         % (tuple(t.shape), s))
    return N, C, H, W, p


def is_nhwc(t):
    try:
        meta(t)
        return True
    except ValueError:
        return False


def to_nhwc(x, dtype=None):
    """Boundary conversion from the reference's NCHW fp32 tensors (fami_nchw_to_nhwc).

    Tensors that already carry channels-last strides (outputs of other fami ops) pass through."""
    _need_cuda(x)
    dtype = dtype or act_dtype()
    if x.dim() == 4 and is_nhwc(x) and x.dtype in (torch.float32, torch.bfloat16, torch.float16):
        return x
    if x.dtype != torch.float32:
        raise TypeError("boundary tensors must be float32 NCHW, got %s" % x.dtype)
    x = x.contiguous()
    N, C, H, W = x.shape
    out = empty_nhwc(N, C, H, W, dtype, x.device)
    _lib.call("fami_nchw_to_nhwc", _ptr(x), C * H * W, _ptr(out), _code(dtype), N, C, H, W, C, _stream())
    return out


def frames_to_nhwc(kf_x, sup_x, dtype=torch.float32):
    """Alignment_V15.py:115-119: [B,3,H,W] + [B,3*ns,H,W] -> frame-major [(1+ns)*B,H,W,3] (float32: the
    stem convolution reads fp32 pixels in both precisions)."""
    _need_cuda(kf_x, sup_x)
    kf_x = kf_x.contiguous()
    sup_x = sup_x.contiguous()
    B, _, H, W = kf_x.shape
    ns = sup_x.shape[1] // 3
    out = empty_nhwc((1 + ns) * B, 3, H, W, dtype, kf_x.device)
    _lib.call("fami_nchw_to_nhwc", _ptr(kf_x), 3 * H * W, _ptr(out[:B]), _code(dtype), B, 3, H, W, 3, _stream())
    for i in range(ns):
        src = sup_x[:, 3 * i:3 * i + 3]
        _lib.call("fami_nchw_to_nhwc", ctypes.c_void_p(src.data_ptr()), 3 * ns * H * W,
                  _ptr(out[(1 + i) * B:(2 + i) * B]), _code(dtype), B, 3, H, W, 3, _stream())
    return out


IMAGENET_MEAN = (0.485, 0.456, 0.406)     # datasets/transforms/build.py:13-14 (RGB)
IMAGENET_STD = (0.229, 0.224, 0.225)


def frames_u8_to_nhwc(kf_u8, sup_u8, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """uint8 RGB frames as a loader holds them -- key frames [B,H,W,3], supporting frames [B,ns,H,W,3] -- to the
    normalised frame-major fp32 NHWC batch [(1+ns)*B, H, W, 3] the backbone consumes: ToTensor + Normalize
    (datasets/transforms/build.py:13-22) fused with the re-batching of Alignment_V15.py:115-119.  A quarter of the
    host->device bytes of the float path."""
    _need_cuda(kf_u8, sup_u8)
    if kf_u8.dtype != torch.uint8 or sup_u8.dtype != torch.uint8:
        raise TypeError("frames_u8_to_nhwc expects uint8 frames")
    kf_u8, sup_u8 = kf_u8.contiguous(), sup_u8.contiguous()
    B, H, W, _ = kf_u8.shape
    ns = sup_u8.shape[1]
    if tuple(sup_u8.shape) != (B, ns, H, W, 3) or kf_u8.shape[3] != 3:
        raise ValueError("expected key frames [B,H,W,3] and supporting frames [B,ns,H,W,3]")
    out = empty_nhwc((1 + ns) * B, 3, H, W, torch.float32, kf_u8.device)
    m = (ctypes.c_float * 3)(*mean)
    s = (ctypes.c_float * 3)(*std)
    fb = H * W * 3
    _lib.call("fami_frames_u8_normalize", _ptr(kf_u8), fb, _ptr(out[:B]), B, H, W, m, s, _stream())
    for i in range(ns):
        src = ctypes.c_void_p(sup_u8.data_ptr() + i * fb)
        _lib.call("fami_frames_u8_normalize", src, ns * fb, _ptr(out[(1 + i) * B:(2 + i) * B]), B, H, W, m, s, _stream())
    return out


def to_nchw(t):
    """NHWC(+pitch) activation -> contiguous float32 NCHW (fami_nhwc_to_nchw)."""
    _need_cuda(t)
    N, C, H, W, p = meta(t)
    out = torch.empty((N, C, H, W), dtype=torch.float32, device=t.device)
    _lib.call("fami_nhwc_to_nchw", _ptr(t), _code(t.dtype), p, _ptr(out), N, C, H, W, _stream())
    return out


# ------------------------------------------------------------------------------------------------
# parameter caches (packed weights, folded BN)
# ------------------------------------------------------------------------------------------------

def _ver(*ts):
    return tuple((t._version, t.data_ptr()) if t is not None else None for t in ts)


_CODE_STORAGE = {F32: torch.float32, TF32: torch.float32, BF16: torch.bfloat16, F16: torch.float16}


def _as_code(dtype):
    return dtype if isinstance(dtype, int) else _code(dtype)


def packed_weight(owner, weight, dtype):
    """Kernel-ready packing of an OIHW weight (fami_pack_conv_weight), cached per parameter version.  `dtype`: a
    torch dtype or a descriptor code (TF32: fp32 storage rounded to tf32, K-major 32-channel rows)."""
    code = _as_code(dtype)
    key = ("w", code)
    cache = owner.__dict__.setdefault("_fami_cache", {})
    ver = _ver(weight)
    hit = cache.get(key)
    if hit is not None and hit[0] == ver:
        return hit[1]
    out = pack_weight(weight, code)
    cache[key] = (ver, out)
    return out


def pack_weight(weight, dtype):
    """Uncached packing of an OIHW weight tensor (training: weights change every step; derived weights)."""
    code = _as_code(dtype)
    Cout, Cin, kh, kw = weight.shape
    n = _lib.load().fami_packed_weight_elems(Cout, Cin, kh, kw, code)
    out = torch.empty(n, dtype=_CODE_STORAGE[code], device=weight.device)
    w = weight.detach().contiguous().float()
    _lib.call("fami_pack_conv_weight", _ptr(w), _ptr(out), Cout, Cin, kh, kw, code, _stream())
    return out


def folded_affine(conv_bias, bn):
    """Per-channel (scale, shift) of eval-mode BatchNorm folded with the conv bias."""
    owner = bn if bn is not None else None
    if bn is None:
        return None, (conv_bias.detach().float() if conv_bias is not None else None)
    cache = owner.__dict__.setdefault("_fami_cache", {})
    ver = _ver(bn.weight, bn.bias, bn.running_mean, bn.running_var, conv_bias)
    hit = cache.get("affine")
    if hit is not None and hit[0] == ver:
        return hit[1], hit[2]
    with torch.no_grad():
        var = bn.running_var.double()
        g = bn.weight.double() if bn.weight is not None else torch.ones_like(var)
        b = bn.bias.double() if bn.bias is not None else torch.zeros_like(var)
        s = g / torch.sqrt(var + bn.eps)
        mu = bn.running_mean.double()
        if conv_bias is not None:
            mu = mu - conv_bias.double()
        scale = s.float().contiguous()
        shift = (b - mu * s).float().contiguous()
    cache["affine"] = (ver, scale, shift)
    return scale, shift


# ------------------------------------------------------------------------------------------------
# convolution (+BN +residual +ReLU +upsample-on-write)
# ------------------------------------------------------------------------------------------------

def _conv_raw(x, w_packed, Cout, k, stride, pad, dil, scale, shift, residual, relu, up, out, stats, out_dtype=None,
              code=None):
    N, Cin, H, W, ip = meta(x)
    Ho = (H + 2 * pad - dil * (k - 1) - 1) // stride + 1
    Wo = (W + 2 * pad - dil * (k - 1) - 1) // stride + 1
    out_dtype = out_dtype or (out.dtype if out is not None else act_dtype())
    if out is None:
        out = empty_nhwc(N, Cout, Ho * up, Wo * up, out_dtype, x.device)
    oN, oC, oH, oW, op = meta(out)
    if (oN, oC, oH, oW) != (N, Cout, Ho * up, Wo * up) or out.dtype != out_dtype:
        raise ValueError("conv output buffer has shape %s, expected %s" % (tuple(out.shape), (N, Cout, Ho * up, Wo * up)))
    rp = 0
    if residual is not None:
        rN, rC, rH, rW, rp = meta(residual)
        if (rN, rC, rH, rW) != (N, Cout, Ho * up, Wo * up) or residual.dtype != x.dtype:
            raise ValueError("residual shape %s does not match conv output %s"
                             % (tuple(residual.shape), (N, Cout, Ho * up, Wo * up)))
    d = ConvDesc(N, H, W, Cin, Cout, k, k, stride, pad, dil, Ho, Wo, up, int(bool(relu)), ip, op, rp,
                 code if code is not None else _code(x.dtype), _code(out_dtype), int(stats is not None), 0)
    _lib.call("fami_conv2d_bn_act_fwd", ctypes.byref(d), _ptr(x), _ptr(w_packed), _ptr(scale), _ptr(shift),
              _ptr(residual), _ptr(out), _ptr(stats), _stream())
    return out


def conv_offsets_blocked(x, conv, G, out=None, layout=2):
    """The fused dcn_offset_k | dcn_mask_k convolution (Alignment_V15.py:144-145; `conv` holds the concatenated,
    tap-major-permuted parameters) writing the blocked buffer the deformable kernel reads (om_to_blocked layout 2 or 3;
    dcn_blocked_layout names the one the deformable kernel of a shape takes).  Returns the flat float32 buffer, tagged
    with its layout for dcn_fwd."""
    _need_cuda(x)
    N, Cin, H, W, ip = meta(x)
    k, pad, dil = conv.kernel_size[0], conv.padding[0], conv.dilation[0]
    Cout = conv.out_channels
    if Cout != 27 * G or k != 3 or conv.stride[0] != 1 or pad != dil:
        raise ValueError("conv_offsets_blocked: 3x3 stride-1 same conv with 27*G output channels expected")
    if out is None:
        out = torch.empty(om_blocked_numel(N, H, W, G), dtype=torch.float32, device=x.device)
    code = conv_code(x, Cin)
    w = packed_weight(conv, conv.weight, code)
    _, shift = folded_affine(conv.bias, None)
    d = ConvDesc(N, H, W, Cin, Cout, k, k, 1, pad, dil, H, W, 1, 0, ip, 0, 0, code, F32, 0, G, layout)
    _lib.call("fami_conv2d_bn_act_fwd", ctypes.byref(d), _ptr(x), _ptr(w), None, _ptr(shift), None, _ptr(out), None,
              _stream())
    out._fami_om_layout = layout
    return out


def conv_bn_act(x, conv, bn=None, relu=False, residual=None, up=1, out=None, out_dtype=None, stream_out=True):
    """conv -> [BatchNorm] -> [+residual] -> [ReLU] -> [nearest x`up`] in one launch (eval-mode BN)
    or conv(+stats) -> finalize -> apply (train-mode BN, batch statistics).

    conv: nn.Conv2d used as a parameter container (square kernel 1/3, groups=1); bn: nn.BatchNorm2d.
    Reference chains: basic_model.py:44-63,83-113; basic_layer.py:55-73; hrnet.py:89-172.
    stream_out=False (fp32-residual-stream mode of the 16-bit arms only): the output is never used as a residual or running
    sum (the inner convolutions of a block), so no float twin is written for it.
    """
    _need_cuda(x)
    if conv.groups != 1:
        raise NotImplementedError("grouped convolutions are not on the FAMI-Pose hot path")
    if x.dtype == torch.float32 and torch.is_grad_enabled() and (
            x.requires_grad or conv.weight.requires_grad
            or (bn is not None and bn.weight is not None and bn.weight.requires_grad)
            or (residual is not None and residual.requires_grad)):
        # differentiable path (fp32 storage: 'fp32' / 'tf32' arms): autograd.ConvBnActFunction.  The 16-bit arms are
        # inference arms: their outputs never carry a grad_fn.
        if out is not None:
            raise NotImplementedError("differentiable conv: caller-provided output slices are inference-only")
        from . import autograd as _ag
        if up != 1:
            # fuse-layer term of an un-frozen backbone (hrnet.py:99-112): conv + BN, then upsample + running sum + ReLU as
            # its own differentiable launch (the inference path replicates on write inside the conv's epilogue)
            t = _ag.conv_bn_act(x, conv, bn, False, None)
            return _ag.UpsampleAddReluFunction.apply(t, residual, up, relu)
        return _ag.conv_bn_act(x, conv, bn, relu, residual)
    k = conv.kernel_size[0]
    stride, pad, dil = conv.stride[0], conv.padding[0], conv.dilation[0]
    code = conv_code(x, conv.in_channels)
    wcode = code
    if _PRECISION == "tf32" and x.dtype == torch.float32 and conv.in_channels == 3:
        # the 3-channel stem of the tf32 arm: descriptor dtype TF32 (the library picks its tensor-core stem kernel when the
        # shape fits, the exact-fp32 kernel otherwise); the weights keep the fp32 packing both kernels read
        code, wcode = TF32, F32
    w = packed_weight(conv, conv.weight, wcode)
    Cout = conv.out_channels
    if bn is not None and bn.training:
        # batch statistics: raw conv (+bias) with fused per-channel sum / sum-of-squares
        bias = conv.bias.detach().float() if conv.bias is not None else None
        stats = torch.zeros(2 * Cout, dtype=torch.float64, device=x.device)
        if code == F32:
            raw = _conv_raw(x, w, Cout, k, stride, pad, dil, None, bias, None, False, 1, None, stats, torch.float32)
            N, _, Ho, Wo, rawp = meta(raw)
        else:
            # tensor-core arms: raw conv output kept in fp32, statistics by a separate column reduction
            raw = _conv_raw(x, w, Cout, k, stride, pad, dil, None, bias, None, False, 1, None, None, torch.float32,
                            code=code)
            N, _, Ho, Wo, rawp = meta(raw)
            _lib.call("fami_bn_stats", _ptr(raw), F32, rawp, N * Ho * Wo, Cout, _ptr(stats), _stream())
        scale = torch.empty(Cout, dtype=torch.float32, device=x.device)
        shift = torch.empty_like(scale)
        mom = bn.momentum if bn.momentum is not None else 0.1
        track = bn.track_running_stats and bn.running_mean is not None
        _lib.call("fami_bn_finalize", _ptr(stats), _ptr(bn.weight), _ptr(bn.bias),
                  _ptr(bn.running_mean if track else None), _ptr(bn.running_var if track else None),
                  _ptr(scale), _ptr(shift), None, None, Cout, N * Ho * Wo, float(bn.eps), float(mom), _stream())
        if track:
            # the running statistics were updated through raw pointers: bump their version counters so the
            # folded eval-mode affine cached by folded_affine() is rebuilt on the next eval forward
            torch.autograd.graph.increment_version((bn.running_mean, bn.running_var))
            if bn.num_batches_tracked is not None:
                bn.num_batches_tracked += 1
        if out is None:
            out = empty_nhwc(N, Cout, Ho * up, Wo * up, out_dtype or act_dtype(), x.device)
        op = meta(out)[4]
        rp = meta(residual)[4] if residual is not None else 0
        _lib.call("fami_bn_apply_act", _ptr(raw), _code(raw.dtype), rawp, _ptr(scale), _ptr(shift), _ptr(residual), rp,
                  _ptr(out), op, _code(out.dtype), N, Ho, Wo, Cout, up, int(bool(relu)), _stream())
        return out
    scale, shift = folded_affine(conv.bias, bn)
    if _STREAM_F32 and x.dtype in (torch.float16, torch.bfloat16) and out is None and out_dtype in (None, x.dtype) and (
            stream_out or residual is not None):
        return _conv_stream(x, w, Cout, k, stride, pad, dil, scale, shift, residual, relu, up, code, need32=stream_out)
    return _conv_raw(x, w, Cout, k, stride, pad, dil, scale, shift, residual, relu, up, out, None, out_dtype, code=code)


def _conv_stream(x, w_packed, Cout, k, stride, pad, dil, scale, shift, residual, relu, up, code, need32=True):
    """fp32-residual-stream form of the fused convolution (16-bit arms, fami_conv2d_bn_act_fwd_stream): the residual is
    taken from the float twin of `residual` when it has one (`_fami_f32`, attached to every output of this function) and
    the output gets a float twin of its own.  Slices / views drop the twin and fall back to the 16-bit values."""
    N, Cin, H, W, ip = meta(x)
    Ho = (H + 2 * pad - dil * (k - 1) - 1) // stride + 1
    Wo = (W + 2 * pad - dil * (k - 1) - 1) // stride + 1
    out = empty_nhwc(N, Cout, Ho * up, Wo * up, x.dtype, x.device)
    y32 = empty_nhwc(N, Cout, Ho * up, Wo * up, torch.float32, x.device) if need32 else None
    res32 = getattr(residual, "_fami_f32", None) if residual is not None else None
    if residual is not None and res32 is None:
        res32 = cast_nhwc(residual, torch.float32)        # a residual without a twin (e.g. a slice): widen it
    rp = meta(res32)[4] if res32 is not None else 0
    if res32 is not None and tuple(res32.shape) != (N, Cout, Ho * up, Wo * up):
        raise ValueError("residual shape %s does not match conv output %s" % (tuple(res32.shape), (N, Cout, Ho * up, Wo * up)))
    d = ConvDesc(N, H, W, Cin, Cout, k, k, stride, pad, dil, Ho, Wo, up, int(bool(relu)), ip, meta(out)[4], rp,
                 code, code, 0, 0)
    _lib.call("fami_conv2d_bn_act_fwd_stream", ctypes.byref(d), _ptr(x), _ptr(w_packed), _ptr(scale), _ptr(shift), _ptr(res32),
              _ptr(out), _ptr(y32), meta(y32)[4] if y32 is not None else 0, _stream())
    if y32 is not None:
        out._fami_f32 = y32
    return out


def upsample_nearest(x, factor):
    """F.interpolate(scale_factor=factor, mode='nearest') (basic_model.py:116-125) as the
    replicate-on-write epilogue with an identity affine."""
    _need_cuda(x)
    if factor not in (1, 2, 4, 8):
        raise NotImplementedError("nearest upsample factor must be 1, 2, 4 or 8")
    N, C, H, W, p = meta(x)
    out = empty_nhwc(N, C, H * factor, W * factor, x.dtype, x.device)
    one = torch.ones(C, dtype=torch.float32, device=x.device)
    zero = torch.zeros(C, dtype=torch.float32, device=x.device)
    _lib.call("fami_bn_apply_act", _ptr(x), _code(x.dtype), p, _ptr(one), _ptr(zero), None, 0, _ptr(out), C,
              _code(x.dtype), N, H, W, C, factor, 0, _stream())
    return out


# ------------------------------------------------------------------------------------------------
# deformable convolution
# ------------------------------------------------------------------------------------------------

def dcn_fused_supported(C, G, dtype):
    """True if the tensor-core DCN kernel (fused tap-major offsets) takes this shape: 16-bit activations, or the fp32
    activations of the 'tf32' arm (cast to fp16 -- the same 11-bit significand as TF32 -- on the way in, fp32 out)."""
    half = dtype in (torch.float16, torch.bfloat16) or (dtype == torch.float32 and _PRECISION == "tf32")
    return half and C % 16 == 0 and (C <= 64 or C % 64 == 0) and C % G == 0 and C // G == 4


def cast_nhwc(x, dtype, out=None):
    """Elementwise storage cast of an NHWC(+pitch) activation (fami_bn_apply_act with an identity affine)."""
    _need_cuda(x)
    N, C, H, W, p = meta(x)
    if out is None:
        out = empty_nhwc(N, C, H, W, dtype, x.device)
    op = meta(out)[4]
    if x.dtype == torch.float32 and dtype in (torch.float16, torch.bfloat16) and C % 8 == 0 and p % 4 == 0 and op % 8 == 0:
        # null scale / shift: the library's plain-cast kernel (8 channels per thread)
        _lib.call("fami_bn_apply_act", _ptr(x), _code(x.dtype), p, None, None, None, 0, _ptr(out), op, _code(dtype),
                  N, H, W, C, 1, 0, _stream())
        return out
    key = (C, x.device)
    ident = _IDENT.get(key)
    if ident is None:
        ident = _IDENT[key] = (torch.ones(C, dtype=torch.float32, device=x.device),
                               torch.zeros(C, dtype=torch.float32, device=x.device))
    _lib.call("fami_bn_apply_act", _ptr(x), _code(x.dtype), p, _ptr(ident[0]), _ptr(ident[1]), None, 0, _ptr(out),
              op, _code(dtype), N, H, W, C, 1, 0, _stream())
    return out


_IDENT = {}


def tap_major_perm(G, k=3):
    """Channel permutation that turns the concatenated [offset(18G) | mask(9G)] torchvision layout into
    the fused tap-major layout [tap][dy(G) | dx(G) | mask(G)]: new[n] = old[perm[n]]."""
    K = k * k
    perm = []
    for t in range(K):
        perm += [g * 2 * K + 2 * t for g in range(G)]
        perm += [g * 2 * K + 2 * t + 1 for g in range(G)]
        perm += [2 * K * G + g * K + t for g in range(G)]
    return perm


DCN_TILE_H, DCN_TILE_W = 16, 8      # pixel tile of the blocked offset|mask layouts (fami_dcn_desc.om_layout 2 / 3)


def dcn_blocked_layout(C, Cout, G):
    """The blocked offset|mask layout the deformable kernel of this shape reads: 3 (k-step-blocked) for the warp-private
    kernel (csrc/dcn_wp.cu: C == Cout in {32, 48}, 4 channels per offset group), else 2 (row-blocked, csrc/dcn_tc.cu)."""
    import os
    wp = os.environ.get("FAMI_DCN_WP", "1") != "0"
    return 3 if wp and C == Cout and C in (32, 48) and 4 * G == C else 2


def om_blocked_numel(B, H, W, G, k=3):
    """Elements of the row-blocked offset|mask buffer (fami_dcn_desc.om_layout = 2)."""
    ty, tx = (H + DCN_TILE_H - 1) // DCN_TILE_H, (W + DCN_TILE_W - 1) // DCN_TILE_W
    return k * k * B * ty * tx * DCN_TILE_H * DCN_TILE_W * 3 * G


def om_to_blocked(om, G, k=3, layout=2):
    """layout 3 (k-step-blocked, the warp-private kernel's): [image tile][tap][row 16][dy | dx | mask][group / 4][pixel 8][group % 4]
    -- tile-major, the nine taps of a tile contiguous.  layout 2:
    tap-major NHWC [B, 27G, H, W] (tap_major_perm order) -> row-blocked layout
    [tap][image tile][row 16][dy | dx | mask][pixel 8][group G] over 16x8-pixel tiles: a gather warp of the deformable kernel
    owns one tile row and walks its 8*G (pixel, group) samples of a tap 32 at a time -- every load instruction of the warp
    reads 128 contiguous bytes.  Host-side converter for tests and tools; in the model the producer convolution writes this
    layout directly (fami_conv_desc.om_groups)."""
    B, FC, H, W = om.shape
    K = k * k
    ty, tx = (H + DCN_TILE_H - 1) // DCN_TILE_H, (W + DCN_TILE_W - 1) // DCN_TILE_W
    t = om.permute(0, 2, 3, 1).reshape(B, H, W, K, 3, G).float()
    if ty * DCN_TILE_H != H or tx * DCN_TILE_W != W:
        pad = torch.zeros((B, ty * DCN_TILE_H, tx * DCN_TILE_W, K, 3, G), dtype=torch.float32, device=om.device)
        pad[:, :H, :W] = t
        t = pad
    if layout == 3:
        t = t.reshape(B, ty, DCN_TILE_H, tx, DCN_TILE_W, K, 3, G // 4, 4).permute(0, 1, 3, 5, 2, 6, 7, 4, 8)   # B,ty,tx,K,row,k,G/4,pix,4
    else:
        t = t.reshape(B, ty, DCN_TILE_H, tx, DCN_TILE_W, K, 3, G).permute(5, 0, 1, 3, 2, 6, 4, 7)   # K,B,ty,tx,row,k,pix,G
    out = t.contiguous().reshape(-1)
    out._fami_om_layout = layout
    return out


def dcn_fwd(x, offset, mask, weight, bias, owner, pad=3, dil=3, out=None, fused_om=None, blocked_om=None, groups=None):
    """torchvision.ops.deform_conv2d(x, offset, weight, bias, stride=1, padding=pad, dilation=dil, mask=mask)
    on NHWC operands (Alignment_V15.py:146,150,154,158).  fused_om: ONE float32 buffer [B, 27G, H, W]
    in tap-major layout (see tap_major_perm) instead of (offset, mask) -- the 16-bit tensor-core kernel."""
    _need_cuda(x, offset, mask, fused_om)
    B, C, H, W, xp = meta(x)
    Cout, Cin, kh, kw = weight.shape
    out_f32 = 0
    if x.dtype == torch.float32 and (blocked_om is not None or fused_om is not None):
        # tf32 arm: the gather and the contraction run on fp16 multiplicands (11-bit significand, as TF32), fp32 out
        x = cast_nhwc(x, torch.float16)
        xp = meta(x)[4]
        out_f32 = 1
    if out is None:
        out = empty_nhwc(B, Cout, H, W, torch.float32 if out_f32 else x.dtype, x.device)
    outp = meta(out)[4]
    b = bias.detach().float() if bias is not None else None
    if blocked_om is not None:
        G = groups
        if Cin != C or not dcn_fused_supported(C, G, x.dtype) or blocked_om.dtype != torch.float32 \
                or blocked_om.numel() != om_blocked_numel(B, H, W, G, kh):
            raise ValueError("row-blocked DCN offsets need 16-bit x, C <= 64 or a multiple of 64, 4 channels per offset group and a "
                             "float32 buffer of om_blocked_numel elements")
        w = packed_weight(owner, weight, x.dtype)
        d = DcnDesc(B, H, W, C, Cout, G, kh, kw, 1, pad, dil, xp, 0, 0, outp, getattr(blocked_om, "_fami_om_layout", 2),
                    _code(x.dtype), out_f32)
        _lib.call("fami_dcn_fwd", ctypes.byref(d), _ptr(x), _ptr(blocked_om), None, _ptr(w), _ptr(b), _ptr(out), _stream())
        return out
    if fused_om is not None:
        fB, FC, fH, fW, fp_ = meta(fused_om)
        if FC % (3 * kh * kw) != 0 or (fB, fH, fW) != (B, H, W) or fused_om.dtype != torch.float32:
            raise RuntimeError("fused offset|mask buffer %s inconsistent with input %s" % (tuple(fused_om.shape), tuple(x.shape)))
        G = FC // (3 * kh * kw)
        if Cin != C or not dcn_fused_supported(C, G, x.dtype):
            raise ValueError("fused tap-major DCN needs 16-bit x, C <= 64 or a multiple of 64, and 4 channels per offset group")
        w = packed_weight(owner, weight, x.dtype)      # UMMA B operand: [CoutPad][9][64] half
        d = DcnDesc(B, H, W, C, Cout, G, kh, kw, 1, pad, dil, xp, fp_, 0, outp, 1, _code(x.dtype), out_f32)
        _lib.call("fami_dcn_fwd", ctypes.byref(d), _ptr(x), _ptr(fused_om), None, _ptr(w), _ptr(b), _ptr(out), _stream())
        return out
    oB, OC, oH, oW, offp = meta(offset)
    if OC % (2 * kh * kw) != 0:
        # torchvision raises RuntimeError for a bad offset channel count (deform_conv.py:85-90)
        raise RuntimeError("offset channels %d not a multiple of 2*kh*kw=%d" % (OC, 2 * kh * kw))
    G = OC // (2 * kh * kw)
    if C % G != 0 or Cin != C:
        raise ValueError("in_channels %d must be divisible by offset groups %d and match the weight (%d)" % (C, G, Cin))
    mB, MC, mH, mW, mp = meta(mask)
    if offset.dtype != torch.float32 or mask.dtype != torch.float32:
        raise TypeError("deformable offsets and masks are float32 in both precisions")
    if MC != G * kh * kw or (oB, oH, oW) != (B, H, W) or (mB, mH, mW) != (B, H, W):
        raise RuntimeError("offset/mask shapes %s %s inconsistent with input %s"
                           % (tuple(offset.shape), tuple(mask.shape), tuple(x.shape)))
    w = packed_weight(owner, weight, torch.float32)   # SIMT contraction: fp32 [K][CoutPad] packing
    d = DcnDesc(B, H, W, C, Cout, G, kh, kw, 1, pad, dil, xp, offp, mp, outp, 0, _code(x.dtype))
    _lib.call("fami_dcn_fwd", ctypes.byref(d), _ptr(x), _ptr(offset), _ptr(mask), _ptr(w), _ptr(b), _ptr(out),
              _stream())
    return out


# ------------------------------------------------------------------------------------------------
# warp / small ops
# ------------------------------------------------------------------------------------------------

def warp_translate(src, txy, out=None):
    """kornia.geometry.warp_affine(src, [[1,0,tx],[0,1,ty]], dsize=(H,W)) (Alignment_V15.py:133-135)."""
    _need_cuda(src, txy)
    B, C, H, W, sp = meta(src)
    txy = txy.detach().float().contiguous()
    if tuple(txy.shape) != (B, 2):
        raise ValueError("txy must be [B,2]")
    if out is None:
        out = empty_nhwc(B, C, H, W, src.dtype, src.device)
    op = meta(out)[4]
    _lib.call("fami_warp_translate_fwd", _ptr(src), sp, _ptr(txy), _ptr(out), op, _code(src.dtype), B, H, W, C,
              _stream())
    return out


def sub_bcast(a, b, rep):
    """out[r] = a[r] - b for r < rep, a = rep stacked blocks shaped like b (dense NHWC)."""
    _need_cuda(a, b)
    Na, C, H, W, pa = meta(a)
    Nb, Cb, Hb, Wb, pb = meta(b)
    if pa != C or pb != Cb or (Nb * rep, Cb, Hb, Wb) != (Na, C, H, W):
        raise ValueError("sub_bcast needs dense NHWC operands with a = rep x b")
    out = empty_nhwc(Na, C, H, W, a.dtype, a.device)
    _lib.call("fami_sub_bcast", _ptr(a), _ptr(b), _ptr(out), _code(a.dtype), Nb * C * H * W, rep, _stream())
    return out


def copy_into(src, dst):
    """dst[:] = src for NHWC operands with independent pitches (one operand of a channel concat)."""
    _need_cuda(src, dst)
    N, C, H, W, sp = meta(src)
    dN, dC, dH, dW, dp = meta(dst)
    if (N, C, H, W) != (dN, dC, dH, dW) or src.dtype != dst.dtype:
        raise ValueError("copy_into shape/dtype mismatch")
    _lib.call("fami_copy2d", _ptr(src), sp, _ptr(dst), dp, _code(src.dtype), N * H * W, C, _stream())
    return dst


def flatten_nchw_order(x):
    """nn.Flatten() of a (small) NHWC activation in the reference's C,H,W order -> float32 [N, C*H*W]."""
    return to_nchw(x).reshape(x.shape[0], -1)


def linear(x, weight, bias):
    """nn.Linear forward on float32 [M,K] (Alignment_V15.py:69-71)."""
    _need_cuda(x)
    x = x.contiguous()
    M, K = x.shape
    N = weight.shape[0]
    y = torch.empty((M, N), dtype=torch.float32, device=x.device)
    w = weight.detach().float().contiguous()
    b = bias.detach().float().contiguous() if bias is not None else None
    _lib.call("fami_linear_fwd", _ptr(x), _ptr(w), _ptr(b), _ptr(y), M, K, N, _stream())
    return y


# ------------------------------------------------------------------------------------------------
# losses / decode
# ------------------------------------------------------------------------------------------------

def joint_mse(pred, target, target_weight, want_grad=False, grad_scale=1.0):
    """JointMSELoss (mse_loss.py:21-40).  pred: NHWC activation or NCHW fp32; target NCHW fp32."""
    _need_cuda(pred, target)
    if not is_nhwc(pred):
        pred = to_nhwc(pred.float(), torch.float32)
    B, J, H, W, pp = meta(pred)
    target = target.float().contiguous()
    w = target_weight.float().reshape(B, J).contiguous() if target_weight is not None else None
    loss = torch.zeros((), dtype=torch.float32, device=pred.device)
    grad = torch.empty((B, H, W, J), dtype=torch.float32, device=pred.device) if want_grad else None
    _lib.call("fami_joint_mse_fwd_bwd", _ptr(pred), _code(pred.dtype), pp, _ptr(target), _ptr(w), _ptr(loss),
              _ptr(grad), float(grad_scale), B, J, H, W, _stream())
    return (loss, grad) if want_grad else loss


def softmax_pkl(a, b, temperature=0.05):
    """kl_div(input=softmax(a/T), target=softmax(b/T), 'mean') with the reference's quirk
    (Alignment_V15.py:250-277); a, b NHWC activations [B,C,H,W]."""
    _need_cuda(a, b)
    B, C, H, W, ap = meta(a)
    B2, C2, H2, W2, bp = meta(b)
    if (B, C, H, W) != (B2, C2, H2, W2) or a.dtype != b.dtype:
        raise ValueError("softmax_pkl operand mismatch")
    out = torch.zeros((), dtype=torch.float32, device=a.device)
    _lib.call("fami_softmax_pkl_fwd", _ptr(a), ap, _ptr(b), bp, _code(a.dtype), _ptr(out), B, H * W, C,
              float(temperature), _stream())
    return out


def argmax_hw(hm):
    """get_max_preds core (heatmaps_process.py:29-30): flat argmax + max per (b,j).  hm NHWC or NCHW."""
    _need_cuda(hm)
    if not is_nhwc(hm):
        hm = to_nhwc(hm.float(), torch.float32)
    B, J, H, W, p = meta(hm)
    idx = torch.empty((B, J), dtype=torch.int32, device=hm.device)
    mx = torch.empty((B, J), dtype=torch.float32, device=hm.device)
    _lib.call("fami_argmax_hw", _ptr(hm), _code(hm.dtype), p, _ptr(idx), _ptr(mx), B, H * W, J, _stream())
    return idx, mx


# ------------------------------------------------------------------------------------------------
# backward of the dense pieces (fp32 arm) -- autograd of nn.Conv2d / nn.BatchNorm2d / nn.Linear and
# of the MI estimator as the reference's training step derives them (alignment_mi_function_term6_1.py:104-156)
# ------------------------------------------------------------------------------------------------

def _conv_desc_f32(x_shape_meta, Cout, k, stride, pad, dil, out_pitch):
    N, Cin, H, W, ip = x_shape_meta
    Ho = (H + 2 * pad - dil * (k - 1) - 1) // stride + 1
    Wo = (W + 2 * pad - dil * (k - 1) - 1) // stride + 1
    return ConvDesc(N, H, W, Cin, Cout, k, k, stride, pad, dil, Ho, Wo, 1, 0, ip, out_pitch, 0, F32, F32, 0, 0), Ho, Wo


def conv_dgrad(grad_y, weight, x_shape, stride=1, pad=0, dil=1, out=None):
    """d loss / d x of y = conv2d(x, weight): grad_y NHWC fp32 [N,Cout,Ho,Wo], weight OIHW; returns NHWC
    fp32 [N,Cin,H,W] (x_shape = logical NCHW shape of x).  On the 'tf32' arm stride-1 dgrad runs on the tensor cores
    (the forward kernels over the flipped / transposed filter)."""
    _need_cuda(grad_y)
    if grad_y.dtype != torch.float32 or not is_nhwc(grad_y):
        raise ValueError("conv_dgrad: grad_y must be an fp32 NHWC activation")
    N, Cin, H, W = x_shape
    Cout, k = weight.shape[0], weight.shape[2]
    gN, gC, gH, gW, gp = meta(grad_y)
    if out is None:
        out = empty_nhwc(N, Cin, H, W, torch.float32, grad_y.device)
    d, Ho, Wo = _conv_desc_f32((N, Cin, H, W, meta(out)[4]), Cout, k, stride, pad, dil, gp)
    if (gN, gC, gH, gW) != (N, Cout, Ho, Wo):
        raise ValueError("conv_dgrad: grad_y shape %s, expected %s" % (tuple(grad_y.shape), (N, Cout, Ho, Wo)))
    code = conv_code(grad_y, Cout) if stride == 1 else F32
    d.dtype = code
    w = weight.detach().float().contiguous()
    scratch = torch.empty_like(w)
    n = _lib.load().fami_packed_weight_elems(Cin, Cout, k, k, code)
    wt = torch.empty(n, dtype=torch.float32, device=grad_y.device)
    _lib.call("fami_pack_conv_weight_dgrad", _ptr(w), _ptr(scratch), _ptr(wt), Cout, Cin, k, k, code, _stream())
    _lib.call("fami_conv2d_dgrad", ctypes.byref(d), _ptr(grad_y), _ptr(wt), _ptr(out), _stream())
    return out


def conv_wgrad(x, grad_y, weight_shape, stride=1, pad=0, dil=1, want_bias=False):
    """(d loss / d weight [OIHW], d loss / d bias or None) of y = conv2d(x, weight) + bias.  Exact fp32 FMAs on the 'fp32'
    arm, TF32 tensor cores with fp32 accumulation on the 'tf32' arm."""
    _need_cuda(x, grad_y)
    if x.dtype != torch.float32 or grad_y.dtype != torch.float32:
        raise ValueError("conv_wgrad: fp32 activations only")
    Cout, Cin, k, _ = weight_shape
    d, Ho, Wo = _conv_desc_f32(meta(x), Cout, k, stride, pad, dil, meta(grad_y)[4])
    if tuple(grad_y.shape) != (x.shape[0], Cout, Ho, Wo) or x.shape[1] != Cin:
        raise ValueError("conv_wgrad: shape mismatch")
    if _PRECISION == "tf32":
        d.dtype = TF32          # 'tf32' arm: the products on mma.sync TF32 (operands rounded to nearest), fp32 accumulation
    gw = torch.zeros(tuple(weight_shape), dtype=torch.float32, device=x.device)
    gb = torch.zeros(Cout, dtype=torch.float32, device=x.device) if want_bias else None
    _lib.call("fami_conv2d_wgrad", ctypes.byref(d), _ptr(x), _ptr(grad_y), _ptr(gw), _ptr(gb), _stream())
    return gw, gb


def bn_bwd(x, grad_y, mean, invstd, gamma, y=None, training=True, want_res=False):
    """BatchNorm2d backward (+ ReLU mask from the post-activation output y, + the residual's gradient).
    x: the raw conv output the BN normalised (fp32 NHWC).  Returns (grad_x, grad_gamma, grad_beta, grad_res)."""
    _need_cuda(x, grad_y)
    N, C, H, W, xp = meta(x)
    gp = meta(grad_y)[4]
    yp = meta(y)[4] if y is not None else 0
    gx = empty_nhwc(N, C, H, W, torch.float32, x.device)
    gres = empty_nhwc(N, C, H, W, torch.float32, x.device) if want_res else None
    dgamma = torch.empty(C, dtype=torch.float32, device=x.device)
    dbeta = torch.empty(C, dtype=torch.float32, device=x.device)
    sums = torch.zeros(2 * C, dtype=torch.float64, device=x.device)
    _lib.call("fami_bn_bwd", _ptr(x), xp, _ptr(grad_y), gp, _ptr(y), yp, _ptr(mean), _ptr(invstd), _ptr(gamma),
              N * H * W, C, int(bool(training)), _ptr(sums), _ptr(gx), meta(gx)[4], _ptr(gres),
              meta(gres)[4] if gres is not None else 0, _ptr(dgamma), _ptr(dbeta), _stream())
    return gx, dgamma, dbeta, gres


def softmax_pkl_bwd(a, b, grad_out, temperature=0.05, want_a=True, want_b=True):
    """Gradients of softmax_pkl(a, b) with respect to a and b (fp32 NHWC), scaled by the device scalar grad_out."""
    _need_cuda(a, b)
    B, C, H, W, ap = meta(a)
    bp = meta(b)[4]
    if a.dtype != torch.float32 or b.dtype != torch.float32:
        raise ValueError("softmax_pkl_bwd: fp32 activations only")
    ga = empty_nhwc(B, C, H, W, torch.float32, a.device) if want_a else None
    gb = empty_nhwc(B, C, H, W, torch.float32, a.device) if want_b else None
    go = grad_out.detach().float().reshape(1).contiguous()
    _lib.call("fami_softmax_pkl_bwd", _ptr(a), ap, _ptr(b), bp, _ptr(go), _ptr(ga), meta(ga)[4] if want_a else 0,
              _ptr(gb), meta(gb)[4] if want_b else 0, B, H * W, C, float(temperature), _stream())
    return ga, gb


def linear_bwd(x, weight, grad_y, want_x=True):
    """(grad_x, grad_w, grad_b) of y = x W^T + b on float32 [M,K]."""
    _need_cuda(x, grad_y)
    x = x.contiguous().float()
    grad_y = grad_y.contiguous().float()
    M, K = x.shape
    N = weight.shape[0]
    w = weight.detach().float().contiguous()
    gx = torch.empty((M, K), dtype=torch.float32, device=x.device) if want_x else None
    gw = torch.empty((N, K), dtype=torch.float32, device=x.device)
    gb = torch.empty(N, dtype=torch.float32, device=x.device)
    _lib.call("fami_linear_bwd", _ptr(x), _ptr(w), _ptr(grad_y), _ptr(gx), _ptr(gw), _ptr(gb), M, K, N, _stream())
    return gx, gw, gb
