"""Drop-in replacements for posetimation/layers (basic_model.py, basic_layer.py) and for
torchvision.ops.DeformConv2d as the reference uses it.

Same class names, constructor signatures, sub-module attribute names (=> identical state_dict keys,
SURVEY.md section 5 "Checkpoint") and forward semantics; the arithmetic runs in libfami_b200.so.
nn.Conv2d / nn.BatchNorm2d objects are kept as PARAMETER CONTAINERS only -- their own forward is
never called.  Inputs may be the reference's NCHW float32 tensors or channels-last activations
produced by other fami modules; outputs are channels-last activations with logical NCHW shape.
"""
import math

import torch
import torch.nn as nn

from . import ops

BN_MOMENTUM = 0.1


def conv3x3(in_planes, out_planes, stride=1, groups=1):
    """3x3 convolution with padding (parameter container; basic_model.py:20-22)."""
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False, groups=groups)


def _only_relu(act):
    if act != 'ReLU':
        raise NotImplementedError("fami_pose_b200 implements the ReLU activation the FAMI-Pose models use; got %r" % act)


class BasicBlock(nn.Module):
    """posetimation/layers/basic_model.py:25-63."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, groups=1, skip_norm=False, act='ReLU'):
        super().__init__()
        assert act in ['ReLU', 'LeakyReLU'], "Not Expectation act function {}".format(act)
        _only_relu(act)
        if groups != 1:
            raise NotImplementedError("grouped BasicBlock is not on the FAMI-Pose hot path")
        self.conv1 = conv3x3(inplanes, planes, stride, groups=groups)
        self.bn1 = nn.BatchNorm2d(planes, momentum=BN_MOMENTUM)
        self.act_fun = nn.ReLU(inplace=True)
        self.conv2 = conv3x3(planes, planes, stride, groups=groups)
        self.bn2 = nn.BatchNorm2d(planes, momentum=BN_MOMENTUM)
        self.downsample = downsample
        self.stride = stride
        self.skip_norm = skip_norm

    def forward(self, x, out=None):
        x = ops.to_nhwc(x)
        bn1 = None if self.skip_norm else self.bn1
        bn2 = None if self.skip_norm else self.bn2
        h = ops.conv_bn_act(x, self.conv1, bn1, relu=True, stream_out=False)
        if self.downsample is not None:
            ds = list(self.downsample.children())
            res = ops.conv_bn_act(x, ds[0], ds[1] if len(ds) > 1 else None, relu=False)
        else:
            res = x
        return ops.conv_bn_act(h, self.conv2, bn2, relu=True, residual=res, out=out)


class Bottleneck(nn.Module):
    """posetimation/layers/basic_model.py:66-113."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, dilation=1):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes, momentum=BN_MOMENTUM)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes, momentum=BN_MOMENTUM)
        self.conv3 = nn.Conv2d(planes, planes * self.expansion, kernel_size=1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * self.expansion, momentum=BN_MOMENTUM)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        x = ops.to_nhwc(x)
        h = ops.conv_bn_act(x, self.conv1, self.bn1, relu=True, stream_out=False)
        h = ops.conv_bn_act(h, self.conv2, self.bn2, relu=True, stream_out=False)
        if self.downsample is not None:
            ds = list(self.downsample.children())
            res = ops.conv_bn_act(x, ds[0], ds[1] if len(ds) > 1 else None, relu=False)
        else:
            res = x
        return ops.conv_bn_act(h, self.conv3, self.bn3, relu=True, residual=res)


class Interpolate(nn.Module):
    """posetimation/layers/basic_model.py:116-125 (nearest, power-of-two factor).  Inside
    HighResolutionModule the upsample is fused into the preceding 1x1 conv's store
    (ops.conv_bn_act(up=...)); standalone it runs the replicate-on-write kernel."""

    def __init__(self, scale_factor, mode='nearest'):
        super().__init__()
        if mode != 'nearest':
            raise NotImplementedError("only nearest-neighbour interpolation is used by FAMI-Pose")
        self.scale_factor = scale_factor
        self.mode = mode

    def forward(self, x):
        return ops.upsample_nearest(ops.to_nhwc(x), int(self.scale_factor))


class ChainOfBasicBlocks(nn.Module):
    """posetimation/layers/basic_model.py:128-148 (argument name 'ouput_channel' [sic] preserved)."""

    def __init__(self, input_channel, ouput_channel, kernel_height=None, kernel_width=None, dilation=None,
                 num_blocks=1, groups=1, skip_norm=False, act='ReLU'):
        super().__init__()
        stride = 1
        if skip_norm:
            downsample = nn.Sequential(
                nn.Conv2d(input_channel, ouput_channel, kernel_size=1, stride=stride, bias=False, groups=groups))
        else:
            downsample = nn.Sequential(
                nn.Conv2d(input_channel, ouput_channel, kernel_size=1, stride=stride, bias=False, groups=groups),
                nn.BatchNorm2d(ouput_channel, momentum=BN_MOMENTUM))
        layers = [BasicBlock(input_channel, ouput_channel, stride, downsample, groups, skip_norm=skip_norm, act=act)]
        for _ in range(1, num_blocks):
            layers.append(BasicBlock(ouput_channel, ouput_channel, stride, downsample=None, groups=groups,
                                     skip_norm=skip_norm, act=act))
        self.layers = nn.Sequential(*layers)

    def forward(self, input, out=None):
        x = input
        n = len(self.layers)
        for i, blk in enumerate(self.layers):
            x = blk(x, out=out if i == n - 1 else None)
        return x


class conv_bn_relu(nn.Module):
    """posetimation/layers/basic_layer.py:13-73."""

    def __init__(self, in_planes, out_planes, kernel_size, stride, padding, dilation,
                 has_bias=True, has_bn=True, has_relu=True, efficient=False, groups=1, act='ReLU'):
        super().__init__()
        assert act in ['ReLU', 'LeakyReLU', 'SiLU'], "Not Expectation act function {}".format(act)
        if has_relu:
            _only_relu(act)
        self.conv = nn.Conv2d(in_planes, out_planes, kernel_size=kernel_size, stride=stride, padding=padding,
                              dilation=dilation, groups=groups, bias=has_bias)
        self.has_bn = has_bn
        self.has_relu = has_relu
        self.efficient = efficient  # activation checkpointing flag of the reference: no effect on values
        self.bn = nn.BatchNorm2d(out_planes, momentum=BN_MOMENTUM) if has_bn else None
        self.relu = nn.ReLU(inplace=True) if has_relu else None

    def forward(self, x, out=None):
        x = ops.to_nhwc(x)
        return ops.conv_bn_act(x, self.conv, self.bn if self.has_bn else None, relu=self.has_relu, out=out)


class DeformConv2d(nn.Module):
    """torchvision.ops.DeformConv2d as constructed at Alignment_V15.py:83,89,95,101
    (kernel 3, stride 1, padding = dilation, weight groups 1).  Same parameters ('weight' [Cout,Cin,3,3],
    'bias'), same default init (kaiming_uniform(a=sqrt(5)) / uniform bias, torchvision deform_conv.py)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True):
        super().__init__()
        if in_channels % groups != 0:
            raise ValueError("in_channels must be divisible by groups")
        if out_channels % groups != 0:
            raise ValueError("out_channels must be divisible by groups")
        ks = kernel_size if isinstance(kernel_size, (tuple, list)) else (kernel_size, kernel_size)
        one = lambda v: v[0] if isinstance(v, (tuple, list)) else v
        if tuple(ks) != (3, 3) or one(stride) != 1 or groups != 1 or one(padding) != one(dilation):
            raise NotImplementedError("fami DeformConv2d supports the reference configuration: 3x3, stride 1, "
                                      "padding == dilation, weight groups 1")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride = tuple(ks), (1, 1)
        self.padding, self.dilation, self.groups = (one(padding),) * 2, (one(dilation),) * 2, groups
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels // groups, 3, 3))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in = self.weight.shape[1] * 9
            bound = 1 / math.sqrt(fan_in)
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, input, offset, mask=None, out=None, fused_om=None, blocked_om=None, groups=None):
        x = ops.to_nhwc(input)
        if blocked_om is not None:
            # fami extension: row-blocked [offset|mask] buffer written by the fused producer conv (ops.om_to_blocked)
            return ops.dcn_fwd(x, None, None, self.weight, self.bias, self, pad=self.padding[0], dil=self.dilation[0],
                               out=out, blocked_om=blocked_om, groups=groups)
        if fused_om is not None:
            # fami extension: one tap-major [offset|mask] buffer from the fused producer conv (ops.tap_major_perm)
            return ops.dcn_fwd(x, None, None, self.weight, self.bias, self, pad=self.padding[0], dil=self.dilation[0],
                               out=out, fused_om=fused_om)
        off = ops.to_nhwc(offset, torch.float32)
        if mask is None:
            raise NotImplementedError("FAMI-Pose always passes a modulation mask (DCNv2)")
        msk = ops.to_nhwc(mask, torch.float32)
        if torch.is_grad_enabled() and out is None and any(t.requires_grad for t in (x, off, msk, self.weight)):
            from .autograd import DeformConvFunction   # differentiable path (fp32 arm): fami_dcn_bwd
            return DeformConvFunction.apply(x, off, msk, self.weight, self.bias, self, self.padding[0], self.dilation[0])
        return ops.dcn_fwd(x, off, msk, self.weight, self.bias, self, pad=self.padding[0], dil=self.dilation[0],
                           out=out)
