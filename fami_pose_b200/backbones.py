"""Drop-in replacements for posetimation/backbones/hrnet.py: HighResolutionModule, HRNetPlus, HRNet.

The module tree (attribute names, Sequential/ModuleList indices) reproduces the reference's so the
state_dict keys are identical ('stage3.1.branches.2.3.conv2.weight', 'transition1.1.0.0.weight',
'stage2.0.fuse_layers.0.1.0.weight', ...), but forward() drives fused launches:
every fuse-layer term is a convolution whose epilogue adds the running sum, applies the final ReLU
and replicates on write for the nearest upsample (hrnet.py:89-146,151-172).
"""
import torch
import torch.nn as nn

from . import ops
from .layers import BasicBlock, Bottleneck, Interpolate

BN_MOMENTUM = 0.1

blocks_dict = {'BASIC': BasicBlock, 'BOTTLENECK': Bottleneck}


def _conv_bn(seq):
    mods = list(seq.children())
    return mods[0], mods[1]


class HighResolutionModule(nn.Module):
    """hrnet.py:17-172."""

    def __init__(self, num_branches, blocks, num_blocks, num_inchannels, num_channels, fuse_method,
                 multi_scale_output=True, name=None):
        super().__init__()
        if num_branches != len(num_blocks):
            raise ValueError('NUM_BRANCHES({}) <> NUM_BLOCKS({})'.format(num_branches, len(num_blocks)))
        self.name = name
        self.num_inchannels = num_inchannels
        self.fuse_method = fuse_method
        self.num_branches = num_branches
        self.multi_scale_output = multi_scale_output
        self.branches = nn.ModuleList(
            [self._make_one_branch(i, blocks, num_blocks, num_channels) for i in range(num_branches)])
        self.fuse_layers = self._make_fuse_layers()
        self.relu = nn.ReLU(True)

    def _make_one_branch(self, branch_index, block, num_blocks, num_channels, stride=1):
        downsample = None
        cin, cout = self.num_inchannels[branch_index], num_channels[branch_index] * block.expansion
        if stride != 1 or cin != cout:
            downsample = nn.Sequential(nn.Conv2d(cin, cout, kernel_size=1, stride=stride, bias=False),
                                       nn.BatchNorm2d(cout, momentum=BN_MOMENTUM))
        layers = [block(cin, num_channels[branch_index], stride, downsample)]
        self.num_inchannels[branch_index] = cout
        for _ in range(1, num_blocks[branch_index]):
            layers.append(block(cout, num_channels[branch_index]))
        return nn.Sequential(*layers)

    def _make_fuse_layers(self):
        if self.num_branches == 1:
            return None
        nb, ch = self.num_branches, self.num_inchannels
        rows = []
        for i in range(nb if self.multi_scale_output else 1):
            row = []
            for j in range(nb):
                if j > i:
                    row.append(nn.Sequential(nn.Conv2d(ch[j], ch[i], 1, 1, 0, bias=False), nn.BatchNorm2d(ch[i]),
                                             Interpolate(scale_factor=2 ** (j - i), mode='nearest')))
                elif j == i:
                    row.append(None)
                else:
                    steps = []
                    for k in range(i - j):
                        last = k == i - j - 1
                        co = ch[i] if last else ch[j]
                        mods = [nn.Conv2d(ch[j], co, 3, 2, 1, bias=False), nn.BatchNorm2d(co)]
                        if not last:
                            mods.append(nn.ReLU(True))
                        steps.append(nn.Sequential(*mods))
                    row.append(nn.Sequential(*steps))
            rows.append(nn.ModuleList(row))
        return nn.ModuleList(rows)

    def get_num_inchannels(self):
        return self.num_inchannels

    def forward(self, x: list):
        if self.num_branches == 1:
            return [self.branches[0](x[0])]
        nb = self.num_branches
        for i in range(nb):
            x[i] = self.branches[i](x[i])
        x_fuse = []
        for i in range(len(self.fuse_layers)):
            # y = x[i] + sum_{j != i} fuse_ij(x[j]); ReLU.  Each term is one launch that adds the running
            # sum in its epilogue; the last one applies the ReLU.
            cur = x[i]
            terms = [j for j in range(nb) if j != i]
            for t, j in enumerate(terms):
                final = t == len(terms) - 1
                if j > i:
                    conv, bn = _conv_bn(self.fuse_layers[i][j])
                    cur = ops.conv_bn_act(x[j], conv, bn, relu=final, residual=cur, up=2 ** (j - i))
                else:
                    h = x[j]
                    steps = list(self.fuse_layers[i][j].children())
                    for k, step in enumerate(steps):
                        conv, bn = _conv_bn(step)
                        if k == len(steps) - 1:
                            h = ops.conv_bn_act(h, conv, bn, relu=final, residual=cur)
                        else:
                            h = ops.conv_bn_act(h, conv, bn, relu=True)
                    cur = h
            x_fuse.append(cur)
        if self.name == 'stage4_module3':
            x_fuse.extend(x[1:])
        return x_fuse


class _HRNetTrunk(nn.Module):
    """Shared constructor/trunk of HRNet (hrnet.py:233-297) and HRNetPlus (hrnet.py:569-629)."""

    def _build(self, cfg, is_train, use_deconv, use_prediction, kwargs):
        extra = cfg.MODEL.EXTRA
        self.pretrained = cfg.MODEL.BACKBONE_PRETRAINED
        self.backbone_pretrained = cfg.MODEL.BACKBONE_PRETRAINED
        self.is_train = is_train
        self.inplanes = 64
        self.use_deconv = use_deconv
        self.use_prediction = use_prediction
        self.vis = kwargs.get("vis", False)
        self.freeze_hrnet_weight = cfg['MODEL']["FREEZE_HRNET_WEIGHTS"]
        self.conv1 = nn.Conv2d(3, 64, kernel_size=3, stride=2, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(64, momentum=BN_MOMENTUM)
        self.conv2 = nn.Conv2d(64, 64, kernel_size=3, stride=2, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(64, momentum=BN_MOMENTUM)
        self.relu = nn.ReLU(inplace=True)
        self.layer1 = self._make_layer(Bottleneck, 64, 4)

        pre = [256]
        for sname in ("STAGE2", "STAGE3", "STAGE4"):
            scfg = cfg['MODEL']['EXTRA'][sname]
            block = blocks_dict[scfg['BLOCK']]
            num_channels = [c * block.expansion for c in scfg['NUM_CHANNELS']]
            idx = sname[-1]
            setattr(self, "stage%s_cfg" % idx, scfg)
            setattr(self, "transition%d" % (int(idx) - 1), self._make_transition_layer(pre, num_channels))
            stage, pre = self._make_stage(scfg, num_channels, multi_scale_output=(sname != "STAGE4"))
            setattr(self, "stage%s" % idx, stage)
        self.pre_stage_channels = pre
        k = extra.FINAL_CONV_KERNEL
        self.final_layer = nn.Conv2d(in_channels=pre[0], out_channels=cfg.MODEL.NUM_JOINTS, kernel_size=k, stride=1,
                                     padding=1 if k == 3 else 0)

    def _make_layer(self, block, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                nn.Conv2d(self.inplanes, planes * block.expansion, kernel_size=1, stride=stride, bias=False),
                nn.BatchNorm2d(planes * block.expansion, momentum=BN_MOMENTUM))
        layers = [block(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes))
        return nn.Sequential(*layers)

    def _make_transition_layer(self, pre, cur):
        out = []
        for i in range(len(cur)):
            if i < len(pre):
                if cur[i] != pre[i]:
                    out.append(nn.Sequential(nn.Conv2d(pre[i], cur[i], 3, 1, 1, bias=False), nn.BatchNorm2d(cur[i]),
                                             nn.ReLU(inplace=True)))
                else:
                    out.append(None)
            else:
                steps = []
                for j in range(i + 1 - len(pre)):
                    cin = pre[-1]
                    cout = cur[i] if j == i - len(pre) else cin
                    steps.append(nn.Sequential(nn.Conv2d(cin, cout, 3, 2, 1, bias=False), nn.BatchNorm2d(cout),
                                               nn.ReLU(inplace=True)))
                out.append(nn.Sequential(*steps))
        return nn.ModuleList(out)

    def _make_stage(self, layer_config, num_inchannels, multi_scale_output=True):
        num_modules = layer_config['NUM_MODULES']
        block = blocks_dict[layer_config['BLOCK']]
        modules = []
        for i in range(num_modules):
            mso = multi_scale_output or i != num_modules - 1
            modules.append(HighResolutionModule(layer_config['NUM_BRANCHES'], block, layer_config['NUM_BLOCKS'],
                                                num_inchannels, layer_config['NUM_CHANNELS'],
                                                layer_config['FUSE_METHOD'], mso))
            num_inchannels = modules[-1].get_num_inchannels()
        return nn.Sequential(*modules), num_inchannels

    def freeze_weight(self):
        """hrnet.py:686-690."""
        for p in self.parameters():
            p.requires_grad = False

    # -- fused trunk ---------------------------------------------------------------------------
    @staticmethod
    def _apply_transition(t, x):
        """transition entry: Sequential(conv,bn,relu) or Sequential(Sequential(conv,bn,relu), ...)."""
        first = list(t.children())[0]
        if isinstance(first, nn.Conv2d):
            conv, bn = _conv_bn(t)
            return ops.conv_bn_act(x, conv, bn, relu=True)
        for step in t.children():
            conv, bn = _conv_bn(step)
            x = ops.conv_bn_act(x, conv, bn, relu=True)
        return x

    def _trunk(self, x, keep_stage4_inputs=False):
        """hrnet.py:651-679 (HRNetPlus) == :302-326 (HRNet)."""
        x = ops.conv_bn_act(x, self.conv1, self.bn1, relu=True)
        x = ops.conv_bn_act(x, self.conv2, self.bn2, relu=True)
        for blk in self.layer1:
            x = blk(x)
        xs = []
        for i in range(self.stage2_cfg['NUM_BRANCHES']):
            xs.append(self._apply_transition(self.transition1[i], x) if self.transition1[i] is not None else x)
        ys = xs
        for m in self.stage2:
            ys = m(ys)
        xs = []
        for i in range(self.stage3_cfg['NUM_BRANCHES']):
            xs.append(self._apply_transition(self.transition2[i], ys[-1]) if self.transition2[i] is not None else ys[i])
        ys = xs
        for m in self.stage3:
            ys = m(ys)
        xs = []
        for i in range(self.stage4_cfg['NUM_BRANCHES']):
            xs.append(self._apply_transition(self.transition3[i], ys[-1]) if self.transition3[i] is not None else ys[i])
        x3_list = xs
        ys = xs
        for m in self.stage4:
            ys = m(ys)
        return ys, x3_list


class HRNetPlus(_HRNetTrunk):
    """hrnet.py:521-869: forward(x) -> (rough_pose_heatmaps, [feat])."""

    def __init__(self, cfg, is_train, use_deconv=False, use_prediction=False, **kwargs):
        super().__init__()
        self._build(cfg, is_train, use_deconv, use_prediction, kwargs)

    def init_weights(self, *args, **kwargs):
        return None

    def forward(self, x, **kwargs):
        if kwargs.get("similar", False):
            raise NotImplementedError("HRNetPlus(similar=True) is not used by any registered model")
        x = ops.to_nhwc(x, torch.float32)   # the stem reads fp32 pixels in both precisions
        ys, _ = self._trunk(x)
        if kwargs.get("heatmap", True) is False:
            return ys[0]
        hm = ops.conv_bn_act(ys[0], self.final_layer, None, relu=False, out_dtype=torch.float32)
        return hm, ys


class HRNet(_HRNetTrunk):
    """hrnet.py:186-333 (single-frame baseline, BASELINE config 1): forward(x) -> (heatmaps, x3_list)."""

    def __init__(self, cfg, is_train, use_deconv=False, use_prediction=False, **kwargs):
        super().__init__()
        self._build(cfg, is_train, use_deconv, use_prediction, kwargs)

    def init_weights(self, *args, **kwargs):
        return None

    def forward(self, x):
        x = ops.to_nhwc(x, torch.float32)
        ys, x3_list = self._trunk(x)
        hm = ops.conv_bn_act(ys[0], self.final_layer, None, relu=False, out_dtype=torch.float32)
        if self.use_deconv or self.use_prediction:
            return x3_list[0], hm
        return hm, x3_list
