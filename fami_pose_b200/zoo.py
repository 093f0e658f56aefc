"""Drop-in replacement for posetimation/zoo/Alignment/Alignment_V15.py (the FAMI-Pose model).

Same constructor (cfg, is_train), sub-module names (state_dict keys), return arity rule (fixed at
construction from `is_train`, Alignment_V15.py:52-56) and forward semantics; BASELINE configs that
the reference cannot express (W32 / 15 joints / 2 supporting frames) are reachable through optional
keyword arguments whose defaults are the reference's literals (SURVEY.md 8a "parametrisation").
"""
import logging

import torch
import torch.nn as nn

from . import ops
from .backbones import HRNetPlus
from .layers import ChainOfBasicBlocks, DeformConv2d, conv_bn_relu

TRAIN_PHASE = "train"  # engine/defaults/__init__.py


class _CatConv:
    """Two convolutions over the same input (dcn_offset_k / dcn_mask_k, Alignment_V15.py:144-145)
    executed as ONE launch writing one [B, 27G, H, W] buffer that the deformable kernel reads in
    place.  Holds a derived, non-trainable copy of the concatenated parameters (refreshed when the
    source parameters change); not an nn.Module, never part of the state_dict."""

    def __init__(self):
        self._ver = None
        self.conv = None

    def get(self, a, b, perm=None):
        ver = (ops._ver(a.weight, a.bias, b.weight, b.bias), perm is not None)
        if self.conv is None or ver != self._ver:
            ca, cb = a.out_channels, b.out_channels
            conv = nn.Conv2d(a.in_channels, ca + cb, a.kernel_size, a.stride, a.padding, a.dilation,
                             bias=a.bias is not None).to(a.weight.device)
            with torch.no_grad():
                wcat = torch.cat([a.weight, b.weight], 0)
                bcat = torch.cat([a.bias, b.bias], 0) if a.bias is not None else None
                if perm is not None:   # tap-major output-channel order for the tensor-core DCN kernel
                    idx = torch.tensor(perm, device=wcat.device)
                    wcat = wcat[idx]
                    bcat = bcat[idx] if bcat is not None else None
                conv.weight.copy_(wcat)
                if bcat is not None:
                    conv.bias.copy_(bcat)
            for p in conv.parameters():
                p.requires_grad = False
            self.conv = conv
            self._ver = ver
        return self.conv


class Alignment_V15(nn.Module):
    """posetimation/zoo/Alignment/Alignment_V15.py:25-277."""

    @classmethod
    def get_model_hyper_parameters(cls, cfg):
        f = cfg.TRAIN.SCALE_FACTOR
        f = f if isinstance(f, list) else [f, f]
        s = "bbox_{}_rot_{}_scale_{}-{}".format(cfg.DATASET.BBOX_ENLARGE_FACTOR, cfg.TRAIN.ROT_FACTOR, 1 - f[0], 1 + f[1])
        if cfg.LOSS.HEATMAP_MSE.USE:
            s += f"_MseLoss_{cfg.LOSS.HEATMAP_MSE.WEIGHT}"
        return s

    def __init__(self, cfg, is_train, width=48, num_sup=4, offset_groups=12, feat_hw=(96, 72), **kwargs):
        super().__init__()
        self.logger = logging.getLogger(__name__)
        self.num_joints = cfg.MODEL.NUM_JOINTS
        self.pretrained = cfg.MODEL.PRETRAINED
        self.is_train = (is_train == TRAIN_PHASE) or (is_train is True)
        self.pretrained_layers = ['*']
        self.width, self.num_sup, self.offset_groups = width, num_sup, offset_groups
        C, J = width, self.num_joints
        if cfg['MODEL']['EXTRA']['STAGE2']['NUM_CHANNELS'][0] != C:
            raise ValueError("cfg backbone width %s != width argument %d"
                             % (cfg['MODEL']['EXTRA']['STAGE2']['NUM_CHANNELS'][0], C))
        if C % offset_groups != 0:
            raise ValueError("offset groups must divide the feature width")
        self.hrnet = HRNetPlus(cfg, self.is_train)
        self.freeze_hrnet_weight = cfg['MODEL']["FREEZE_HRNET_WEIGHTS"]

        h, w = feat_hw
        for _ in range(5):
            h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        self.feat_global_offset_layers = nn.Sequential(
            ChainOfBasicBlocks(C, 16, num_blocks=1),
            conv_bn_relu(16, 16, 3, 2, 1, 1),
            conv_bn_relu(16, 16, 3, 2, 1, 1),
            conv_bn_relu(16, 16, 3, 2, 1, 1),
            conv_bn_relu(16, 16, 3, 2, 1, 1),
            conv_bn_relu(16, 16, 3, 2, 1, 1),
            nn.Flatten(),
            nn.Linear(16 * h * w, 64),
            nn.Linear(64, 64),
            nn.Linear(64, 2),
        )
        n_off = 2 * 3 * 3 * offset_groups
        n_msk = 3 * 3 * offset_groups
        self.combined_feat_layers = ChainOfBasicBlocks(C * 2, C, (3, 3), (1, 1), (1, 1), num_blocks=1)
        for k in (1, 2, 3, 4):
            setattr(self, "dcn_offset_%d" % k, conv_bn_relu(C, n_off, 3, 1, padding=3, dilation=3, has_bn=False,
                                                            has_relu=False))
            setattr(self, "dcn_mask_%d" % k, conv_bn_relu(C, n_msk, 3, 1, padding=3, dilation=3, has_bn=False,
                                                          has_relu=False))
            setattr(self, "dcn_%d" % k, DeformConv2d(C, C, 3, padding=3, dilation=3))
        self.sup_agg_block = ChainOfBasicBlocks(input_channel=C * num_sup, ouput_channel=C, num_blocks=2)
        self.init_feature_agg_block = ChainOfBasicBlocks(input_channel=C * 2, ouput_channel=C, num_blocks=3)
        self.agg_final_layer = nn.Conv2d(C, J, 3, 1, 1)
        self.softmax = torch.nn.Softmax(dim=1)
        self._offmask = [_CatConv() for _ in range(4)]  # plain list: not registered, not in state_dict
        self.init_weights()
        if self.freeze_hrnet_weight:
            self.hrnet.freeze_weight()

    # -- init (Alignment_V15.py:185-248) ---------------------------------------------------------
    def init_weights(self, *args, **kwargs):
        import os.path as osp
        hrnet_names = set()
        for name, m in self.named_modules():
            if name.split('.')[0] == "hrnet":
                hrnet_names.add(name)
            if isinstance(m, nn.Conv2d):
                nn.init.normal_(m.weight, std=0.001)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
            else:
                # Linear / DeformConv2d: bias 0, weight keeps its constructor default (the reference's
                # 'weights' key never matches)
                for pname, p in m.named_parameters(recurse=False):
                    if pname == 'bias':
                        nn.init.constant_(p, 0)
        if self.pretrained and osp.isfile(self.pretrained):
            sd = torch.load(self.pretrained, map_location="cpu")
            sd = sd.get('state_dict', sd)
            if list(sd.keys())[0].startswith('module.'):
                sd = {k[7:]: v for k, v in sd.items()}
            need = {}
            for name, v in sd.items():
                layer = name.split('.')[0]
                if layer in hrnet_names:
                    need[name] = v
                elif "hrnet.{}".format(layer) in hrnet_names:
                    need["hrnet.{}".format(name)] = v
            self.load_state_dict(need, strict=False)
        elif self.pretrained:
            self.logger.error('=> please download pre-trained models first!')

    # -- forward (Alignment_V15.py:113-183) -------------------------------------------------------
    def _global_offset(self, diff):
        L = self.feat_global_offset_layers
        t = L[0](diff)
        for i in range(1, 6):
            t = L[i](t)
        v = ops.flatten_nchw_order(t)
        for i in (7, 8, 9):
            v = ops.linear(v, L[i].weight, L[i].bias)
        return v

    def _dcn(self, k, feat_for_offsets, x, out=None):
        """dcn_offset_k + dcn_mask_k (one fused conv, Alignment_V15.py:144-145) then dcn_k (:146).  On the
        16-bit arm the fused conv writes the tap-major layout the tensor-core DCN kernel streams."""
        off_m, msk_m = getattr(self, "dcn_offset_%d" % k), getattr(self, "dcn_mask_%d" % k)
        dcn = getattr(self, "dcn_%d" % k)
        fused = ops.dcn_fused_supported(self.width, self.offset_groups, feat_for_offsets.dtype)
        perm = ops.tap_major_perm(self.offset_groups) if fused else None
        conv = self._offmask[k - 1].get(off_m.conv, msk_m.conv, perm)
        if fused:
            # the producer writes the blocked layout the deformable kernel of this shape streams (ops.om_to_blocked)
            buf = ops.conv_offsets_blocked(feat_for_offsets, conv, self.offset_groups,
                                           layout=ops.dcn_blocked_layout(self.width, dcn.out_channels, self.offset_groups))
            return dcn(x, None, None, out=out, blocked_om=buf, groups=self.offset_groups)
        buf = ops.conv_bn_act(feat_for_offsets, conv, None, relu=False, out_dtype=torch.float32)   # sub-pixel offsets stay fp32
        n_off = off_m.conv.out_channels
        return dcn(x, buf[:, :n_off], buf[:, n_off:], out=out)

    def forward(self, kf_x, sup_x, **kwargs):
        B, ns = kf_x.shape[0], sup_x.shape[1] // 3
        if ns != self.num_sup:
            raise ValueError("model was built for %d supporting frames, got %d" % (self.num_sup, ns))
        x = ops.frames_to_nhwc(kf_x, sup_x)                       # :117-119, frame-major batch
        return self._forward_frames(x, B, ns)

    def forward_u8(self, kf_u8, sup_u8, **kwargs):
        """fami extension: the same forward from uint8 RGB frames as a loader holds them (key frames [B,H,W,3],
        supporting frames [B,ns,H,W,3]); ToTensor + Normalize of datasets/transforms/build.py:13-22 run on the
        device (ops.frames_u8_to_nhwc).  Bit-identical to forward() on the normalised float tensors."""
        B, ns = kf_u8.shape[0], sup_u8.shape[1]
        if ns != self.num_sup:
            raise ValueError("model was built for %d supporting frames, got %d" % (self.num_sup, ns))
        return self._forward_frames(ops.frames_u8_to_nhwc(kf_u8, sup_u8), B, ns)

    def _forward_frames(self, x, B, ns):
        C = self.width
        graph = (torch.is_grad_enabled() and ops.get_precision() in ("fp32", "tf32")
                 and any(p.requires_grad for p in self.parameters()))
        bp = getattr(self, "backbone_precision", None)
        if graph and bp is not None and bp != ops.get_precision() and not any(
                p.requires_grad for p in self.hrnet.parameters()):
            # opt-in for training: the frozen backbone (96 % of the FLOPs) runs on the tensor-core arm, its
            # outputs are widened to fp32 for the differentiable head
            prev = ops.get_precision()
            ops.set_precision(bp)
            try:
                with torch.no_grad():
                    hm_all, feats = self.hrnet(x)
            finally:
                ops.set_precision(prev)
            hm_all, feat_all = hm_all.float(), feats[0].float()
        else:
            hm_all, feats = self.hrnet(x)                         # :120
            feat_all = feats[0]
        kf_bb_hm, kf_feat, sup_feat = hm_all[:B], feat_all[:B], feat_all[B:]
        _, _, H, W, _ = ops.meta(kf_feat)
        if graph:
            return self._forward_graph(B, ns, kf_bb_hm, kf_feat, sup_feat)

        # :130-137 global translation per supporting frame (shared weights).  Eval-mode BN: all frames
        # in one batch.  Train-mode BN uses per-call batch statistics, so keep the reference's loop.
        if self.feat_global_offset_layers.training:
            txy = torch.cat([self._global_offset(ops.sub_bcast(sup_feat[i * B:(i + 1) * B], kf_feat, 1))
                             for i in range(ns)], 0)
        else:
            txy = self._global_offset(ops.sub_bcast(sup_feat, kf_feat, ns))
        agg_in = ops.empty_nhwc(B, C * ns, H, W, kf_feat.dtype, kf_feat.device)
        for i in range(ns):
            ops.warp_translate(sup_feat[i * B:(i + 1) * B], txy[i * B:(i + 1) * B], out=agg_in[:, C * i:C * (i + 1)])

        cat1 = ops.empty_nhwc(B, 2 * C, H, W, kf_feat.dtype, kf_feat.device)      # [agg_sup_feat | kf_feat] :143
        agg_sup_feat = self.sup_agg_block(agg_in, out=cat1[:, :C])                # :140
        ops.copy_into(kf_feat, cat1[:, C:])
        combined = self.combined_feat_layers(cat1)                                 # :143

        combined = self._dcn(1, combined, combined)                                # :144-146
        combined = self._dcn(2, combined, combined)                                # :148-150
        aligned = self._dcn(3, combined, agg_sup_feat)                             # :152-154 (offsets from the refined stream)
        cat2 = ops.empty_nhwc(B, 2 * C, H, W, kf_feat.dtype, kf_feat.device)       # [kf_feat | aligned] :160
        ops.copy_into(kf_feat, cat2[:, :C])
        self._dcn(4, aligned, aligned, out=cat2[:, C:])                            # :156-158
        all_agg = self.init_feature_agg_block(cat2)                                # :161
        final_hm = ops.conv_bn_act(all_agg, self.agg_final_layer, None, relu=False, out_dtype=torch.float32)  # :163

        final_out, kf_out = ops.to_nchw(final_hm), ops.to_nchw(kf_bb_hm)
        self._last = {"final_hm_nhwc": final_hm, "txy": txy}
        if not self.is_train:
            return final_out, kf_out
        mi = [self.feat_label_mi_estimation(all_agg, final_hm),      # :167
              self.feat_feat_mi_estimation(kf_feat, all_agg),        # :169
              self.feat_label_mi_estimation(agg_sup_feat, final_hm),  # :171
              self.feat_feat_mi_estimation(agg_sup_feat, all_agg),   # :173
              self.feat_label_mi_estimation(kf_feat, final_hm),      # :175
              None]
        mi[5] = mi[1]                                               # :177 identical arguments to mi_2
        return final_out, kf_out, mi

    # -- differentiable forward (training step, fp32 arm) ------------------------------------------
    def _forward_graph(self, B, ns, kf_bb_hm, kf_feat, sup_feat):
        """Same computation as forward() with every head op recorded for autograd (autograd.py): the
        reference's concatenations are torch.cat here so their gradients slice back.  The backbone is
        expected frozen (FREEZE_HRNET_WEIGHTS, the reference default): its fused upsample-on-write
        layers have no backward in this round."""
        from . import autograd as ag
        C = self.width
        if kf_feat.dtype != torch.float32:
            raise NotImplementedError("training runs on fp32 storage: fami.set_precision('fp32' | 'tf32')")
        L = self.feat_global_offset_layers
        txys, warped = [], []
        for i in range(ns):                                                        # :130-137
            sup_i = sup_feat[i * B:(i + 1) * B]
            t = L[0](ag.SubFunction.apply(sup_i, kf_feat))
            for j in range(1, 6):
                t = L[j](t)
            v = t.reshape(t.shape[0], -1)                                          # nn.Flatten in C,H,W order
            for j in (7, 8, 9):
                v = ag.LinearFunction.apply(v, L[j].weight, L[j].bias)
            txys.append(v)
            warped.append(ag.WarpTranslateFunction.apply(sup_i, v))
        txy = torch.cat(txys, 0)
        agg_sup_feat = self.sup_agg_block(ag.cat_channels(warped))                 # :139-140
        combined = self.combined_feat_layers(ag.cat_channels([agg_sup_feat, kf_feat]))   # :143
        combined = self._dcn_graph(1, combined, combined)                          # :144-146
        combined = self._dcn_graph(2, combined, combined)                          # :148-150
        aligned = self._dcn_graph(3, combined, agg_sup_feat)                       # :152-154
        aligned = self._dcn_graph(4, aligned, aligned)                             # :156-158
        all_agg = self.init_feature_agg_block(ag.cat_channels([kf_feat, aligned]))  # :160-161
        final_hm = ops.conv_bn_act(all_agg, self.agg_final_layer, None, relu=False, out_dtype=torch.float32)  # :163
        final_out, kf_out = final_hm.contiguous(), ops.to_nchw(kf_bb_hm)
        # detached: a live grad_fn here would keep the whole autograd graph (and its AccumulateGrad nodes) alive until the
        # next forward, which breaks CUDA-graph capture of the training step on another stream
        self._last = {"final_hm_nhwc": final_hm.detach(), "txy": txy.detach()}
        if not self.is_train:
            return final_out, kf_out
        mi = [self.feat_label_mi_estimation(all_agg, final_hm),
              self.feat_feat_mi_estimation(kf_feat, all_agg),
              self.feat_label_mi_estimation(agg_sup_feat, final_hm),
              self.feat_feat_mi_estimation(agg_sup_feat, all_agg),
              self.feat_label_mi_estimation(kf_feat, final_hm),
              None]
        mi[5] = mi[1]
        return final_out, kf_out, mi

    def _dcn_graph(self, k, feat_for_offsets, x):
        """Differentiable _dcn: the offset and mask convolutions still run as one launch over the
        concatenated (differentiable) parameters; torchvision channel order."""
        from . import autograd as ag
        off_m, msk_m = getattr(self, "dcn_offset_%d" % k), getattr(self, "dcn_mask_%d" % k)
        w = torch.cat([off_m.conv.weight, msk_m.conv.weight], 0)
        b = torch.cat([off_m.conv.bias, msk_m.conv.bias], 0) if off_m.conv.bias is not None else None
        buf = ag.conv_bn_act(feat_for_offsets, off_m.conv, None, False, None, weight=w, bias=b)
        n_off = off_m.conv.out_channels
        return getattr(self, "dcn_%d" % k)(x, buf[:, :n_off], buf[:, n_off:])

    def feat_label_mi_estimation(self, Feat, Y):
        """Alignment_V15.py:250-263."""
        if torch.is_grad_enabled() and Y.requires_grad:
            # the reference detaches the prediction branch (:259): only Y receives a gradient
            from . import autograd as ag
            with torch.no_grad():
                pred_Y = ops.conv_bn_act(ops.to_nhwc(Feat.detach()), self.hrnet.final_layer, None, relu=False,
                                         out_dtype=torch.float32)
            return ag.SoftmaxPklFunction.apply(pred_Y, ops.to_nhwc(Y), 0.05)
        pred_Y = ops.conv_bn_act(ops.to_nhwc(Feat), self.hrnet.final_layer, None, relu=False, out_dtype=torch.float32)
        return ops.softmax_pkl(pred_Y, ops.to_nhwc(Y), 0.05)

    def feat_feat_mi_estimation(self, F1, F2):
        """Alignment_V15.py:265-277."""
        if torch.is_grad_enabled() and F2.requires_grad:
            from . import autograd as ag
            return ag.SoftmaxPklFunction.apply(ops.to_nhwc(F1.detach()), ops.to_nhwc(F2), 0.05)
        return ops.softmax_pkl(ops.to_nhwc(F1), ops.to_nhwc(F2), 0.05)
