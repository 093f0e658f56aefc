"""Run-time proof of the drop-in boundary: the fami modules driven through the REFERENCE's own calling
sequence, i.e. what posetimation/zoo/Alignment/Alignment_V15.py:113-183 executes after fami_pose_b200.patch_reference()
has rebound conv_bn_relu / ChainOfBasicBlocks / HRNetPlus / DeformConv2d / kornia.geometry.warp_affine.

/root/reference does not exist on the GPU box, so the reference's forward() is RESTATED below line by line (same torch
calls, same argument conventions: NCHW float32 inputs, torch.chunk / torch.cat, `sup - kf` through torch, the affine
matrix built with in-place assignment, kornia.geometry.warp_affine(src, M, dsize=(H, W)), separate torchvision-order
offset and mask tensors into DeformConv2d.forward(input, offset, mask), plain nn.Conv2d / nn.Linear / nn.Flatten modules
called through nn.Module.__call__, the 3-tuple in train phase).  Nothing in this forward knows about channels-last
strides, fused offset|mask buffers or output slices.  On this container (CPU, reference present)
tests/test_abi_and_host.py additionally constructs the UNMODIFIED reference class with the patched names.

Outputs are compared with the goldens the unmodified reference produced (tests/golden/model_reference.npz)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import fami_oracle as fo  # noqa: E402  (checker only)
from oracle import ref_harness as rh  # noqa: E402

DEV = "cuda"
SEED = 19970808


def _reference_calling_sequence_model(phase):
    import fami_pose_b200 as fp
    from fami_pose_b200 import kornia_shim
    import types
    kornia = types.SimpleNamespace(geometry=types.SimpleNamespace(warp_affine=kornia_shim.warp_affine))   # Alignment_V15.py:12

    class Alignment_V15_RefForward(fp.Alignment_V15):
        """Same sub-modules / state_dict as fami_pose_b200.Alignment_V15; forward() and the MI estimators are the
        reference's (Alignment_V15.py:113-183, 250-277) restated verbatim."""

        def forward(self, kf_x, sup_x, **kwargs):
            batch_size, num_sup = kf_x.shape[0], sup_x.shape[1] // 3                           # :115
            sup_x = torch.cat(torch.chunk(sup_x, num_sup, dim=1), dim=0)                       # :117
            x = torch.cat([kf_x, sup_x], dim=0)                                                # :119
            x_bb_hm, x_bb_feat = self.hrnet(x)                                                 # :120
            x_bb_hm_list = torch.chunk(x_bb_hm, num_sup + 1, dim=0)                            # :121
            x_bb_feat_list = torch.chunk(x_bb_feat[0], num_sup + 1, dim=0)                     # :122
            kf_bb_hm, kf_bb_feat = x_bb_hm_list[0], x_bb_feat_list[0]                          # :124
            sup_bb_hm_list, sup_bb_feat_list = x_bb_hm_list[1:], x_bb_feat_list[1:]            # :125
            aligned_sup_feat_list = []
            B, _, H, W = kf_bb_hm.shape                                                        # :128
            for i in range(num_sup):                                                           # :130
                sup_bb_hm, sup_bb_feat = sup_bb_hm_list[i], sup_bb_feat_list[i]
                feat_offset = self.feat_global_offset_layers(sup_bb_feat - kf_bb_feat)         # :132  [B,2]
                offset_params = torch.eye(3)[0:2].view(1, 2, 3).repeat(B, 1, 1).to(sup_bb_feat.device)          # :133
                offset_params[:, 0, 2], offset_params[:, 1, 2] = feat_offset[:, 0], feat_offset[:, 1]           # :134
                global_aligned_feat = kornia.geometry.warp_affine(sup_bb_feat, offset_params, dsize=(H, W))     # :135
                aligned_sup_feat_list.append(global_aligned_feat)
            agg_sup_feat = torch.cat(aligned_sup_feat_list, dim=1)                             # :139
            agg_sup_feat = self.sup_agg_block(agg_sup_feat)                                    # :140
            combined_feat = self.combined_feat_layers(torch.cat([agg_sup_feat, kf_bb_feat], dim=1))            # :143
            dcn_offset = self.dcn_offset_1(combined_feat)                                      # :144
            dcn_mask = self.dcn_mask_1(combined_feat)                                          # :145
            combined_feat = self.dcn_1(combined_feat, dcn_offset, dcn_mask)                    # :146
            dcn_offset = self.dcn_offset_2(combined_feat)
            dcn_mask = self.dcn_mask_2(combined_feat)
            combined_feat = self.dcn_2(combined_feat, dcn_offset, dcn_mask)                    # :150
            dcn_offset = self.dcn_offset_3(combined_feat)
            dcn_mask = self.dcn_mask_3(combined_feat)
            aligned_sup_feat = self.dcn_3(agg_sup_feat, dcn_offset, dcn_mask)                  # :154
            dcn_offset = self.dcn_offset_4(aligned_sup_feat)
            dcn_mask = self.dcn_mask_4(aligned_sup_feat)
            aligned_sup_feat = self.dcn_4(aligned_sup_feat, dcn_offset, dcn_mask)              # :158
            kf_sup_feat = torch.cat([kf_bb_feat, aligned_sup_feat], dim=1)                     # :160
            all_agg_features = self.init_feature_agg_block(kf_sup_feat)                        # :161
            final_hm = self.agg_final_layer(all_agg_features)                                  # :163  plain nn.Conv2d.__call__
            if self.is_train:                                                                  # :165-181
                mi = [self.feat_label_mi_estimation(all_agg_features, final_hm),
                      self.feat_feat_mi_estimation(kf_bb_feat, all_agg_features),
                      self.feat_label_mi_estimation(agg_sup_feat, final_hm),
                      self.feat_feat_mi_estimation(agg_sup_feat, all_agg_features),
                      self.feat_label_mi_estimation(kf_bb_feat, final_hm),
                      self.feat_feat_mi_estimation(kf_bb_feat, all_agg_features)]
                return final_hm, kf_bb_hm, mi
            return final_hm, kf_bb_hm

        def feat_label_mi_estimation(self, Feat, Y):                                           # :250-263
            batch_size = Feat.shape[0]
            temperature = 0.05
            pred_Y = self.hrnet.final_layer(Feat)                                              # nn.Conv2d.__call__
            pred_Y = pred_Y.reshape(batch_size * self.num_joints, -1)
            Y = Y.reshape(batch_size * self.num_joints, -1)
            return F.kl_div(input=self.softmax(pred_Y.detach() / temperature), target=self.softmax(Y / temperature),
                            reduction='mean')

        def feat_feat_mi_estimation(self, F1, F2):                                             # :265-277
            batch_size = F1.shape[0]
            temperature = 0.05
            F1 = F1.reshape(batch_size * 48, -1)
            F2 = F2.reshape(batch_size * 48, -1)
            return F.kl_div(input=self.softmax(F1.detach() / temperature), target=self.softmax(F2 / temperature),
                            reduction='mean')

    cfg = rh.make_cfg(48, 17)
    m = Alignment_V15_RefForward(cfg, phase)
    sd = fo.seeded_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, SEED)
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).eval()


@pytest.mark.parametrize("prec,tol", [("fp32", 1e-3), ("tf32", 1e-3)])
def test_reference_calling_sequence_eval(golden_dir, prec, tol):
    """Eval phase, B=1: 2-tuple of NCHW float32 heatmaps equal to the reference golden within the arm's tolerance."""
    import fami_pose_b200 as fp
    gold = np.load(os.path.join(golden_dir, "model_reference.npz"))
    fp.set_precision(prec)
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = prec == "tf32"     # agg_final_layer runs through torch's own nn.Conv2d here
    try:
        m = _reference_calling_sequence_model("validate")
        kf, sup, _, _ = fo.synthetic_clip(1, seed=SEED)
        with torch.no_grad():
            out = m(kf.to(DEV), sup.to(DEV))
        assert isinstance(out, tuple) and len(out) == 2
        hm, kfhm = out
        assert tuple(hm.shape) == (1, 17, 96, 72) and tuple(kfhm.shape) == (1, 17, 96, 72) and hm.dtype == torch.float32
        e1 = float(np.abs(hm.float().cpu().numpy() - gold["v15_eval_final_hm"]).max())
        e2 = float(np.abs(kfhm.float().cpu().numpy() - gold["v15_eval_kf_hm"]).max())
        print("reference calling sequence (%s): final %.3e kf %.3e" % (prec, e1, e2))
        assert e1 <= tol and e2 <= tol
    finally:
        torch.backends.cudnn.allow_tf32 = prev
        fp.set_precision("fp32")


def test_reference_calling_sequence_train_phase_tuple(golden_dir):
    """Train phase (BN in eval mode as in the golden), B=2: 3-tuple with the six MI terms computed by the reference's
    own torch code on the tensors the fami modules return."""
    import fami_pose_b200 as fp
    gold = np.load(os.path.join(golden_dir, "model_reference.npz"))
    fp.set_precision("fp32")
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        m = _reference_calling_sequence_model("train")
        kf, sup, _, _ = fo.synthetic_clip(2, seed=SEED + 1)
        with torch.no_grad():
            out = m(kf.to(DEV), sup.to(DEV))
        assert len(out) == 3 and len(out[2]) == 6
        hm, kfhm, mi = out
        assert float(np.abs(hm.float().cpu().numpy() - gold["v15_train_final_hm"]).max()) <= 1e-3
        got = np.array([float(v) for v in mi])
        ref = gold["v15_train_mi"]
        print("mi", got, ref)
        assert np.all(np.abs(got - ref) <= 1e-6 + 1e-3 * np.abs(ref))
    finally:
        torch.backends.cudnn.allow_tf32 = prev
        fp.set_precision("fp32")


def test_deform_conv2d_reference_argument_convention():
    """DeformConv2d.forward(input, offset, mask) with plain contiguous NCHW float32 tensors in torchvision's channel order
    (offset [B, 2*9*G, H, W] as (dy, dx) pairs per tap, mask [B, 9*G, H, W]) -- what Alignment_V15.py:146 passes."""
    import fami_pose_b200 as fp
    fp.set_precision("fp32")
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 48, 20, 14, generator=g)
    off = 2 * torch.randn(2, 216, 20, 14, generator=g)
    msk = torch.randn(2, 108, 20, 14, generator=g)
    dcn = fp.DeformConv2d(48, 48, 3, padding=3, dilation=3).to(DEV)
    with torch.no_grad():
        y = dcn(x.to(DEV), off.to(DEV), msk.to(DEV))
    assert tuple(y.shape) == (2, 48, 20, 14)
    ref = fo.dcn_fwd(x.numpy(), off.numpy(), msk.numpy(), dcn.weight.detach().cpu().numpy(), dcn.bias.detach().cpu().numpy())
    assert float(np.abs(y.float().cpu().numpy() - ref).max()) <= 1e-4
