"""Keypoint decode (get_final_preds), PCK accuracy and gaussian target generation:
the numpy oracle against the pins produced by the UNMODIFIED reference functions (tests/golden/decode_reference.npz,
make_golden.py decode), and the CUDA kernels against both through the C ABI."""
import os

import numpy as np
import pytest
import torch

from oracle import fami_oracle as fo  # checker only

SEED = 19970808


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "decode_reference.npz"))


def test_oracle_decode_vs_reference_golden(gold):
    hm, tgt, center, scale, joints, vis = fo.synthetic_decode_case(SEED)
    preds, maxvals = fo.get_final_preds(hm, center, scale)
    assert np.abs(preds - gold["final_preds"]).max() <= 1e-4          # float32 image coordinates up to ~400 px
    assert np.array_equal(maxvals, gold["maxvals"])
    acc, avg, cnt, pred = fo.accuracy(hm, tgt)
    assert np.array_equal(acc, gold["acc"]) and avg == float(gold["avg_acc"]) and cnt == int(gold["cnt"])
    assert np.array_equal(pred, gold["acc_pred"])
    acc2, avg2, _, _ = fo.accuracy(hm, tgt, thr=0.2)
    assert np.array_equal(acc2, gold["acc_thr02"]) and avg2 == float(gold["avg_acc_thr02"])
    for b in range(joints.shape[0]):
        t, w = fo.generate_heatmaps(joints[b], vis[b], 3, np.array([288, 384]), np.array([72, 96]), joints.shape[1])
        assert np.array_equal(t, gold["targets"][b]) and np.array_equal(w, gold["target_weight"][b])
    assert 0 < gold["target_weight"].sum() < gold["target_weight"].size     # both branches pinned


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["nchw", "nhwc_f16"])
def test_final_preds_cuda_vs_reference_golden(gold, layout):
    import fami_pose_b200 as fp
    from fami_pose_b200 import ops
    hm, tgt, center, scale, joints, vis = fo.synthetic_decode_case(SEED)
    t = torch.from_numpy(hm).cuda()
    if layout == "nhwc_f16":
        # 16-bit channels-last activation as the tensor-core arm would hand it over: compare with the oracle run
        # on the same rounded values
        t = ops.to_nhwc(t, torch.float16)
        hm = ops.to_nchw(t).cpu().numpy()
        ref, refmax = fo.get_final_preds(hm, center, scale)
    else:
        ref, refmax = gold["final_preds"], gold["maxvals"]
    preds, maxvals = fp.get_final_preds(t, center, scale)
    assert preds.shape == (4, 17, 2) and maxvals.shape == (4, 17, 1)
    assert np.abs(preds.cpu().numpy() - ref).max() <= 1e-4
    assert np.array_equal(maxvals.cpu().numpy(), refmax)


@pytest.mark.gpu
def test_accuracy_cuda_vs_reference_golden(gold):
    import fami_pose_b200 as fp
    hm, tgt, center, scale, joints, vis = fo.synthetic_decode_case(SEED)
    acc, avg, cnt, pred = fp.accuracy(torch.from_numpy(hm).cuda(), torch.from_numpy(tgt).cuda())
    assert acc.is_cuda and avg.is_cuda                                  # no host round trip inside
    np.testing.assert_allclose(acc.cpu().numpy(), gold["acc"], rtol=0, atol=1e-15)
    assert abs(float(avg) - float(gold["avg_acc"])) < 1e-15 and int(cnt) == int(gold["cnt"])
    assert np.array_equal(pred.cpu().numpy(), gold["acc_pred"])
    acc2, avg2, _, _ = fp.accuracy(torch.from_numpy(hm).cuda(), torch.from_numpy(tgt).cuda(), thr=0.2)
    np.testing.assert_allclose(acc2.cpu().numpy(), gold["acc_thr02"], rtol=0, atol=1e-15)


@pytest.mark.gpu
def test_gaussian_targets_cuda_vs_reference_golden(gold):
    import fami_pose_b200 as fp
    hm, tgt, center, scale, joints, vis = fo.synthetic_decode_case(SEED)
    t, w = fp.generate_heatmaps(joints, vis, 3, (288, 384), (72, 96), 17)
    assert t.shape == (4, 17, 96, 72) and w.shape == (4, 17, 1)
    assert np.array_equal(w.cpu().numpy(), gold["target_weight"])
    got, ref = t.cpu().numpy(), gold["targets"]
    assert np.array_equal(got != 0, ref != 0)                            # identical support (patch placement, clipping)
    assert np.abs(got - ref).max() <= 2e-7                               # expf vs numpy's float32 exp: <= 2 ulp at 1.0
    assert np.array_equal(got.reshape(4, 17, -1).argmax(2), ref.reshape(4, 17, -1).argmax(2))


@pytest.mark.gpu
def test_frames_u8_normalize_bit_exact_and_forward_u8():
    """uint8 frames -> normalised fp32 NHWC on the device == torchvision's ToTensor + Normalize
    (datasets/transforms/build.py:13-22: x/255, then (x - mean)/std) bit for bit, in the frame-major order of
    Alignment_V15.py:115-119; and the model's forward_u8 equals forward() on the normalised float tensors."""
    import fami_pose_b200 as fp
    from fami_pose_b200 import ops
    from oracle import ref_harness as rh
    g = torch.Generator().manual_seed(5)
    B, ns, H, W = 2, 4, 64, 64
    kf = torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8)
    sup = torch.randint(0, 256, (B, ns, H, W, 3), generator=g, dtype=torch.uint8)
    mean = torch.tensor(ops.IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(ops.IMAGENET_STD).view(1, 3, 1, 1)

    def tv(x_hwc):                                         # ToTensor + Normalize on [N,H,W,3] uint8
        t = x_hwc.permute(0, 3, 1, 2).float().div(255)
        return t.sub(mean).div(std)
    kf_f = tv(kf)                                           # [B,3,H,W]
    sup_f = torch.cat([tv(sup[:, i]) for i in range(ns)], 1)  # [B,3*ns,H,W] as the loader stacks the window
    x = ops.frames_u8_to_nhwc(kf.cuda(), sup.cuda())
    ref = ops.frames_to_nhwc(kf_f.cuda(), sup_f.cuda())
    assert x.shape == ref.shape == ((1 + ns) * B, 3, H, W)
    assert torch.equal(x, ref)
    fp.set_precision("fp32")
    m = fp.Alignment_V15(rh.make_cfg(48, 17), "validate", feat_hw=(16, 16))
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(fo.seeded_state_dict(shapes, SEED), strict=True)
    m = m.cuda().eval()
    with torch.no_grad():
        a = m(kf_f.cuda(), sup_f.cuda())
        b = m.forward_u8(kf.cuda(), sup.cuda())
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
