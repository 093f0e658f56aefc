"""Input pipeline: the affine crop of datasets/zoo/posetrack/PoseTrack_Alignment.py:233-241.
CPU: the oracle's restatement of cv2.warpAffine and the host-side get_affine_transform against the goldens produced by the
unmodified reference + cv2 (tests/golden/crop_reference.npz).  GPU: fami_crop_affine_u8 bit for bit against the same."""
import os

import numpy as np
import pytest
import torch

from oracle import fami_oracle as fo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "crop_reference.npz")
NCASE = 4


def test_oracle_warp_affine_matches_cv2_golden():
    g = np.load(GOLD)
    for i in range(NCASE):
        p = g["case%d/params" % i]
        for f in range(g["frames"].shape[0]):
            got = fo.warp_affine_u8(g["frames"][f], g["case%d/trans" % i], (int(p[5]), int(p[6])))
            assert np.array_equal(got, g["case%d/out" % i][f]), (i, f)


def test_get_affine_transform_matches_reference_golden():
    from fami_pose_b200 import pipeline
    g = np.load(GOLD)
    for i in range(NCASE):
        p = g["case%d/params" % i]
        for fn in (pipeline.get_affine_transform, fo.get_affine_transform_rot):
            t = fn(np.array(p[0:2], np.float32), np.array(p[2:4], np.float32), float(p[4]), (int(p[5]), int(p[6])))
            assert np.abs(t - g["case%d/trans" % i]).max() <= 1e-9 * max(1.0, np.abs(g["case%d/trans" % i]).max())
    # inverse = the map cv2.warpAffine resamples with
    m = pipeline.invert_affine(g["case1/trans"])[0].reshape(2, 3)
    fwd = np.vstack([g["case1/trans"], [0, 0, 1]])
    assert np.abs(np.vstack([m, [0, 0, 1]]) @ fwd - np.eye(3)).max() < 1e-9


@pytest.mark.gpu
def test_crop_affine_u8_bit_exact_vs_cv2_golden():
    """fami_crop_affine_u8 == cv2.warpAffine(INTER_LINEAR) bit for bit, with the golden's matrices and with matrices rebuilt
    by pipeline.get_affine_transform; the fused ToTensor + Normalize output equals normalising those bytes."""
    import fami_pose_b200 as fp
    from fami_pose_b200 import ops, pipeline
    g = np.load(GOLD)
    frames = torch.from_numpy(g["frames"]).cuda()
    for i in range(NCASE):
        p = g["case%d/params" % i]
        osz = (int(p[5]), int(p[6]))
        ref = g["case%d/out" % i]
        out = pipeline.crop_affine_u8(frames, g["case%d/trans" % i], osz)
        assert out.dtype == torch.uint8 and tuple(out.shape) == ref.shape
        assert np.array_equal(out.cpu().numpy(), ref), "case %d: %d bytes differ" % (i, int((out.cpu().numpy() != ref).sum()))
        t = pipeline.get_affine_transform(np.array(p[0:2], np.float32), np.array(p[2:4], np.float32), float(p[4]), osz)
        out2 = pipeline.crop_affine_u8(frames, t, osz)
        assert int((out2.cpu().numpy() != ref).sum()) <= 3 * ref.shape[0]     # a last-bit difference in the solve may move a coordinate
        nrm = pipeline.crop_affine_u8(frames, g["case%d/trans" % i], osz, normalize=True)
        want = (torch.from_numpy(ref).float() / 255.0 - torch.tensor(ops.IMAGENET_MEAN)) / torch.tensor(ops.IMAGENET_STD)
        assert tuple(nrm.shape) == (ref.shape[0], 3, osz[1], osz[0])
        assert float((nrm.permute(0, 2, 3, 1).cpu() - want).abs().max()) <= 1e-6


@pytest.mark.gpu
def test_clip_from_raw_frames_feeds_the_model():
    """Raw uint8 frames -> device crop + normalise -> Alignment_V15 forward equals the same forward on crops made by the
    oracle's cv2 restatement and normalised on the host."""
    import fami_pose_b200 as fp
    from fami_pose_b200 import ops, pipeline
    from oracle import ref_harness as rh
    g = np.load(GOLD)
    raw = np.concatenate([g["frames"], g["frames"][::-1], g["frames"][:1]])     # a 5-frame window
    center, scale, rot = (128.0, 96.0), (0.9, 1.2), 11.0
    x = pipeline.clip_from_frames(torch.from_numpy(raw).cuda(), center, scale, rot)
    t = pipeline.get_affine_transform(np.array(center, np.float32), np.array(scale, np.float32), rot, (288, 384))
    crops = np.stack([fo.warp_affine_u8(fr, t, (288, 384)) for fr in raw])
    want = (torch.from_numpy(crops).float() / 255.0 - torch.tensor(ops.IMAGENET_MEAN)) / torch.tensor(ops.IMAGENET_STD)
    assert float((x.permute(0, 2, 3, 1).cpu() - want).abs().max()) <= 1e-6
    fp.set_precision("fp32")
    m = fp.Alignment_V15(rh.make_cfg(48, 17), "validate")
    m.load_state_dict(fo.seeded_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}))
    m = m.cuda().eval()
    with torch.no_grad():
        hm_dev, _ = m._forward_frames(x, 1, 4)
        nchw = want.permute(0, 3, 1, 2).contiguous().cuda()
        hm_host, _ = m(nchw[:1], nchw[1:].reshape(1, 12, 384, 288))
    assert float((hm_dev - hm_host).abs().max()) <= 1e-6
