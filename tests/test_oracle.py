"""CPU: pins the oracle (oracle/fami_oracle.py) against (a) torchvision's CPU deform_conv2d goldens,
(b) outputs of the UNMODIFIED reference model run in the build container (tests/golden), and
(c) torch's own functional ops for the loss restatements.  The reference ships no test vectors of
its own; tests/golden/make_golden.py is the generating script."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import fami_oracle as fo
from oracle import ref_harness as rh
from tests_support import SEED, dcn_cases, dcn_inputs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("case", dcn_cases(), ids=lambda c: c[0])
def test_dcn_fwd_oracle_vs_torchvision_golden(case):
    gold = np.load(os.path.join(GOLD, "dcn_torchvision.npz"))
    name = case[0]
    x, off, msk, w, b, go = dcn_inputs(*case, dtype=torch.float64)
    out = fo.dcn_fwd(x.numpy(), off.numpy(), msk.numpy(), w.numpy(), b.numpy())
    assert np.abs(out - gold["%s_f64_out" % name]).max() <= 1e-12
    x, off, msk, w, b, go = dcn_inputs(*case, dtype=torch.float32)
    out32 = fo.dcn_fwd(x.numpy(), off.numpy(), msk.numpy(), w.numpy(), b.numpy())
    assert np.abs(out32 - gold["%s_f32_out" % name]).max() <= 2e-5


@pytest.mark.parametrize("case", dcn_cases()[:3], ids=lambda c: c[0])
def test_dcn_bwd_oracle_vs_torchvision_autograd_golden(case):
    """analytic backward == torchvision autograd, fp64, incl. OOB offsets."""
    gold = np.load(os.path.join(GOLD, "dcn_torchvision.npz"))
    name = case[0]
    x, off, msk, w, b, go = dcn_inputs(*case, dtype=torch.float64)
    gx, goff, gmask, gw, gb = fo.dcn_bwd(x.numpy(), off.numpy(), msk.numpy(), w.numpy(), go.numpy())
    for got, key in ((gx, "gx"), (goff, "goff"), (gmask, "gmask"), (gw, "gw"), (gb, "gb")):
        ref = gold["%s_f64_%s" % (name, key)]
        assert np.abs(got - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max()), key


def test_dcn_live_torchvision_ragged_and_empty_offsets():
    """live check against the installed torchvision (the un-vendored dependency): tiny ragged shape,
    all samples out of bounds -> bias only."""
    import torchvision
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 8, 3, 5, generator=g, dtype=torch.float64)
    off = torch.full((1, 18 * 2, 3, 5), 100.0, dtype=torch.float64)
    msk = torch.randn(1, 9 * 2, 3, 5, generator=g, dtype=torch.float64)
    w = torch.randn(4, 8, 3, 3, generator=g, dtype=torch.float64)
    b = torch.randn(4, generator=g, dtype=torch.float64)
    ref = torchvision.ops.deform_conv2d(x, off, w, b, padding=3, dilation=3, mask=msk).numpy()
    out = fo.dcn_fwd(x.numpy(), off.numpy(), msk.numpy(), w.numpy(), b.numpy())
    assert np.abs(out - ref).max() <= 1e-12
    assert np.abs(out - b.numpy()[None, :, None, None]).max() <= 1e-12


def test_warp_translate_closed_form_equals_kornia_restatement():
    g = torch.Generator().manual_seed(3)
    src = torch.randn(4, 6, 12, 9, generator=g, dtype=torch.float64)
    txy = torch.tensor([[0.0, 0.0], [1.5, -2.25], [-0.75, 3.0], [30.0, 1.0]], dtype=torch.float64)
    M = torch.eye(3, dtype=torch.float64)[0:2].view(1, 2, 3).repeat(4, 1, 1)
    M[:, 0, 2], M[:, 1, 2] = txy[:, 0], txy[:, 1]
    a = fo.warp_affine_kornia(src, M, (12, 9)).numpy()
    b = fo.warp_translate(src.numpy(), txy.numpy())
    assert np.abs(a - b).max() <= 1e-12
    assert np.abs(b[0] - src.numpy()[0]).max() == 0.0     # identity
    assert np.abs(b[3]).max() == 0.0                       # shifted fully out of view


def test_losses_vs_torch():
    g = torch.Generator().manual_seed(9)
    B, J, H, W = 3, 17, 12, 9
    pred = torch.randn(B, J, H, W, generator=g, dtype=torch.float64)
    tgt = torch.rand(B, J, H, W, generator=g, dtype=torch.float64)
    tw = (torch.rand(B, J, 1, generator=g) < 0.85).double()
    # reference formula, mse_loss.py:21-40
    loss = 0
    for j in range(J):
        loss = loss + F.mse_loss(pred[:, j].reshape(B, -1) * tw[:, j], tgt[:, j].reshape(B, -1) * tw[:, j])
    assert abs(fo.joint_mse(pred.numpy(), tgt.numpy(), tw.numpy()) - float(loss / J)) <= 1e-12
    a = torch.randn(B * J, H * W, generator=g, dtype=torch.float64)
    b = torch.randn(B * J, H * W, generator=g, dtype=torch.float64)
    ref = F.kl_div(input=F.softmax(a / 0.05, dim=1), target=F.softmax(b / 0.05, dim=1), reduction="mean")
    assert abs(fo.softmax_pkl(a.numpy(), b.numpy()) - float(ref)) <= 1e-12


def test_get_max_preds_ties_and_negative():
    hm = np.zeros((1, 3, 4, 5), np.float32)
    hm[0, 1, 2, 3] = 2.0
    hm[0, 1, 3, 1] = 2.0
    hm[0, 2] = -1.0
    preds, maxv, idx = fo.get_max_preds(hm)
    assert idx.tolist() == [[0, 13, 0]]
    assert preds[0, 1].tolist() == [3.0, 2.0] and preds[0, 2].tolist() == [0.0, 0.0]
    assert maxv[0, :, 0].tolist() == [0.0, 2.0, -1.0]


@pytest.fixture(scope="module")
def gold_model():
    return np.load(os.path.join(GOLD, "model_reference.npz"))


@pytest.fixture(scope="module")
def sd48():
    import json
    keys = json.load(open(os.path.join(GOLD, "state_dict_keys.json")))["alignment_v15_w48"]
    return fo.seeded_state_dict({k: tuple(s) for k, s in keys}, SEED)


def test_functional_model_vs_reference_golden_eval(gold_model, sd48):
    """whole-model restatement == unmodified reference (eval-mode BN), fp32, same machine class:
    tolerance 1e-5 (identical torch CPU kernels, possibly different thread counts)."""
    kf, sup, tgt, tw = fo.synthetic_clip(1, seed=SEED)
    with torch.no_grad():
        hm, kfhm = fo.FunctionalFami(sd48).alignment(kf, sup)
    assert float(np.abs(hm.numpy() - gold_model["v15_eval_final_hm"]).max()) <= 1e-5
    assert float(np.abs(kfhm.numpy() - gold_model["v15_eval_kf_hm"]).max()) <= 1e-5
    assert np.array_equal(fo.get_max_preds(hm.numpy())[2].astype(np.int32), gold_model["v15_eval_final_argmax"])
    assert abs(fo.joint_mse(hm.numpy(), tgt.numpy(), tw.numpy()) - float(gold_model["v15_eval_mse"])) <= 1e-6


def test_functional_model_vs_reference_golden_train_mi(gold_model, sd48):
    kf, sup, tgt, tw = fo.synthetic_clip(2, seed=SEED + 1)
    with torch.no_grad():
        hm, kfhm, mi = fo.FunctionalFami(sd48).alignment(kf, sup, with_mi=True)
    assert float(np.abs(hm.numpy() - gold_model["v15_train_final_hm"]).max()) <= 1e-5
    got = np.array([float(v) for v in mi])
    assert np.abs(got - gold_model["v15_train_mi"]).max() <= 1e-8


def test_functional_model_vs_reference_golden_bn_train(gold_model, sd48):
    kf, sup, tgt, tw = fo.synthetic_clip(1, seed=SEED)
    with torch.no_grad():
        hm, kfhm, mi = fo.FunctionalFami(sd48, bn_train=True).alignment(kf, sup, with_mi=True)
    assert float(np.abs(hm.numpy() - gold_model["v15_bntrain_final_hm"]).max()) <= 1e-4
    assert float(np.abs(kfhm.numpy() - gold_model["v15_bntrain_kf_hm"]).max()) <= 1e-4


def test_functional_hrnet_w32_vs_reference_golden(gold_model):
    import json
    keys = json.load(open(os.path.join(GOLD, "state_dict_keys.json")))["hrnet_w32"]
    sd = fo.seeded_state_dict({k: tuple(s) for k, s in keys}, SEED)
    g = torch.Generator().manual_seed(SEED)
    x = torch.randn(1, 3, 256, 192, generator=g)
    with torch.no_grad():
        hm, _ = fo.FunctionalFami(sd, width=32).hrnet_trunk(x, "")
    assert float(np.abs(hm.numpy() - gold_model["hrnet_w32_hm"]).max()) <= 1e-5


@pytest.mark.skipif(not rh.reference_available(), reason="reference tree only exists in the build container")
def test_live_reference_matches_golden(gold_model, sd48):
    """build container only: the unmodified reference reproduces the committed golden."""
    ref = rh.load_reference()
    m = ref.Alignment_V15(rh.make_cfg(48, 17), 'validate').eval()
    m.load_state_dict(sd48, strict=True)
    kf, sup, _, _ = fo.synthetic_clip(1, seed=SEED)
    with torch.no_grad():
        hm, kfhm = m(kf, sup)
    assert float(np.abs(hm.numpy() - gold_model["v15_eval_final_hm"]).max()) <= 1e-6
