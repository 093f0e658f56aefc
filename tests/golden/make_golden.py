"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference,
build container only) and torchvision's CPU deform_conv2d on seeded inputs.

    python tests/golden/make_golden.py

The reference ships no golden vectors of its own; these files are the pins.
Inputs/weights are not stored: they are regenerated from the seeds by oracle.fami_oracle
(seeded_state_dict / synthetic_clip, CPU torch.Generator), only outputs are stored.
"""
import os
import sys

import numpy as np
import torch
import torchvision

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import fami_oracle as fo  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SEED = 19970808  # the reference's seed, tools/run.py:32-34


def dcn_cases():
    """(name, B, C, Cout, G, H, W, offset sigma) -- includes out-of-bounds offsets and G=1."""
    return [("c48g12", 2, 48, 48, 12, 12, 9, 3.0), ("c32g8", 1, 32, 32, 8, 10, 7, 2.0),
            ("c16g1_oob", 1, 16, 32, 1, 6, 5, 8.0), ("c64g16", 1, 64, 48, 16, 8, 8, 1.0)]


def dcn_inputs(name, B, C, Cout, G, H, W, sig, dtype=torch.float32):
    g = torch.Generator().manual_seed(SEED + sum(map(ord, name)))
    x = torch.randn(B, C, H, W, generator=g, dtype=torch.float64).to(dtype)
    off = (sig * torch.randn(B, 18 * G, H, W, generator=g, dtype=torch.float64)).to(dtype)
    msk = torch.randn(B, 9 * G, H, W, generator=g, dtype=torch.float64).to(dtype)
    w = (0.05 * torch.randn(Cout, C, 3, 3, generator=g, dtype=torch.float64)).to(dtype)
    b = (0.1 * torch.randn(Cout, generator=g, dtype=torch.float64)).to(dtype)
    go = torch.randn(B, Cout, H, W, generator=g, dtype=torch.float64).to(dtype)
    return x, off, msk, w, b, go


def make_dcn():
    out = {}
    for case in dcn_cases():
        name = case[0]
        for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
            x, off, msk, w, b, go = dcn_inputs(*case, dtype=dt)
            x.requires_grad_(True); off.requires_grad_(True); msk.requires_grad_(True)
            w.requires_grad_(True); b.requires_grad_(True)
            y = torchvision.ops.deform_conv2d(x, off, w, b, stride=1, padding=3, dilation=3, mask=msk)
            y.backward(go)
            out["%s_%s_out" % (name, tag)] = y.detach().numpy()
            if tag == "f64":
                for nm, t in (("gx", x), ("goff", off), ("gmask", msk), ("gw", w), ("gb", b)):
                    out["%s_%s_%s" % (name, tag, nm)] = t.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "dcn_torchvision.npz"), **out)
    print("dcn_torchvision.npz", len(out))


def make_model():
    ref = rh.load_reference()
    out = {}
    # config 2 shape at B=1: Alignment_V15 W48 384x288, 5 frames, 17 joints -- eval-mode BN
    cfg = rh.make_cfg(48, 17)
    m = ref.Alignment_V15(cfg, 'validate').eval()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = fo.seeded_state_dict(shapes, SEED)
    m.load_state_dict(sd, strict=True)
    kf, sup, tgt, tw = fo.synthetic_clip(1, seed=SEED)
    with torch.no_grad():
        hm, kfhm = m(kf, sup)
    out["v15_eval_final_hm"] = hm.numpy()
    out["v15_eval_kf_hm"] = kfhm.numpy()
    out["v15_eval_final_argmax"] = hm.flatten(2).argmax(2).numpy().astype(np.int32)
    out["v15_eval_kf_argmax"] = kfhm.flatten(2).argmax(2).numpy().astype(np.int32)
    out["v15_eval_mse"] = np.float64(ref.JointMSELoss()(hm, tgt, tw).item())

    # train-phase model (3-tuple incl. the six MI terms), BN in eval mode for determinism, B=2
    mt = ref.Alignment_V15(cfg, 'train').eval()
    mt.load_state_dict(sd, strict=True)
    kf2, sup2, tgt2, tw2 = fo.synthetic_clip(2, seed=SEED + 1)
    with torch.no_grad():
        hm2, kfhm2, mi = mt(kf2, sup2)
    out["v15_train_final_hm"] = hm2.numpy()
    out["v15_train_mi"] = np.array([float(v) for v in mi], dtype=np.float64)
    out["v15_train_mse"] = np.float64(ref.JointMSELoss()(hm2, tgt2, tw2).item())

    # train-mode BatchNorm (batch statistics over the 5B frames), B=1
    mb = ref.Alignment_V15(cfg, 'train').train()
    mb.load_state_dict(sd, strict=True)
    with torch.no_grad():
        hm3, kfhm3, mi3 = mb(kf, sup)
    out["v15_bntrain_final_hm"] = hm3.numpy()
    out["v15_bntrain_kf_hm"] = kfhm3.numpy()
    out["v15_bntrain_mi"] = np.array([float(v) for v in mi3], dtype=np.float64)
    out["v15_bntrain_hrnet_bn1_running_mean"] = mb.state_dict()["hrnet.bn1.running_mean"].numpy()
    out["v15_bntrain_hrnet_bn1_running_var"] = mb.state_dict()["hrnet.bn1.running_var"].numpy()

    # config 1: HRNet-W32 256x192 single-frame forward, batch 1
    cfg32 = rh.make_cfg(32, 17)
    h = ref.HRNet(cfg32, False).eval()
    shapes32 = {k: tuple(v.shape) for k, v in h.state_dict().items()}
    sd32 = fo.seeded_state_dict(shapes32, SEED)
    h.load_state_dict(sd32, strict=True)
    g = torch.Generator().manual_seed(SEED)
    x = torch.randn(1, 3, 256, 192, generator=g)
    with torch.no_grad():
        hm32, _ = h(x)
    out["hrnet_w32_hm"] = hm32.numpy()
    out["hrnet_w32_argmax"] = hm32.flatten(2).argmax(2).numpy().astype(np.int32)
    np.savez_compressed(os.path.join(OUT, "model_reference.npz"), **out)
    print("model_reference.npz", {k: v.shape for k, v in out.items()})


def make_decode():
    """get_final_preds / accuracy / generate_heatmaps of the UNMODIFIED reference (numpy + cv2)."""
    r = rh.load_reference_decode()
    hm, tgt, center, scale, joints, vis = fo.synthetic_decode_case(SEED)
    preds, maxvals = r.get_final_preds(hm.copy(), center, scale)
    acc, avg, cnt, pred = r.accuracy(hm, tgt)
    out = {"final_preds": preds, "maxvals": maxvals, "acc": acc, "avg_acc": np.float64(avg), "cnt": np.int64(cnt),
           "acc_pred": pred}
    acc2, avg2, cnt2, _ = r.accuracy(hm, tgt, thr=0.2)
    out.update({"acc_thr02": acc2, "avg_acc_thr02": np.float64(avg2)})
    T, Wt = [], []
    for b in range(joints.shape[0]):
        t, w = r.generate_heatmaps(joints[b], vis[b], 3, np.array([288, 384]), np.array([72, 96]), joints.shape[1])
        T.append(t); Wt.append(w)
    out["targets"] = np.stack(T); out["target_weight"] = np.stack(Wt)
    np.savez_compressed(os.path.join(OUT, "decode_reference.npz"), **out)
    print("decode_reference.npz", {k: np.shape(v) for k, v in out.items()})


def synthetic_frames(seed, F=3, H=200, W=260):
    """Smooth, textured uint8 RGB frames (compress well, exercise every interpolation weight)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    out = np.zeros((F, H, W, 3), np.uint8)
    for f in range(F):
        for c in range(3):
            a, b, ph = rng.uniform(0.02, 0.2, 3)
            v = 127 + 60 * np.sin(a * xx + ph * 10 + f) + 50 * np.cos(b * yy - ph * 7) + 15 * np.sin(0.9 * xx + 1.3 * yy)
            out[f, :, :, c] = np.clip(v, 0, 255).astype(np.uint8)
        out[f, 40:60, 50:90] = rng.integers(0, 256, (20, 40, 3))          # a noisy patch
    return out


def make_crop():
    """The dataset's affine crop (datasets/zoo/posetrack/PoseTrack_Alignment.py:233-241): the UNMODIFIED reference's
    get_affine_transform (datasets/process/affine_transform.py:13-45) + cv2.warpAffine(INTER_LINEAR) on synthetic frames:
    rotation / scale augmentation, a crop reaching far outside the image (constant-0 border), an up-scaling crop."""
    import cv2
    import importlib
    rh.load_reference_decode()          # installs the bare `datasets.process` package stub
    at = importlib.import_module("datasets.process.affine_transform")
    cases = [  # center (x, y), scale (w, h)/200, rot, output (w, h)
        ((130.0, 100.0), (0.9, 1.2), 0.0, (288, 384)),
        ((110.5, 92.25), (0.75, 1.0), 33.7, (144, 192)),
        ((20.0, 180.0), (1.3, 1.7333), -51.2, (144, 192)),
        ((128.0, 96.0), (0.3, 0.4), 7.0, (96, 128)),
    ]
    out = {"frames": synthetic_frames(SEED, F=2)}
    for i, (c, sc, r, osz) in enumerate(cases):
        trans = at.get_affine_transform(np.array(c, np.float32), np.array(sc, np.float32), r, osz)
        out["case%d/params" % i] = np.array([c[0], c[1], sc[0], sc[1], r, osz[0], osz[1]], np.float64)
        out["case%d/trans" % i] = np.asarray(trans, np.float64)
        out["case%d/out" % i] = np.stack([cv2.warpAffine(fr, trans, (int(osz[0]), int(osz[1])), flags=cv2.INTER_LINEAR)
                                          for fr in out["frames"]])
    out["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(OUT, "crop_reference.npz"), **out)
    print("crop_reference.npz", {k: np.shape(v) for k, v in out.items()})


def grad_digest(g):
    """Compact pin of one gradient tensor: L2 norm, sum, and 48 evenly strided samples."""
    f = g.detach().double().flatten()
    idx = torch.linspace(0, f.numel() - 1, steps=min(48, f.numel())).long()
    return np.concatenate([[float(f.norm()), float(f.sum())], f[idx].numpy()])


def make_train():
    """One training-step backward of the unmodified reference (train phase, train-mode BatchNorm, HRNet
    frozen as in configs/Alignment/Base_PoseTrack17.yaml), B=2: the loss of
    alignment_mi_function_term6_1.py:119-148 and a digest of every trainable parameter's gradient.
    Run twice: in float64 (the pin) and in the reference's own float32 (stored beside it so the test can
    state how far fp32 rounding alone moves these heavily cancelling sums)."""
    ref = rh.load_reference()
    cfg = rh.make_cfg(48, 17)
    out = {}
    for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        m = ref.Alignment_V15(cfg, 'train').train()
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        m.load_state_dict(fo.seeded_state_dict(shapes, SEED), strict=True)
        m = m.to(dt)
        kf, sup, tgt, tw = fo.synthetic_clip(2, seed=SEED + 2)
        hm, kfhm, mi = m(kf.to(dt), sup.to(dt))
        mse = ref.JointMSELoss()(hm, tgt.to(dt), tw.to(dt))
        loss = mse * 1.0 + 0.5 * (-0.1 * mi[0] + 0.1 * mi[1] + mi[2] - mi[3] + mi[4] - mi[5])
        loss.backward()
        out.update({tag + "/loss": np.float64(loss.item()), tag + "/mse": np.float64(mse.item()),
                    tag + "/mi": np.array([float(v) for v in mi], dtype=np.float64)})
        if tag == "f64":
            out["final_hm"] = hm.detach().float().numpy()
        names = []
        for name, p_ in m.named_parameters():
            if p_.requires_grad:
                assert p_.grad is not None, name
                names.append(name)
                out[tag + "/grad/" + name] = grad_digest(p_.grad)
        out["names"] = np.array(names)
        print(tag, "loss", out[tag + "/loss"], flush=True)
    np.savez_compressed(os.path.join(OUT, "train_reference.npz"), **out)
    print("train_reference.npz", len(names), "trainable tensors")


def make_train_unfrozen():
    """The same training-step backward with the backbone UN-frozen (FREEZE_HRNET_WEIGHTS: false,
    Alignment_V15.py:110-111 / hrnet.py:686-690): a digest of the gradient of every one of the model's parameters
    (HRNet-W48 included: the 258.6 MB all-reduce case of SURVEY.md 8e).  B=1 (5 frames through train-mode BatchNorm),
    float64 pin + the reference's own float32 run."""
    ref = rh.load_reference()
    cfg = rh.make_cfg(48, 17, freeze_hrnet=False)
    out = {}
    for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        m = ref.Alignment_V15(cfg, 'train').train()
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        m.load_state_dict(fo.seeded_state_dict(shapes, SEED), strict=True)
        m = m.to(dt)
        kf, sup, tgt, tw = fo.synthetic_clip(1, seed=SEED + 3)
        hm, kfhm, mi = m(kf.to(dt), sup.to(dt))
        mse = ref.JointMSELoss()(hm, tgt.to(dt), tw.to(dt))
        loss = mse * 1.0 + 0.5 * (-0.1 * mi[0] + 0.1 * mi[1] + mi[2] - mi[3] + mi[4] - mi[5])
        loss.backward()
        out.update({tag + "/loss": np.float64(loss.item()), tag + "/mse": np.float64(mse.item())})
        if tag == "f64":
            out["final_hm"] = hm.detach().float().numpy()
        names = []
        for name, p_ in m.named_parameters():
            assert p_.requires_grad, name
            if p_.grad is None:        # parameters the loss does not reach (e.g. unused heads)
                continue
            names.append(name)
            out[tag + "/grad/" + name] = grad_digest(p_.grad)
        out["names"] = np.array(names)
        out["n_params_total"] = np.int64(sum(1 for _ in m.named_parameters()))
        print(tag, "loss", out[tag + "/loss"], len(names), "tensors with gradients", flush=True)
    np.savez_compressed(os.path.join(OUT, "train_unfrozen_reference.npz"), **out)
    print("train_unfrozen_reference.npz", len(names), "tensors")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    if len(sys.argv) > 1 and sys.argv[1] == "crop":
        make_crop()
    elif len(sys.argv) > 1 and sys.argv[1] == "train_unfrozen":
        make_train_unfrozen()
    elif len(sys.argv) > 1 and sys.argv[1] == "train":
        make_train()
    elif len(sys.argv) > 1 and sys.argv[1] == "decode":
        make_decode()
    else:
        make_decode()
        make_crop()
        make_dcn()
        make_model()
        make_train()
