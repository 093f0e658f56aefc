"""Training-step parity: one forward + backward of fami_pose_b200.Alignment_V15 (train phase, train-mode
BatchNorm, HRNet frozen -- the reference default) against the gradients the UNMODIFIED reference's autograd
produced on the same seeded inputs (tests/golden/train_reference.npz, made by make_golden.py train: the
reference run in float64 is the pin, its float32 run is stored beside it to show what rounding alone does).
Loss of alignment_mi_function_term6_1.py:119-148.  fp32 arm."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import fami_oracle as fo  # noqa: E402  (checker only)
from oracle import ref_harness as rh  # noqa: E402  (cfg helper only)

DEV = "cuda"
SEED = 19970808


def _digest(g):
    f = g.detach().double().flatten().cpu()
    idx = torch.linspace(0, f.numel() - 1, steps=min(48, f.numel())).long()
    return np.concatenate([[float(f.norm()), float(f.sum())], f[idx].numpy()])


def test_training_step_gradients_vs_reference_golden(golden_dir):
    import fami_pose_b200 as fp
    from fami_pose_b200.loss import JointMSELoss, combine_losses
    fp.set_precision("fp32")
    gold = np.load(os.path.join(golden_dir, "train_reference.npz"))
    cfg = rh.make_cfg(48, 17)
    m = fp.Alignment_V15(cfg, "train")
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(fo.seeded_state_dict(shapes, SEED), strict=True)
    m = m.to(DEV).train()
    kf, sup, tgt, tw = fo.synthetic_clip(2, seed=SEED + 2)
    launches0 = fp._lib.launch_count()
    hm, kfhm, mi = m(kf.to(DEV), sup.to(DEV))
    mse = JointMSELoss()(hm, tgt.to(DEV), tw.to(DEV))
    loss = combine_losses(mse, mi)
    loss.backward()
    torch.cuda.synchronize()
    assert fp._lib.launch_count() - launches0 > 400          # forward + backward ran on our kernels
    e_hm = float(np.abs(hm.detach().cpu().numpy() - gold["final_hm"]).max())
    print("train fwd: final_hm max-abs err %.3e  loss %.6f (ref %.6f)" % (e_hm, float(loss), float(gold["f64/loss"])))
    assert e_hm <= 2e-3
    assert abs(float(mse) - float(gold["f64/mse"])) <= 1e-4 * max(1.0, abs(float(gold["f64/mse"])))
    np.testing.assert_allclose(np.array([float(v) for v in mi]), gold["f64/mi"], rtol=2e-3, atol=1e-6)
    assert abs(float(loss) - float(gold["f64/loss"])) <= 1e-3 * max(1.0, abs(float(gold["f64/loss"])))

    names = [str(n) for n in gold["names"]]
    params = dict(m.named_parameters())
    trainable = [n for n, p in params.items() if p.requires_grad]
    assert sorted(trainable) == sorted(names)                 # same frozen / trainable split as the reference
    rows, bad = [], []
    for n in names:
        g = params[n].grad
        assert g is not None, "no gradient for %s" % n
        got, ref, ref32 = _digest(g), gold["f64/grad/" + n], gold["f32/grad/" + n]
        if ref[0] < 1e-7:
            # a conv bias feeding a train-mode BatchNorm: the exact gradient is 0, both sides hold rounding noise
            assert got[0] < 1e-5, "%s: expected a numerically zero gradient, norm %.3e" % (n, got[0])
            continue
        scale = max(np.abs(ref[2:]).max(), ref[0] / np.sqrt(max(g.numel(), 1)))
        err = float(np.abs(got[2:] - ref[2:]).max() / scale)
        nerr = abs(got[0] - ref[0]) / ref[0]
        err32 = float(np.abs(ref32[2:] - ref[2:]).max() / scale)      # the reference's own fp32 rounding
        rows.append((max(err, nerr), nerr, err, err32, ref[0], n))
        if nerr > 1e-2 or err > max(1e-2, 4 * err32):
            bad.append(rows[-1])
    rows.sort(reverse=True)
    for r in rows[:8]:
        print("grad dev %.3e (norm rel %.3e, samples %.3e; reference fp32 vs fp64 %.3e; |g| %.3e)  %s" % r)
    assert not bad, "gradients deviate from the reference: %s" % ", ".join("%s (%.2e)" % (b[5], b[0]) for b in bad)


def test_adam_step_vs_torch_optim():
    """fami_adam_step == torch.optim.Adam (the optimizer of posetimation/optimizer/optimizer.py:66-72)."""
    import fami_pose_b200 as fp
    from fami_pose_b200 import ops
    g = torch.Generator().manual_seed(9)
    n = 10007
    p0 = torch.randn(n, generator=g)
    ref = torch.nn.Parameter(p0.clone().double())
    opt = torch.optim.Adam([ref], lr=1e-3)
    p = p0.clone().to(DEV)
    m = torch.zeros(n, device=DEV)
    v = torch.zeros(n, device=DEV)
    for step in range(1, 6):
        grad = torch.randn(n, generator=g) * (0.1 if step % 2 else 3.0)
        ref.grad = grad.double()
        opt.step()
        gd = grad.to(DEV)
        fp._lib.call("fami_adam_step", ops._ptr(p), ops._ptr(gd), ops._ptr(m), ops._ptr(v), n, 1e-3, 0.9, 0.999, 1e-8, step,
                     ops._stream())
        assert float((p.double().cpu() - ref.detach()).abs().max()) < 2e-6


def test_train_step_updates_parameters_like_adam_and_keeps_backbone_frozen():
    """One TrainStep: loss is a device scalar, every trainable tensor moves by exactly the Adam update of its own
    gradient (|delta| = lr at step 1 wherever the gradient is non-zero), frozen HRNet weights do not move, and
    a following inference forward sees the updated weights (cache invalidation)."""
    import fami_pose_b200 as fp
    from fami_pose_b200.train import TrainStep
    fp.set_precision("fp32")
    cfg = rh.make_cfg(48, 17)
    m = fp.Alignment_V15(cfg, "train")
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(fo.seeded_state_dict(shapes, SEED), strict=True)
    m = m.to(DEV).train()
    kf, sup, tgt, tw = (t.to(DEV) for t in fo.synthetic_clip(1, seed=SEED + 3))
    before = {n: p.detach().clone() for n, p in m.named_parameters()}
    m.eval()
    with torch.no_grad():
        hm_before = m(kf, sup)[0].clone()
    m.train()
    step = TrainStep(m, lr=1e-3)
    loss, hm = step(kf, sup, tgt, tw)
    assert loss.is_cuda and loss.dim() == 0 and torch.isfinite(loss)
    lr = 1e-3
    for n, p in m.named_parameters():
        d = (p.detach() - before[n])
        if not p.requires_grad:
            assert float(d.abs().max()) == 0.0, n
            continue
        g = p.grad
        nz = g.abs() > 1e-12
        if nz.any():
            # Adam step 1: delta = -lr * g / (|g| + eps)
            expect = -lr * g / (g.abs() + 1e-8)
            assert float((d - expect).abs().max()) < 1e-7, n
    m.eval()
    with torch.no_grad():
        hm_after = m(kf, sup)[0]
    assert float((hm_after - hm_before).abs().max()) > 1e-6
