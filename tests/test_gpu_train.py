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


def test_bn_eval_after_train_sees_updated_running_stats():
    """eval -> train -> eval (the trainer's train-epoch / validate-epoch loop): the folded eval-mode BatchNorm affine
    is rebuilt from the running statistics the train-mode forward updated through raw pointers (the cache
    key is the version counter, which fami_bn_finalize cannot bump itself).  The second eval forward must equal a
    freshly built model loaded from the current state_dict."""
    import fami_pose_b200 as fp
    fp.set_precision("fp32")
    cfg = rh.make_cfg(48, 17)
    m = fp.Alignment_V15(cfg, "validate")
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(fo.seeded_state_dict(shapes, SEED), strict=True)
    m = m.to(DEV)
    kf, sup, _, _ = (t.to(DEV) for t in fo.synthetic_clip(1, seed=SEED + 5))
    with torch.no_grad():
        m.eval()
        hm0 = m(kf, sup)[0].clone()
        m.train()
        m(kf, sup)                       # updates every running_mean / running_var
        m.eval()
        hm1 = m(kf, sup)[0].clone()
        fresh = fp.Alignment_V15(cfg, "validate")
        fresh.load_state_dict({k: v.detach().cpu() for k, v in m.state_dict().items()}, strict=True)
        fresh = fresh.to(DEV).eval()
        hm2 = fresh(kf, sup)[0]
    assert float((hm1 - hm0).abs().max()) > 1e-4           # the statistics did move
    assert float((hm1 - hm2).abs().max()) <= 1e-6          # and the cached affine followed them


def test_upsample_add_relu_backward_vs_torch():
    """fami_upsample_add_bwd (backward of nearest upsample + running sum + ReLU, hrnet.py:99-112,151-172)."""
    import fami_pose_b200 as fp
    from fami_pose_b200 import autograd as ag, ops
    fp.set_precision("fp32")
    g = torch.Generator().manual_seed(3)
    for up in (2, 4, 8):
        t = torch.randn(2, 48, 5, 4, generator=g, dtype=torch.float64, requires_grad=True)
        r = torch.randn(2, 48, 5 * up, 4 * up, generator=g, dtype=torch.float64, requires_grad=True)
        y = torch.relu(r + torch.nn.functional.interpolate(t, scale_factor=up, mode="nearest"))
        go = torch.randn(y.shape, generator=g, dtype=torch.float64)
        y.backward(go)
        td = ops.to_nhwc(t.detach().float().to(DEV), torch.float32).requires_grad_(True)
        rd = ops.to_nhwc(r.detach().float().to(DEV), torch.float32).requires_grad_(True)
        yd = ag.UpsampleAddReluFunction.apply(td, rd, up, True)
        assert float((ops.to_nchw(yd.detach()).cpu().double() - y.detach()).abs().max()) < 1e-6
        yd.backward(go.float().to(DEV))
        assert float((ops.to_nchw(td.grad).cpu().double() - t.grad).abs().max()) < 1e-5
        assert float((ops.to_nchw(rd.grad).cpu().double() - r.grad).abs().max()) < 1e-6


def _grad_check(m, gold, names, norm_tol, band):
    params = dict(m.named_parameters())
    rows, bad = [], []
    for n in names:
        g = params[n].grad
        assert g is not None, "no gradient for %s" % n
        got, ref, ref32 = _digest(g), gold["f64/grad/" + n], gold["f32/grad/" + n]
        if ref[0] < 1e-7:
            assert got[0] < 1e-4, "%s: expected a numerically zero gradient, norm %.3e" % (n, got[0])
            continue
        scale = max(np.abs(ref[2:]).max(), ref[0] / np.sqrt(max(g.numel(), 1)))
        err = float(np.abs(got[2:] - ref[2:]).max() / scale)
        nerr = abs(got[0] - ref[0]) / ref[0]
        err32 = float(np.abs(ref32[2:] - ref[2:]).max() / scale)
        rows.append((max(err, nerr), nerr, err, err32, ref[0], n))
        if nerr > norm_tol or err > max(band, 4 * err32):
            bad.append(rows[-1])
    rows.sort(reverse=True)
    for r in rows[:6]:
        print("grad dev %.3e (norm rel %.3e, samples %.3e; reference fp32 vs fp64 %.3e; |g| %.3e)  %s" % r)
    return bad


def _tf32_rna(t):
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def test_training_step_tf32_arm(golden_dir):
    """The configuration bench.py's `train` record times: everything on the 'tf32' arm (frozen backbone with train-mode
    BatchNorm; head forward convs and stride-1 dgrad on tcgen05 kind::tf32; fp32 wgrad / BatchNorm / DCN backward), B=2.

    Train-mode BatchNorm over tiny batches (B=2: 10 frames, down to 18 samples per channel in the global-offset head)
    through ~300 layers amplifies operand rounding ~100x compared with eval mode, chaotically: the CPU oracle itself, run
    with TF32-rounded conv multiplicands, moves its heatmaps by 5.7e-2 against its own fp32 run (measured; that is what
    the reference does on a GPU, where cuDNN's TF32 is PyTorch's default), and two TF32 runs that differ only in
    accumulation order land equally far apart (a value near a rounding boundary flips a whole TF32 ulp).  Heatmap parity
    beyond that scale is not defined for this arithmetic, so the forward is held to 1.5e-1 of BOTH the reference's float64
    golden and the oracle with TF32 multiplicands (heatmap magnitude ~3), while the LOSS -- an average, and what training
    consumes -- must match the float64 golden to 2e-3 relative (measured 6e-5).  Gradients: every trainable tensor OUTSIDE
    the global-offset head must keep its norm within 25 % and point the same way (cosine of the sampled digest >= 0.85);
    the offset head (feat_global_offset_layers: 3x3 maps x 2 clips = 18 samples per BatchNorm channel, feeding the warp
    translation that the oracle's own TF32 run moves by 0.1 px of 1.7) is the chaotic part -- its gradients are reported and
    only required to be finite and of the right order of magnitude (norm within 4x).  The TF32 kernels themselves are
    pinned per op (test_conv_tc_tf32, test_conv_dgrad_tf32_vs_autograd: 2e-5) and the exact-fp32 arm above carries the
    tight whole-model gradient pin."""
    import fami_pose_b200 as fp
    from fami_pose_b200.loss import JointMSELoss, combine_losses
    gold = np.load(os.path.join(golden_dir, "train_reference.npz"))
    cfg = rh.make_cfg(48, 17)
    fp.set_precision("tf32")
    try:
        m = fp.Alignment_V15(cfg, "train")
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        sd = fo.seeded_state_dict(shapes, SEED)
        m.load_state_dict(sd, strict=True)
        m = m.to(DEV).train()
        kf, sup, tgt, tw = fo.synthetic_clip(2, seed=SEED + 2)
        hm, kfhm, mi = m(kf.to(DEV), sup.to(DEV))
        mse = JointMSELoss()(hm, tgt.to(DEV), tw.to(DEV))
        loss = combine_losses(mse, mi)
        loss.backward()
        torch.cuda.synchronize()
    finally:
        fp.set_precision("fp32")
    with torch.no_grad():
        ohm, okf = fo.FunctionalFami(sd, bn_train=True, conv_hook=lambda x, w: (_tf32_rna(x), _tf32_rna(w))).alignment(kf, sup)
    e_model = float((hm.detach().cpu() - ohm).abs().max())
    e_kf = float((kfhm.detach().cpu() - okf).abs().max())
    e_gold = float(np.abs(hm.detach().cpu().numpy() - gold["final_hm"]).max())
    print("tf32 arm train fwd: vs oracle with TF32 multiplicands final %.3e kf %.3e; vs the fp64 golden %.3e; loss %.6f (ref %.6f)"
          % (e_model, e_kf, e_gold, float(loss), float(gold["f64/loss"])))
    assert e_model <= 1.5e-1 and e_kf <= 1.5e-1 and e_gold <= 1.5e-1
    assert abs(float(loss) - float(gold["f64/loss"])) <= 2e-3 * max(1.0, abs(float(gold["f64/loss"])))
    params = dict(m.named_parameters())
    worst_norm, worst_cos, n_checked = 0.0, 1.0, 0
    for n in [str(x) for x in gold["names"]]:
        got, ref = _digest(params[n].grad), gold["f64/grad/" + n]
        if ref[0] < 1e-7:
            continue
        nerr = abs(got[0] - ref[0]) / ref[0]
        cos = float(np.dot(got[2:], ref[2:]) / (np.linalg.norm(got[2:]) * np.linalg.norm(ref[2:]) + 1e-30))
        if n.startswith("feat_global_offset_layers."):
            print("tf32 gradient (offset head, reported): %s norm off by %.3f, cosine %.4f (|g| %.3e)" % (n, nerr, cos, ref[0]))
            assert np.isfinite(got).all() and 0.25 <= got[0] / ref[0] <= 4.0, n
            continue
        worst_norm, worst_cos, n_checked = max(worst_norm, nerr), min(worst_cos, cos), n_checked + 1
        if nerr > 0.25 or cos < 0.85:
            print("tf32 gradient outlier: %s norm off by %.3f, cosine %.4f (|g| %.3e)" % (n, nerr, cos, ref[0]))
    print("tf32 arm gradients: %d tensors, worst norm deviation %.3f, worst cosine %.4f" % (n_checked, worst_norm, worst_cos))
    assert worst_norm <= 0.25 and worst_cos >= 0.85


def test_training_step_unfrozen_backbone_gradients(golden_dir):
    """FREEZE_HRNET_WEIGHTS: false (Alignment_V15.py:110-111): backward through the whole HRNet-W48 -- fuse layers with
    nearest upsampling, stride-2 chains, Bottlenecks, the stem -- against the reference's float64 autograd for EVERY
    parameter that receives a gradient (exact-fp32 arm, B=1)."""
    import fami_pose_b200 as fp
    from fami_pose_b200.loss import JointMSELoss, combine_losses
    path = os.path.join(golden_dir, "train_unfrozen_reference.npz")
    gold = np.load(path)
    fp.set_precision("fp32")
    cfg = rh.make_cfg(48, 17, freeze_hrnet=False)
    m = fp.Alignment_V15(cfg, "train")
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(fo.seeded_state_dict(shapes, SEED), strict=True)
    m = m.to(DEV).train()
    assert all(p.requires_grad for p in m.parameters())
    kf, sup, tgt, tw = fo.synthetic_clip(1, seed=SEED + 3)
    hm, kfhm, mi = m(kf.to(DEV), sup.to(DEV))
    mse = JointMSELoss()(hm, tgt.to(DEV), tw.to(DEV))
    loss = combine_losses(mse, mi)
    loss.backward()
    torch.cuda.synchronize()
    print("unfrozen train: loss %.6f (ref %.6f)" % (float(loss), float(gold["f64/loss"])))
    assert abs(float(loss) - float(gold["f64/loss"])) <= 2e-3 * max(1.0, abs(float(gold["f64/loss"])))
    names = [str(n) for n in gold["names"]]
    got_names = [n for n, p in m.named_parameters() if p.grad is not None]
    assert sorted(got_names) == sorted(names)                 # the same set of parameters receives gradients
    assert len(names) > 900 and int(gold["n_params_total"]) >= len(names)
    # band: sampled entries within max(1e-1, 4x the reference's own fp32-vs-fp64 deviation, which reaches 3e-2 here) of the
    # tensor's largest sampled gradient, norms within 5e-2: sums over up to 110k pixels through ~300 layers
    bad = _grad_check(m, gold, names, 5e-2, 1e-1)
    assert not bad, "%d gradients deviate from the reference: %s" % (len(bad), ", ".join("%s (%.2e)" % (b[5], b[0]) for b in bad[:8]))


def test_train_step_cuda_graph_matches_eager_and_checkpoint_roundtrip(tmp_path):
    """(i) A CUDA-graph-captured TrainStep (forward + loss + backward + all-reduce + Adam, live lr / bias corrections)
    follows the eager steps: same losses, same gradients, same Adam moments; (ii) save_checkpoint / resume in the
    reference's format (engine/defaults/checkpoints.py:45-107: {begin_epoch, state_dict, optimizer: [Adam.state_dict()]})
    restores parameters AND Adam moments bit for bit; (iii) the optimizer entry loads into a real torch.optim.Adam built as
    posetimation/optimizer/optimizer.py:66-72 does.
    lr = 1e-6: Adam's first steps move every element by ~lr * sign(g), so elements whose gradient is rounding noise
    (fp32 atomics order) walk apart by +-lr per step between ANY two runs; a small lr keeps the two trajectories on the same
    parameters so that losses, gradients and moments can be compared tightly."""
    import fami_pose_b200 as fp
    from fami_pose_b200.train import TrainStep
    fp.set_precision("fp32")
    cfg = rh.make_cfg(48, 17)

    def build():
        m = fp.Alignment_V15(cfg, "train")
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        m.load_state_dict(fo.seeded_state_dict(shapes, SEED), strict=True)
        return m.to(DEV).train()

    def rel(a, b):
        return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))

    kf, sup, tgt, tw = (t.to(DEV) for t in fo.synthetic_clip(1, seed=SEED + 7))
    me, mg = build(), build()
    se, sg = TrainStep(me, lr=1e-6), TrainStep(mg, lr=1e-6)
    p0 = [p.detach().clone() for p in se.params]
    for _ in range(4):
        le, _ = se(kf, sup, tgt, tw)
    sg.capture(kf, sup, tgt, tw, warmup=2)          # two eager warm-up steps inside
    for _ in range(2):
        lg, _ = sg.replay()
    torch.cuda.synchronize()
    assert sg.step == se.step == 4
    g_rel = max(rel(a, b) for a, b in zip(sg.buckets.flat, se.buckets.flat))
    m_rel = max(rel(a, b) for a, b in zip(sg.exp_avg, se.exp_avg))
    v_rel = max(rel(a, b) for a, b in zip(sg.exp_avg_sq, se.exp_avg_sq))
    moved = max(float((p.detach() - q).abs().max()) for p, q in zip(sg.params, p0))
    print("graph vs eager after 4 steps: loss %.7f vs %.7f, grad rel %.2e, exp_avg rel %.2e, exp_avg_sq rel %.2e, max |dp| %.2e"
          % (float(lg), float(le), g_rel, m_rel, v_rel, moved))
    assert abs(float(lg) - float(le)) <= 1e-5 * max(1.0, abs(float(le)))
    assert g_rel <= 2e-3 and m_rel <= 2e-3 and v_rel <= 4e-3
    assert 3e-6 <= moved <= 4.5e-6                    # four Adam steps of ~lr each, with the live bias corrections
    # ---- checkpoint round trip
    path = se.save_checkpoint(3, str(tmp_path))
    assert os.path.basename(path) == "epoch_3_state.pth"
    ck = torch.load(path, map_location="cpu")
    assert sorted(ck.keys()) == ["begin_epoch", "optimizer", "state_dict"] and ck["begin_epoch"] == 3
    assert list(ck["state_dict"].keys()) == list(me.state_dict().keys())
    mr = build()
    sr = TrainStep(mr, lr=5e-4)
    assert sr.resume(path) == 4 and sr.step == 4 and sr.lr == 1e-6
    for a, b in zip(me.state_dict().values(), mr.state_dict().values()):
        assert torch.equal(a, b)
    for a, b in zip(se.exp_avg + se.exp_avg_sq, sr.exp_avg + sr.exp_avg_sq):
        assert torch.equal(a, b)
    l1, _ = se(kf, sup, tgt, tw)
    l2, _ = sr(kf, sup, tgt, tw)
    assert abs(float(l1) - float(l2)) <= 1e-6 * max(1.0, abs(float(l1)))
    # ---- the optimizer entry is a genuine torch.optim.Adam state_dict
    opt = torch.optim.Adam([p for p in build().parameters() if p.requires_grad], lr=1e-4)
    opt.load_state_dict(ck["optimizer"][0])
    assert len(opt.state) == len(se.params) and opt.param_groups[0]["lr"] == 1e-6
    st0 = opt.state[opt.param_groups[0]["params"][0]]
    assert float(st0["step"]) == 4.0 and st0["exp_avg"].shape == se.params[0].shape
