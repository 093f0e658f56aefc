"""Max-abs heatmap error of every precision arm against the CPU oracle on one config-2 clip (B=1), plus the eager step time.
usage: python tools/arm_errors.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fami_pose_b200 as fp
from fami_pose_b200 import synth
from oracle import fami_oracle as fo

cfg = synth.make_cfg(48, 17)
m = fp.Alignment_V15(cfg, "validate")
sd = synth.seeded_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()})
m.load_state_dict(sd)
m = m.cuda().eval()
kf, sup, tgt, tw = synth.synthetic_clip(2)
with torch.no_grad():
    rhm, rkf = fo.FunctionalFami(sd).alignment(kf, sup)
for arm, stream in (("fp32", None), ("tf32", None), ("fp16", False), ("fp16", True), ("bf16", False), ("bf16", True)):
    fp.set_precision(arm, stream_f32=stream)
    with torch.no_grad():
        hm, kfhm = m(kf.cuda(), sup.cuda())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            hm, kfhm = m(kf.cuda(), sup.cuda())
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
    print("%-5s stream_f32=%-5s final %.2e  kf %.2e   eager B=2 step %.1f ms" % (
        arm, stream, float((hm.float().cpu() - rhm).abs().max()), float((kfhm.float().cpu() - rkf).abs().max()), dt * 1e3), flush=True)
fp.set_precision("fp32")
