import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import fami_pose_b200 as fp
from fami_pose_b200 import synth
from oracle import fami_oracle as fo
cfg = synth.make_cfg(48, 17)
m = fp.Alignment_V15(cfg, "train")
sd = synth.seeded_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()})
m.load_state_dict(sd); m = m.cuda().train()
for B in (1, 2, 4):
    kf, sup, tgt, tw = synth.synthetic_clip(B, seed=7)
    with torch.no_grad():
        rhm, rkf, rmi = fo.FunctionalFami(sd, bn_train=True).alignment(kf, sup, with_mi=True)
    for arm in ("fp32", "tf32", "fp16", "bf16"):
        fp.set_precision(arm)
        m2 = fp.Alignment_V15(cfg, "train"); m2.load_state_dict(sd); m2 = m2.cuda().train()
        with torch.no_grad():
            hm, kfhm, mi = m2(kf.cuda(), sup.cuda())
        print("B=%d %-5s final %.2e kf %.2e |ref| max %.2f" % (B, arm, float((hm.float().cpu() - rhm).abs().max()), float((kfhm.float().cpu() - rkf).abs().max()), float(rhm.abs().max())), flush=True)
fp.set_precision("fp32")
