"""Times the training step (fami_pose_b200.train.TrainStep) at BASELINE config 2 shape.
usage: python tools/time_train.py [batch] [backbone_precision|none] [steps]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fami_pose_b200 as fp
from fami_pose_b200.train import TrainStep
from oracle import ref_harness as rh          # cfg helper only
from oracle import fami_oracle as fo          # seeded weights / synthetic clip (measurement tool, not product)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
bp = sys.argv[2] if len(sys.argv) > 2 else "fp16"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
fp.set_precision("fp32")
m = fp.Alignment_V15(rh.make_cfg(48, 17), "train")
shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
m.load_state_dict(fo.seeded_state_dict(shapes, 19970808), strict=True)
m = m.cuda().train()
if bp != "none":
    m.backbone_precision = bp
kf, sup, tgt, tw = (t.cuda() for t in fo.synthetic_clip(B, seed=1))
step = TrainStep(m)
losses = []
for _ in range(2):
    losses.append(step(kf, sup, tgt, tw)[0])
torch.cuda.synchronize()
l0 = fp._lib.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    losses.append(step(kf, sup, tgt, tw)[0])
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(json.dumps({"what": "train step (fwd + loss + bwd + adam), HRNet frozen", "batch": B, "backbone_precision": bp,
                  "ms_per_step": ms, "clips_per_s": B / ms * 1e3, "launches_per_step": (fp._lib.launch_count() - l0) / steps,
                  "loss": [float(l) for l in losses]}))
