"""Per-call device timing of one training step (CUDA events around every C-ABI call), aggregated by entry point.
usage: python tools/profile_train.py [batch] [backbone_precision|none] [head_precision]   (bench.py's train record: 32 tf32 tf32)"""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fami_pose_b200 as fp
from fami_pose_b200 import _lib, ops, autograd as ag, train as tr
from oracle import fami_oracle as fo, ref_harness as rh

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
bp = sys.argv[2] if len(sys.argv) > 2 else "tf32"
head = sys.argv[3] if len(sys.argv) > 3 else "tf32"
fp.set_precision(head)
m = fp.Alignment_V15(rh.make_cfg(48, 17), "train")
m.load_state_dict(fo.seeded_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 19970808))
m = m.cuda().train()
if bp != "none" and bp != head:
    m.backbone_precision = bp
kf, sup, tgt, tw = (t.cuda() for t in fo.synthetic_clip(B, seed=1))
step = tr.TrainStep(m)
step(kf, sup, tgt, tw); step(kf, sup, tgt, tw)
torch.cuda.synchronize()
records = []
orig = _lib.call
def timed(name, *args):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    key = name
    if name in ("fami_conv2d_bn_act_fwd", "fami_conv2d_dgrad", "fami_conv2d_wgrad"):
        d = args[0]._obj
        key = "%s Cin=%d Cout=%d k=%d s=%d d=%d %dx%d N=%d dt=%d" % (name[5:], d.Cin, d.Cout, d.kh, d.stride, d.dil, d.H, d.W, d.N, d.dtype)
    if name == "fami_bn_apply_act" and os.environ.get("BN_BY_SHAPE"):
        # (x, x_dtype, x_pitch, scale, shift, res, res_pitch, y, y_pitch, y_dtype, N, Ho, Wo, C, up, relu, stream)
        key = "bn_apply_act N=%d %dx%d C=%d res=%d" % (args[10], args[11], args[12], args[13], int(args[5] is not None and bool(args[5])))
    if name == "fami_bn_stats" and os.environ.get("BN_BY_SHAPE"):
        key = "bn_stats rows=%d C=%d" % (args[3], args[4])
    e0.record(); orig(name, *args); e1.record()
    records.append((key, e0, e1))
for mod in (_lib, ops, ag, tr):
    mod._lib.call = timed if mod is not _lib else None
_lib.call = timed
w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
w0.record(); step(kf, sup, tgt, tw); w1.record()
torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for k, e0, e1 in records:
    agg[k][0] += 1; agg[k][1] += e0.elapsed_time(e1)
tot = sum(v[1] for v in agg.values())
print("step %.1f ms wall on device; %.2f ms inside %d C-ABI calls (eager, event-timed)" % (w0.elapsed_time(w1), tot, len(records)))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(os.environ.get('TOPN', 40))]:
    print("%8.3f ms %5.1f%% n=%3d avg %7.1f us  %s" % (t, 100 * t / tot, n, 1000 * t / n, k))
