"""Per-call device timing of one eager forward step (CUDA events around every C-ABI call), aggregated
by entry point and shape.  usage: python tools/profile_calls.py [fp16|fp32|bf16] [batch]"""
import collections, ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fami_pose_b200 as fp
from fami_pose_b200 import _lib, ops
from oracle import fami_oracle as fo, ref_harness as rh

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
fp.set_precision(prec)
m = fp.Alignment_V15(rh.make_cfg(48, 17), "validate")
m.load_state_dict(fo.seeded_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}))
m = m.cuda().eval()
kf, sup, tgt, tw = fo.synthetic_clip(B)
kf, sup = kf.cuda(), sup.cuda()
with torch.no_grad():
    m(kf, sup); m(kf, sup)
torch.cuda.synchronize()
records = []
orig = _lib.call
def timed(name, *args):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    key = name
    if name == "fami_conv2d_bn_act_fwd":
        d = args[0]._obj
        key = "conv Cin=%d Cout=%d k=%d s=%d d=%d %dx%d N=%d up=%d res=%d" % (d.Cin, d.Cout, d.kh, d.stride, d.dil, d.H, d.W, d.N, d.up, int(args[5] is not None and args[5].value is not None))
    elif name == "fami_dcn_fwd":
        d = args[0]._obj
        key = "dcn C=%d G=%d %dx%d B=%d" % (d.C, d.G, d.H, d.W, d.B)
    e0.record(); orig(name, *args); e1.record()
    records.append((key, e0, e1))
_lib.call = timed
ops._lib.call = timed
with torch.no_grad():
    m(kf, sup)
torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for k, e0, e1 in records:
    agg[k][0] += 1; agg[k][1] += e0.elapsed_time(e1)
tot = sum(v[1] for v in agg.values())
print("total %.2f ms in %d calls (eager, event-timed; includes launch gaps)" % (tot, len(records)))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print("%8.3f ms %5.1f%% n=%3d avg %7.1f us  %s" % (t, 100 * t / tot, n, 1000 * t / n, k))
