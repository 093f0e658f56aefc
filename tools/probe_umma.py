"""Probe: row-shifted SWIZZLE_128B UMMA descriptors (see fami_debug_umma_rowshift)."""
import os
os.environ["FAMI_PROBES"] = "1"   # fami_debug_* live in libfami_b200_probes.so (csrc/build.py --probes)
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fami_pose_b200 import _lib, ops
R = 160
g = torch.Generator().manual_seed(0)
x = torch.randn(R, 64, generator=g).half().cuda()
w = torch.randn(16, 64, generator=g).half().cuda()
for mode in (0, 1):
    for shift in (0, 1, 2, 3, 5, 7, 8, 9, 16, 19):
        out = torch.zeros(128, 16, device="cuda")
        _lib.call("fami_debug_umma_rowshift", ops._ptr(x), ops._ptr(w), ops._ptr(out), R, shift, mode, ops._stream())
        torch.cuda.synchronize()
        ref = x[shift:shift + 128].float() @ w.float().t()
        err = float((out - ref).abs().max())
        # which shift would the result correspond to, if any?
        best = min(range(0, R - 128), key=lambda s2: float((out - x[s2:s2 + 128].float() @ w.float().t()).abs().max()))
        print("mode %d shift %2d: max err %.4f  (best-matching shift %d)" % (mode, shift, err, best))
