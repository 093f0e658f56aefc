"""Probe: what does the TMA unit do to fp32 data under a TFLOAT32 tensor map?  (fami_debug_tma_tf32)

Prints, for FLOAT32 / TFLOAT32 / TFLOAT32_FTZ maps, whether the shared-memory image equals the source bits, the
source truncated to 19 bits, round-to-nearest-even or round-to-nearest-ties-away (cvt.rna.tf32.f32)."""
import os
os.environ["FAMI_PROBES"] = "1"   # fami_debug_* live in libfami_b200_probes.so (csrc/build.py --probes)
import json
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fami_pose_b200 import _lib, ops

rows = 256
g = torch.Generator().manual_seed(0)
x = torch.randn(rows, 32, generator=g)
bits = x.view(torch.int32).clone()
bits[0, :8] = torch.tensor([0x3F801000, 0x3F800FFF, 0x3F801001, 0x3F803000, 0x3F802FFF, 0x00000001, 0x7F7FFFFF, 0x3F7FF000],
                           dtype=torch.int32)   # ties, just below / above, denormal, max, carry into the exponent
x = bits.view(torch.float32)
src = bits.numpy().view(np.uint32)
trunc = src & np.uint32(0xFFFFE000)
rna = (src + np.uint32(0x1000)) & np.uint32(0xFFFFE000)
rne = (src + np.uint32(0x0FFF) + ((src >> np.uint32(13)) & np.uint32(1))) & np.uint32(0xFFFFE000)
res = {}
for mode, name in ((0, "FLOAT32"), (1, "TFLOAT32"), (2, "TFLOAT32_FTZ")):
    out = torch.zeros(rows, 32, dtype=torch.int32, device="cuda")
    _lib.call("fami_debug_tma_tf32", ops._ptr(x.cuda()), ops._ptr(out), rows, mode, ops._stream())
    torch.cuda.synchronize()
    o = out.cpu().numpy().view(np.uint32)
    fin = np.isfinite(src.view(np.float32)) & (np.abs(src.view(np.float32)) > 1e-30) & (np.abs(src.view(np.float32)) < 1e30)
    res[name] = {"identity": bool((o == src).all()), "truncate": bool((o[fin] == trunc[fin]).all()),
                 "rna": bool((o[fin] == rna[fin]).all()), "rne": bool((o[fin] == rne[fin]).all()),
                 "special_first8_hex": ["%08x" % v for v in o[0, :8]]}
print(json.dumps(res, indent=1))
