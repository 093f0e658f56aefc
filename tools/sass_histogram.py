"""Per-kernel counts of the SASS mnemonics that prove the tensor-core / TMA / TMEM path (cuobjdump -sass of the in-tree
library):  python tools/sass_histogram.py [libfami_b200.so] > profiles/r2_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                        "fami_pose_b200", "libfami_b200.so")
KEYS = ["UTCHMMA", "UTCMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKPF", "UBLKCP", "SYNCS", "LDGSTS",
        "LDS", "STS", "LDG", "STG", "RED", "HFMA2", "FFMA", "HMMA"]
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
filt = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
names = iter(filt)
counts = collections.OrderedDict()
cur = None
for line in out.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = re.sub(r"fami::\(anonymous namespace\)::|fami::", "", next(names))
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                counts[cur][k] += 1
                if k == "UTMALDG" and "IM2COL" in op:
                    counts[cur]["UTMALDG.IM2COL"] += 1
cols = ["_total", "UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMALDG.IM2COL", "UTMASTG", "UBLKPF", "SYNCS", "LDS", "STS", "LDG", "STG", "HFMA2", "FFMA"]
print("SASS mnemonic counts per kernel of %s (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor load, "
      "UTCBAR = tcgen05.commit, UBLKPF = cp.async.bulk.prefetch.L2, SYNCS = mbarrier ops)" % os.path.basename(so))
print(" ".join("%14s" % c for c in cols) + "  kernel")
tot = collections.Counter()
for k, c in counts.items():
    if c["UTCHMMA"] or c["UTMALDG"]:
        print(" ".join("%14d" % c[x] for x in cols) + "  " + k)
    tot.update(c)
print(" ".join("%14d" % tot[x] for x in cols) + "  TOTAL (all %d kernels)" % len(counts))
