"""SASS opcode mix + shared-memory wavefronts + stall-sample shares of the (single) kernel in an ncu report:
    python tools/ncu_opmix.py report.ncu-rep [top]
Reads `ncu -i report --page source --csv` (needs -lineinfo / --import-source on at capture time)."""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]
iS, iE, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
iW, iWi = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
ops, samp, wf, wfi = (collections.Counter() for _ in range(4))
tot = totS = 0
for r in rows[hi + 1:]:
    if len(r) <= iE:
        continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iS].strip())
    if not m:
        continue
    op = m.group(2)
    key = op if op.split(".")[0] in ("F2F", "F2I", "FRND", "I2F", "F2FP", "I2FP", "LDS", "STS", "LDG", "STG") else op.split(".")[0]
    n, s = int(r[iE] or 0), int(r[iSamp] or 0)
    ops[key] += n; samp[key] += s; tot += n; totS += s
    wf[key] += int(r[iW] or 0); wfi[key] += int(r[iWi] or 0)
print(rows[0][1] if len(rows[0]) > 1 else "")
print("warp instructions %d, stall samples %d" % (tot, totS))
for k, v in ops.most_common(top):
    print("%-26s %10d %5.1f%%  samples %5.1f%%%s" % (k, v, 100 * v / tot, 100 * samp[k] / max(totS, 1),
          ("  smem wavefronts %d (ideal %d, x%.2f)" % (wf[k], wfi[k], wf[k] / max(wfi[k], 1))) if wf[k] else ""))
