"""Hot spots of one kernel from an ncu report taken with --set full --import-source on:
    python tools/ncu_hotspots.py report.ncu-rep [units]
headline counters, executed warp instructions per opcode (per `units`, default 124416 = the warp-taps of the config-2
deformable launch), warp-state samples, the instructions with the most samples and the shared-memory wavefront totals."""
import csv, collections, re, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum","smsp__inst_executed.sum","sm__inst_executed.avg.per_cycle_elapsed","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","dram__bytes_read.sum","dram__bytes_write.sum","l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum","lts__t_sectors_srcunit_tex_op_read.sum","l1tex__t_sector_hit_rate.pct","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed","launch__registers_per_thread","lts__t_sector_hit_rate.pct","smsp__issue_active.avg.pct","sm__warps_active.avg.pct_of_peak_sustained_active"]
for i,h in enumerate(hdr):
    if h in want: print(h, units[i], vals[i])
src = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]; ix = {h:i for i,h in enumerate(hdr)}; data = rows[2:]
NWT = float(sys.argv[2]) if len(sys.argv) > 2 else 124416.0
tot_exec = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
tot_samp = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("total exec", tot_exec, "per warp-tap", tot_exec/NWT, "samples", tot_samp)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
st = {h: sum(int(r[ix[h]] or 0) for r in data) for h in stalls}
for k,v in sorted(st.items(), key=lambda kv:-kv[1])[:9]: print(k, v, "%.1f%%" % (100*v/tot_samp))
op = collections.Counter(); ops = collections.Counter()
for r in data:
    s = r[ix["Source"]].strip()
    m = re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_]+)", s)
    if not m: continue
    o = m.group(2)
    op[o] += int(r[ix["Instructions Executed"]] or 0); ops[o] += int(r[ix["# Samples"]] or 0)
for k,v in op.most_common(24): print("%-8s exec/warp-tap %6.1f (%.1f%%)  samples %.1f%%" % (k, v/NWT, 100*v/tot_exec, 100*ops[k]/tot_samp))
top = sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:16]
print("samples long_sb wait short_sb  source")
for r in top: print(r[ix["# Samples"]], r[ix["stall_long_sb"]], r[ix["stall_wait"]], r[ix["stall_short_sb"]], r[ix["Source"]][:90])
w = sum(int(r[ix["L1 Wavefronts Shared"]] or 0) for r in data); wi = sum(int(r[ix["L1 Wavefronts Shared Ideal"]] or 0) for r in data)
print("smem wavefronts", w, "ideal", wi)
for r in sorted(data, key=lambda r: -int(r[ix["L1 Wavefronts Shared"]] or 0))[:6]: print(r[ix["L1 Wavefronts Shared"]], r[ix["L1 Wavefronts Shared Ideal"]], r[ix["Source"]][:70])
