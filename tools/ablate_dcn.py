"""Timing ablations of the tensor-core DCN kernel (FAMI_DCN_ABLATE bits; results are wrong by construction, only the time is
read): which part of the tile period each mechanism accounts for.  One process per setting (the knob is read per launch)."""
import os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
for name, bits in (("baseline", 0), ("no epilogue stores", 1), ("no window reload", 2), ("no corner loads / blends", 4), ("no far path", 8),
                   ("no offset loads", 16), ("no window reload + no far", 10), ("no loads/blends + no window + no far", 14),
                   ("only barriers + MMA (everything off)", 31)):
    env = dict(os.environ, FAMI_DCN_ABLATE=str(bits), BLOCKED="1", FAMI_DCN_WP="0")   # the tcgen05 kernel (csrc/dcn_tc.cu)
    out = subprocess.run([sys.executable, os.path.join(HERE, "time_dcn.py")], env=env, capture_output=True, text=True).stdout
    lines = [l for l in out.split("\n") if l.startswith("sigma")]
    print("%-42s %s" % (name, " | ".join(l.split(":")[1].split("->")[0].strip() for l in lines)), flush=True)
