"""Timing ablations of the warp-private DCN kernel (FAMI_DCN_ABLATE bits of csrc/dcn_wp.cu; results are wrong by construction,
only the time is read) and its A/B switches.  One process per setting (the knobs are read per launch)."""
import os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
for name, env in (("shipped kernel", {}), ("no epilogue stores", {"FAMI_DCN_ABLATE": "1"}), ("no far-sample path", {"FAMI_DCN_ABLATE": "8"}),
                  ("no offset loads (constants: regular sample positions, no far samples)", {"FAMI_DCN_ABLATE": "16"}),
                  ("segments of one tile (no ring reuse between tiles)", {"FAMI_DCN_WP_SEG": "1"}),
                  ("tcgen05 kernel (dcn_tc.cu, row-blocked offsets)", {"FAMI_DCN_WP": "0"})):
    e = dict(os.environ, BLOCKED="1", **env)
    out = subprocess.run([sys.executable, os.path.join(HERE, "time_dcn.py")], env=e, capture_output=True, text=True).stdout
    lines = [l for l in out.split("\n") if l.startswith("sigma")]
    print("%-75s %s" % (name, " | ".join(l.split(":")[1].split("->")[0].strip() for l in lines)), flush=True)
