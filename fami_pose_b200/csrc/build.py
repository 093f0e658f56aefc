"""Builds libfami_b200.so in-tree with nvcc for sm_100a (no torch dependency in the library).

    python fami_pose_b200/csrc/build.py [-v] [-f] [--probes]

--probes additionally builds libfami_b200_probes.so: the same sources with -DFAMI_DEBUG_PROBES, which adds the
fami_debug_* hardware probes / kernel timelines used by tools/ (never loaded by the package itself)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["api.cu", "conv_simt.cu", "conv_tc.cu", "conv_halo.cu", "dcn_tc.cu", "dcn_wp.cu", "misc.cu", "bwd.cu", "bwd_dense.cu",
           "decode.cu", "stem_tc.cu", "probes.cu"]
OUT = os.path.join(os.path.dirname(HERE), "libfami_b200.so")
OUT_PROBES = os.path.join(os.path.dirname(HERE), "libfami_b200_probes.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def _stale(obj, src):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    deps = [src, os.path.join(HERE, "common.cuh"), os.path.join(HERE, "..", "..", "include", "fami_b200.h"),
            os.path.abspath(__file__)]
    deps += [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(".cuh")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(verbose=False, force=False, probes=False):
    objdir = os.path.join(HERE, "build_probes" if probes else "build")
    out = OUT_PROBES if probes else OUT
    flags = FLAGS + (["-DFAMI_DEBUG_PROBES"] if probes else [])
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for s in SOURCES:
        src = os.path.join(HERE, s)
        obj = os.path.join(objdir, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, src):
            cmd = [NVCC] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        o, _ = p.communicate()
        if p.returncode != 0 or verbose:
            print("==== %s ====\n%s" % (s, o))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if procs or not os.path.exists(out):
        subprocess.check_call([NVCC, "-shared", "-o", out] + objs + ["-lcudart"])
    return out


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
    if "--probes" in sys.argv:
        print(build(verbose="-v" in sys.argv, force="-f" in sys.argv, probes=True))
