"""ctypes binding of libfami_b200.so (the C ABI declared in include/fami_b200.h).

The library is built in-tree by fami_pose_b200/csrc/build.py (nvcc, sm_100a).  There is no CPU or
PyTorch fallback: if the shared object is missing, loading fails loudly, and every op raises if a
call returns non-zero.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfami_b200.so")

F32 = 0
BF16 = 1
F16 = 2
TF32 = 3   # conv / dcn descriptors only: float storage, tcgen05 kind::tf32 arithmetic


class ConvDesc(Structure):
    _fields_ = [(n, c_int32) for n in (
        "N", "H", "W", "Cin", "Cout", "kh", "kw", "stride", "pad", "dil", "Ho", "Wo", "up", "relu",
        "in_pitch", "out_pitch", "res_pitch", "dtype", "out_dtype", "stats", "om_groups", "om_layout")]


class DcnDesc(Structure):
    _fields_ = [(n, c_int32) for n in (
        "B", "H", "W", "C", "Cout", "G", "kh", "kw", "stride", "pad", "dil",
        "x_pitch", "off_pitch", "mask_pitch", "out_pitch", "om_layout", "dtype", "out_f32")]


# name -> (restype, argtypes); mirrors include/fami_b200.h exactly (tests check every symbol)
SIGNATURES = {
    "fami_last_error": (c_char_p, []),
    "fami_abi_version": (c_int, []),
    "fami_launch_count": (c_int64, []),
    "fami_workspace_bytes": (c_int64, [c_int, c_void_p]),
    "fami_nchw_to_nhwc": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "fami_nhwc_to_nchw": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "fami_conv_cout_pad": (c_int, [c_int]),
    "fami_packed_weight_elems": (c_int64, [c_int, c_int, c_int, c_int, c_int]),
    "fami_pack_conv_weight": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "fami_conv2d_bn_act_fwd": (c_int, [POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p]),
    "fami_conv2d_bn_act_fwd_stream": (c_int, [POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                              c_void_p, c_int, c_void_p]),
    "fami_bn_finalize": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_int, c_int64, c_float, c_float, c_void_p]),
    "fami_bn_apply_act": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int,
                                  c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "fami_bn_stats": (c_int, [c_void_p, c_int, c_int, c_int64, c_int, c_void_p, c_void_p]),
    "fami_dcn_fwd": (c_int, [POINTER(DcnDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fami_dcn_bwd": (c_int, [POINTER(DcnDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_void_p, c_void_p, c_void_p, c_void_p]),
    "fami_warp_translate_fwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                        c_int, c_void_p]),
    "fami_warp_translate_bwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                        c_int, c_int, c_int, c_int, c_void_p]),
    "fami_sub_bcast": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p]),
    "fami_copy2d": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int64, c_int, c_void_p]),
    "fami_linear_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "fami_joint_mse_fwd_bwd": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                       c_int, c_int, c_int, c_int, c_void_p]),
    "fami_softmax_pkl_fwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int,
                                     c_float, c_void_p]),
    "fami_pack_conv_weight_dgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "fami_conv2d_dgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fami_conv2d_wgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fami_bn_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int64,
                            c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "fami_softmax_pkl_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int,
                                     c_int, c_int, c_int, c_float, c_void_p]),
    "fami_linear_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                c_void_p]),
    "fami_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float,
                               c_int, c_void_p]),
    "fami_adam_step_graph": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_float, c_float, c_float,
                                     c_void_p]),
    "fami_upsample_add_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                      c_int, c_int, c_void_p]),
    "fami_final_preds": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                 c_int, c_int, c_void_p]),
    "fami_pck_accuracy": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float,
                                  c_void_p]),
    "fami_gaussian_targets": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                      c_int, c_void_p]),
    "fami_crop_affine_u8": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                    c_void_p]),
    "fami_frames_u8_normalize": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "fami_argmax_hw": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
}

# hardware probes / kernel timelines: exported only by libfami_b200_probes.so (csrc/build.py --probes), which the
# scripts under tools/ select with FAMI_PROBES=1 in the environment before importing the package
PROBE_SIGNATURES = {
    "fami_debug_read_trace": (c_int, [c_void_p, c_int]),
    "fami_debug_umma_rate": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p]),
    "fami_debug_umma_rowshift": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "fami_debug_tma_tf32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
}
PROBES = os.environ.get("FAMI_PROBES", "0") not in ("", "0")
if PROBES:
    LIB_PATH = os.path.join(_HERE, "libfami_b200_probes.so")

_lib = None


class FamiLibraryError(RuntimeError):
    pass


def load():
    """Loads the shared library (once).  Raises FamiLibraryError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FamiLibraryError(
            "libfami_b200.so not found at %s -- run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or python fami_pose_b200/csrc/build.py).  There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    sigs = dict(SIGNATURES)
    if PROBES:
        sigs.update(PROBE_SIGNATURES)
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.fami_abi_version() != 3:
        raise FamiLibraryError("libfami_b200.so ABI version mismatch")
    _lib = lib
    return lib


def call(name, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (name, rc, lib.fami_last_error().decode()))


def launch_count():
    return int(load().fami_launch_count())
