"""fami_pose_b200 -- B200 (sm_100a) implementation of the FAMI-Pose forward/backward hot path.

Python mirror of the reference's operator interface (posetimation/layers, posetimation/backbones,
posetimation/zoo/Alignment, posetimation/loss, torchvision DeformConv2d, kornia warp_affine) over
the C ABI in include/fami_b200.h.  There is no CPU / PyTorch fallback: ops raise if the CUDA library
is missing or if they are given CPU tensors.
"""
from . import _lib, ops  # noqa: F401
from .backbones import HighResolutionModule, HRNet, HRNetPlus  # noqa: F401
from .decode import accuracy, argmax_indices, generate_heatmaps, get_final_preds, get_max_preds  # noqa: F401
from .layers import (BasicBlock, Bottleneck, ChainOfBasicBlocks, DeformConv2d, Interpolate,  # noqa: F401
                     conv_bn_relu)
from .loss import JointMSELoss, combine_losses  # noqa: F401
from .ops import get_precision, set_precision  # noqa: F401
from .zoo import Alignment_V15  # noqa: F401

__all__ = ["Alignment_V15", "HRNetPlus", "HRNet", "HighResolutionModule", "BasicBlock", "Bottleneck",
           "ChainOfBasicBlocks", "Interpolate", "conv_bn_relu", "DeformConv2d", "JointMSELoss",
           "combine_losses", "get_max_preds", "argmax_indices", "set_precision", "get_precision",
           "patch_reference"]


def patch_reference():
    """Rebinds the names the reference resolves at import time to the fami modules, so
    the reference's own model definitions / engine call the B200 kernels unchanged.  Call after the
    reference packages are importable and before models are constructed.  Returns the patched names."""
    import importlib
    import sys
    import types
    from . import kornia_shim
    patched = []

    def rebind(modname, **names):
        try:
            mod = importlib.import_module(modname)
        except Exception:
            return
        for k, v in names.items():
            setattr(mod, k, v)
            patched.append("%s.%s" % (modname, k))

    layer_names = dict(BasicBlock=BasicBlock, Bottleneck=Bottleneck, Interpolate=Interpolate,
                       ChainOfBasicBlocks=ChainOfBasicBlocks)
    rebind("posetimation.layers.basic_model", **layer_names)
    rebind("posetimation.layers.basic_layer", conv_bn_relu=conv_bn_relu)
    rebind("posetimation.layers", conv_bn_relu=conv_bn_relu, **layer_names)
    rebind("posetimation.backbones.hrnet", BasicBlock=BasicBlock, Bottleneck=Bottleneck, Interpolate=Interpolate,
           HighResolutionModule=HighResolutionModule, HRNetPlus=HRNetPlus, HRNet=HRNet,
           blocks_dict={'BASIC': BasicBlock, 'BOTTLENECK': Bottleneck})
    rebind("posetimation.loss.mse_loss", JointMSELoss=JointMSELoss)
    # kornia.geometry.warp_affine: the translation-only shim is bound to the name `kornia` INSIDE the Alignment_V15 module
    # only (its single call site, Alignment_V15.py:133-135, builds a pure translation) -- a real kornia install keeps its own
    # warp_affine for every other caller (rotation / scale augmentation code).  Without kornia installed a stub module is
    # registered so that the reference's `import kornia` succeeds; the stub has nothing but the shim.
    try:
        importlib.import_module("kornia")
    except Exception:
        k = types.ModuleType("kornia")
        k.geometry = types.ModuleType("kornia.geometry")
        k.geometry.warp_affine = kornia_shim.warp_affine
        sys.modules["kornia"] = k
        sys.modules["kornia.geometry"] = k.geometry
        patched.append("kornia (stub module)")
    proxy = types.SimpleNamespace(geometry=types.SimpleNamespace(warp_affine=kornia_shim.warp_affine))
    rebind("posetimation.zoo.Alignment.Alignment_V15", conv_bn_relu=conv_bn_relu,
           ChainOfBasicBlocks=ChainOfBasicBlocks, HRNetPlus=HRNetPlus, DeformConv2d=DeformConv2d, kornia=proxy)
    # engine plug-in registries (engine/defaults/constant.py:9-11; looked up by cfg.CORE_FUNCTION, engine/core/base.py:65,
    # and cfg.MODEL.NAME, posetimation/zoo/build.py:65): the entries are REPLACED under the reference's own names, so a
    # config naming AlignmentMIFunction_Term6_V1 / Alignment_V15 runs the fami versions with no config change
    try:
        const = importlib.import_module("engine.defaults.constant")
        from .train import AlignmentMIFunction_Term6_V1
        for reg_name, obj in (("CORE_FUNCTION_REGISTRY", AlignmentMIFunction_Term6_V1), ("MODEL_REGISTRY", Alignment_V15)):
            reg = getattr(const, reg_name, None)
            if reg is not None and hasattr(reg, "_obj_map"):
                reg._obj_map[obj.__name__] = obj
                patched.append("engine.defaults.constant.%s[%s]" % (reg_name, obj.__name__))
    except Exception:
        pass
    return patched
