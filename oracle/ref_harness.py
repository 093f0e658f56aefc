"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the *unmodified* reference model classes from /root/reference (read-only, exists only in
the build container, never on the GPU box) so that golden vectors can be generated from the
reference itself.  Follows SURVEY.md Appendix A:

  * posetimation/__init__.py (yacs) and posetimation/loss/__init__.py (missing integral_loss,
    reference posetimation/loss/base.py:11) are bypassed by pre-seeding sys.modules with bare
    package objects whose __path__ points at the reference directories;
  * engine / engine.defaults / engine.defaults.constant are faked (three Registry objects from the
    reference's own utils/utils_registry.py) so engine/__init__.py (tensorboardX, pycocotools) is
    not executed;
  * kornia is not installed: `kornia.geometry.warp_affine` is shimmed with kornia>=0.6 semantics
    (align_corners=True, bilinear, zeros padding) -- see oracle/fami_oracle.py::warp_affine_kornia.

Nothing here is copied from the reference; the reference code is executed where it lies.
"""
import os
import sys
import types

import torch

REF = os.environ.get("FAMI_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REF, "posetimation"))


class AttrDict(dict):
    """cfg stand-in: dict with attribute access (the model only reads ~12 keys)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def _ad(d):
    if isinstance(d, dict):
        return AttrDict({k: _ad(v) for k, v in d.items()})
    return d


def make_cfg(width=48, num_joints=17, freeze_hrnet=True):
    """Same keys/values as configs/Alignment/Base_PoseTrack17.yaml:45-87 (W32: 32/64/128/256)."""
    c = width
    return _ad({
        "MODEL": {
            "NUM_JOINTS": num_joints, "PRETRAINED": "", "BACKBONE_PRETRAINED": "",
            "FREEZE_HRNET_WEIGHTS": freeze_hrnet,
            "EXTRA": {
                "FINAL_CONV_KERNEL": 1,
                "PRETRAINED_LAYERS": ["*"],
                "STAGE2": {"NUM_MODULES": 1, "NUM_BRANCHES": 2, "BLOCK": "BASIC",
                           "NUM_BLOCKS": [4, 4], "NUM_CHANNELS": [c, 2 * c], "FUSE_METHOD": "SUM"},
                "STAGE3": {"NUM_MODULES": 4, "NUM_BRANCHES": 3, "BLOCK": "BASIC",
                           "NUM_BLOCKS": [4, 4, 4], "NUM_CHANNELS": [c, 2 * c, 4 * c], "FUSE_METHOD": "SUM"},
                "STAGE4": {"NUM_MODULES": 3, "NUM_BRANCHES": 4, "BLOCK": "BASIC",
                           "NUM_BLOCKS": [4, 4, 4, 4], "NUM_CHANNELS": [c, 2 * c, 4 * c, 8 * c],
                           "FUSE_METHOD": "SUM"},
            },
        },
    })


_loaded = {}


def load_reference():
    """Returns a namespace with the unmodified reference classes."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF)
    from . import fami_oracle

    if REF not in sys.path:
        sys.path.insert(0, REF)

    def ns(name, path=None, **kw):
        m = types.ModuleType(name)
        m.__dict__.update(kw)
        if path:
            m.__path__ = [path]
        sys.modules[name] = m
        return m

    # our own repo has no top-level `utils`/`posetimation`/`engine`, so these names are free
    ns("posetimation", REF + "/posetimation")
    ns("posetimation.loss", REF + "/posetimation/loss")
    from utils.utils_registry import Registry  # reference's own registry

    ns("engine")
    ns("engine.defaults", TRAIN_PHASE="train", VAL_PHASE="validate", TEST_PHASE="test")
    ns("engine.defaults.constant", MODEL_REGISTRY=Registry("MODEL"),
       CORE_FUNCTION_REGISTRY=Registry("CORE_FUNCTION"), DATASET_REGISTRY=Registry("DATASET"))
    k = ns("kornia")
    k.geometry = ns("kornia.geometry", warp_affine=fami_oracle.warp_affine_kornia)

    from posetimation.zoo.Alignment.Alignment_V15 import Alignment_V15
    from posetimation.backbones.hrnet import HRNet, HRNetPlus
    from posetimation.loss.mse_loss import JointMSELoss
    from posetimation.layers.basic_model import BasicBlock, Bottleneck, ChainOfBasicBlocks, Interpolate
    from posetimation.layers.basic_layer import conv_bn_relu

    _loaded.update(Alignment_V15=Alignment_V15, HRNet=HRNet, HRNetPlus=HRNetPlus,
                   JointMSELoss=JointMSELoss, BasicBlock=BasicBlock, Bottleneck=Bottleneck,
                   ChainOfBasicBlocks=ChainOfBasicBlocks, Interpolate=Interpolate,
                   conv_bn_relu=conv_bn_relu)
    return types.SimpleNamespace(**_loaded)


_decode = {}


def load_reference_decode():
    """The unmodified reference post-/pre-processing functions (numpy + cv2; build container only):
    datasets/process/heatmaps_process.py (get_max_preds, get_final_preds, generate_heatmaps) and
    engine/core/utils/evaluate.py (accuracy).  The packages' __init__ files pull in the whole dataset /
    engine stack, so the modules are loaded under bare package stubs whose __path__ points at the
    reference directories (same recipe as load_reference)."""
    if _decode:
        return types.SimpleNamespace(**_decode)
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF)
    import importlib
    import importlib.util

    def pkg(name, path):
        m = sys.modules.get(name)
        if m is None or not hasattr(m, "__path__"):
            m = types.ModuleType(name)
            m.__path__ = [path]
            sys.modules[name] = m
        return m

    pkg("datasets", REF + "/datasets")
    pkg("datasets.process", REF + "/datasets/process")
    hp = importlib.import_module("datasets.process.heatmaps_process")
    # evaluate.py does `from datasets.process.heatmaps_process import get_max_preds`; load it by path so the
    # faked `engine` package of load_reference() does not matter
    spec = importlib.util.spec_from_file_location("_fami_ref_evaluate", REF + "/engine/core/utils/evaluate.py")
    ev = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ev)
    _decode.update(get_max_preds=hp.get_max_preds, get_final_preds=hp.get_final_preds,
                   generate_heatmaps=hp.generate_heatmaps, accuracy=ev.accuracy)
    return types.SimpleNamespace(**_decode)
