"""Synthetic inputs and seeded weights for benchmarks and parity tests (pure torch CPU, no kernels).

SURVEY.md 8(c)/(d): the reference ships no pretrained weights and its own init gives |heatmap| ~ 1e-4, which would
make a 1e-3 parity check vacuous, so weights are re-initialised to O(1) output scale by a seeded, key-hashed recipe;
clips are N(0,1) frames (ImageNet-normalised images, datasets/transforms/build.py:13-22) with sigma = 3 gaussian
targets (datasets/process/heatmaps_process.py:146-203) and Bernoulli(0.85) target weights; the base seed is the
reference's own (tools/run.py:32-34).  make_cfg reproduces configs/Alignment/Base_PoseTrack17.yaml:45-87."""
import math

import torch

SEED = 19970808


class AttrDict(dict):
    """cfg.MODEL.EXTRA and cfg['MODEL']['EXTRA'] both work, as with the reference's yacs CfgNode."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


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


def _key_seed(key, seed):
    h = 1469598103934665603
    for ch in key.encode():
        h = ((h ^ ch) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return (h ^ (seed * 0x9E3779B97F4A7C15)) & 0x7FFFFFFFFFFFFFFF


def seeded_state_dict(shapes, seed=19970808, dtype=torch.float32):
    """shapes: {key: shape} with reference state_dict keys.  Deterministic per key (order
    independent).  Conv/Linear/DCN weights: N(0, gain^2/fan_in); biases N(0,0.05^2);
    BN gamma U(0.5,1.5), beta N(0,0.1^2), running_mean N(0,0.1^2), running_var U(0.5,1.5).
    dcn_offset_* convs are scaled so raw offsets are ~N(0, 1.5 px) and dcn_mask_* ~N(0.5,0.5)."""
    out = {}
    for key, shape in shapes.items():
        g = torch.Generator().manual_seed(_key_seed(key, seed))
        shape = tuple(shape)
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            out[key] = torch.zeros(shape, dtype=torch.long)
        elif leaf == "running_var":
            out[key] = (torch.rand(shape, generator=g) + 0.5).to(dtype)
        elif leaf == "running_mean":
            out[key] = (0.1 * torch.randn(shape, generator=g)).to(dtype)
        elif leaf == "weight" and len(shape) == 1:
            t = torch.rand(shape, generator=g) * 0.5 + 0.5
            if _is_block_last_bn(key):
                t = t * 0.4          # keep residual branches modest so 100+ stacked blocks stay O(1)
            out[key] = t.to(dtype)
        elif leaf == "bias":
            t = 0.1 * torch.randn(shape, generator=g) if _is_bn_bias(key, shapes) else 0.05 * torch.randn(shape, generator=g)
            if "dcn_mask_" in key:
                t = t + 0.5
            out[key] = t.to(dtype)
        elif leaf == "weight":
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            gain = 1.0
            if "dcn_offset_" in key:
                gain = 1.5
            elif "dcn_mask_" in key:
                gain = 0.5
            elif "feat_global_offset_layers.9" in key:
                gain = 3.0
            elif "hrnet.final_layer" in key or key.startswith("final_layer"):
                gain = 0.3   # rough heatmaps with |max| ~ 1
            elif len(shape) == 2 or "final_layer" in key or key.startswith("dcn_"):
                gain = 1.0
            out[key] = (gain / math.sqrt(fan_in) * torch.randn(shape, generator=g)).to(dtype)
        else:
            raise KeyError(key)
    return out


def _is_block_last_bn(key):
    """bn2 of a BasicBlock / bn3 of a Bottleneck / BN of a fuse or downsample path."""
    mod = key.rsplit(".", 1)[0]
    last = mod.rsplit(".", 1)[-1]
    if "layer1." in key:
        return last in ("bn3", "1") and "downsample" in key or last == "bn3"
    return last == "bn2" or "fuse_layers" in key or "downsample" in key


def _is_bn_bias(key, shapes):
    return (key.rsplit(".", 1)[0] + ".running_mean") in shapes


def synthetic_clip(B, H=384, W=288, num_sup=4, J=17, seed=19970808):
    """SURVEY.md 8(d) synthetic inputs: N(0,1) frames; sigma=3 gaussian targets at random joint
    centres (datasets/process/heatmaps_process.py:146-203); Bernoulli(0.85) target weights."""
    g = torch.Generator().manual_seed(seed)
    kf = torch.randn(B, 3, H, W, generator=g)
    sup = torch.randn(B, 3 * num_sup, H, W, generator=g)
    hh, ww = H // 4, W // 4
    cx = torch.rand(B, J, generator=g) * (ww - 1)
    cy = torch.rand(B, J, generator=g) * (hh - 1)
    ys = torch.arange(hh).view(1, 1, hh, 1).float()
    xs = torch.arange(ww).view(1, 1, 1, ww).float()
    tgt = torch.exp(-((xs - cx.round().view(B, J, 1, 1)) ** 2 + (ys - cy.round().view(B, J, 1, 1)) ** 2) / (2 * 3.0 ** 2))
    tw = (torch.rand(B, J, 1, generator=g) < 0.85).float()
    return kf, sup, tgt, tw
