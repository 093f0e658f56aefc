"""On-device keypoint decode: datasets/process/heatmaps_process.py:16-44 (get_max_preds)."""
import torch

from . import ops


def get_max_preds(batch_heatmaps):
    """Same contract as the reference (preds [B,J,2] float32 (x,y), maxvals [B,J,1]) but computed on
    the GPU from a CUDA tensor (NCHW float32 or a channels-last activation); returns torch tensors."""
    if not isinstance(batch_heatmaps, torch.Tensor) or batch_heatmaps.dim() != 4:
        raise AssertionError('batch_heatmaps should be a 4-D torch.Tensor on the GPU')
    W = batch_heatmaps.shape[3]
    idx, maxvals = ops.argmax_hw(batch_heatmaps)
    preds = torch.stack([(idx % W).float(), torch.div(idx, W, rounding_mode="floor").float()], dim=2)
    preds = preds * (maxvals > 0.0).unsqueeze(2).float()
    return preds, maxvals.unsqueeze(2)


def argmax_indices(batch_heatmaps):
    """Flat argmax indices [B,J] int32 (the bit-exact acceptance quantity)."""
    return ops.argmax_hw(batch_heatmaps)[0]
