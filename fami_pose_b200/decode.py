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


def _as_f32_cuda(a, device):
    t = torch.as_tensor(a)
    return t.to(device=device, dtype=torch.float32).contiguous()


def get_final_preds(batch_heatmaps, center, scale):
    """datasets/process/heatmaps_process.py:47-73 on the device: argmax, +-0.25 px refinement, inverse affine
    back to image coordinates.  Same argument order and return contract as the reference
    ((preds [B,J,2], maxvals [B,J,1])) but takes / returns torch tensors on the GPU; center, scale: [B,2]
    (numpy arrays or tensors, scale in units of 200 px)."""
    if not isinstance(batch_heatmaps, torch.Tensor) or batch_heatmaps.dim() != 4:
        raise AssertionError('batch_heatmaps should be a 4-D torch.Tensor on the GPU')
    hm = batch_heatmaps if ops.is_nhwc(batch_heatmaps) else ops.to_nhwc(batch_heatmaps.float(), torch.float32)
    B, J, H, W, pitch = ops.meta(hm)
    idx, maxvals = ops.argmax_hw(hm)
    c = _as_f32_cuda(center, hm.device).reshape(B, 2)
    s = _as_f32_cuda(scale, hm.device).reshape(B, 2)
    preds = torch.empty((B, J, 2), dtype=torch.float32, device=hm.device)
    ops._lib.call("fami_final_preds", ops._ptr(hm), ops._code(hm.dtype), pitch, ops._ptr(idx), ops._ptr(maxvals), ops._ptr(c),
                  ops._ptr(s), ops._ptr(preds), B, H, W, J, ops._stream())
    return preds, maxvals.unsqueeze(2)


def accuracy(output, target, hm_type='gaussian', thr=0.5):
    """engine/core/utils/evaluate.py:39-75 without the host round trip the reference pays every iteration
    (alignment_mi_function_term6_1.py:159-174): returns (acc [J+1] float64, avg_acc, cnt, pred [B,J,2]) as
    DEVICE tensors (avg_acc, cnt 0-d)."""
    if hm_type != 'gaussian':
        raise NotImplementedError("only hm_type='gaussian' is used by FAMI-Pose")
    pidx, pmax = ops.argmax_hw(output)
    tidx, tmax = ops.argmax_hw(target)
    B, J, H, W = output.shape
    out = torch.empty(J + 3, dtype=torch.float64, device=output.device)
    ops._lib.call("fami_pck_accuracy", ops._ptr(pidx), ops._ptr(pmax), ops._ptr(tidx), ops._ptr(tmax), ops._ptr(out), B, H, W,
                  J, float(thr), ops._stream())
    pred = torch.stack([(pidx % W).float(), torch.div(pidx, W, rounding_mode="floor").float()], dim=2)
    pred = pred * (pmax > 0.0).unsqueeze(2).float()
    return out[:J + 1], out[J + 1], out[J + 2].to(torch.int64), pred


def generate_heatmaps(joints, joints_vis, sigma, image_size, heatmap_size, num_joints=None):
    """datasets/process/heatmaps_process.py:146-203 batched on the device: joints / joints_vis [B,J,3] ->
    (target [B,J,h,w] float32, target_weight [B,J,1]).  image_size / heatmap_size are (width, height)."""
    j = torch.as_tensor(joints)
    if j.dim() == 2:
        j = j.unsqueeze(0)
    v = torch.as_tensor(joints_vis).reshape(j.shape)
    if not j.is_cuda:
        j, v = j.cuda(), v.cuda()
    j = j.float().contiguous()
    v = v.float().contiguous()
    B, J = j.shape[:2]
    if num_joints is not None and num_joints != J:
        raise ValueError("num_joints does not match the joints array")
    if int(sigma) != sigma:
        raise NotImplementedError("integer sigma (Base_PoseTrack17.yaml: SIGMA 3)")
    iw, ih = int(image_size[0]), int(image_size[1])
    hw, hh = int(heatmap_size[0]), int(heatmap_size[1])
    target = torch.empty((B, J, hh, hw), dtype=torch.float32, device=j.device)
    weight = torch.empty((B, J), dtype=torch.float32, device=j.device)
    ops._lib.call("fami_gaussian_targets", ops._ptr(j), ops._ptr(v), ops._ptr(target), ops._ptr(weight), B, J, int(sigma), iw, ih,
                  hw, hh, ops._stream())
    return target, weight.unsqueeze(2)
