"""Input pipeline on the device: the per-frame affine crop the reference's dataset performs on the CPU
with cv2 (datasets/zoo/posetrack/PoseTrack_Alignment.py:199-241, datasets/process/affine_transform.py:13-82), followed by
ToTensor + Normalize (datasets/transforms/build.py:13-22) and the window re-batching of Alignment_V15.py:115-119.

get_affine_transform is host arithmetic on six numbers per frame (float64, as in the reference); the pixel work --
cv2.warpAffine's fixed-point bilinear resampling, bit for bit -- runs in fami_crop_affine_u8."""
import ctypes

import numpy as np
import torch

from . import _lib, ops


def _third(a, b):
    d = a - b
    return b + np.array([-d[1], d[0]], np.float32)


def affine_from_points(src_pts, dst_pts):
    """cv2.getAffineTransform(src, dst): the 2x3 matrix mapping three points src -> dst (float64 solve)."""
    A = np.zeros((6, 6), np.float64)
    b = np.zeros(6, np.float64)
    for i in range(3):
        x, y = float(src_pts[i][0]), float(src_pts[i][1])
        A[2 * i] = [x, y, 1, 0, 0, 0]
        A[2 * i + 1] = [0, 0, 0, x, y, 1]
        b[2 * i], b[2 * i + 1] = float(dst_pts[i][0]), float(dst_pts[i][1])
    return np.linalg.solve(A, b).reshape(2, 3)


def get_affine_transform(center, scale, rot, output_size, shift=(0.0, 0.0), inv=0):
    """datasets/process/affine_transform.py:13-45 (same arguments, same float32 control points)."""
    if not isinstance(scale, (np.ndarray, list, tuple)):
        scale = np.array([scale, scale])
    scale_tmp = np.asarray(scale) * 200.0          # keeps the caller's dtype (float32 in the dataset), as the reference does
    src_w = scale_tmp[0]
    dst_w, dst_h = output_size[0], output_size[1]
    rot_rad = np.pi * rot / 180
    sn, cs = np.sin(rot_rad), np.cos(rot_rad)
    p = [0.0, src_w * -0.5]
    src_dir = [p[0] * cs - p[1] * sn, p[0] * sn + p[1] * cs]
    dst_dir = np.array([0, dst_w * -0.5], np.float32)
    src = np.zeros((3, 2), np.float32)
    dst = np.zeros((3, 2), np.float32)
    sh = scale_tmp * np.array(shift, dtype=np.float32)
    src[0, :] = np.asarray(center) + sh
    src[1, :] = np.asarray(center) + src_dir + sh
    dst[0, :] = [dst_w * 0.5, dst_h * 0.5]
    dst[1, :] = np.array([dst_w * 0.5, dst_h * 0.5]) + dst_dir
    src[2, :] = _third(src[0], src[1])
    dst[2, :] = _third(dst[0], dst[1])
    return affine_from_points(dst, src) if inv else affine_from_points(src, dst)


def invert_affine(trans):
    """cv::invertAffineTransform on [F,2,3] float64 matrices, operation for operation (cv2.warpAffine inverts the forward
    matrix it is given before resampling)."""
    m = np.array(trans, dtype=np.float64).reshape(-1, 6).copy()
    D = m[:, 0] * m[:, 4] - m[:, 1] * m[:, 3]
    D = np.where(D != 0, 1.0 / np.where(D != 0, D, 1.0), 0.0)
    A11, A22 = m[:, 4] * D, m[:, 0] * D
    m[:, 0] = A11
    m[:, 1] *= -D
    m[:, 3] *= -D
    m[:, 4] = A22
    b1 = -m[:, 0] * m[:, 2] - m[:, 1] * m[:, 5]
    b2 = -m[:, 3] * m[:, 2] - m[:, 4] * m[:, 5]
    m[:, 2], m[:, 5] = b1, b2
    return m


def crop_affine_u8(frames_u8, trans, output_size, normalize=False, mean=ops.IMAGENET_MEAN, std=ops.IMAGENET_STD, out=None):
    """cv2.warpAffine(frame, trans[f], output_size, flags=cv2.INTER_LINEAR) for every frame of a uint8 CUDA tensor
    [F, Hs, Ws, 3] (bit-exact).  trans: [F,2,3] (or [2,3] shared by all frames) forward matrices as get_affine_transform
    returns them.  normalize=False -> uint8 [F, H, W, 3]; True -> float32 NHWC frames [F, 3(logical), H, W] already
    ToTensor'ed + Normalize'd, ready for the backbone (what Alignment_V15.forward_u8 builds from crops)."""
    ops._need_cuda(frames_u8)
    if frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4 or frames_u8.shape[3] != 3:
        raise TypeError("crop_affine_u8 expects uint8 frames [F, Hs, Ws, 3]")
    frames_u8 = frames_u8.contiguous()
    F, Hs, Ws, _ = frames_u8.shape
    t = np.asarray(trans, np.float64)
    if t.shape == (2, 3):
        t = np.broadcast_to(t, (F, 2, 3))
    if t.shape != (F, 2, 3):
        raise ValueError("trans must be [F,2,3] or [2,3]")
    minv = torch.from_numpy(invert_affine(t)).to(frames_u8.device)
    Wd, Hd = int(output_size[0]), int(output_size[1])
    if normalize:
        if out is None:
            out = ops.empty_nhwc(F, 3, Hd, Wd, torch.float32, frames_u8.device)
        m = (ctypes.c_float * 3)(*mean)
        s = (ctypes.c_float * 3)(*std)
    else:
        if out is None:
            out = torch.empty((F, Hd, Wd, 3), dtype=torch.uint8, device=frames_u8.device)
        m = s = None
    _lib.call("fami_crop_affine_u8", ops._ptr(frames_u8), Hs * Ws * 3, Hs, Ws, ops._ptr(minv), ops._ptr(out), F, Hd, Wd, m, s,
              ops._stream())
    return out


def clip_from_frames(frames_u8, center, scale, rot=0.0, image_size=(288, 384)):
    """One clip's network input from its raw uint8 frames [1+ns, Hs, Ws, 3] (key frame first): the dataset's
    get_affine_transform(center, scale, rot, image_size) (PoseTrack_Alignment.py:233) applied to every frame of the window,
    then ToTensor + Normalize -- float32 NHWC frames [1+ns, 3, H, W] on the device."""
    trans = get_affine_transform(np.asarray(center, np.float64), np.asarray(scale, np.float64), rot, image_size)
    return crop_affine_u8(frames_u8, trans, image_size, normalize=True)
