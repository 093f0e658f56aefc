"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the FAMI-Pose hot path (the parity oracle).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product path (fami_pose_b200) never does and fails loudly without its CUDA
library.

Every function cites the reference file:line it restates (paths relative to /root/reference, or
the third-party call the reference makes).  Pinning status (see DESIGN.md "Oracle"):
the reference ships NO golden vectors/tests for this path, so the restatement
is pinned against outputs of the reference itself run in the build container
(tests/golden/make_golden.py -> tests/golden/*.npz) and against torchvision's CPU deform_conv2d
(the un-vendored dependency the reference calls, Alignment_V15.py:11,83,146).

Third-party arithmetic restated here:
  * torchvision.ops.deform_conv2d (reference does not pin a version; container has 0.26.0):
    modulated deformable convolution v2 -- dcn_fwd / dcn_bwd below.
  * kornia.geometry.warp_affine (unpinned, not installed): restated with kornia>=0.6 semantics
    (normalised homography, affine_grid/grid_sample align_corners=True, bilinear, zeros) --
    warp_affine_kornia; and its closed form for pure translations -- warp_translate.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------
# Modulated deformable convolution (torchvision.ops.deform_conv2d as called at
# posetimation/zoo/Alignment/Alignment_V15.py:83,146,150,154,158; spec: SURVEY.md Appendix B)
# --------------------------------------------------------------------------------------------


def _bilinear_parts(py, px, H, W):
    """Corner indices, weights and validity for sample positions (numpy arrays, any shape).

    torchvision semantics: the sample is 0 if py<=-1 or py>=H or px<=-1 or px>=W; each of the four
    corners contributes only when it lies inside [0,H-1]x[0,W-1].
    """
    inside = (py > -1) & (py < H) & (px > -1) & (px < W)
    y0 = np.floor(py)
    x0 = np.floor(px)
    ly = py - y0
    lx = px - x0
    hy = 1.0 - ly
    hx = 1.0 - lx
    y0 = y0.astype(np.int64)
    x0 = x0.astype(np.int64)
    y1 = y0 + 1
    x1 = x0 + 1
    v00 = inside & (y0 >= 0) & (x0 >= 0)
    v01 = inside & (y0 >= 0) & (x1 <= W - 1)
    v10 = inside & (y1 <= H - 1) & (x0 >= 0)
    v11 = inside & (y1 <= H - 1) & (x1 <= W - 1)
    return (y0, x0, y1, x1), (ly, lx, hy, hx), (v00, v01, v10, v11)


def _gather(xb, yy, xx, valid, H, W):
    """xb [C,H,W]; yy/xx/valid [...]; returns [C, ...] with zeros where invalid."""
    yc = np.clip(yy, 0, H - 1)
    xc = np.clip(xx, 0, W - 1)
    v = xb[:, yc, xc]
    return v * valid[None].astype(xb.dtype)


def dcn_columns(x, offset, mask, kh=3, kw=3, stride=1, pad=3, dil=3):
    """Sampled, modulated columns col[B, C, kh*kw, Ho, Wo] (numpy).  SURVEY.md Appendix B."""
    x = np.asarray(x)
    offset = np.asarray(offset)
    B, C, H, W = x.shape
    K = kh * kw
    G = offset.shape[1] // (2 * K)
    assert offset.shape[1] == 2 * K * G and C % G == 0
    cpg = C // G
    Ho = (H + 2 * pad - dil * (kh - 1) - 1) // stride + 1
    Wo = (W + 2 * pad - dil * (kw - 1) - 1) // stride + 1
    assert offset.shape[2:] == (Ho, Wo)
    col = np.zeros((B, C, K, Ho, Wo), dtype=x.dtype)
    ys = (np.arange(Ho) * stride - pad)[:, None].astype(x.dtype)
    xs = (np.arange(Wo) * stride - pad)[None, :].astype(x.dtype)
    for b in range(B):
        for g in range(G):
            xb = x[b, g * cpg:(g + 1) * cpg]
            for t in range(K):
                i, j = divmod(t, kw)
                py = ys + i * dil + offset[b, g * 2 * K + 2 * t]
                px = xs + j * dil + offset[b, g * 2 * K + 2 * t + 1]
                (y0, x0, y1, x1), (ly, lx, hy, hx), (v00, v01, v10, v11) = _bilinear_parts(py, px, H, W)
                val = (hy * hx)[None] * _gather(xb, y0, x0, v00, H, W) \
                    + (hy * lx)[None] * _gather(xb, y0, x1, v01, H, W) \
                    + (ly * hx)[None] * _gather(xb, y1, x0, v10, H, W) \
                    + (ly * lx)[None] * _gather(xb, y1, x1, v11, H, W)
                if mask is not None:
                    val = val * np.asarray(mask)[b, g * K + t][None]
                col[b, g * cpg:(g + 1) * cpg, t] = val
    return col


def dcn_fwd(x, offset, mask, weight, bias=None, stride=1, pad=3, dil=3):
    """out[b,o,y,x] = bias[o] + sum_{c,t} W[o,c,t] * col[b,c,t,y,x]   (weight groups = 1)."""
    weight = np.asarray(weight)
    Co, C, kh, kw = weight.shape
    col = dcn_columns(x, offset, mask, kh, kw, stride, pad, dil)
    out = np.einsum("oct,bctyx->boyx", weight.reshape(Co, C, kh * kw), col, optimize=True)
    if bias is not None:
        out = out + np.asarray(bias)[None, :, None, None]
    return out.astype(np.asarray(x).dtype)


def dcn_bwd(x, offset, mask, weight, grad_out, stride=1, pad=3, dil=3):
    """Analytic gradients (g_x, g_offset, g_mask, g_weight, g_bias); SURVEY.md Appendix B backward.

    Follows torchvision's deformable_col2im / col2im_coord definitions: floor() is treated as
    locally constant, out-of-range corners contribute 0 to both value and coordinate derivative.
    """
    x = np.asarray(x)
    offset = np.asarray(offset)
    mask = np.asarray(mask)
    weight = np.asarray(weight)
    grad_out = np.asarray(grad_out)
    B, C, H, W = x.shape
    Co, _, kh, kw = weight.shape
    K = kh * kw
    G = offset.shape[1] // (2 * K)
    cpg = C // G
    Ho, Wo = grad_out.shape[2:]
    g_col = np.einsum("oct,boyx->bctyx", weight.reshape(Co, C, K), grad_out, optimize=True)
    g_x = np.zeros_like(x)
    g_off = np.zeros_like(offset)
    g_mask = np.zeros_like(mask)
    ys = (np.arange(Ho) * stride - pad)[:, None].astype(x.dtype)
    xs = (np.arange(Wo) * stride - pad)[None, :].astype(x.dtype)
    col_unmod = np.zeros((B, C, K, Ho, Wo), dtype=x.dtype)
    for b in range(B):
        for g in range(G):
            cs = slice(g * cpg, (g + 1) * cpg)
            xb = x[b, cs]
            for t in range(K):
                i, j = divmod(t, kw)
                py = ys + i * dil + offset[b, g * 2 * K + 2 * t]
                px = xs + j * dil + offset[b, g * 2 * K + 2 * t + 1]
                (y0, x0, y1, x1), (ly, lx, hy, hx), (v00, v01, v10, v11) = _bilinear_parts(py, px, H, W)
                a00 = _gather(xb, y0, x0, v00, H, W)
                a01 = _gather(xb, y0, x1, v01, H, W)
                a10 = _gather(xb, y1, x0, v10, H, W)
                a11 = _gather(xb, y1, x1, v11, H, W)
                val = (hy * hx)[None] * a00 + (hy * lx)[None] * a01 + (ly * hx)[None] * a10 + (ly * lx)[None] * a11
                col_unmod[b, cs, t] = val
                m = mask[b, g * K + t]
                gc = g_col[b, cs, t]                           # [cpg,Ho,Wo]
                g_mask[b, g * K + t] = (gc * val).sum(0)
                dval_dy = hx[None] * (a10 - a00) + lx[None] * (a11 - a01)
                dval_dx = hy[None] * (a01 - a00) + ly[None] * (a11 - a10)
                g_off[b, g * 2 * K + 2 * t] = (gc * dval_dy).sum(0) * m
                g_off[b, g * 2 * K + 2 * t + 1] = (gc * dval_dx).sum(0) * m
                gm = gc * m[None]
                for (yy, xx, wgt, vv) in ((y0, x0, hy * hx, v00), (y0, x1, hy * lx, v01),
                                          (y1, x0, ly * hx, v10), (y1, x1, ly * lx, v11)):
                    yc = np.clip(yy, 0, H - 1)
                    xc = np.clip(xx, 0, W - 1)
                    contrib = gm * (wgt * vv)[None]
                    for c in range(cpg):
                        np.add.at(g_x[b, g * cpg + c], (yc, xc), contrib[c])
    col = col_unmod * np.repeat(mask.reshape(B, G, K, Ho, Wo), cpg, axis=1)
    g_w = np.einsum("boyx,bctyx->oct", grad_out, col, optimize=True).reshape(weight.shape)
    g_b = grad_out.sum((0, 2, 3))
    return g_x, g_off, g_mask, g_w, g_b


# --------------------------------------------------------------------------------------------
# Global translation warp (kornia.geometry.warp_affine as called at Alignment_V15.py:133-135)
# --------------------------------------------------------------------------------------------


def _normal_transform_pixel(h, w, dtype):
    """kornia.geometry.conversions.normal_transform_pixel: pixel -> [-1,1] (align_corners=True)."""
    n = torch.tensor([[1.0, 0.0, -1.0], [0.0, 1.0, -1.0], [0.0, 0.0, 1.0]], dtype=dtype)
    n[0, 0] = n[0, 0] * 2.0 / max(w - 1, 1e-14)
    n[1, 1] = n[1, 1] * 2.0 / max(h - 1, 1e-14)
    return n


def warp_affine_kornia(src, M, dsize, mode="bilinear", padding_mode="zeros", align_corners=True):
    """Restatement of kornia>=0.6 `warp_affine(src[B,C,H,W], M[B,2,3], dsize=(H,W))`.

    M3 = [M; 0 0 1]; theta = inverse(N_dst @ M3 @ inverse(N_src))[:, :2];
    grid = affine_grid(theta, align_corners=True); grid_sample(bilinear, zeros, align_corners=True).
    """
    B, C, H, W = src.shape
    M3 = torch.zeros(B, 3, 3, dtype=src.dtype, device=src.device)
    M3[:, :2] = M
    M3[:, 2, 2] = 1.0
    n_src = _normal_transform_pixel(H, W, src.dtype).to(src.device)
    n_dst = _normal_transform_pixel(dsize[0], dsize[1], src.dtype).to(src.device)
    dst_norm_trans_src_norm = n_dst @ M3 @ torch.inverse(n_src)
    src_norm_trans_dst_norm = torch.inverse(dst_norm_trans_src_norm)
    grid = F.affine_grid(src_norm_trans_dst_norm[:, :2, :], [B, C, dsize[0], dsize[1]],
                         align_corners=align_corners)
    return F.grid_sample(src, grid, mode=mode, padding_mode=padding_mode, align_corners=align_corners)


def warp_translate(src, txy):
    """Closed form of the above for M=[[1,0,tx],[0,1,ty]]: out[b,c,y,x] = bilinear_zeropad(src, y-ty, x-tx).

    grid_sample(zeros) semantics: each corner contributes iff inside the image (no (-1,H) cut-off
    subtlety: a corner outside contributes 0, which is the same thing).  numpy, float64/32.
    """
    src = np.asarray(src)
    txy = np.asarray(txy)
    B, C, H, W = src.shape
    out = np.zeros_like(src)
    for b in range(B):
        py = np.arange(H, dtype=src.dtype)[:, None] - txy[b, 1] + np.zeros((1, W), src.dtype)
        px = np.arange(W, dtype=src.dtype)[None, :] - txy[b, 0] + np.zeros((H, 1), src.dtype)
        y0 = np.floor(py)
        x0 = np.floor(px)
        ly, lx = py - y0, px - x0
        y0 = y0.astype(np.int64)
        x0 = x0.astype(np.int64)
        for (yy, xx, wgt) in ((y0, x0, (1 - ly) * (1 - lx)), (y0, x0 + 1, (1 - ly) * lx),
                              (y0 + 1, x0, ly * (1 - lx)), (y0 + 1, x0 + 1, ly * lx)):
            valid = (yy >= 0) & (yy <= H - 1) & (xx >= 0) & (xx <= W - 1)
            out[b] += _gather(src[b], yy, xx, valid, H, W) * wgt[None]
    return out


# --------------------------------------------------------------------------------------------
# Losses (posetimation/loss/mse_loss.py:21-40; Alignment_V15.py:250-277)
# --------------------------------------------------------------------------------------------


def joint_mse(output, target, target_weight, use_target_weight=True, divided_num_joints=True):
    """JointMSELoss.forward: sum_j mean_{b,p}((pred_j*w_bj - gt_j*w_bj)^2) [/ J]."""
    output = np.asarray(output, dtype=np.float64)
    target = np.asarray(target, dtype=np.float64)
    B, J = output.shape[:2]
    p = output.reshape(B, J, -1)
    t = target.reshape(B, J, -1)
    if use_target_weight:
        w = np.asarray(target_weight, dtype=np.float64).reshape(B, J, 1)
        p = p * w
        t = t * w
    loss = ((p - t) ** 2).mean(axis=(0, 2)).sum()
    if divided_num_joints:
        loss = loss / J
    return loss


def softmax_pkl(inp_rows, tgt_rows, temperature=0.05):
    """The reference's MI estimator core: kl_div(input=softmax(a/T), target=softmax(b/T), 'mean').

    QUIRK reproduced on purpose: F.kl_div expects log-probabilities as `input`
    but the reference passes probabilities, so the value is mean_{all elems}( t*(log t - p) ).
    Rows: [R, L]; softmax over L (dim=1).  float64 numpy.
    """
    a = np.asarray(inp_rows, dtype=np.float64) / temperature
    b = np.asarray(tgt_rows, dtype=np.float64) / temperature
    a = a - a.max(1, keepdims=True)
    b = b - b.max(1, keepdims=True)
    p = np.exp(a)
    p /= p.sum(1, keepdims=True)
    logt = b - np.log(np.exp(b).sum(1, keepdims=True))
    t = np.exp(logt)
    # torch xlogy semantics: t==0 -> 0
    term = np.where(t > 0, t * logt, 0.0) - t * p
    return term.mean()


def get_max_preds(batch_heatmaps):
    """datasets/process/heatmaps_process.py:16-44: flat argmax (first max wins) -> (x,y), maxvals;
    coordinates zeroed where maxval <= 0.  Returns (preds[B,J,2] float32, maxvals[B,J,1], idx[B,J])."""
    hm = np.asarray(batch_heatmaps)
    B, J, H, W = hm.shape
    r = hm.reshape(B, J, -1)
    idx = np.argmax(r, 2)
    maxvals = np.amax(r, 2).reshape(B, J, 1)
    preds = np.zeros((B, J, 2), np.float32)
    preds[:, :, 0] = idx % W
    preds[:, :, 1] = np.floor(idx / W)
    preds *= (maxvals > 0.0).astype(np.float32)
    return preds, maxvals, idx


def refine_quarter_pixel(batch_heatmaps, coords):
    """heatmaps_process.py:47-62 (the +-0.25 px shift toward the higher neighbour), before the
    inverse affine.  coords modified copy returned."""
    hm = np.asarray(batch_heatmaps)
    coords = np.array(coords, copy=True)
    B, J, H, W = hm.shape
    for n in range(B):
        for p in range(J):
            px = int(math.floor(coords[n][p][0] + 0.5))
            py = int(math.floor(coords[n][p][1] + 0.5))
            if 1 < px < W - 1 and 1 < py < H - 1:
                diff = np.array([hm[n][p][py][px + 1] - hm[n][p][py][px - 1],
                                 hm[n][p][py + 1][px] - hm[n][p][py - 1][px]])
                coords[n][p] += np.sign(diff) * .25
    return coords


def affine_from_points(src_pts, dst_pts):
    """cv2.getAffineTransform(src, dst) (called at datasets/process/affine_transform.py:40-43): the 2x3
    matrix mapping three float32 points src -> dst, solved in float64."""
    A = np.zeros((6, 6), np.float64)
    b = np.zeros(6, np.float64)
    for i in range(3):
        x, y = float(src_pts[i][0]), float(src_pts[i][1])
        A[2 * i] = [x, y, 1, 0, 0, 0]
        A[2 * i + 1] = [0, 0, 0, x, y, 1]
        b[2 * i], b[2 * i + 1] = float(dst_pts[i][0]), float(dst_pts[i][1])
    return np.linalg.solve(A, b).reshape(2, 3)


def get_affine_transform(center, scale, output_size, inv=0):
    """datasets/process/affine_transform.py:13-45 for rot = 0, shift = 0 (the only values the decode path
    passes, heatmaps_process.py:76): three float32 point pairs, then cv2.getAffineTransform."""
    scale_tmp = np.asarray(scale, np.float64) * 200.0
    src_w = scale_tmp[0]
    dst_w, dst_h = output_size[0], output_size[1]
    src = np.zeros((3, 2), np.float32)
    dst = np.zeros((3, 2), np.float32)
    src[0, :] = np.asarray(center)
    src[1, :] = np.asarray(center) + np.array([0.0, src_w * -0.5])
    dst[0, :] = [dst_w * 0.5, dst_h * 0.5]
    dst[1, :] = np.array([dst_w * 0.5, dst_h * 0.5]) + np.array([0, dst_w * -0.5], np.float32)

    def third(a, b):
        d = a - b
        return b + np.array([-d[1], d[0]], np.float32)
    src[2, :] = third(src[0], src[1])
    dst[2, :] = third(dst[0], dst[1])
    return affine_from_points(dst, src) if inv else affine_from_points(src, dst)


def get_affine_transform_rot(center, scale, rot, output_size, shift=(0.0, 0.0), inv=0):
    """datasets/process/affine_transform.py:13-45 with rotation and shift (the training-time augmentation of
    datasets/zoo/posetrack/PoseTrack_Alignment.py:213-233): three float32 point pairs, then cv2.getAffineTransform."""
    scale_tmp = np.asarray(scale) * 200.0          # dtype of the caller's scale (float32 in the dataset) is kept
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

    def third(a, b):
        d = a - b
        return b + np.array([-d[1], d[0]], np.float32)
    src[2, :] = third(src[0], src[1])
    dst[2, :] = third(dst[0], dst[1])
    return affine_from_points(dst, src) if inv else affine_from_points(src, dst)


def warp_affine_u8(img, M, dsize):
    """cv2.warpAffine(img_u8[H,W,C], M[2,3], (w, h), flags=cv2.INTER_LINEAR) with the default constant-0 border, as called
    at datasets/zoo/posetrack/PoseTrack_Alignment.py:235-241,417-423 and affine_transform.py:76-82 -- a bit-exact restatement
    of OpenCV's fixed-point path (pinned against cv2 4.13 in tests/golden/crop_reference.npz): the forward matrix is inverted
    in double; source coordinates are 10-bit fixed point (rounded products per column / row, + 16, >> 5 -> 5 fractional
    bits); the four bilinear weights are (32-fy|fy)*(32-fx|fx)*32 (sum 2^15, exact); result = (sum w*p + 2^14) >> 15."""
    Wd, Hd = int(dsize[0]), int(dsize[1])
    m = np.array(M, dtype=np.float64).reshape(6).copy()
    D = m[0] * m[4] - m[1] * m[3]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = m[4] * D, m[0] * D
    m[0] = A11; m[1] *= -D; m[3] *= -D; m[4] = A22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    xs = np.arange(Wd, dtype=np.float64)
    adelta = np.rint(m[0] * xs * 1024).astype(np.int64)
    bdelta = np.rint(m[3] * xs * 1024).astype(np.int64)
    img = np.asarray(img)
    Hs, Ws, C = img.shape
    out = np.zeros((Hd, Wd, C), np.uint8)
    for y in range(Hd):
        X0 = int(np.rint((m[1] * y + m[2]) * 1024)) + 16
        Y0 = int(np.rint((m[4] * y + m[5]) * 1024)) + 16
        X, Y = (X0 + adelta) >> 5, (Y0 + bdelta) >> 5
        ix, iy, fx, fy = X >> 5, Y >> 5, X & 31, Y & 31
        acc = np.zeros((Wd, C), np.int64)
        for dy, dx, w in ((0, 0, (32 - fy) * (32 - fx)), (0, 1, (32 - fy) * fx), (1, 0, fy * (32 - fx)), (1, 1, fy * fx)):
            yy, xx = iy + dy, ix + dx
            ok = (yy >= 0) & (yy < Hs) & (xx >= 0) & (xx < Ws)
            p = np.zeros((Wd, C), np.int64)
            p[ok] = img[yy[ok], xx[ok]]
            acc += (w * 32)[:, None] * p
        out[y] = ((acc + 16384) >> 15).astype(np.uint8)
    return out


def get_final_preds(batch_heatmaps, center, scale):
    """datasets/process/heatmaps_process.py:47-73: argmax, +-0.25 px refinement, inverse affine back to image
    coordinates (transform_preds :76-81).  Returns (preds [B,J,2] float32, maxvals [B,J,1])."""
    hm = np.asarray(batch_heatmaps)
    coords, maxvals, _ = get_max_preds(hm)
    coords = refine_quarter_pixel(hm, coords)
    B, J, H, W = hm.shape
    preds = coords.copy()
    for i in range(B):
        t = get_affine_transform(center[i], scale[i], [W, H], inv=1)
        for p in range(J):
            v = t @ np.array([coords[i, p, 0], coords[i, p, 1], 1.0])
            preds[i, p] = v[:2]
    return preds, maxvals


def accuracy(output, target, thr=0.5):
    """engine/core/utils/evaluate.py:13-75 (hm_type='gaussian'): PCK on heatmap argmaxes, distances
    normalised by (h, w)/10 -- in that order against (x, y), as the reference does -- ignoring joints whose
    target argmax has x <= 1 or y <= 1.  Returns (acc [J+1], avg_acc, cnt, pred)."""
    pred, _, _ = get_max_preds(output)
    tgt, _, _ = get_max_preds(target)
    B, J = pred.shape[:2]
    h, w = output.shape[2], output.shape[3]
    norm = np.ones((B, 2)) * np.array([h, w]) / 10
    dists = np.zeros((J, B))
    for n in range(B):
        for c in range(J):
            if tgt[n, c, 0] > 1 and tgt[n, c, 1] > 1:
                dists[c, n] = np.linalg.norm(pred[n, c, :] / norm[n] - tgt[n, c, :] / norm[n])
            else:
                dists[c, n] = -1
    acc = np.zeros(J + 1)
    avg, cnt = 0.0, 0
    for i in range(J):
        valid = dists[i] != -1
        acc[i + 1] = (dists[i][valid] < thr).sum() * 1.0 / valid.sum() if valid.sum() > 0 else -1
        if acc[i + 1] >= 0:
            avg += acc[i + 1]
            cnt += 1
    avg = avg / cnt if cnt != 0 else 0
    if cnt != 0:
        acc[0] = avg
    return acc, avg, cnt, pred


def generate_heatmaps(joints, joints_vis, sigma, image_size, heatmap_size, num_joints):
    """datasets/process/heatmaps_process.py:146-203: unnormalised gaussian (centre value 1) of radius 3*sigma
    pasted at round(joint / stride); weight 0 for joints whose patch misses the map entirely."""
    target_weight = np.ones((num_joints, 1), np.float32)
    target_weight[:, 0] = joints_vis[:, 0]
    target = np.zeros((num_joints, heatmap_size[1], heatmap_size[0]), np.float32)
    tmp = sigma * 3
    stride = np.asarray(image_size, np.float64) / np.asarray(heatmap_size, np.float64)
    size = 2 * tmp + 1
    ax = np.arange(0, size, 1, np.float32)
    g = np.exp(-((ax - size // 2) ** 2 + (ax[:, None] - size // 2) ** 2) / (2 * sigma ** 2))
    for j in range(num_joints):
        mu_x = int(joints[j][0] / stride[0] + 0.5)
        mu_y = int(joints[j][1] / stride[1] + 0.5)
        ul = [int(mu_x - tmp), int(mu_y - tmp)]
        br = [int(mu_x + tmp + 1), int(mu_y + tmp + 1)]
        if ul[0] >= heatmap_size[0] or ul[1] >= heatmap_size[1] or br[0] < 0 or br[1] < 0:
            target_weight[j] = 0
            continue
        gx = max(0, -ul[0]), min(br[0], heatmap_size[0]) - ul[0]
        gy = max(0, -ul[1]), min(br[1], heatmap_size[1]) - ul[1]
        ix = max(0, ul[0]), min(br[0], heatmap_size[0])
        iy = max(0, ul[1]), min(br[1], heatmap_size[1])
        if target_weight[j] > 0.5:
            target[j][iy[0]:iy[1], ix[0]:ix[1]] = g[gy[0]:gy[1], gx[0]:gx[1]]
    return target, target_weight


# --------------------------------------------------------------------------------------------
# Whole-model functional restatement over a state_dict (torch CPU fp32/fp64).
# Keys are the reference's state_dict keys (SURVEY.md section 5 "Checkpoint").
# --------------------------------------------------------------------------------------------


class FunctionalFami:
    """Functional forward of HRNetPlus / HRNet / Alignment_V15 over a reference-keyed state_dict.

    bn_train=False: BatchNorm uses running stats (module.eval()); True: batch statistics (biased
    variance, eps 1e-5), which is what the reference's default training does even with frozen
    HRNet weights (Alignment_V15.py:110-111, hrnet.py:686-690 only clear requires_grad).
    """

    def __init__(self, sd, width=48, num_joints=17, num_sup=4, offset_groups=12, bn_train=False,
                 conv_hook=None):
        self.sd = sd
        self.C = width
        self.J = num_joints
        self.num_sup = num_sup
        self.G = offset_groups
        self.bn_train = bn_train
        self.conv_hook = conv_hook  # optional fn(x, w) -> (x, w), e.g. to emulate tf32/bf16 operands

    # -- primitives ---------------------------------------------------------------------------
    def conv(self, x, name, stride=1, pad=0, dil=1):
        w = self.sd[name + ".weight"]
        b = self.sd.get(name + ".bias")
        if self.conv_hook is not None:
            x, w = self.conv_hook(x, w)
        return F.conv2d(x, w, b, stride=stride, padding=pad, dilation=dil)

    def bn(self, x, name):
        sd = self.sd
        if self.bn_train:
            return F.batch_norm(x, None, None, sd[name + ".weight"], sd[name + ".bias"], True, 0.1, 1e-5)
        return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"],
                            sd[name + ".weight"], sd[name + ".bias"], False, 0.1, 1e-5)

    def basic_block(self, x, p):
        """posetimation/layers/basic_model.py:44-63"""
        out = F.relu(self.bn(self.conv(x, p + "conv1", 1, 1), p + "bn1"))
        out = self.bn(self.conv(out, p + "conv2", 1, 1), p + "bn2")
        if (p + "downsample.0.weight") in self.sd:
            res = self.bn(self.conv(x, p + "downsample.0"), p + "downsample.1")
        else:
            res = x
        return F.relu(out + res)

    def bottleneck(self, x, p):
        """basic_model.py:83-113"""
        out = F.relu(self.bn(self.conv(x, p + "conv1"), p + "bn1"))
        out = F.relu(self.bn(self.conv(out, p + "conv2", 1, 1), p + "bn2"))
        out = self.bn(self.conv(out, p + "conv3"), p + "bn3")
        if (p + "downsample.0.weight") in self.sd:
            res = self.bn(self.conv(x, p + "downsample.0"), p + "downsample.1")
        else:
            res = x
        return F.relu(out + res)

    def chain(self, x, p, n):
        """ChainOfBasicBlocks, basic_model.py:128-148"""
        for i in range(n):
            x = self.basic_block(x, "%slayers.%d." % (p, i))
        return x

    def cbr(self, x, p, stride, pad, dil, has_bn=True, has_relu=True):
        """conv_bn_relu, basic_layer.py:13-73"""
        x = self.conv(x, p + "conv", stride, pad, dil)
        if has_bn:
            x = self.bn(x, p + "bn")
        if has_relu:
            x = F.relu(x)
        return x

    # -- HRNet --------------------------------------------------------------------------------
    def hr_module(self, xs, p, nb, multi_scale_output=True):
        """HighResolutionModule.forward, hrnet.py:151-172 (+ _make_fuse_layers :89-146)."""
        xs = [self._branch(xs[i], "%sbranches.%d." % (p, i)) for i in range(nb)]
        outs = []
        for i in range(nb if multi_scale_output else 1):
            y = None
            for j in range(nb):
                if j == i:
                    t = xs[j]
                elif j > i:
                    q = "%sfuse_layers.%d.%d." % (p, i, j)
                    t = self.bn(self.conv(xs[j], q + "0"), q + "1")
                    t = F.interpolate(t, scale_factor=2 ** (j - i), mode="nearest")
                else:
                    t = xs[j]
                    for k in range(i - j):
                        q = "%sfuse_layers.%d.%d.%d." % (p, i, j, k)
                        t = self.bn(self.conv(t, q + "0", 2, 1), q + "1")
                        if k != i - j - 1:
                            t = F.relu(t)
                y = t if y is None else y + t
            outs.append(F.relu(y))
        return outs

    def _branch(self, x, p):
        for b in range(4):
            x = self.basic_block(x, "%s%d." % (p, b))
        return x

    def hrnet_trunk(self, x, p):
        """HRNetPlus.forward, hrnet.py:651-680 (identical trunk in HRNet.forward :302-333)."""
        x = F.relu(self.bn(self.conv(x, p + "conv1", 2, 1), p + "bn1"))
        x = F.relu(self.bn(self.conv(x, p + "conv2", 2, 1), p + "bn2"))
        for i in range(4):
            x = self.bottleneck(x, "%slayer1.%d." % (p, i))
        xs = [F.relu(self.bn(self.conv(x, p + "transition1.0.0", 1, 1), p + "transition1.0.1")),
              F.relu(self.bn(self.conv(x, p + "transition1.1.0.0", 2, 1), p + "transition1.1.0.1"))]
        xs = self.hr_module(xs, p + "stage2.0.", 2)
        xs = xs + [F.relu(self.bn(self.conv(xs[-1], p + "transition2.2.0.0", 2, 1), p + "transition2.2.0.1"))]
        for m in range(4):
            xs = self.hr_module(xs, "%sstage3.%d." % (p, m), 3)
        xs = xs + [F.relu(self.bn(self.conv(xs[-1], p + "transition3.3.0.0", 2, 1), p + "transition3.3.0.1"))]
        for m in range(3):
            xs = self.hr_module(xs, "%sstage4.%d." % (p, m), 4, multi_scale_output=(m != 2))
        feat = xs[0]
        hm = self.conv(feat, p + "final_layer")
        return hm, feat

    # -- Alignment_V15 ------------------------------------------------------------------------
    def global_offset(self, d):
        """feat_global_offset_layers, Alignment_V15.py:61-72."""
        p = "feat_global_offset_layers."
        x = self.chain(d, p + "0.", 1)
        for i in range(1, 6):
            x = self.cbr(x, "%s%d." % (p, i), 2, 1, 1)
        x = x.flatten(1)
        for i in (7, 8, 9):
            x = F.linear(x, self.sd["%s%d.weight" % (p, i)], self.sd["%s%d.bias" % (p, i)])
        return x

    def dcn(self, x, off, msk, name):
        import torchvision
        return torchvision.ops.deform_conv2d(x, off, self.sd[name + ".weight"], self.sd[name + ".bias"],
                                             stride=1, padding=3, dilation=3, mask=msk)

    def mi_feat_label(self, feat, y):
        """Alignment_V15.py:250-263"""
        B = feat.shape[0]
        pred = self.conv(feat, "hrnet.final_layer").reshape(B * self.J, -1)
        yy = y.reshape(B * self.J, -1)
        return F.kl_div(input=F.softmax(pred.detach() / 0.05, dim=1), target=F.softmax(yy / 0.05, dim=1),
                        reduction="mean")

    def mi_feat_feat(self, f1, f2):
        """Alignment_V15.py:265-277"""
        B, C = f1.shape[:2]
        a = f1.reshape(B * C, -1)
        b = f2.reshape(B * C, -1)
        return F.kl_div(input=F.softmax(a.detach() / 0.05, dim=1), target=F.softmax(b / 0.05, dim=1),
                        reduction="mean")

    def alignment(self, kf_x, sup_x, with_mi=False, return_intermediates=False):
        """Alignment_V15.forward, Alignment_V15.py:113-183."""
        B = kf_x.shape[0]
        ns = sup_x.shape[1] // 3
        sup = torch.cat(torch.chunk(sup_x, ns, dim=1), dim=0)
        x = torch.cat([kf_x, sup], dim=0)
        hm_all, feat_all = self.hrnet_trunk(x, "hrnet.")
        kf_hm = hm_all[:B]
        feats = torch.chunk(feat_all, ns + 1, dim=0)
        kf_feat = feats[0]
        H, W = kf_feat.shape[2:]
        warped = []
        txys = []
        for i in range(ns):
            sf = feats[1 + i]
            txy = self.global_offset(sf - kf_feat)
            txys.append(txy)
            M = torch.eye(3, dtype=sf.dtype, device=sf.device)[0:2].view(1, 2, 3).repeat(B, 1, 1)
            M[:, 0, 2], M[:, 1, 2] = txy[:, 0], txy[:, 1]
            warped.append(warp_affine_kornia(sf, M, (H, W)))
        agg = self.chain(torch.cat(warped, dim=1), "sup_agg_block.", 2)
        comb = self.chain(torch.cat([agg, kf_feat], dim=1), "combined_feat_layers.", 1)
        inter = {"kf_feat": kf_feat, "txy": torch.stack(txys, 0), "agg_sup_feat": agg, "combined0": comb}
        cur = comb
        src = [None, None, agg, None]
        for k in range(1, 5):
            off = self.cbr(cur, "dcn_offset_%d." % k, 1, 3, 3, False, False)
            msk = self.cbr(cur, "dcn_mask_%d." % k, 1, 3, 3, False, False)
            inp = cur if src[k - 1] is None else src[k - 1]
            cur = self.dcn(inp, off, msk, "dcn_%d" % k)
            inter["dcn%d" % k] = cur
        allf = self.chain(torch.cat([kf_feat, cur], dim=1), "init_feature_agg_block.", 3)
        final = self.conv(allf, "agg_final_layer", 1, 1)
        inter["all_agg"] = allf
        out = [final, kf_hm]
        if with_mi:
            mi = [self.mi_feat_label(allf, final), self.mi_feat_feat(kf_feat, allf),
                  self.mi_feat_label(agg, final), self.mi_feat_feat(agg, allf),
                  self.mi_feat_label(kf_feat, final), self.mi_feat_feat(kf_feat, allf)]
            out.append(mi)
        if return_intermediates:
            out.append(inter)
        return tuple(out)


def combine_losses(mse, mi, w_mse=1.0, alpha=0.5, beta=0.1):
    """engine/core/functions/alignment_mi_function_term6_1.py:119-148:
    loss = MSE*w + alpha*( -beta*mi1 + beta*mi2 + mi3 - mi4 + mi5 - mi6 )."""
    return mse * w_mse + alpha * (-beta * mi[0] + beta * mi[1] + mi[2] - mi[3] + mi[4] - mi[5])


# --------------------------------------------------------------------------------------------
# Seeded O(1)-scale initialisation (SURVEY.md 8c caveat (i)): the reference's own init gives
# |heatmap| ~ 1e-4 which would make a 1e-3 parity check vacuous.
# --------------------------------------------------------------------------------------------


# The generators themselves (seeded_state_dict, synthetic_clip) are input/weight recipes shared by bench.py and the
# tests, not part of the checker: they live in fami_pose_b200/synth.py and are re-exported here for the tests.
from fami_pose_b200.synth import _is_block_last_bn, _is_bn_bias, _key_seed, seeded_state_dict, synthetic_clip  # noqa: E402,F401


def synthetic_decode_case(seed=19970808):
    """Seeded inputs of the decode / accuracy / target-generation pins (regenerated by the tests)."""
    rng = np.random.default_rng(seed)
    B, J, H, W = 4, 17, 96, 72
    hm = rng.standard_normal((B, J, H, W)).astype(np.float32)
    hm[0, 0] = -1.0                                   # no positive value: coordinates are zeroed
    hm[0, 1, 0, 0] = 6; hm[0, 2, H - 1, W - 1] = 6    # peaks on the border: no refinement
    hm[0, 3, 1, 1] = 6; hm[0, 4, 2, 2] = 6            # first positions outside / inside the refinement band
    hm[0, 5, H - 2, W - 2] = 6; hm[0, 6, H - 3, W - 3] = 6
    hm[1, 0, 40, 30] = 6; hm[1, 0, 40, 31] = 6        # tie: first maximum wins; equal neighbours -> sign 0
    tgt = np.zeros((B, J, H, W), np.float32)
    for b in range(B):
        for j in range(J):
            if (b + j) % 5 == 0:
                continue                              # all-zero target: argmax (0,0) -> ignored by accuracy
            y, x = rng.integers(0, H), rng.integers(0, W)
            tgt[b, j, y, x] = 1.0
            if (b * J + j) % 3 == 0:
                hm[b, j, y, min(x + int(rng.integers(0, 6)), W - 1)] = 7   # prediction close to the target
    center = rng.uniform(80, 400, (B, 2)).astype(np.float32)
    scale = rng.uniform(0.4, 2.6, (B, 2)).astype(np.float32)
    joints = rng.uniform(-60, 440, (B, J, 3)).astype(np.float32)
    joints[0, 0, :2] = [0, 0]; joints[0, 1, :2] = [287.9, 383.9]; joints[0, 2, :2] = [-11.9, 100]; joints[0, 3, :2] = [330, 100]
    vis = (rng.random((B, J, 3)) > 0.2).astype(np.float32)
    vis[..., 1] = vis[..., 0]; vis[..., 2] = 0
    return hm, tgt, center, scale, joints, vis
