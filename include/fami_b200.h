/*
 * fami_b200.h -- C ABI of libfami_b200.so: the B200 (sm_100a) implementation of the FAMI-Pose
 * forward/backward hot path.
 *
 * This is the drop-in boundary.  The reference has no FFI of its own: its hot path
 * is Python nn.Modules that bottom out in ATen / cuDNN / torchvision's _C.so.  Each entry point
 * below names the reference call (file:line, relative to the reference tree) whose arithmetic it
 * replaces; the Python shim in fami_pose_b200/ binds these symbols with ctypes and re-exposes them
 * as nn.Module / autograd.Function objects with the reference's own names and signatures
 * (INTEGRATION.md shows the binding a reference maintainer would add).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; fami_last_error() gives the text
 *     (thread-local).  No entry point allocates device memory or synchronises the host.
 *   - pointers are DEVICE pointers unless the name ends in _host.  `stream` is a cudaStream_t
 *     passed as void* (0 = legacy default stream).
 *   - activations are NHWC ("channels-last": element (n,y,x,c) at ((n*H+y)*W+x)*pitch + c) where
 *     `pitch` >= C is the per-pixel element pitch, so channel slices of a wider buffer (the
 *     reference's torch.cat along dim=1) can be read/written in place.
 *   - dtype: FAMI_F32 activations are float (exact-fp32 SIMT arm); FAMI_F16 / FAMI_BF16 activations
 *     are __half / __nv_bfloat16 (tcgen05 tensor-core arm; "half" below means either 16-bit type).
 *     FAMI_TF32 (convolution / deformable descriptors only) = float STORAGE of x, residual and y with
 *     the contraction on tcgen05.mma.kind::tf32: multiplicands rounded to TF32 (weights at pack time,
 *     activations by the TMA load), fp32 accumulation -- the arithmetic cuDNN applies to the
 *     reference's fp32 nn.Conv2d on a GPU (torch.backends.cudnn.allow_tf32 defaults to True).
 *     Per-channel scale/shift vectors, biases, loss outputs are always float.
 */
#ifndef FAMI_B200_H_
#define FAMI_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FAMI_ABI_VERSION 3

enum { FAMI_F32 = 0, FAMI_BF16 = 1, FAMI_F16 = 2, FAMI_TF32 = 3 };

/* Caller-provided scratch.  No entry point allocates device memory: where an operation needs scratch beyond its
 * inputs and outputs, the caller passes it (stats_out / sums / scratch_oihw / w_packed_t / grad_w_packed arguments
 * below) and sizes it with fami_workspace_bytes(op, desc), desc = the operation's fami_conv_desc / fami_dcn_desc.
 *   FAMI_OP_CONV_FWD    : 16*Cout if desc->stats (double stats_out[2*Cout], zeroed by the caller), else 0
 *   FAMI_OP_CONV_DGRAD  : scratch_oihw (4*Cout*Cin*kh*kw) + w_packed_t (fami_packed_weight_elems(Cin,Cout,..) floats)
 *   FAMI_OP_CONV_WGRAD  : 0 (accumulates into grad_w_oihw / grad_bias, which the caller zeroes)
 *   FAMI_OP_BN_BWD      : 16*Cout (double sums[2*C], zeroed by the caller)
 *   FAMI_OP_DCN_FWD     : 0
 *   FAMI_OP_DCN_BWD     : grad_w_packed, 4 * kh*kw * C * fami_conv_cout_pad(Cout)
 * Returns -1 for an unknown op or a null descriptor.                                                           */
enum { FAMI_OP_CONV_FWD = 0, FAMI_OP_CONV_DGRAD = 1, FAMI_OP_CONV_WGRAD = 2, FAMI_OP_BN_BWD = 3, FAMI_OP_DCN_FWD = 4,
       FAMI_OP_DCN_BWD = 5 };
int64_t fami_workspace_bytes(int op, const void* desc);

/* error text of the last failing call on this thread ("" if none) */
const char* fami_last_error(void);
int fami_abi_version(void);
/* number of kernel launches issued through this library since load (bench.py gpu_launches) */
int64_t fami_launch_count(void);

/* ---- layout -------------------------------------------------------------------------------
 * The reference keeps NCHW fp32 tensors (SURVEY.md 8b "Tensors").  These convert at the boundary.
 * nchw_to_nhwc also performs the frame re-batching of Alignment_V15.py:117-119 when called per
 * frame slice (src_n_stride lets sup_x[B,12,H,W] be read as 4 frame-major [B,3,H,W] slabs). */
int fami_nchw_to_nhwc(const float* src, int64_t src_n_stride, void* dst, int dst_dtype, int N, int C,
                      int H, int W, int dst_pitch, void* stream);
int fami_nhwc_to_nchw(const void* src, int src_dtype, int src_pitch, float* dst, int N, int C, int H,
                      int W, void* stream);

/* ---- convolution + folded BN + residual + activation (+ nearest upsample on write) ---------
 * Replaces nn.Conv2d -> nn.BatchNorm2d(eval) -> [+residual] -> nn.ReLU chains:
 *   BasicBlock.forward   posetimation/layers/basic_model.py:44-63
 *   Bottleneck.forward   posetimation/layers/basic_model.py:83-113
 *   conv_bn_relu.forward posetimation/layers/basic_layer.py:55-73
 *   HighResolutionModule fuse layers posetimation/backbones/hrnet.py:89-146,151-172
 *   (conv1x1+BN+Interpolate(nearest x2^k) is `up` = 2^k; the running sum is `residual`)
 *   HRNetPlus stem/transition/final_layer posetimation/backbones/hrnet.py:651-680
 *   y[n, yo*up+dy, xo*up+dx, o] = act( scale[o]*conv(x)[n,yo,xo,o] + shift[o] + residual[same] )
 * Weights are pre-packed by fami_pack_conv_weight to [kh*kw][Cin][CoutPad] (CoutPad = Cout rounded
 * up to 16), float for the F32 path, bf16 UMMA-tiled for the BF16 tensor-core path.              */
typedef struct fami_conv_desc {
  int32_t N, H, W, Cin;          /* input  [N,H,W,Cin], pitch in_pitch          */
  int32_t Cout, kh, kw;          /* square kernels 1 or 3                        */
  int32_t stride, pad, dil;
  int32_t Ho, Wo;                /* conv output size (before `up`)               */
  int32_t up;                    /* 1,2,4,8: nearest-neighbour replication on write */
  int32_t relu;                  /* 0/1                                          */
  int32_t in_pitch, out_pitch, res_pitch;
  int32_t dtype;                 /* FAMI_F32 or FAMI_BF16 (x, residual, w)       */
  int32_t out_dtype;             /* dtype of y (and residual): equal to dtype; or FAMI_F32 with dtype
                                    FAMI_BF16 (offset/mask and heatmap convs keep fp32 outputs, no
                                    residual); or FAMI_BF16 with dtype FAMI_F32 (stem conv only)   */
  int32_t stats;                 /* 1: also accumulate per-channel sum / sum-of-squares of the RAW
                                    (post scale/shift, pre residual/act) output into the double
                                    array stats_out[2*Cout], which the caller zeroes (train-mode BN) */
  int32_t om_groups;             /* 0: y is an NHWC activation (out_pitch).  G > 0: y is the row-blocked offset|mask
                                    buffer of the tensor-core deformable kernel (fami_dcn_desc.om_layout = 2) for G
                                    offset groups: Cout = 27*G channels in tap-major order, out_dtype FAMI_F32, up = 1,
                                    no residual, 3x3 stride-1 "same" convolution (the fused dcn_offset_k | dcn_mask_k
                                    producer, Alignment_V15.py:144-145)                                              */
  int32_t om_layout;             /* with om_groups > 0: 0 or 2 = row-blocked (fami_dcn_desc.om_layout 2); 3 = k-step-blocked
                                    (fami_dcn_desc.om_layout 3, the warp-private deformable kernel's)                  */
} fami_conv_desc;

int fami_conv_cout_pad(int Cout);
int64_t fami_packed_weight_elems(int Cout, int Cin, int kh, int kw, int dtype);
int fami_pack_conv_weight(const float* w_oihw, void* w_packed, int Cout, int Cin, int kh, int kw,
                          int dtype, void* stream);
int fami_conv2d_bn_act_fwd(const fami_conv_desc* d, const void* x, const void* w_packed,
                           const float* scale, const float* shift, const void* residual, void* y,
                           double* stats_out, void* stream);

/* fp32 residual stream for the 16-bit arms (dtype = out_dtype = FAMI_F16 / FAMI_BF16): the same fused convolution, but
 * the residual operand is FLOAT (residual_f32, pitch d->res_pitch, may be NULL) and the result is additionally stored,
 * before rounding, into the float tensor y_f32 (pitch y32_pitch, may be NULL).  With every block output kept this way the
 * ~100 sequential residual additions of the HRNet trunk (basic_model.py:60,110; hrnet.py:151-172) accumulate in fp32 while
 * all tensor-core operands stay 16-bit: bf16 then meets north_star's 1e-2 (fp16: 1e-3) -- see DESIGN.md section 3.      */
int fami_conv2d_bn_act_fwd_stream(const fami_conv_desc* d, const void* x, const void* w_packed, const float* scale,
                                  const float* shift, const float* residual_f32, void* y, float* y_f32, int y32_pitch,
                                  void* stream);

/* ---- backward of the dense pieces (fp32 storage; autograd of the modules above) --------------
 * nn.Conv2d backward as autograd derives it for basic_model.py:44-63 / basic_layer.py:55-73 /
 * Alignment_V15.py:79-106.  `d` is the FORWARD descriptor (up = 1, dtype = out_dtype = FAMI_F32).
 * dgrad consumes the filter flipped and transposed (fami_pack_conv_weight_dgrad: scratch holds
 * Cout*Cin*kh*kw floats; the packing has fami_packed_weight_elems(Cin, Cout, kh, kw, dtype)
 * elements); stride-1 dgrad runs on the forward kernels -- d->dtype = FAMI_F32: exact SIMT,
 * FAMI_TF32: tcgen05 kind::tf32 with a FAMI_TF32 packing -- other strides on an fp32 gather kernel.
 * wgrad ACCUMULATES into grad_w_oihw [Cout][Cin][kh][kw] and grad_bias [Cout] (may be NULL):
 * the caller zeroes them (fp32 atomics over pixel chunks).  d->dtype = FAMI_F32: exact fp32 FMAs;
 * FAMI_TF32: the products on mma.sync TF32 tensor cores (operands rounded to nearest TF32 on their
 * way into shared memory, fp32 accumulation).                                                        */
int fami_pack_conv_weight_dgrad(const float* w_oihw, float* scratch_oihw, void* w_packed_t, int Cout, int Cin,
                                int kh, int kw, int dtype, void* stream);
int fami_conv2d_dgrad(const fami_conv_desc* d, const float* grad_y, const float* w_packed_t, float* grad_x,
                      void* stream);
int fami_conv2d_wgrad(const fami_conv_desc* d, const float* x, const float* grad_y, float* grad_w_oihw,
                      float* grad_bias, void* stream);
/* nn.BatchNorm2d backward fused with the ReLU mask of the block epilogue (basic_model.py:44-63):
 * g' = grad_y * [y > 0] (y = the block's post-activation output, NULL = no ReLU);
 * training = 1: batch statistics (mean / invstd saved by fami_bn_finalize), the full three-term
 * formula; training = 0: running statistics, grad_x = gamma * invstd * g'.
 * grad_gamma = sum g' * xhat, grad_beta = sum g' (may be NULL); grad_res (may be NULL) receives g',
 * the gradient of the residual operand.  sums: double[2*C] workspace, zeroed by the caller.       */
int fami_bn_bwd(const float* x, int x_pitch, const float* grad_y, int gy_pitch, const float* y, int y_pitch,
                const float* mean, const float* invstd, const float* gamma, int64_t rows, int C, int training,
                double* sums, float* grad_x, int gx_pitch, float* grad_res, int gres_pitch, float* grad_gamma,
                float* grad_beta, void* stream);

/* train-mode BatchNorm (batch statistics over N*H*W, biased variance, eps, momentum update):
 * nn.BatchNorm2d(momentum=0.1) as used at basic_model.py:29,39 and hrnet.py:53,106,124,137.
 * fami_bn_finalize turns (sum, sumsq) into scale/shift and updates running stats;
 * fami_bn_apply_act is y = act(scale*x + shift + residual) with the same `up` write semantics; scale == shift == null
 * is a plain fp32 -> 16-bit storage cast (no residual, up = 1, no relu).  */
int fami_bn_finalize(const double* stats, const float* gamma, const float* beta, float* running_mean,
                     float* running_var, float* scale, float* shift, float* save_mean,
                     float* save_invstd, int C, int64_t count, float eps, float momentum, void* stream);
int fami_bn_apply_act(const void* x, int x_dtype, int x_pitch, const float* scale, const float* shift,
                      const void* residual, int res_pitch, void* y, int y_pitch, int dtype, int N, int Ho,
                      int Wo, int C, int up, int relu, void* stream);
/* per-channel (sum, sum of squares) of an NHWC activation [rows, C] accumulated into the double array
 * stats[2*C] (caller zeroes): batch statistics for the tensor-core conv path, whose epilogue does not
 * fuse them.                                                                                      */
int fami_bn_stats(const void* x, int dtype, int pitch, int64_t rows, int C, double* stats, void* stream);

/* ---- modulated deformable convolution v2 (the north-star kernel) ---------------------------
 * Replaces torchvision.ops.deform_conv2d as constructed/called at
 * posetimation/zoo/Alignment/Alignment_V15.py:83,89,95,101 / :146,150,154,158
 * (DeformConv2d(C,Cout,3,padding=3,dilation=3), offset groups G = offset_channels/18, mask raw).
 * x [B,H,W,C]; offset [B,H,W,18G] (channel g*18+2t = dy, +1 = dx); mask [B,H,W,9G];
 * w_packed: fami_pack_conv_weight(dtype FAMI_F32) for fp32 x, or (half dtype) for the 16-bit kernel;
 * out [B,H,W,Cout].  Fused gather -> on-chip columns -> contraction; the [C*9, B*H*W] im2col buffer
 * of the reference never exists in HBM.                                                          */
typedef struct fami_dcn_desc {
  int32_t B, H, W, C, Cout, G;
  int32_t kh, kw, stride, pad, dil; /* 3,3,1,3,3 in the reference; stride must be 1 */
  int32_t x_pitch, off_pitch, mask_pitch, out_pitch;
  int32_t om_layout;                /* 0: torchvision layout -- `offset` [.,18G] (channel g*18+2t = dy, +1 = dx)
                                       and `mask` [.,9G] (channel g*9+t) are separate operands;
                                       1: fused tap-major -- `offset` points at ONE NHWC buffer holding, per pixel,
                                       [9 taps][dy(G) | dx(G) | mask(G)] (off_pitch >= 27G), `mask` is ignored;
                                       2: fused row-blocked -- `offset` points at ONE dense buffer
                                       [9 taps][image tile][row 16][dy | dx | mask][pixel 8][group G] over 16x8-pixel
                                       tiles (tile = (b * ceil(H/16) + y/16) * ceil(W/8) + x/8, row = y%16, pixel = x%8);
                                       pitches and `mask` ignored.
                                       Layouts 1 - 3 are the 16-bit tensor-core kernels' (1: both; 2: the tcgen05 kernel,
                                       csrc/dcn_tc.cu; 3: the warp-private kernel, csrc/dcn_wp.cu); the alignment head's
                                       fused offset|mask convolution writes layout 2 or 3 (fami_conv_desc.om_groups /
                                       om_layout).  Layout 2: a gather warp of the tcgen05 kernel owns one tile row and walks its 8*G (pixel, group)
                                       samples of a tap 32 at a time, every load instruction of the warp reads 128
                                       contiguous bytes and the warp's reads of a (row, tap) are one contiguous run of
                                       24*G floats.
                                       3: fused k-step-blocked -- ONE dense buffer
                                       [image tile][9 taps][row 16][dy | dx | mask][group / 4][pixel 8][group % 4] over the
                                       same 16x8-pixel tiles (tile-major: the 9 * 128 * 3G floats of a tile are one
                                       contiguous run, a strip of tiles is a handful of sequential DRAM streams instead
                                       of nine per tile): the warp-private kernel (C == Cout in {32, 48})
                                       maps lane (pixel, group % 4) of an mma.sync fragment to one sample per k-step
                                       (= group / 4), so each of its load instructions reads one contiguous 128-byte
                                       line.  Layout 3 is accepted by that kernel only.  */
  int32_t dtype;                    /* storage of x/out: FAMI_F32 or FAMI_F16 / FAMI_BF16; offset, mask, packed
                                       weights and bias are always float (sub-pixel precision)        */
  int32_t out_f32;                  /* 1 (16-bit dtype, layouts 1 - 3 only): `out` is float while x stays 16-bit -- the
                                       tf32 arm's deformable convolutions: x is cast to fp16 (same 11-bit significand as
                                       TF32), the contraction runs on 16-bit tensor-core MMAs with fp32 accumulation, the result
                                       returns to the fp32 activation stream.  0: out has the storage type of x.       */
} fami_dcn_desc;

int fami_dcn_fwd(const fami_dcn_desc* d, const void* x, const void* offset, const void* mask,
                 const void* w_packed, const float* bias, void* out, void* stream);
/* backward (fp32 storage, om_layout 0): grad wrt input (atomic scatter), offset+mask, weight+bias.
 * torchvision: deformable_col2im / deformable_col2im_coord + GEMMs.  Gradient buffers
 * are DENSE (grad_x [B,H,W,C], grad_offset [B,H,W,18G], grad_mask [B,H,W,9G], grad_w_packed in the
 * fp32 packed layout [9*C][CoutPad], grad_bias [Cout] or NULL) and are zero-filled by the call.
 * d->dtype = FAMI_F32: exact fp32 throughout; FAMI_TF32: the weight gradient's products (sampled columns x
 * grad_out over 64-pixel chunks) run on mma.sync TF32 with fp32 accumulation, everything else stays exact.  */
int fami_dcn_bwd(const fami_dcn_desc* d, const float* x, const float* offset, const float* mask,
                 const float* w_packed, const float* grad_out, float* grad_x, float* grad_offset,
                 float* grad_mask, float* grad_w_packed, float* grad_bias, void* stream);

/* ---- global translation warp ---------------------------------------------------------------
 * Replaces kornia.geometry.warp_affine(src, [[1,0,tx],[0,1,ty]], dsize=(H,W)) at
 * Alignment_V15.py:133-135: out[b,y,x,c] = bilinear_zero_pad(src[b], y - ty_b, x - tx_b).
 * txy [B,2] = (tx, ty).  out may be a channel slice of the 4-frame concat buffer (:139).        */
int fami_warp_translate_fwd(const void* src, int src_pitch, const float* txy, void* out, int out_pitch,
                            int dtype, int B, int H, int W, int C, void* stream);
/* grad_src [B,H,W,C] dense (gs_pitch == C) and grad_txy [B,2] are zero-filled by the call; either may be NULL. */
int fami_warp_translate_bwd(const float* src, int src_pitch, const float* txy, const float* grad_out,
                            int go_pitch, float* grad_src, int gs_pitch, float* grad_txy, int B, int H,
                            int W, int C, void* stream);

/* ---- small dense pieces --------------------------------------------------------------------
 * a - b (Alignment_V15.py:132 `sup_bb_feat - kf_bb_feat`), b broadcast over `rep` groups of the
 * leading dimension: out[r*n + i] = a[r*n + i] - b[i].                                          */
int fami_sub_bcast(const void* a, const void* b, void* out, int dtype, int64_t n, int rep, void* stream);
/* dst[r, 0:cols] = src[r, 0:cols] for r < rows with independent row pitches: writes one operand of
 * the reference's torch.cat(dim=1) (Alignment_V15.py:143,160) into its channel slice.           */
int fami_copy2d(const void* src, int src_pitch, void* dst, int dst_pitch, int dtype, int64_t rows, int cols,
                void* stream);
/* nn.Linear chain of feat_global_offset_layers[7..9] (Alignment_V15.py:69-71): y = x W^T + b.
 * x [M,K] float, w [N,K] float (torch layout), y [M,N].                                         */
int fami_linear_fwd(const float* x, const float* w, const float* b, float* y, int M, int K, int N,
                    void* stream);
/* backward of the same nn.Linear: grad_x [M,K] = grad_y W, grad_w [N,K] = grad_y^T x, grad_b [N] = column
 * sums of grad_y; any output may be NULL.                                                          */
int fami_linear_bwd(const float* x, const float* w, const float* grad_y, float* grad_x, float* grad_w,
                    float* grad_b, int M, int K, int N, void* stream);
/* Backward of the HRNet fuse-layer tail `y = ReLU(residual + nearest_up(t))` (Interpolate, basic_model.py:116-125, inside
 * hrnet.py:99-112,151-172): g' = grad_y * [y > 0] (y NULL = no ReLU); grad_small [N,Ho,Wo,C] = sum of g' over the up x up
 * replicas; grad_res (may be NULL) = g'.  grad_y / y / grad_res are [N, Ho*up, Wo*up, C] (pitches in elements).            */
int fami_upsample_add_bwd(const float* grad_y, int gy_pitch, const float* y, int y_pitch, float* grad_small, int gs_pitch,
                          float* grad_res, int gr_pitch, int N, int Ho, int Wo, int C, int up, void* stream);
/* fami_adam_step for a CUDA-graph-captured training step: hyper_dev = device float[3] {lr, 1 - beta1^t, sqrt(1 - beta2^t)},
 * refreshed by the host before every replay (MultiStepLR, scheduler.py:14-26, and the step count stay live).               */
int fami_adam_step_graph(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                         const float* hyper_dev, float beta1, float beta2, float eps, void* stream);
/* One Adam update over a flat fp32 parameter bucket (torch.optim.Adam as built by
 * posetimation/optimizer/optimizer.py:66-72: betas, eps, no weight decay / amsgrad); `step` counts from 1. */
int fami_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                   float beta1, float beta2, float eps, int step, void* stream);

/* ---- losses --------------------------------------------------------------------------------
 * JointMSELoss.forward, posetimation/loss/mse_loss.py:21-40 (use_target_weight, divided by J):
 * loss = 1/(J*B*HW) * sum w_bj^2 (pred-gt)^2 ; pred NHWC [B,H,W,J] (pitch), target NCHW fp32
 * [B,J,H,W] as the reference's loader produces it, weight [B,J].  grad_pred (NHWC float, may be
 * NULL) = d loss / d pred * grad_scale.  loss_out is a single float accumulated with atomics:
 * caller zeroes it.                                                                             */
int fami_joint_mse_fwd_bwd(const void* pred, int pred_dtype, int pred_pitch, const float* target_nchw,
                           const float* weight, float* loss_out, float* grad_pred, float grad_scale,
                           int B, int J, int H, int W, void* stream);
/* MI estimator core, Alignment_V15.py:250-277: kl_div(input=softmax(a/T), target=softmax(b/T),
 * reduction='mean') with the reference's quirk (probabilities passed as log-probs), i.e.
 * mean_{r,l}( t*(log t - p) ), rows r = (sample, channel), l over H*W.  a, b NHWC [B,HW,C].
 * out: single float accumulated with atomics (caller zeroes).                                   */
int fami_softmax_pkl_fwd(const void* a, int a_pitch, const void* b, int b_pitch, int dtype, float* out,
                         int B, int HW, int C, float temperature, void* stream);
/* gradient of that scalar with respect to a and b (fp32 NHWC, either output may be NULL), scaled by
 * the device scalar *grad_out: d/da_j = -(1/T) p_j (t_j - sum t p), d/db_j = (1/T) t_j ((log t_j - p_j) - L_row),
 * both divided by B*C*HW.                                                                          */
int fami_softmax_pkl_bwd(const float* a, int a_pitch, const float* b, int b_pitch, const float* grad_out,
                         float* grad_a, int ga_pitch, float* grad_b, int gb_pitch, int B, int HW, int C,
                         float temperature, void* stream);

/* ---- keypoint argmax -----------------------------------------------------------------------
 * get_max_preds, datasets/process/heatmaps_process.py:16-44: flat argmax over H*W per (b,j),
 * first maximum wins; idx_out [B,J] int32, maxval_out [B,J] float.  hm NHWC.                    */
int fami_argmax_hw(const void* hm, int dtype, int pitch, int32_t* idx_out, float* maxval_out, int B,
                   int HW, int J, void* stream);
/* get_final_preds, datasets/process/heatmaps_process.py:47-81: from the flat argmax (idx, maxvals of
 * fami_argmax_hw) -> (x, y), zeroed where maxval <= 0, moved +-0.25 px toward the higher neighbour when the
 * peak is at least 2 px inside the map, then mapped back to image coordinates by the inverse of
 * get_affine_transform(center, scale, rot = 0, [W, H]) (affine_transform.py:13-45; scale in units of 200 px).
 * hm NHWC [B,H,W,J] (pitch); center, scale [B,2] float; preds [B,J,2] float.                        */
int fami_final_preds(const void* hm, int dtype, int pitch, const int32_t* idx, const float* maxvals,
                     const float* center, const float* scale, float* preds, int B, int H, int W, int J,
                     void* stream);
/* accuracy(output, target, hm_type='gaussian', thr), engine/core/utils/evaluate.py:13-75, from the argmaxes of
 * the predicted and the ground-truth heat maps.  out: double[J+3] = acc[0..J] (acc[0] = average over joints
 * with at least one valid sample, -1 marks joints without), avg_acc, cnt.                            */
int fami_pck_accuracy(const int32_t* pred_idx, const float* pred_max, const int32_t* target_idx,
                      const float* target_max, double* out, int B, int H, int W, int J, float thr, void* stream);
/* generate_heatmaps, datasets/process/heatmaps_process.py:146-203, batched: joints / joints_vis [B,J,3]
 * (image pixels; visibility in column 0) -> target [B,J,hm_h,hm_w] (NCHW float, as the loader produces it)
 * and target_weight [B,J].                                                                           */
int fami_gaussian_targets(const float* joints, const float* joints_vis, float* target, float* target_weight,
                          int B, int J, int sigma, int img_w, int img_h, int hm_w, int hm_h, void* stream);
/* The affine crop of the input pipeline on the device: cv2.warpAffine(frame, trans, (Wd, Hd), flags=cv2.INTER_LINEAR) with the
 * default constant-0 border, bit for bit (datasets/zoo/posetrack/PoseTrack_Alignment.py:233-241,415-423; crop(),
 * datasets/process/affine_transform.py:76-82).  `nframes` uint8 HWC frames of Hs x Ws, `src_frame_stride` BYTES apart;
 * inv_trans: DEVICE double[nframes][6], the inverse of each frame's 2x3 `trans` exactly as cv::invertAffineTransform forms it
 * (the caller inverts in float64: fami_pose_b200.pipeline.invert_affine).  out: uint8 [nframes][Hd][Wd][3] when mean3 / std3
 * (HOST pointers) are NULL, else float [nframes][Hd][Wd][3] = ((u8 / 255) - mean) / std -- ToTensor + Normalize of
 * datasets/transforms/build.py:13-22 applied to that 8-bit value, i.e. the frame the backbone consumes.                      */
int fami_crop_affine_u8(const uint8_t* frames, int64_t src_frame_stride, int Hs, int Ws, const double* inv_trans, void* out,
                        int nframes, int Hd, int Wd, const float* mean3, const float* std3, void* stream);
/* torchvision ToTensor + Normalize(mean, std) of datasets/transforms/build.py:13-22 on the device, fused with the
 * frame re-batching of Alignment_V15.py:115-119: `nframes` uint8 HWC (RGB) frames, `src_frame_stride` BYTES apart,
 * -> dense fp32 NHWC frames out[f][y][x][c] = ((u8/255) - mean[c]) / std[c].  mean3 / std3 are HOST pointers.   */
int fami_frames_u8_normalize(const uint8_t* frames, int64_t src_frame_stride, float* out, int nframes, int H, int W,
                             const float* mean3, const float* std3, void* stream);

/* ---- hardware probes (tooling only) ------------------------------------------------------------
 * NOT exported by the product library: compiled only with -DFAMI_DEBUG_PROBES into
 * libfami_b200_probes.so (python fami_pose_b200/csrc/build.py --probes), which tools/probe_*.py and
 * tools/trace_*.py load instead of libfami_b200.so.                                                */
#ifdef FAMI_DEBUG_PROBES
/* out[128][16] = x[shift:shift+128][64] @ w[16][64]^T through one tcgen05 tile whose A descriptor
 * starts `shift` 128-byte rows into a TMA-written SWIZZLE_128B tile (mode 0: base_offset 0,
 * mode 1: base_offset = (addr >> 7) & 7).  Establishes the row-shift rule the halo conv relies on. */
int fami_debug_umma_rowshift(const void* x_f16, const void* w_f16, float* out, int R, int shift, int mode,
                             void* stream);
/* globaltimer timeline (ns) of CTA 0 of the last halo conv launched with FAMI_HALO_TRACE=1:
 * [role 0..2][tile iteration 0..63][event 0..7] (tools/trace_halo.py).  Synchronises the device.   */
int fami_debug_read_trace(uint64_t* host_out, int n);
/* out[0] = SM clocks for `iters` back-to-back M=128,K=16 f16 tcgen05.mma of width N (issue + drain),
 * out[1] = clocks of the issue loop alone (tools/probe_umma.py).                                   */
int fami_debug_umma_rate(int64_t* out, int N, int iters, int variant, void* stream);
/* bit patterns that land in shared memory when [rows][32] floats are TMA-loaded through a tensor map of
 * element type FLOAT32 (mode 0), TFLOAT32 (1) or TFLOAT32_FTZ (2) (tools/probe_tma_tf32.py).           */
int fami_debug_tma_tf32(const float* x, uint32_t* out_bits, int rows, int mode, void* stream);
#endif

#ifdef __cplusplus
}
#endif
#endif /* FAMI_B200_H_ */
