// extern "C" surface of libfami_b200.so (see include/fami_b200.h).  Argument validation lives
// here; kernels live in conv_simt.cu / conv_tc.cu / dcn.cu / misc.cu.
#include <math.h>
#include <atomic>
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace fami {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms() {
  static std::atomic<int> sms[64];   // per device (zero-initialised)
  int dev = 0;
  cudaGetDevice(&dev);
  int v = sms[dev & 63].load(std::memory_order_relaxed);
  if (!v) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
    sms[dev & 63].store(v, std::memory_order_relaxed);
  }
  return v;
}

// kernels implemented in other translation units
int conv_f32_launch(const fami_conv_desc* d, const float* x, const float* w, const float* scale,
                    const float* shift, const void* res, void* y, double* stats, cudaStream_t st);
int pack_w_f32_launch(const float* w, float* out, int Cout, int Cin, int kh, int kw, cudaStream_t st);
int dcn_simt_launch(const fami_dcn_desc* d, const void* x, const float* off, const float* mask,
                    const float* w, const float* bias, void* out, cudaStream_t st);
int conv_bf16_tc_supported(const fami_conv_desc* d);
int conv_bf16_tc_launch(const fami_conv_desc* d, const void* x, const void* w, const float* scale,
                        const float* shift, const void* res, void* y, double* stats, cudaStream_t st,
                        const float* res32 = nullptr, float* y32 = nullptr, int y32_pitch = 0);
int64_t pack_w_bf16_elems(int Cout, int Cin, int kh, int kw, int dtype);
int stem_tc_supported(const fami_conv_desc* d, const void* y);
int stem_tc_launch(const fami_conv_desc* d, const float* x, const float* w, const float* scale, const float* shift, void* y,
                   cudaStream_t st);
int conv_halo_supported(const fami_conv_desc* d);
int conv_halo_launch(const fami_conv_desc* d, const void* x, const void* w, const float* scale, const float* shift,
                     const void* res, void* y, cudaStream_t st, const float* res32 = nullptr, float* y32 = nullptr,
                     int y32_pitch = 0);
int pack_w_bf16_launch(const float* w, void* out, int Cout, int Cin, int kh, int kw, int dtype, cudaStream_t st);
int flip_transpose_launch(const float* w, float* wt, int Cout, int Cin, int kh, int kw, cudaStream_t st);
int conv_dgrad_launch(const fami_conv_desc* d, const float* gy, const float* wt_packed, float* gx, cudaStream_t st);
int conv_wgrad_launch(const fami_conv_desc* d, const float* x, const float* gy, float* dw, float* dbias, cudaStream_t st);
int bn_bwd_launch(const float* x, int xp, const float* gy, int gp, const float* y, int yp, const float* mean,
                  const float* invstd, const float* gamma, int64_t rows, int C, int training, double* sums, float* dx,
                  int dxp, float* g_res, int grp, float* dgamma, float* dbeta, cudaStream_t st);
int softmax_pkl_bwd_launch(const float* a, int ap, const float* b, int bp, const float* gout, float* ga, int gap, float* gb,
                           int gbp, int B, int HW, int C, float temperature, cudaStream_t st);
int linear_bwd_launch(const float* x, const float* w, const float* gy, float* gx, float* gw, float* gb, int M, int K, int N,
                      cudaStream_t st);
int adam_step_launch(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                     float bc1, float bc2_sqrt, cudaStream_t st);
int adam_step_dev_launch(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, float beta1, float beta2,
                         float eps, cudaStream_t st);
int upsample_add_bwd_launch(const float* gy, int gp, const float* y, int yp, float* gsmall, int sp, float* gres, int rp, int N,
                            int Ho, int Wo, int C, int up, cudaStream_t st);
int final_preds_launch(const void* hm, int dt, int pitch, const int32_t* idx, const float* maxv, const float* center,
                       const float* scale, float* preds, int B, int H, int W, int J, cudaStream_t st);
int pck_accuracy_launch(const int32_t* pidx, const float* pmax, const int32_t* tidx, const float* tmax, double* out, int B,
                        int H, int W, int J, float thr, cudaStream_t st);
int gaussian_targets_launch(const float* joints, const float* vis, float* target, float* weight, int B, int J, int sigma,
                            int img_w, int img_h, int hm_w, int hm_h, cudaStream_t st);
int frames_u8_normalize_launch(const uint8_t* src, int64_t src_frame_stride, float* dst, int nframes, int64_t px_per_frame,
                               const float* mean, const float* std, cudaStream_t st);
int crop_affine_u8_launch(const uint8_t* src, int64_t src_frame_stride, int Hs, int Ws, const double* minv, void* dst,
                          int nframes, int Hd, int Wd, const float* mean, const float* std, cudaStream_t st);
int dcn_tc_supported(const fami_dcn_desc* d);
int dcn_wp_supported(const fami_dcn_desc* d);
int dcn_wp_launch(const fami_dcn_desc* d, const void* x, const float* om, const void* w, const float* bias, void* out,
                  cudaStream_t st);
int dcn_tc_launch(const fami_dcn_desc* d, const void* x, const float* om, const void* w, const float* bias, void* out,
                  cudaStream_t st);
int dcn_bwd_launch(const fami_dcn_desc* d, const float* x, const float* off, const float* mask, const float* w,
                   const float* go, float* gx, float* goff, float* gmask, float* gw, float* gb, cudaStream_t st);
int warp_translate_bwd_launch(const float* src, int sp, const float* txy, const float* go, int gop, float* gs,
                              int gsp, float* gtxy, int B, int H, int W, int C, cudaStream_t st);
int nchw_to_nhwc_launch(const float*, int64_t, void*, int, int, int, int, int, int, cudaStream_t);
int nhwc_to_nchw_launch(const void*, int, int, float*, int, int, int, int, cudaStream_t);
int warp_translate_fwd_launch(const void*, int, const float*, void*, int, int, int, int, int, int, cudaStream_t);
int sub_bcast_launch(const void*, const void*, void*, int, int64_t, int, cudaStream_t);
int linear_fwd_launch(const float*, const float*, const float*, float*, int, int, int, cudaStream_t);
int copy2d_launch(const void*, int, void*, int, int, int64_t, int, cudaStream_t);
int bn_finalize_launch(const double*, const float*, const float*, float*, float*, float*, float*, float*, float*, int,
                       int64_t, float, float, cudaStream_t);
int bn_apply_act_launch(const void*, int, int, const float*, const float*, const void*, int, void*, int, int, int, int,
                        int, int, int, int, cudaStream_t);
int bn_stats_launch(const void*, int, int, int64_t, int, double*, cudaStream_t);
int joint_mse_launch(const void*, int, int, const float*, const float*, float*, float*, float, int, int, int, int,
                     cudaStream_t);
int softmax_pkl_launch(const void*, int, const void*, int, int, float*, int, int, int, float, cudaStream_t);
int argmax_hw_launch(const void*, int, int, int32_t*, float*, int, int, int, cudaStream_t);
int debug_read_trace(unsigned long long* host_out, int n);
int debug_read_dcn_trace(unsigned long long* host_out, int n);
int debug_umma_rate_launch(long long* out, int N, int iters, int variant, cudaStream_t st);
int debug_umma_rowshift_launch(const void*, const void*, float*, int, int, int, cudaStream_t);
int debug_tma_tf32_launch(const float* x, uint32_t* out, int rows, int mode, cudaStream_t st);

}  // namespace fami

using namespace fami;

static inline bool valid_dtype(int dt) { return dt == FAMI_F32 || dt == FAMI_BF16 || dt == FAMI_F16; }   /* storage types */
static inline size_t esize(int dt) { return (dt == FAMI_F32 || dt == FAMI_TF32) ? 4 : 2; }
static inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

extern "C" {

const char* fami_last_error(void) { return g_err; }
int fami_abi_version(void) { return FAMI_ABI_VERSION; }
int64_t fami_launch_count(void) { return g_launches.load(); }

int64_t fami_workspace_bytes(int op, const void* desc) {
  if (!desc) return -1;
  const fami_conv_desc* c = static_cast<const fami_conv_desc*>(desc);
  const fami_dcn_desc* d = static_cast<const fami_dcn_desc*>(desc);
  switch (op) {
    case FAMI_OP_CONV_FWD: return c->stats ? 16ll * c->Cout : 0;
    case FAMI_OP_CONV_DGRAD:
      return 4ll * c->Cout * c->Cin * c->kh * c->kw + 4ll * fami_packed_weight_elems(c->Cin, c->Cout, c->kh, c->kw, FAMI_F32);
    case FAMI_OP_CONV_WGRAD: return 0;
    case FAMI_OP_BN_BWD: return 16ll * c->Cout;
    case FAMI_OP_DCN_FWD: return 0;
    case FAMI_OP_DCN_BWD: return 4ll * d->kh * d->kw * d->C * fami_conv_cout_pad(d->Cout);
    default: return -1;
  }
}

int fami_nchw_to_nhwc(const float* src, int64_t src_n_stride, void* dst, int dst_dtype, int N, int C, int H, int W,
                      int dst_pitch, void* stream) {
  FAMI_CHECK_ARG(src && dst, "fami_nchw_to_nhwc: null pointer");
  FAMI_CHECK_ARG(valid_dtype(dst_dtype), "fami_nchw_to_nhwc: bad dtype %d", dst_dtype);
  FAMI_CHECK_ARG(N > 0 && C > 0 && H > 0 && W > 0 && dst_pitch >= C, "fami_nchw_to_nhwc: bad shape");
  return nchw_to_nhwc_launch(src, src_n_stride, dst, dst_dtype, N, C, H, W, dst_pitch, (cudaStream_t)stream);
}

int fami_nhwc_to_nchw(const void* src, int src_dtype, int src_pitch, float* dst, int N, int C, int H, int W,
                      void* stream) {
  FAMI_CHECK_ARG(src && dst, "fami_nhwc_to_nchw: null pointer");
  FAMI_CHECK_ARG(valid_dtype(src_dtype), "fami_nhwc_to_nchw: bad dtype %d", src_dtype);
  FAMI_CHECK_ARG(N > 0 && C > 0 && H > 0 && W > 0 && src_pitch >= C, "fami_nhwc_to_nchw: bad shape");
  return nhwc_to_nchw_launch(src, src_dtype, src_pitch, dst, N, C, H, W, (cudaStream_t)stream);
}

int fami_conv_cout_pad(int Cout) { return ((Cout + 15) / 16) * 16; }

int64_t fami_packed_weight_elems(int Cout, int Cin, int kh, int kw, int dtype) {
  if (is_tc_dtype(dtype)) return pack_w_bf16_elems(Cout, Cin, kh, kw, dtype);
  int64_t kpad = ((int64_t)kh * kw * Cin + 15) / 16 * 16;
  return kpad * fami_conv_cout_pad(Cout);
}

int fami_pack_conv_weight(const float* w_oihw, void* w_packed, int Cout, int Cin, int kh, int kw, int dtype,
                          void* stream) {
  FAMI_CHECK_ARG(w_oihw && w_packed, "fami_pack_conv_weight: null pointer");
  FAMI_CHECK_ARG(valid_dtype(dtype) || dtype == FAMI_TF32, "fami_pack_conv_weight: bad dtype %d", dtype);
  FAMI_CHECK_ARG(Cout > 0 && Cin > 0 && kh > 0 && kw > 0, "fami_pack_conv_weight: bad shape");
  if (is_tc_dtype(dtype)) return pack_w_bf16_launch(w_oihw, w_packed, Cout, Cin, kh, kw, dtype, (cudaStream_t)stream);
  return pack_w_f32_launch(w_oihw, (float*)w_packed, Cout, Cin, kh, kw, (cudaStream_t)stream);
}

int fami_conv2d_bn_act_fwd(const fami_conv_desc* d, const void* x, const void* w_packed, const float* scale,
                           const float* shift, const void* residual, void* y, double* stats_out, void* stream) {
  FAMI_CHECK_ARG(d && x && w_packed && y, "fami_conv2d_bn_act_fwd: null pointer");
  FAMI_CHECK_ARG(valid_dtype(d->dtype) || d->dtype == FAMI_TF32, "fami_conv2d_bn_act_fwd: bad dtype %d", d->dtype);
  FAMI_CHECK_ARG(d->N > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "fami_conv2d_bn_act_fwd: bad shape");
  FAMI_CHECK_ARG(d->kh == d->kw && (d->kh == 1 || d->kh == 3), "fami_conv2d_bn_act_fwd: kernel %dx%d unsupported",
                 d->kh, d->kw);
  FAMI_CHECK_ARG(d->stride >= 1 && d->dil >= 1 && d->pad >= 0, "fami_conv2d_bn_act_fwd: bad stride/dil/pad");
  int Ho = (d->H + 2 * d->pad - d->dil * (d->kh - 1) - 1) / d->stride + 1;
  int Wo = (d->W + 2 * d->pad - d->dil * (d->kw - 1) - 1) / d->stride + 1;
  FAMI_CHECK_ARG(Ho == d->Ho && Wo == d->Wo, "fami_conv2d_bn_act_fwd: Ho/Wo (%d,%d) inconsistent, expected (%d,%d)",
                 d->Ho, d->Wo, Ho, Wo);
  FAMI_CHECK_ARG(d->up == 1 || d->up == 2 || d->up == 4 || d->up == 8, "fami_conv2d_bn_act_fwd: up=%d", d->up);
  FAMI_CHECK_ARG(d->in_pitch >= d->Cin && (d->om_groups != 0 || d->out_pitch >= d->Cout), "fami_conv2d_bn_act_fwd: pitch < channels");
  FAMI_CHECK_ARG(!residual || d->res_pitch >= d->Cout, "fami_conv2d_bn_act_fwd: res_pitch < Cout");
  FAMI_CHECK_ARG(!d->stats || stats_out, "fami_conv2d_bn_act_fwd: stats requested without stats_out");
  FAMI_CHECK_ARG((int64_t)d->N * d->Ho * d->Wo * d->up * d->up < (1ll << 31),
                 "fami_conv2d_bn_act_fwd: too many output pixels");
  FAMI_CHECK_ARG(valid_dtype(d->out_dtype), "fami_conv2d_bn_act_fwd: bad out_dtype %d", d->out_dtype);
  /* fp32 input with bf16 output: only the stem (Cin not a multiple of 16), which runs the SIMT kernel */
  FAMI_CHECK_ARG(!(d->dtype == FAMI_F32 && is_half_dtype(d->out_dtype)) || (d->Cin % 16 != 0 && !d->stats),
                 "fami_conv2d_bn_act_fwd: fp32-in/bf16-out is only supported for the stem convolution");
  if (d->om_groups != 0) {
    FAMI_CHECK_ARG(d->om_groups > 0 && d->om_groups % 4 == 0 && d->Cout == 27 * d->om_groups && d->out_dtype == FAMI_F32 &&
                       is_tc_dtype(d->dtype) && d->up == 1 && !residual && !d->stats && d->kh == 3 && d->stride == 1 &&
                       d->pad == d->dil,
                   "fami_conv2d_bn_act_fwd: om_groups needs a 16-bit 3x3 stride-1 same conv with Cout = 27*G, fp32 output, "
                   "no residual / upsample / statistics");
    FAMI_CHECK_ARG((reinterpret_cast<uintptr_t>(y) & 15) == 0, "fami_conv2d_bn_act_fwd: om_groups: y must be 16-byte aligned");
    FAMI_CHECK_ARG(d->om_layout == 0 || d->om_layout == 2 || d->om_layout == 3, "fami_conv2d_bn_act_fwd: bad om_layout %d", d->om_layout);
    static const bool halo_off_om = getenv("FAMI_DISABLE_HALO") != nullptr;
    if (!halo_off_om && conv_halo_supported(d))
      return conv_halo_launch(d, x, w_packed, scale, shift, residual, y, (cudaStream_t)stream);
    FAMI_CHECK_ARG(conv_bf16_tc_supported(d), "fami_conv2d_bn_act_fwd: om_groups: shape not supported by the tensor path");
    return conv_bf16_tc_launch(d, x, w_packed, scale, shift, residual, y, stats_out, (cudaStream_t)stream);
  }
  if (d->dtype == FAMI_TF32 && d->Cin == 3) {
    // the 3-channel stem of the tf32 arm: fp32 pixels and fp32 output; on the tensor cores with fp16 multiplicands (the same
    // 11-bit significand as TF32, formed in the CTA while it builds the im2col tile) when the shape is the HRNet stem's,
    // else on the exact-fp32 kernel
    fami_conv_desc e = *d;
    e.dtype = FAMI_F32;
    static const bool stem_tc_off = getenv("FAMI_DISABLE_STEM_TC") != nullptr;
    if (!stem_tc_off && !residual && !d->stats && e.out_dtype == FAMI_F32 && stem_tc_supported(&e, y))
      return stem_tc_launch(&e, (const float*)x, (const float*)w_packed, scale, shift, y, (cudaStream_t)stream);
    return conv_f32_launch(&e, (const float*)x, (const float*)w_packed, scale, shift, residual, y, stats_out, (cudaStream_t)stream);
  }
  if (is_tc_dtype(d->dtype)) {
    FAMI_CHECK_ARG(d->out_dtype == d->dtype || d->out_dtype == FAMI_F32,
                   "fami_conv2d_bn_act_fwd: tensor-core conv output must be the input's storage type or fp32");
    FAMI_CHECK_ARG(conv_bf16_tc_supported(d), "fami_conv2d_bn_act_fwd: shape not supported by the tensor-core path "
                                              "(Cin multiple of 16 (16-bit) / 8 (tf32), 16-byte aligned pixel pitch)");
    static const bool halo_off = getenv("FAMI_DISABLE_HALO") != nullptr;
    if (!halo_off && !d->stats && conv_halo_supported(d))
      return conv_halo_launch(d, x, w_packed, scale, shift, residual, y, (cudaStream_t)stream);
    return conv_bf16_tc_launch(d, x, w_packed, scale, shift, residual, y, stats_out, (cudaStream_t)stream);
  }
  return conv_f32_launch(d, (const float*)x, (const float*)w_packed, scale, shift, residual, y, stats_out,
                         (cudaStream_t)stream);
}

int fami_conv2d_bn_act_fwd_stream(const fami_conv_desc* d, const void* x, const void* w_packed, const float* scale,
                                  const float* shift, const float* residual_f32, void* y, float* y_f32, int y32_pitch,
                                  void* stream) {
  FAMI_CHECK_ARG(d && x && w_packed && y, "fami_conv2d_bn_act_fwd_stream: null pointer");
  FAMI_CHECK_ARG(is_half_dtype(d->dtype) && d->out_dtype == d->dtype, "fami_conv2d_bn_act_fwd_stream: 16-bit arms only");
  FAMI_CHECK_ARG(residual_f32 || y_f32, "fami_conv2d_bn_act_fwd_stream: neither a float residual nor a float output given");
  FAMI_CHECK_ARG(d->om_groups == 0 && !d->stats, "fami_conv2d_bn_act_fwd_stream: no om_groups / stats in stream mode");
  FAMI_CHECK_ARG(d->kh == d->kw && (d->kh == 1 || d->kh == 3) && d->stride >= 1 && d->dil >= 1 && d->pad >= 0,
                 "fami_conv2d_bn_act_fwd_stream: bad kernel geometry");
  int Ho = (d->H + 2 * d->pad - d->dil * (d->kh - 1) - 1) / d->stride + 1;
  int Wo = (d->W + 2 * d->pad - d->dil * (d->kw - 1) - 1) / d->stride + 1;
  FAMI_CHECK_ARG(Ho == d->Ho && Wo == d->Wo, "fami_conv2d_bn_act_fwd_stream: Ho/Wo inconsistent");
  FAMI_CHECK_ARG(d->up == 1 || d->up == 2 || d->up == 4 || d->up == 8, "fami_conv2d_bn_act_fwd_stream: up=%d", d->up);
  FAMI_CHECK_ARG(d->in_pitch >= d->Cin && d->out_pitch >= d->Cout && (!residual_f32 || d->res_pitch >= d->Cout) &&
                     (!y_f32 || y32_pitch >= d->Cout),
                 "fami_conv2d_bn_act_fwd_stream: pitch < channels");
  FAMI_CHECK_ARG(conv_bf16_tc_supported(d), "fami_conv2d_bn_act_fwd_stream: shape not supported by the tensor-core path");
  static const bool halo_off = getenv("FAMI_DISABLE_HALO") != nullptr;
  if (!halo_off && conv_halo_supported(d))
    return conv_halo_launch(d, x, w_packed, scale, shift, nullptr, y, (cudaStream_t)stream, residual_f32, y_f32, y32_pitch);
  return conv_bf16_tc_launch(d, x, w_packed, scale, shift, nullptr, y, nullptr, (cudaStream_t)stream, residual_f32, y_f32,
                             y32_pitch);
}

static int check_conv_bwd(const fami_conv_desc* d, const char* who) {
  FAMI_CHECK_ARG(d, "%s: null desc", who);
  FAMI_CHECK_ARG((d->dtype == FAMI_F32 || d->dtype == FAMI_TF32) && d->out_dtype == FAMI_F32, "%s: fp32 storage only", who);
  FAMI_CHECK_ARG(d->N > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "%s: bad shape", who);
  FAMI_CHECK_ARG(d->kh == d->kw && (d->kh == 1 || d->kh == 3), "%s: kernel %dx%d unsupported", who, d->kh, d->kw);
  FAMI_CHECK_ARG(d->stride >= 1 && d->dil >= 1 && d->pad >= 0 && d->up == 1, "%s: bad stride/dil/pad/up", who);
  int Ho = (d->H + 2 * d->pad - d->dil * (d->kh - 1) - 1) / d->stride + 1;
  int Wo = (d->W + 2 * d->pad - d->dil * (d->kw - 1) - 1) / d->stride + 1;
  FAMI_CHECK_ARG(Ho == d->Ho && Wo == d->Wo, "%s: Ho/Wo (%d,%d) inconsistent, expected (%d,%d)", who, d->Ho, d->Wo, Ho, Wo);
  FAMI_CHECK_ARG(d->in_pitch >= d->Cin && d->out_pitch >= d->Cout, "%s: pitch < channels", who);
  return 0;
}

int fami_pack_conv_weight_dgrad(const float* w_oihw, float* scratch_oihw, void* w_packed_t, int Cout, int Cin, int kh,
                                int kw, int dtype, void* stream) {
  FAMI_CHECK_ARG(w_oihw && scratch_oihw && w_packed_t, "fami_pack_conv_weight_dgrad: null pointer");
  FAMI_CHECK_ARG(dtype == FAMI_F32 || dtype == FAMI_TF32, "fami_pack_conv_weight_dgrad: fp32 storage only (FAMI_F32 / FAMI_TF32)");
  if (int e = flip_transpose_launch(w_oihw, scratch_oihw, Cout, Cin, kh, kw, (cudaStream_t)stream)) return e;
  return fami_pack_conv_weight(scratch_oihw, w_packed_t, Cin, Cout, kh, kw, dtype, stream);
}

int fami_conv2d_dgrad(const fami_conv_desc* d, const float* grad_y, const float* w_packed_t, float* grad_x, void* stream) {
  if (int e = check_conv_bwd(d, "fami_conv2d_dgrad")) return e;
  FAMI_CHECK_ARG(grad_y && w_packed_t && grad_x, "fami_conv2d_dgrad: null pointer");
  FAMI_CHECK_ARG(d->stride > 1 || d->dil * (d->kh - 1) - d->pad >= 0, "fami_conv2d_dgrad: pad larger than the filter reach");
  FAMI_CHECK_ARG(d->dtype == FAMI_F32 || d->stride == 1, "fami_conv2d_dgrad: the tf32 tensor-core path takes stride-1 convolutions "
                                                         "(strided dgrad runs the fp32 gather kernel: pass FAMI_F32)");
  return conv_dgrad_launch(d, grad_y, w_packed_t, grad_x, (cudaStream_t)stream);
}

int fami_conv2d_wgrad(const fami_conv_desc* d, const float* x, const float* grad_y, float* grad_w_oihw, float* grad_bias,
                      void* stream) {
  if (int e = check_conv_bwd(d, "fami_conv2d_wgrad")) return e;
  FAMI_CHECK_ARG(x && grad_y && grad_w_oihw, "fami_conv2d_wgrad: null pointer");
  FAMI_CHECK_ARG(d->dtype == FAMI_F32 || d->dtype == FAMI_TF32, "fami_conv2d_wgrad: fp32 storage (FAMI_F32: exact FMAs; FAMI_TF32: "
                                                                  "TF32 tensor cores, fp32 accumulation)");
  return conv_wgrad_launch(d, x, grad_y, grad_w_oihw, grad_bias, (cudaStream_t)stream);
}

int fami_bn_bwd(const float* x, int x_pitch, const float* grad_y, int gy_pitch, const float* y, int y_pitch,
                const float* mean, const float* invstd, const float* gamma, int64_t rows, int C, int training,
                double* sums, float* grad_x, int gx_pitch, float* grad_res, int gres_pitch, float* grad_gamma,
                float* grad_beta, void* stream) {
  FAMI_CHECK_ARG(x && grad_y && mean && invstd && sums, "fami_bn_bwd: null pointer");
  FAMI_CHECK_ARG(rows > 0 && C > 0 && C <= 1024 && x_pitch >= C && gy_pitch >= C, "fami_bn_bwd: bad shape");
  return bn_bwd_launch(x, x_pitch, grad_y, gy_pitch, y, y_pitch, mean, invstd, gamma, rows, C, training, sums, grad_x,
                       gx_pitch, grad_res, gres_pitch, grad_gamma, grad_beta, (cudaStream_t)stream);
}

int fami_softmax_pkl_bwd(const float* a, int a_pitch, const float* b, int b_pitch, const float* grad_out, float* grad_a,
                         int ga_pitch, float* grad_b, int gb_pitch, int B, int HW, int C, float temperature, void* stream) {
  FAMI_CHECK_ARG(a && b && grad_out && (grad_a || grad_b), "fami_softmax_pkl_bwd: null pointer");
  FAMI_CHECK_ARG(B > 0 && HW > 0 && C > 0 && C <= 1024 && temperature > 0.f, "fami_softmax_pkl_bwd: bad shape");
  return softmax_pkl_bwd_launch(a, a_pitch, b, b_pitch, grad_out, grad_a, ga_pitch, grad_b, gb_pitch, B, HW, C, temperature,
                                (cudaStream_t)stream);
}

int fami_linear_bwd(const float* x, const float* w, const float* grad_y, float* grad_x, float* grad_w, float* grad_b, int M,
                    int K, int N, void* stream) {
  FAMI_CHECK_ARG(x && w && grad_y && M > 0 && K > 0 && N > 0, "fami_linear_bwd: bad arguments");
  return linear_bwd_launch(x, w, grad_y, grad_x, grad_w, grad_b, M, K, N, (cudaStream_t)stream);
}

int fami_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                   float beta2, float eps, int step, void* stream) {
  FAMI_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && n > 0 && step >= 1, "fami_adam_step: bad arguments");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  return adam_step_launch(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, (float)bc1, (float)sqrt(bc2),
                          (cudaStream_t)stream);
}

int fami_adam_step_graph(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, const float* hyper_dev,
                         float beta1, float beta2, float eps, void* stream) {
  FAMI_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && hyper_dev && n > 0, "fami_adam_step_graph: bad arguments");
  return adam_step_dev_launch(param, grad, exp_avg, exp_avg_sq, n, hyper_dev, beta1, beta2, eps, (cudaStream_t)stream);
}

int fami_upsample_add_bwd(const float* grad_y, int gy_pitch, const float* y, int y_pitch, float* grad_small, int gs_pitch,
                          float* grad_res, int gr_pitch, int N, int Ho, int Wo, int C, int up, void* stream) {
  FAMI_CHECK_ARG(grad_y && grad_small, "fami_upsample_add_bwd: null pointer");
  FAMI_CHECK_ARG(N > 0 && Ho > 0 && Wo > 0 && C > 0 && C % 4 == 0 && (up == 1 || up == 2 || up == 4 || up == 8),
                 "fami_upsample_add_bwd: bad shape (C multiple of 4, up in 1/2/4/8)");
  FAMI_CHECK_ARG(gy_pitch % 4 == 0 && gs_pitch % 4 == 0 && (!y || y_pitch % 4 == 0) && (!grad_res || gr_pitch % 4 == 0) &&
                     aligned(grad_y, 16) && aligned(grad_small, 16) && (!y || aligned(y, 16)) && (!grad_res || aligned(grad_res, 16)),
                 "fami_upsample_add_bwd: 16-byte aligned operands with pitches multiple of 4");
  return upsample_add_bwd_launch(grad_y, gy_pitch, y, y_pitch, grad_small, gs_pitch, grad_res, gr_pitch, N, Ho, Wo, C, up,
                                 (cudaStream_t)stream);
}

int fami_frames_u8_normalize(const uint8_t* frames, int64_t src_frame_stride, float* out, int nframes, int H, int W,
                             const float* mean3, const float* std3, void* stream) {
  FAMI_CHECK_ARG(frames && out && mean3 && std3 && nframes > 0 && H > 0 && W > 0, "fami_frames_u8_normalize: bad arguments");
  FAMI_CHECK_ARG(src_frame_stride >= (int64_t)H * W * 3, "fami_frames_u8_normalize: frame stride smaller than a frame");
  return frames_u8_normalize_launch(frames, src_frame_stride, out, nframes, (int64_t)H * W, mean3, std3, (cudaStream_t)stream);
}

int fami_crop_affine_u8(const uint8_t* frames, int64_t src_frame_stride, int Hs, int Ws, const double* inv_trans, void* out,
                        int nframes, int Hd, int Wd, const float* mean3, const float* std3, void* stream) {
  FAMI_CHECK_ARG(frames && inv_trans && out && nframes > 0 && Hs > 0 && Ws > 0 && Hd > 0 && Wd > 0, "fami_crop_affine_u8: bad arguments");
  FAMI_CHECK_ARG(src_frame_stride >= (int64_t)Hs * Ws * 3, "fami_crop_affine_u8: frame stride smaller than a frame");
  FAMI_CHECK_ARG((mean3 == nullptr) == (std3 == nullptr), "fami_crop_affine_u8: pass both mean3 and std3, or neither");
  FAMI_CHECK_ARG(Hs < 32768 && Ws < 32768, "fami_crop_affine_u8: source larger than OpenCV's 16-bit coordinate range");
  return crop_affine_u8_launch(frames, src_frame_stride, Hs, Ws, inv_trans, out, nframes, Hd, Wd, mean3, std3, (cudaStream_t)stream);
}

int fami_final_preds(const void* hm, int dtype, int pitch, const int32_t* idx, const float* maxvals, const float* center,
                     const float* scale, float* preds, int B, int H, int W, int J, void* stream) {
  FAMI_CHECK_ARG(hm && idx && maxvals && center && scale && preds, "fami_final_preds: null pointer");
  FAMI_CHECK_ARG(valid_dtype(dtype) && B > 0 && H > 0 && W > 0 && J > 0 && pitch >= J, "fami_final_preds: bad arguments");
  return final_preds_launch(hm, dtype, pitch, idx, maxvals, center, scale, preds, B, H, W, J, (cudaStream_t)stream);
}

int fami_pck_accuracy(const int32_t* pred_idx, const float* pred_max, const int32_t* target_idx, const float* target_max,
                      double* out, int B, int H, int W, int J, float thr, void* stream) {
  FAMI_CHECK_ARG(pred_idx && pred_max && target_idx && target_max && out, "fami_pck_accuracy: null pointer");
  FAMI_CHECK_ARG(B > 0 && H > 0 && W > 0 && J > 0 && J <= 1024, "fami_pck_accuracy: bad shape");
  return pck_accuracy_launch(pred_idx, pred_max, target_idx, target_max, out, B, H, W, J, thr, (cudaStream_t)stream);
}

int fami_gaussian_targets(const float* joints, const float* joints_vis, float* target, float* target_weight, int B, int J,
                          int sigma, int img_w, int img_h, int hm_w, int hm_h, void* stream) {
  FAMI_CHECK_ARG(joints && joints_vis && target && target_weight, "fami_gaussian_targets: null pointer");
  FAMI_CHECK_ARG(B > 0 && J > 0 && sigma > 0 && img_w > 0 && img_h > 0 && hm_w > 0 && hm_h > 0,
                 "fami_gaussian_targets: bad arguments");
  return gaussian_targets_launch(joints, joints_vis, target, target_weight, B, J, sigma, img_w, img_h, hm_w, hm_h,
                                 (cudaStream_t)stream);
}

int fami_bn_finalize(const double* stats, const float* gamma, const float* beta, float* running_mean,
                     float* running_var, float* scale, float* shift, float* save_mean, float* save_invstd, int C,
                     int64_t count, float eps, float momentum, void* stream) {
  FAMI_CHECK_ARG(stats && scale && shift && C > 0 && count > 0, "fami_bn_finalize: bad arguments");
  return bn_finalize_launch(stats, gamma, beta, running_mean, running_var, scale, shift, save_mean, save_invstd, C,
                            count, eps, momentum, (cudaStream_t)stream);
}

int fami_bn_apply_act(const void* x, int x_dtype, int x_pitch, const float* scale, const float* shift,
                      const void* residual, int res_pitch, void* y, int y_pitch, int dtype, int N, int Ho, int Wo, int C,
                      int up, int relu, void* stream) {
  FAMI_CHECK_ARG(x && y && ((scale && shift) || (!scale && !shift)), "fami_bn_apply_act: null pointer");
  FAMI_CHECK_ARG(valid_dtype(dtype) && valid_dtype(x_dtype), "fami_bn_apply_act: bad dtype");
  FAMI_CHECK_ARG(up == 1 || up == 2 || up == 4 || up == 8, "fami_bn_apply_act: up=%d", up);
  return bn_apply_act_launch(x, x_dtype, x_pitch, scale, shift, residual, res_pitch, y, y_pitch, dtype, N, Ho, Wo, C, up,
                             relu, (cudaStream_t)stream);
}

int fami_bn_stats(const void* x, int dtype, int pitch, int64_t rows, int C, double* stats, void* stream) {
  FAMI_CHECK_ARG(x && stats, "fami_bn_stats: null pointer");
  FAMI_CHECK_ARG(valid_dtype(dtype), "fami_bn_stats: bad dtype");
  FAMI_CHECK_ARG(rows > 0 && C > 0 && C <= 1024 && pitch >= C, "fami_bn_stats: bad shape");
  return bn_stats_launch(x, dtype, pitch, rows, C, stats, (cudaStream_t)stream);
}

static int check_dcn(const fami_dcn_desc* d, const char* who) {
  FAMI_CHECK_ARG(d, "%s: null desc", who);
  FAMI_CHECK_ARG(d->B > 0 && d->H > 0 && d->W > 0 && d->C > 0 && d->Cout > 0 && d->G > 0, "%s: bad shape", who);
  FAMI_CHECK_ARG(d->kh == 3 && d->kw == 3 && d->stride == 1, "%s: only 3x3 stride-1 deformable kernels", who);
  FAMI_CHECK_ARG(d->pad == d->dil, "%s: pad (%d) must equal dilation (%d) (same-size output)", who, d->pad, d->dil);
  /* torchvision raises for channels not divisible by groups (deform_conv.py:129-132) */
  FAMI_CHECK_ARG(d->C % d->G == 0, "%s: in_channels %d not divisible by offset groups %d", who, d->C, d->G);
  FAMI_CHECK_ARG((d->C / d->G) % 4 == 0 && d->C % 16 == 0,
                 "%s: channels per offset group must be a multiple of 4 and C a multiple of 16 (C=%d G=%d)", who,
                 d->C, d->G);
  FAMI_CHECK_ARG(d->om_layout >= 0 && d->om_layout <= 3, "%s: bad om_layout %d", who, d->om_layout);
  FAMI_CHECK_ARG(d->x_pitch >= d->C && d->out_pitch >= d->Cout, "%s: pitch too small", who);
  FAMI_CHECK_ARG(d->om_layout >= 2 || (d->om_layout == 1 ? d->off_pitch >= 27 * d->G : (d->off_pitch >= 18 * d->G && d->mask_pitch >= 9 * d->G)),
                 "%s: offset/mask pitch too small", who);
  FAMI_CHECK_ARG((int64_t)d->B * d->H * d->W < (1ll << 31), "%s: too many pixels", who);
  return 0;
}

int fami_dcn_fwd(const fami_dcn_desc* d, const void* x, const void* offset, const void* mask, const void* w_packed,
                 const float* bias, void* out, void* stream) {
  if (int e = check_dcn(d, "fami_dcn_fwd")) return e;
  FAMI_CHECK_ARG(x && offset && (mask || d->om_layout >= 1) && w_packed && out, "fami_dcn_fwd: null pointer");
  FAMI_CHECK_ARG(valid_dtype(d->dtype), "fami_dcn_fwd: bad dtype %d", d->dtype);
  FAMI_CHECK_ARG(!d->out_f32 || (d->om_layout >= 1 && is_half_dtype(d->dtype)),
                 "fami_dcn_fwd: out_f32 applies to the 16-bit tensor-core kernel (om_layout 1 / 2)");
  if (d->om_layout >= 1) {
    /* C == Cout <= 64: the warp-private kernel (dcn_wp.cu); everything else: the tcgen05 kernel (dcn_tc.cu) */
    if (dcn_wp_supported(d)) return dcn_wp_launch(d, x, (const float*)offset, w_packed, bias, out, (cudaStream_t)stream);
    FAMI_CHECK_ARG(d->om_layout != 3, "fami_dcn_fwd: the k-step-blocked offset layout (3) needs the warp-private kernel "
                                      "(16-bit x, C == Cout in {32, 48}, 4 channels per offset group, 3x3, pad == dil <= 4)");
    FAMI_CHECK_ARG(dcn_tc_supported(d), "fami_dcn_fwd: fused tap-major offsets need the 16-bit tensor-core kernel "
                                        "(C <= 64, 4 channels per offset group, 3x3, pad == dil)");
    return dcn_tc_launch(d, x, (const float*)offset, w_packed, bias, out, (cudaStream_t)stream);
  }
  FAMI_CHECK_ARG(aligned(x, 4 * esize(d->dtype)) && d->x_pitch % 4 == 0,
                 "fami_dcn_fwd: x must be aligned to 4 elements with pitch %% 4 == 0");
  return dcn_simt_launch(d, x, (const float*)offset, (const float*)mask, (const float*)w_packed, bias, out,
                         (cudaStream_t)stream);
}

int fami_dcn_bwd(const fami_dcn_desc* d, const float* x, const float* offset, const float* mask,
                 const float* w_packed, const float* grad_out, float* grad_x, float* grad_offset, float* grad_mask,
                 float* grad_w_packed, float* grad_bias, void* stream) {
  if (int e = check_dcn(d, "fami_dcn_bwd")) return e;
  FAMI_CHECK_ARG(x && offset && mask && w_packed && grad_out, "fami_dcn_bwd: null pointer");
  FAMI_CHECK_ARG(d->dtype == FAMI_F32 || d->dtype == FAMI_TF32, "fami_dcn_bwd: fp32 storage only (FAMI_F32: exact; FAMI_TF32: the "
                                                                  "weight gradient's products on TF32 tensor cores)");
  return dcn_bwd_launch(d, x, offset, mask, w_packed, grad_out, grad_x, grad_offset, grad_mask, grad_w_packed,
                        grad_bias, (cudaStream_t)stream);
}

int fami_warp_translate_fwd(const void* src, int src_pitch, const float* txy, void* out, int out_pitch, int dtype,
                            int B, int H, int W, int C, void* stream) {
  FAMI_CHECK_ARG(src && txy && out, "fami_warp_translate_fwd: null pointer");
  FAMI_CHECK_ARG(valid_dtype(dtype), "fami_warp_translate_fwd: bad dtype");
  FAMI_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "fami_warp_translate_fwd: C must be a multiple of 4");
  FAMI_CHECK_ARG(src_pitch % 4 == 0 && out_pitch % 4 == 0 && src_pitch >= C && out_pitch >= C,
                 "fami_warp_translate_fwd: pitches must be multiples of 4 and >= C");
  FAMI_CHECK_ARG(aligned(src, 4 * esize(dtype)) && aligned(out, 4 * esize(dtype)),
                 "fami_warp_translate_fwd: pointers must be aligned to 4 elements");
  return warp_translate_fwd_launch(src, src_pitch, txy, out, out_pitch, dtype, B, H, W, C, (cudaStream_t)stream);
}

int fami_warp_translate_bwd(const float* src, int src_pitch, const float* txy, const float* grad_out, int go_pitch,
                            float* grad_src, int gs_pitch, float* grad_txy, int B, int H, int W, int C,
                            void* stream) {
  FAMI_CHECK_ARG(src && txy && grad_out && (grad_src || grad_txy), "fami_warp_translate_bwd: null pointer");
  FAMI_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0, "fami_warp_translate_bwd: bad shape");
  return warp_translate_bwd_launch(src, src_pitch, txy, grad_out, go_pitch, grad_src, gs_pitch, grad_txy, B, H, W, C,
                                   (cudaStream_t)stream);
}

int fami_sub_bcast(const void* a, const void* b, void* out, int dtype, int64_t n, int rep, void* stream) {
  FAMI_CHECK_ARG(a && b && out, "fami_sub_bcast: null pointer");
  FAMI_CHECK_ARG(valid_dtype(dtype), "fami_sub_bcast: bad dtype");
  FAMI_CHECK_ARG(n > 0 && n % 4 == 0 && rep > 0, "fami_sub_bcast: n must be a positive multiple of 4");
  FAMI_CHECK_ARG(aligned(a, 4 * esize(dtype)) && aligned(b, 4 * esize(dtype)) && aligned(out, 4 * esize(dtype)),
                 "fami_sub_bcast: pointers must be aligned to 4 elements");
  return sub_bcast_launch(a, b, out, dtype, n, rep, (cudaStream_t)stream);
}

int fami_copy2d(const void* src, int src_pitch, void* dst, int dst_pitch, int dtype, int64_t rows, int cols,
                void* stream) {
  FAMI_CHECK_ARG(src && dst, "fami_copy2d: null pointer");
  FAMI_CHECK_ARG(valid_dtype(dtype), "fami_copy2d: bad dtype");
  FAMI_CHECK_ARG(rows > 0 && cols > 0 && cols % 4 == 0 && src_pitch % 4 == 0 && dst_pitch % 4 == 0 &&
                     src_pitch >= cols && dst_pitch >= cols,
                 "fami_copy2d: cols and pitches must be multiples of 4");
  FAMI_CHECK_ARG(aligned(src, 4 * esize(dtype)) && aligned(dst, 4 * esize(dtype)),
                 "fami_copy2d: pointers must be aligned to 4 elements");
  return copy2d_launch(src, src_pitch, dst, dst_pitch, dtype, rows, cols, (cudaStream_t)stream);
}

int fami_linear_fwd(const float* x, const float* w, const float* b, float* y, int M, int K, int N, void* stream) {
  FAMI_CHECK_ARG(x && w && y && M > 0 && K > 0 && N > 0, "fami_linear_fwd: bad arguments");
  return linear_fwd_launch(x, w, b, y, M, K, N, (cudaStream_t)stream);
}

int fami_joint_mse_fwd_bwd(const void* pred, int pred_dtype, int pred_pitch, const float* target_nchw,
                           const float* weight, float* loss_out, float* grad_pred, float grad_scale, int B, int J,
                           int H, int W, void* stream) {
  FAMI_CHECK_ARG(pred && target_nchw && loss_out, "fami_joint_mse_fwd_bwd: null pointer");
  FAMI_CHECK_ARG(valid_dtype(pred_dtype), "fami_joint_mse_fwd_bwd: bad dtype");
  FAMI_CHECK_ARG(B > 0 && J > 0 && H > 0 && W > 0 && pred_pitch >= J, "fami_joint_mse_fwd_bwd: bad shape");
  return joint_mse_launch(pred, pred_dtype, pred_pitch, target_nchw, weight, loss_out, grad_pred, grad_scale, B, J, H,
                          W, (cudaStream_t)stream);
}

int fami_softmax_pkl_fwd(const void* a, int a_pitch, const void* b, int b_pitch, int dtype, float* out, int B, int HW,
                         int C, float temperature, void* stream) {
  FAMI_CHECK_ARG(a && b && out, "fami_softmax_pkl_fwd: null pointer");
  FAMI_CHECK_ARG(valid_dtype(dtype), "fami_softmax_pkl_fwd: bad dtype");
  FAMI_CHECK_ARG(B > 0 && HW > 0 && C > 0 && C <= 1024 && a_pitch >= C && b_pitch >= C && temperature > 0.f,
                 "fami_softmax_pkl_fwd: bad shape");
  return softmax_pkl_launch(a, a_pitch, b, b_pitch, dtype, out, B, HW, C, temperature, (cudaStream_t)stream);
}

int fami_argmax_hw(const void* hm, int dtype, int pitch, int32_t* idx_out, float* maxval_out, int B, int HW, int J,
                   void* stream) {
  FAMI_CHECK_ARG(hm && idx_out && maxval_out, "fami_argmax_hw: null pointer");
  FAMI_CHECK_ARG(valid_dtype(dtype), "fami_argmax_hw: bad dtype");
  FAMI_CHECK_ARG(B > 0 && HW > 0 && J > 0 && J <= 1024 && pitch >= J, "fami_argmax_hw: bad shape");
  return argmax_hw_launch(hm, dtype, pitch, idx_out, maxval_out, B, HW, J, (cudaStream_t)stream);
}

#ifdef FAMI_DEBUG_PROBES
/* per-role timeline of CTA 0 of the last halo conv launched with FAMI_HALO_TRACE=1 (tools/trace_halo.py) */
int fami_debug_read_trace(uint64_t* host_out, int n) {
  FAMI_CHECK_ARG(host_out && n != 0, "fami_debug_read_trace: bad arguments");
  if (n < 0) return debug_read_dcn_trace(reinterpret_cast<unsigned long long*>(host_out), -n);   /* n < 0: DCN kernel trace */
  return debug_read_trace(reinterpret_cast<unsigned long long*>(host_out), n);
}

/* hardware probe: tcgen05.mma issue/execution rate (tools/probe_umma.py) */
int fami_debug_umma_rate(int64_t* out, int N, int iters, int variant, void* stream) {
  FAMI_CHECK_ARG(out, "fami_debug_umma_rate: null pointer");
  return debug_umma_rate_launch(reinterpret_cast<long long*>(out), N, iters, variant, (cudaStream_t)stream);
}

/* hardware probe used by tools/probe_umma.py (not part of the product path) */
int fami_debug_umma_rowshift(const void* x_f16, const void* w_f16, float* out, int R, int shift, int mode,
                             void* stream) {
  FAMI_CHECK_ARG(x_f16 && w_f16 && out, "fami_debug_umma_rowshift: null pointer");
  return debug_umma_rowshift_launch(x_f16, w_f16, out, R, shift, mode, (cudaStream_t)stream);
}

/* hardware probe: element conversion performed by a TMA load with a TFLOAT32 tensor map (tools/probe_tma_tf32.py) */
int fami_debug_tma_tf32(const float* x, uint32_t* out_bits, int rows, int mode, void* stream) {
  FAMI_CHECK_ARG(x && out_bits, "fami_debug_tma_tf32: null pointer");
  return debug_tma_tf32_launch(x, out_bits, rows, mode, (cudaStream_t)stream);
}
#endif  /* FAMI_DEBUG_PROBES */

}  // extern "C"
