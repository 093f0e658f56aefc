// Keypoint decode, PCK accuracy and gaussian target generation on the device -- the steps immediately after /
// before the hot path that the reference does in numpy on the host every iteration:
//   get_final_preds      datasets/process/heatmaps_process.py:47-81 (+ affine_transform.py:13-45, rot = 0)
//   accuracy             engine/core/utils/evaluate.py:13-75 (hm_type = 'gaussian')
//   generate_heatmaps    datasets/process/heatmaps_process.py:146-203
#include "common.cuh"

namespace fami {

namespace {

// cv2.getAffineTransform(src, dst) for three point pairs, solved in double (Cramer): rows (a, b, c) with
// X = a x + b y + c, Y likewise.
__device__ void affine_from_points(const float (&s)[3][2], const float (&d)[3][2], double (&t)[2][3]) {
  const double x0 = s[0][0], y0 = s[0][1], x1 = s[1][0], y1 = s[1][1], x2 = s[2][0], y2 = s[2][1];
  const double det = x0 * (y1 - y2) - y0 * (x1 - x2) + (x1 * y2 - x2 * y1);
  for (int k = 0; k < 2; ++k) {
    const double u0 = d[0][k], u1 = d[1][k], u2 = d[2][k];
    t[k][0] = (u0 * (y1 - y2) - y0 * (u1 - u2) + (u1 * y2 - u2 * y1)) / det;
    t[k][1] = (x0 * (u1 - u2) - u0 * (x1 - x2) + (x1 * u2 - x2 * u1)) / det;
    t[k][2] = (x0 * (y1 * u2 - y2 * u1) - y0 * (x1 * u2 - x2 * u1) + u0 * (x1 * y2 - x2 * y1)) / det;
  }
}

template <typename T>
__global__ void final_preds_kernel(const T* __restrict__ hm, int pitch, const int32_t* __restrict__ idx,
                                   const float* __restrict__ maxv, const float* __restrict__ center,
                                   const float* __restrict__ scale, float* __restrict__ preds, int B, int H, int W, int J) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * J) return;
  const int b = i / J, j = i - b * J;
  const int id = idx[i];
  float cx = (float)(id % W), cy = floorf((float)id / (float)W);
  if (!(maxv[i] > 0.f)) { cx = 0.f; cy = 0.f; }
  // +-0.25 px toward the higher neighbour (heatmaps_process.py:56-62)
  const int px = (int)floorf(cx + 0.5f), py = (int)floorf(cy + 0.5f);
  if (1 < px && px < W - 1 && 1 < py && py < H - 1) {
    const T* base = hm + (int64_t)b * H * W * pitch + j;
    const float dx = to_f<T>(base[((int64_t)py * W + px + 1) * pitch]) - to_f<T>(base[((int64_t)py * W + px - 1) * pitch]);
    const float dy = to_f<T>(base[((int64_t)(py + 1) * W + px) * pitch]) - to_f<T>(base[((int64_t)(py - 1) * W + px) * pitch]);
    cx += (dx > 0.f ? 0.25f : (dx < 0.f ? -0.25f : 0.f));
    cy += (dy > 0.f ? 0.25f : (dy < 0.f ? -0.25f : 0.f));
  }
  // inverse affine heat map -> image (transform_preds :76-81; get_affine_transform with rot = 0, shift = 0, inv = 1)
  const double c0 = center[2 * b], c1 = center[2 * b + 1];
  const double src_w = (double)scale[2 * b] * 200.0;
  float s[3][2], d[3][2];
  s[0][0] = (float)c0; s[0][1] = (float)c1;
  s[1][0] = (float)(c0 + 0.0); s[1][1] = (float)(c1 + src_w * -0.5);
  d[0][0] = (float)(W * 0.5); d[0][1] = (float)(H * 0.5);
  d[1][0] = (float)((double)(W * 0.5) + 0.0); d[1][1] = (float)((double)(H * 0.5) + (double)(float)(W * -0.5));
  s[2][0] = s[1][0] - (s[0][1] - s[1][1]); s[2][1] = s[1][1] + (s[0][0] - s[1][0]);
  d[2][0] = d[1][0] - (d[0][1] - d[1][1]); d[2][1] = d[1][1] + (d[0][0] - d[1][0]);
  double t[2][3];
  affine_from_points(d, s, t);
  preds[2 * i] = (float)(t[0][0] * (double)cx + t[0][1] * (double)cy + t[0][2]);
  preds[2 * i + 1] = (float)(t[1][0] * (double)cx + t[1][1] * (double)cy + t[1][2]);
}

// one block; thread j = joint.  out[0..J] = acc (acc[0] = average), out[J+1] = avg_acc, out[J+2] = cnt
__global__ void pck_accuracy_kernel(const int32_t* __restrict__ pidx, const float* __restrict__ pmax,
                                    const int32_t* __restrict__ tidx, const float* __restrict__ tmax,
                                    double* __restrict__ out, int B, int H, int W, int J, double thr) {
  extern __shared__ double s_acc[];
  const int j = threadIdx.x;
  if (j < J) {
    int valid = 0, hit = 0;
    const double nx = (double)H / 10.0, ny = (double)W / 10.0;   // (h, w)/10 against (x, y): the reference's order
    for (int b = 0; b < B; ++b) {
      const int i = b * J + j;
      float px = (float)(pidx[i] % W), py = floorf((float)pidx[i] / (float)W);
      float tx = (float)(tidx[i] % W), ty = floorf((float)tidx[i] / (float)W);
      if (!(pmax[i] > 0.f)) { px = 0.f; py = 0.f; }
      if (!(tmax[i] > 0.f)) { tx = 0.f; ty = 0.f; }
      if (tx > 1.f && ty > 1.f) {
        const double dx = (double)px / nx - (double)tx / nx, dy = (double)py / ny - (double)ty / ny;
        ++valid;
        if (sqrt(dx * dx + dy * dy) < thr) ++hit;
      }
    }
    s_acc[j] = valid > 0 ? (double)hit / (double)valid : -1.0;
  }
  __syncthreads();
  if (j == 0) {
    double avg = 0.0;
    int cnt = 0;
    for (int k = 0; k < J; ++k) {
      out[k + 1] = s_acc[k];
      if (s_acc[k] >= 0.0) { avg += s_acc[k]; ++cnt; }
    }
    avg = cnt != 0 ? avg / cnt : 0.0;
    out[0] = cnt != 0 ? avg : 0.0;
    out[J + 1] = avg;
    out[J + 2] = (double)cnt;
  }
}

// block = one (sample, joint); target zero-filled by the launcher
__global__ void gaussian_targets_kernel(const float* __restrict__ joints, const float* __restrict__ vis,
                                        float* __restrict__ target, float* __restrict__ weight, int J, int sigma,
                                        double stride_x, double stride_y, int hw, int hh) {
  const int bj = blockIdx.x;
  const int tmp = sigma * 3;
  const int mu_x = (int)((double)joints[3 * bj] / stride_x + 0.5);        // int(): truncation toward zero
  const int mu_y = (int)((double)joints[3 * bj + 1] / stride_y + 0.5);
  const int ul_x = mu_x - tmp, ul_y = mu_y - tmp, br_x = mu_x + tmp + 1, br_y = mu_y + tmp + 1;
  float w = vis[3 * bj];
  const bool outside = ul_x >= hw || ul_y >= hh || br_x < 0 || br_y < 0;
  if (outside) w = 0.f;
  if (threadIdx.x == 0) weight[bj] = w;
  if (outside || !(w > 0.5f)) return;
  const int size = 2 * tmp + 1;
  const int x_lo = ul_x > 0 ? ul_x : 0, x_hi = br_x < hw ? br_x : hw;
  const int y_lo = ul_y > 0 ? ul_y : 0, y_hi = br_y < hh ? br_y : hh;
  const float denom = (float)(2 * sigma * sigma);
  float* t = target + (int64_t)bj * hh * hw;
  const int nx = x_hi - x_lo, n = nx * (y_hi - y_lo);
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int yy = y_lo + e / nx, xx = x_lo + e % nx;
    const float gx = (float)(xx - ul_x - size / 2), gy = (float)(yy - ul_y - size / 2);
    t[(int64_t)yy * hw + xx] = expf(-(gx * gx + gy * gy) / denom);
  }
}

// uint8 HWC frames -> normalised fp32 NHWC (torchvision ToTensor + Normalize, datasets/transforms/build.py:13-22):
// ((u8 / 255) - mean[c]) / std[c], in that order of IEEE operations.  Frames are gathered with a source frame stride so
// the (batch, window) order of a loader becomes the frame-major order of Alignment_V15.py:115-119.
__global__ void frames_u8_normalize_kernel(const uint8_t* __restrict__ src, int64_t src_frame_stride, float* __restrict__ dst,
                                           int nframes, int64_t px_per_frame, float m0, float m1, float m2, float s0,
                                           float s1, float s2) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)nframes * px_per_frame) return;
  const int f = (int)(i / px_per_frame);
  const int64_t pix = i - (int64_t)f * px_per_frame;
  const uint8_t* s = src + (int64_t)f * src_frame_stride + pix * 3;
  float* d = dst + i * 3;
  d[0] = ((float)s[0] / 255.f - m0) / s0;
  d[1] = ((float)s[1] / 255.f - m1) / s1;
  d[2] = ((float)s[2] / 255.f - m2) / s2;
}

// cv2.warpAffine(frame_u8, trans, (Wd, Hd), flags=INTER_LINEAR), constant-0 border, bit for bit (OpenCV's fixed-point path;
// datasets/zoo/posetrack/PoseTrack_Alignment.py:235-241, affine_transform.py:76-82).  minv[f][6]: the INVERSE affine map of
// frame f in double, as cv::invertAffineTransform computes it (done by the caller in float64).  Source coordinates:
// X = (rint((m1*y + m2)*1024) + 16 + rint(m0*x*1024)) >> 5 (5 fractional bits), likewise Y; weights (32-fy|fy)*(32-fx|fx)*32
// sum to 2^15; out = (sum w*p + 2^14) >> 15.  The double products are formed with explicit round-to-nearest multiplies and
// adds (no fused contraction), as the scalar C++ does.  One thread per destination pixel (3 channels).
// kNorm: write ToTensor + Normalize(mean, std) of that 8-bit value as float (datasets/transforms/build.py:13-22) instead.
template <bool kNorm>
__global__ void crop_affine_u8_kernel(const uint8_t* __restrict__ src, int64_t src_frame_stride, int Hs, int Ws,
                                      const double* __restrict__ minv, void* __restrict__ dst, int nframes, int Hd, int Wd,
                                      float m0, float m1, float m2, float s0, float s1, float s2) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)nframes * Hd * Wd) return;
  const int x = (int)(i % Wd);
  const int y = (int)((i / Wd) % Hd);
  const int f = (int)(i / ((int64_t)Wd * Hd));
  const double* m = minv + f * 6;
  const long long ad = __double2ll_rn(__dmul_rn(__dmul_rn(m[0], (double)x), 1024.0));
  const long long bd = __double2ll_rn(__dmul_rn(__dmul_rn(m[3], (double)x), 1024.0));
  const long long X0 = __double2ll_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[1], (double)y), m[2]), 1024.0)) + 16;
  const long long Y0 = __double2ll_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[4], (double)y), m[5]), 1024.0)) + 16;
  // cv::saturate_cast<int> of the rounded products, then int arithmetic
  auto sat = [](long long v) { return (int)(v > 2147483647ll ? 2147483647ll : (v < -2147483648ll ? -2147483648ll : v)); };
  const int X = (sat(X0 - 16) + 16 + sat(ad)) >> 5, Y = (sat(Y0 - 16) + 16 + sat(bd)) >> 5;
  int ix = X >> 5, iy = Y >> 5;
  // cv::saturate_cast<short> of the integer coordinates (remap stores them as 16-bit)
  ix = ix > 32767 ? 32767 : (ix < -32768 ? -32768 : ix);
  iy = iy > 32767 ? 32767 : (iy < -32768 ? -32768 : iy);
  const int fx = X & 31, fy = Y & 31;
  const int w00 = (32 - fy) * (32 - fx) * 32, w01 = (32 - fy) * fx * 32, w10 = fy * (32 - fx) * 32, w11 = fy * fx * 32;
  const uint8_t* s = src + (int64_t)f * src_frame_stride;
  int acc[3] = {16384, 16384, 16384};
  auto tap = [&](int yy, int xx, int w) {
    if (w != 0 && yy >= 0 && yy < Hs && xx >= 0 && xx < Ws) {
      const uint8_t* p = s + ((int64_t)yy * Ws + xx) * 3;
      acc[0] += w * p[0]; acc[1] += w * p[1]; acc[2] += w * p[2];
    }
  };
  tap(iy, ix, w00); tap(iy, ix + 1, w01); tap(iy + 1, ix, w10); tap(iy + 1, ix + 1, w11);
  const int v0 = acc[0] >> 15, v1 = acc[1] >> 15, v2 = acc[2] >> 15;
  if (kNorm) {
    float* d = reinterpret_cast<float*>(dst) + i * 3;
    d[0] = ((float)v0 / 255.f - m0) / s0;
    d[1] = ((float)v1 / 255.f - m1) / s1;
    d[2] = ((float)v2 / 255.f - m2) / s2;
  } else {
    uint8_t* d = reinterpret_cast<uint8_t*>(dst) + i * 3;
    d[0] = (uint8_t)v0; d[1] = (uint8_t)v1; d[2] = (uint8_t)v2;
  }
}

}  // namespace

int crop_affine_u8_launch(const uint8_t* src, int64_t src_frame_stride, int Hs, int Ws, const double* minv, void* dst,
                          int nframes, int Hd, int Wd, const float* mean, const float* std, cudaStream_t st) {
  const int64_t tot = (int64_t)nframes * Hd * Wd;
  if (mean && std)
    crop_affine_u8_kernel<true><<<cdiv(tot, 256), 256, 0, st>>>(src, src_frame_stride, Hs, Ws, minv, dst, nframes, Hd, Wd, mean[0],
                                                                mean[1], mean[2], std[0], std[1], std[2]);
  else
    crop_affine_u8_kernel<false><<<cdiv(tot, 256), 256, 0, st>>>(src, src_frame_stride, Hs, Ws, minv, dst, nframes, Hd, Wd, 0.f, 0.f,
                                                                 0.f, 1.f, 1.f, 1.f);
  FAMI_CHECK_LAUNCH("crop_affine_u8_kernel");
  return 0;
}

int frames_u8_normalize_launch(const uint8_t* src, int64_t src_frame_stride, float* dst, int nframes, int64_t px_per_frame,
                               const float* mean, const float* std, cudaStream_t st) {
  const int64_t tot = (int64_t)nframes * px_per_frame;
  frames_u8_normalize_kernel<<<cdiv(tot, 256), 256, 0, st>>>(src, src_frame_stride, dst, nframes, px_per_frame, mean[0], mean[1],
                                                             mean[2], std[0], std[1], std[2]);
  FAMI_CHECK_LAUNCH("frames_u8_normalize_kernel");
  return 0;
}

#define DISPATCH_T(dtype, ...)                                                 \
  if ((dtype) == FAMI_F32) { using T = float; __VA_ARGS__ }                    \
  else if ((dtype) == FAMI_F16) { using T = __half; __VA_ARGS__ }              \
  else { using T = __nv_bfloat16; __VA_ARGS__ }

int final_preds_launch(const void* hm, int dt, int pitch, const int32_t* idx, const float* maxv, const float* center,
                       const float* scale, float* preds, int B, int H, int W, int J, cudaStream_t st) {
  DISPATCH_T(dt, final_preds_kernel<T><<<cdiv(B * J, 128), 128, 0, st>>>((const T*)hm, pitch, idx, maxv, center, scale,
                                                                          preds, B, H, W, J);)
  FAMI_CHECK_LAUNCH("final_preds_kernel");
  return 0;
}

int pck_accuracy_launch(const int32_t* pidx, const float* pmax, const int32_t* tidx, const float* tmax, double* out, int B,
                        int H, int W, int J, float thr, cudaStream_t st) {
  const int threads = ((J + 31) / 32) * 32;
  pck_accuracy_kernel<<<1, threads, (size_t)J * sizeof(double), st>>>(pidx, pmax, tidx, tmax, out, B, H, W, J, (double)thr);
  FAMI_CHECK_LAUNCH("pck_accuracy_kernel");
  return 0;
}

int gaussian_targets_launch(const float* joints, const float* vis, float* target, float* weight, int B, int J, int sigma,
                            int img_w, int img_h, int hm_w, int hm_h, cudaStream_t st) {
  cudaMemsetAsync(target, 0, sizeof(float) * (size_t)B * J * hm_h * hm_w, st);
  gaussian_targets_kernel<<<B * J, 128, 0, st>>>(joints, vis, target, weight, J, sigma, (double)img_w / (double)hm_w,
                                                 (double)img_h / (double)hm_h, hm_w, hm_h);
  FAMI_CHECK_LAUNCH("gaussian_targets_kernel");
  return 0;
}

}  // namespace fami
