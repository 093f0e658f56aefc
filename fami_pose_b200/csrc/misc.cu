// Bandwidth-bound pieces of the FAMI-Pose hot path: layout conversion, global translation warp,
// frame difference, tiny linear head, train-mode BN finalize/apply, losses, keypoint argmax.
// Reference call sites are cited at each launcher (paths relative to the reference tree).
#include "common.cuh"

namespace fami {

// ---------------------------------------------------------------------------------------------
// NCHW fp32 <-> NHWC {fp32,bf16}
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, int64_t src_n_stride, T* __restrict__ dst,
                                    int N, int C, int HW, int pitch) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // pixel index over N*HW
  if (i >= (int64_t)N * HW) return;
  int n = (int)(i / HW);
  int pix = (int)(i - (int64_t)n * HW);
  const float* s = src + n * src_n_stride + pix;
  T* d = dst + i * pitch;
  for (int c = 0; c < C; ++c) d[c] = from_f<T>(__ldg(s + (int64_t)c * HW));
}

// C >= 8: the per-pixel loop above stores 4 (or 2) bytes per lane at a stride of `pitch` elements -- 32 sectors per store
// instruction, an eighth of each used (229 us for a [32,48,96,72] gradient).  Here a block moves 64 pixels x 32 channels
// through shared memory: loads run along the pixels of a channel plane, stores along the channels of a pixel.
template <typename T>
__global__ void __launch_bounds__(256) nchw_to_nhwc_tiled_kernel(const float* __restrict__ src, int64_t src_n_stride,
                                                                 T* __restrict__ dst, int N, int C, int HW, int pitch) {
  __shared__ float tile[32][65];
  const int64_t tot = (int64_t)N * HW;
  const int64_t pix0 = (int64_t)blockIdx.x * 64;
  const int tid = threadIdx.x;
  const int lp = tid & 63;                         // pixel of this thread's loads
  const int64_t li = pix0 + lp;
  const int ln = li < tot ? (int)(li / HW) : 0;
  const float* lsrc = src + (int64_t)ln * src_n_stride + (li - (int64_t)ln * HW);
  for (int c0 = 0; c0 < C; c0 += 32) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int cl = (tid >> 6) + 4 * k;
      tile[cl][lp] = (li < tot && c0 + cl < C) ? __ldg(lsrc + (int64_t)(c0 + cl) * HW) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int e = tid + 256 * k;
      const int cl = e & 31, pl = e >> 5;
      if (pix0 + pl < tot && c0 + cl < C) dst[(pix0 + pl) * pitch + c0 + cl] = from_f<T>(tile[cl][pl]);
    }
    __syncthreads();
  }
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ src, int pitch, float* __restrict__ dst, int N, int C,
                                    int HW) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)N * HW) return;
  int n = (int)(i / HW);
  int pix = (int)(i - (int64_t)n * HW);
  const T* s = src + i * pitch;
  float* d = dst + (int64_t)n * C * HW + pix;
  for (int c = 0; c < C; ++c) d[(int64_t)c * HW] = to_f<T>(s[c]);
}

// ---------------------------------------------------------------------------------------------
// translation warp: out[b,y,x,:] = bilinear_zero_pad(src[b], y - ty, x - tx)
// (kornia.geometry.warp_affine with M=[[1,0,tx],[0,1,ty]], Alignment_V15.py:133-135)
// one thread per (pixel, channel quad)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void warp_translate_fwd_kernel(const T* __restrict__ src, int src_pitch, const float* __restrict__ txy,
                                          T* __restrict__ out, int out_pitch, int B, int H, int W, int C4) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t tot = (int64_t)B * H * W * C4;
  if (i >= tot) return;
  int q = (int)(i % C4);
  int64_t pix = i / C4;
  int x = (int)(pix % W);
  int64_t t = pix / W;
  int y = (int)(t % H);
  int b = (int)(t / H);
  float tx = __ldg(txy + 2 * b), ty = __ldg(txy + 2 * b + 1);
  float px = (float)x - tx, py = (float)y - ty;
  float fx = floorf(px), fy = floorf(py);
  float lx = px - fx, ly = py - fy;
  // clamp before the int conversion so absurd translations cannot overflow
  fx = fminf(fmaxf(fx, -2.f), (float)W);
  fy = fminf(fmaxf(fy, -2.f), (float)H);
  int x0 = (int)fx, y0 = (int)fy;
  const T* sb = src + (int64_t)b * H * W * src_pitch + q * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int yy = y0 + (k >> 1), xx = x0 + (k & 1);
    float w = ((k >> 1) ? ly : 1.f - ly) * ((k & 1) ? lx : 1.f - lx);
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
      float4 v = ld4<T>(sb + ((int64_t)yy * W + xx) * src_pitch);
      acc.x += w * v.x; acc.y += w * v.y; acc.z += w * v.z; acc.w += w * v.w;
    }
  }
  st4<T>(out + pix * out_pitch + q * 4, acc);
}

// ---------------------------------------------------------------------------------------------
// out[r*n + i] = a[r*n + i] - b[i]   (Alignment_V15.py:132, all supporting frames at once)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void sub_bcast_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, int64_t n4,
                                 int rep) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4 * rep) return;
  int64_t j = i % n4;
  float4 va = ld4<T>(a + i * 4), vb = ld4<T>(b + j * 4);
  st4<T>(out + i * 4, make_float4(va.x - vb.x, va.y - vb.y, va.z - vb.z, va.w - vb.w));
}

// rows x cols strided copy (channel-slice writes of the reference's torch.cat along dim=1)
template <typename T>
__global__ void copy2d_kernel(const T* __restrict__ src, int sp, T* __restrict__ dst, int dp, int64_t rows, int c4) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * c4) return;
  int64_t r = i / c4;
  int q = (int)(i - r * c4);
  st4<T>(dst + r * dp + q * 4, ld4<T>(src + r * sp + q * 4));
}

// ---------------------------------------------------------------------------------------------
// y[M,N] = x[M,K] w[N,K]^T + b   (nn.Linear, Alignment_V15.py:69-71); one warp per output
// ---------------------------------------------------------------------------------------------
__global__ void linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                  const float* __restrict__ b, float* __restrict__ y, int M, int K, int N) {
  int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (gw >= M * N) return;
  int m = gw / N, n = gw - m * N;
  float s = 0.f;
  for (int k = lane; k < K; k += 32) s = fmaf(__ldg(x + (int64_t)m * K + k), __ldg(w + (int64_t)n * K + k), s);
  s = warp_sum(s);
  if (lane == 0) y[gw] = s + (b ? __ldg(b + n) : 0.f);
}

// ---------------------------------------------------------------------------------------------
// train-mode BN
// ---------------------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(const double* __restrict__ stats, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* running_mean, float* running_var,
                                   float* scale, float* shift, float* save_mean, float* save_invstd, int C,
                                   double count, float eps, float momentum) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double mean = stats[c] / count;
  double var = stats[C + c] / count - mean * mean;
  if (var < 0) var = 0;
  float invstd = (float)(1.0 / sqrt(var + (double)eps));
  float g = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
  scale[c] = g * invstd;
  shift[c] = be - (float)mean * g * invstd;
  if (save_mean) save_mean[c] = (float)mean;
  if (save_invstd) save_invstd[c] = invstd;
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
  if (running_var) {
    double unb = count > 1 ? var * count / (count - 1.0) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
  }
}

// per-channel sum / sum-of-squares over rows of an NHWC activation (train-mode BN on the tensor-core path)
template <typename T>
__global__ void __launch_bounds__(1024) bn_stats_kernel(const T* __restrict__ x, int pitch, int64_t rows, int C,
                                                        double* __restrict__ stats, int rows_per_block) {
  extern __shared__ float sm[];
  const int RG = blockDim.x / C;
  const int tid = threadIdx.x, c = tid % C, rg = tid / C;
  int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  // shifted sums (shift k = the channel's value in row 0, the same in every block): the float partial sums of (v - k) and
  // (v - k)^2 stay well conditioned when |mean| >> std (post-ReLU features, large conv biases); they are converted back
  // to sum / sum-of-squares of v in double per block
  float s = 0.f, q = 0.f;
  const float k = rg < RG ? to_f<T>(x[c]) : 0.f;
  if (rg < RG)
    for (int64_t r = r0 + rg; r < r1; r += RG) {
      float v = to_f<T>(x[r * pitch + c]) - k;
      s += v;
      q = fmaf(v, v, q);
    }
  if (rg < RG) { sm[rg * C + c] = s; sm[(RG + rg) * C + c] = q; }
  __syncthreads();
  if (tid < C && r1 > r0) {
    double ds = 0, dq = 0;
    for (int r = 0; r < RG; ++r) { ds += sm[r * C + tid]; dq += sm[(RG + r) * C + tid]; }
    const double kd = (double)k, n = (double)(r1 - r0);
    atomicAdd(stats + tid, ds + n * kd);
    atomicAdd(stats + C + tid, dq + 2.0 * kd * ds + n * kd * kd);
  }
}

// fp32 input, 4 channels per thread, two independent row streams per thread
__global__ void __launch_bounds__(1024) bn_stats_vec4_kernel(const float* __restrict__ x, int pitch, int64_t rows, int C,
                                                             double* __restrict__ stats, int rows_per_block) {
  extern __shared__ float sm[];
  const int cq = C >> 2;
  const int RG = blockDim.x / cq;
  const int tid = threadIdx.x, q = tid % cq, rg = tid / cq;
  int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  // per-thread sums of the shifted values in DOUBLE: the result does not depend on how the rows are cut into blocks and
  // row groups (to ~1e-16), so the grid can follow the SM count; the shared-memory hand-over stays double as well
  double s[4] = {0., 0., 0., 0.}, t[4] = {0., 0., 0., 0.};
  double* smd = reinterpret_cast<double*>(sm);
  if (rg < RG) {
    // shifted sums, see bn_stats_kernel
    const float4 k = __ldg(reinterpret_cast<const float4*>(x) + q);
    int64_t r = r0 + rg;
    for (; r + 3 * (int64_t)RG < r1; r += 4 * RG) {       // four row streams in flight per thread
      float4 a = __ldg(reinterpret_cast<const float4*>(x + r * pitch) + q);
      float4 b = __ldg(reinterpret_cast<const float4*>(x + (r + RG) * pitch) + q);
      float4 c = __ldg(reinterpret_cast<const float4*>(x + (r + 2 * RG) * pitch) + q);
      float4 d = __ldg(reinterpret_cast<const float4*>(x + (r + 3 * RG) * pitch) + q);
      a.x -= k.x; a.y -= k.y; a.z -= k.z; a.w -= k.w;
      b.x -= k.x; b.y -= k.y; b.z -= k.z; b.w -= k.w;
      c.x -= k.x; c.y -= k.y; c.z -= k.z; c.w -= k.w;
      d.x -= k.x; d.y -= k.y; d.z -= k.z; d.w -= k.w;
      s[0] += ((double)a.x + (double)b.x) + ((double)c.x + (double)d.x); s[1] += ((double)a.y + (double)b.y) + ((double)c.y + (double)d.y);
      s[2] += ((double)a.z + (double)b.z) + ((double)c.z + (double)d.z); s[3] += ((double)a.w + (double)b.w) + ((double)c.w + (double)d.w);
      t[0] += ((double)a.x * a.x + (double)b.x * b.x) + ((double)c.x * c.x + (double)d.x * d.x);
      t[1] += ((double)a.y * a.y + (double)b.y * b.y) + ((double)c.y * c.y + (double)d.y * d.y);
      t[2] += ((double)a.z * a.z + (double)b.z * b.z) + ((double)c.z * c.z + (double)d.z * d.z);
      t[3] += ((double)a.w * a.w + (double)b.w * b.w) + ((double)c.w * c.w + (double)d.w * d.w);
    }
    for (; r + RG < r1; r += 2 * RG) {
      float4 a = __ldg(reinterpret_cast<const float4*>(x + r * pitch) + q);
      float4 b = __ldg(reinterpret_cast<const float4*>(x + (r + RG) * pitch) + q);
      a.x -= k.x; a.y -= k.y; a.z -= k.z; a.w -= k.w;
      b.x -= k.x; b.y -= k.y; b.z -= k.z; b.w -= k.w;
      s[0] += (double)a.x + (double)b.x; s[1] += (double)a.y + (double)b.y; s[2] += (double)a.z + (double)b.z; s[3] += (double)a.w + (double)b.w;
      t[0] += (double)a.x * a.x + (double)b.x * b.x; t[1] += (double)a.y * a.y + (double)b.y * b.y;
      t[2] += (double)a.z * a.z + (double)b.z * b.z; t[3] += (double)a.w * a.w + (double)b.w * b.w;
    }
    if (r < r1) {
      float4 a = __ldg(reinterpret_cast<const float4*>(x + r * pitch) + q);
      a.x -= k.x; a.y -= k.y; a.z -= k.z; a.w -= k.w;
      s[0] += a.x; s[1] += a.y; s[2] += a.z; s[3] += a.w;
      t[0] += (double)a.x * a.x; t[1] += (double)a.y * a.y; t[2] += (double)a.z * a.z; t[3] += (double)a.w * a.w;
    }
    double* ps = smd + (rg * C + 4 * q);
    double* pq = smd + ((RG + rg) * C + 4 * q);
#pragma unroll
    for (int e = 0; e < 4; ++e) { ps[e] = s[e]; pq[e] = t[e]; }
  }
  __syncthreads();
  if (tid < C && r1 > r0) {
    double ds = 0, dq = 0;
    for (int r = 0; r < RG; ++r) { ds += smd[r * C + tid]; dq += smd[(RG + r) * C + tid]; }
    const double kd = (double)__ldg(x + tid), n = (double)(r1 - r0);
    atomicAdd(stats + tid, ds + n * kd);
    atomicAdd(stats + C + tid, dq + 2.0 * kd * ds + n * kd * kd);
  }
}

template <typename TI, typename T>
__global__ void bn_apply_act_kernel(const TI* __restrict__ x, int x_pitch, const float* __restrict__ scale,
                                    const float* __restrict__ shift, const T* __restrict__ res, int res_pitch,
                                    T* __restrict__ y, int y_pitch, int N, int Ho, int Wo, int C, int up, int relu) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t tot = (int64_t)N * Ho * Wo * C;
  if (i >= tot) return;
  int c = (int)(i % C);
  int64_t pix = i / C;
  float v = to_f<TI>(x[pix * x_pitch + c]) * __ldg(scale + c) + __ldg(shift + c);
  int xo = (int)(pix % Wo);
  int64_t t = pix / Wo;
  int yo = (int)(t % Ho);
  int n = (int)(t / Ho);
  int Hout = Ho * up, Wout = Wo * up;
  for (int dy = 0; dy < up; ++dy)
    for (int dx = 0; dx < up; ++dx) {
      int64_t op = ((int64_t)n * Hout + yo * up + dy) * Wout + xo * up + dx;
      float o = v;
      if (res) o += to_f<T>(res[op * res_pitch + c]);
      if (relu) o = fmaxf(o, 0.f);
      y[op * y_pitch + c] = from_f<T>(o);
    }
}

// 4 channels per thread (C, pitches multiples of 4, 16-byte-aligned bases): 16-byte loads of the raw conv
// output, 8/16-byte residual loads and stores
template <typename TI, typename T>
__global__ void bn_apply_act_vec4_kernel(const TI* __restrict__ x, int x_pitch, const float* __restrict__ scale,
                                         const float* __restrict__ shift, const T* __restrict__ res, int res_pitch,
                                         T* __restrict__ y, int y_pitch, int N, int Ho, int Wo, int C, int up, int relu) {
  const int cq = C >> 2;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t tot = (int64_t)N * Ho * Wo * cq;
  if (i >= tot) return;
  const int c = (int)(i % cq) << 2;
  const int64_t pix = i / cq;
  const float4 xv = ld4<TI>(x + pix * x_pitch + c);
  const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + c));
  const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + c));
  const float4 v = make_float4(fmaf(xv.x, sc.x, sh.x), fmaf(xv.y, sc.y, sh.y), fmaf(xv.z, sc.z, sh.z), fmaf(xv.w, sc.w, sh.w));
  if (up == 1) {
    float4 o = v;
    if (res) {
      const float4 r = ld4<T>(res + pix * res_pitch + c);
      o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    st4<T>(y + pix * y_pitch + c, o);
    return;
  }
  const int xo = (int)(pix % Wo);
  const int64_t t = pix / Wo;
  const int yo = (int)(t % Ho);
  const int n = (int)(t / Ho);
  const int Hout = Ho * up, Wout = Wo * up;
  for (int dy = 0; dy < up; ++dy)
    for (int dx = 0; dx < up; ++dx) {
      const int64_t op = ((int64_t)n * Hout + yo * up + dy) * Wout + xo * up + dx;
      float4 o = v;
      if (res) {
        const float4 r = ld4<T>(res + op * res_pitch + c);
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
      st4<T>(y + op * y_pitch + c, o);
    }
}

// up == 1: four 16-byte pieces per thread, all loads issued before the first use -- one piece per thread leaves 32 KB in
// flight per SM, which is 4.0 TB/s at HBM latency (measured: 159 us for the 637 MB of a [160,48,96,72] apply with residual)
template <typename TI, typename T>
__global__ void __launch_bounds__(256) bn_apply_act_vec4x4_kernel(const TI* __restrict__ x, int x_pitch,
                                                                  const float* __restrict__ scale, const float* __restrict__ shift,
                                                                  const T* __restrict__ res, int res_pitch, T* __restrict__ y,
                                                                  int y_pitch, int64_t tot, int C, int relu) {
  const int cq = C >> 2;
  const int64_t base = (int64_t)blockIdx.x * (256 * 4) + threadIdx.x;
  float4 xv[4], rv[4];
  int c[4];
  int64_t pix[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t i = base + k * 256;
    pix[k] = i / cq;
    c[k] = (int)(i - pix[k] * cq) << 2;
    rv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < tot) {
      xv[k] = ld4<TI>(x + pix[k] * x_pitch + c[k]);
      if (res) rv[k] = ld4<T>(res + pix[k] * res_pitch + c[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (base + k * 256 < tot) {
      const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + c[k]));
      const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + c[k]));
      float4 o = make_float4(fmaf(xv[k].x, sc.x, sh.x) + rv[k].x, fmaf(xv[k].y, sc.y, sh.y) + rv[k].y,
                             fmaf(xv[k].z, sc.z, sh.z) + rv[k].z, fmaf(xv[k].w, sc.w, sh.w) + rv[k].w);
      if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
      st4<T>(y + pix[k] * y_pitch + c[k], o);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// JointMSELoss (mse_loss.py:21-40): loss = 1/(J*B*HW) sum_{b,j,p} w_bj^2 (pred - gt)^2
// pred NHWC, target NCHW.  one thread per (b, pixel), loop over joints.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void joint_mse_kernel(const T* __restrict__ pred, int pitch, const float* __restrict__ target,
                                 const float* __restrict__ weight, float* loss_out, float* grad_pred,
                                 float grad_scale, int B, int J, int HW, float inv_count) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float s = 0.f;
  if (i < (int64_t)B * HW) {
    int b = (int)(i / HW);
    int pix = (int)(i - (int64_t)b * HW);
    for (int j = 0; j < J; ++j) {
      float w = weight ? __ldg(weight + b * J + j) : 1.f;
      float p = to_f<T>(pred[i * pitch + j]);
      float g = __ldg(target + ((int64_t)b * J + j) * HW + pix);
      float d = p * w - g * w;
      s = fmaf(d, d, s);
      if (grad_pred) grad_pred[i * J + j] = 2.f * d * w * inv_count * grad_scale;
    }
  }
  s = warp_sum(s);
  __shared__ float red[32];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) red[wid] = s;
  __syncthreads();
  if (wid == 0) {
    s = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
    s = warp_sum(s);
    if (lane == 0) atomicAdd(loss_out, s * inv_count);
  }
}

// ---------------------------------------------------------------------------------------------
// MI estimator core (Alignment_V15.py:250-277): per row r=(b,c) over L=HW positions
//   t = softmax(b/T), p = softmax(a/T);  value = mean_{r,l} t*(log t - p)   [reference quirk]
// One block per image; thread (rg, c): channel c = tid % C, row-group rg = tid / C strides over
// pixels.  Pass 1 column maxima, pass 2 partition sums (data re-read from L2).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(1024) softmax_pkl_kernel(const T* __restrict__ a, int a_pitch,
                                                           const T* __restrict__ bsrc, int b_pitch, float* out,
                                                           int HW, int C, float inv_temp, float inv_count) {
  extern __shared__ float sm[];  // [6][RG*C] scratch
  const int RG = blockDim.x / C;
  const int tid = threadIdx.x;
  const int c = tid % C, rg = tid / C;
  const bool active = rg < RG;
  const T* ab = a + (int64_t)blockIdx.x * HW * a_pitch;
  const T* bb = bsrc + (int64_t)blockIdx.x * HW * b_pitch;
  float ma = -INFINITY, mb = -INFINITY;
  if (active)
    for (int l = rg; l < HW; l += RG) {
      ma = fmaxf(ma, to_f<T>(ab[(int64_t)l * a_pitch + c]) * inv_temp);
      mb = fmaxf(mb, to_f<T>(bb[(int64_t)l * b_pitch + c]) * inv_temp);
    }
  float* s_ma = sm;
  float* s_mb = sm + RG * C;
  if (active) { s_ma[rg * C + c] = ma; s_mb[rg * C + c] = mb; }
  __syncthreads();
  if (active) {
    for (int r = 0; r < RG; ++r) { ma = fmaxf(ma, s_ma[r * C + c]); mb = fmaxf(mb, s_mb[r * C + c]); }
  }
  __syncthreads();
  float za = 0.f, zb = 0.f, sb = 0.f, xab = 0.f;
  if (active)
    for (int l = rg; l < HW; l += RG) {
      float va = to_f<T>(ab[(int64_t)l * a_pitch + c]) * inv_temp - ma;
      float vb = to_f<T>(bb[(int64_t)l * b_pitch + c]) * inv_temp - mb;
      float ea = expf(va), eb = expf(vb);
      za += ea; zb += eb; sb = fmaf(eb, vb, sb); xab = fmaf(ea, eb, xab);
    }
  float* s0 = sm; float* s1 = sm + RG * C; float* s2 = sm + 2 * RG * C; float* s3 = sm + 3 * RG * C;
  if (active) { s0[rg * C + c] = za; s1[rg * C + c] = zb; s2[rg * C + c] = sb; s3[rg * C + c] = xab; }
  __syncthreads();
  float val = 0.f;
  if (tid < C) {
    za = zb = sb = xab = 0.f;
    for (int r = 0; r < RG; ++r) { za += s0[r * C + tid]; zb += s1[r * C + tid]; sb += s2[r * C + tid]; xab += s3[r * C + tid]; }
    // sum_l t*(log t - p) = sb/zb - log zb - xab/(za*zb)
    val = sb / zb - logf(zb) - xab / (za * zb);
  }
  __syncthreads();
  // block reduce val over the first C threads
  val = warp_sum(val);
  if ((tid & 31) == 0) sm[tid >> 5] = val;
  __syncthreads();
  if (tid < 32) {
    float v = tid < (blockDim.x >> 5) ? sm[tid] : 0.f;
    v = warp_sum(v);
    if (tid == 0) atomicAdd(out, v * inv_count);
  }
}

// ---------------------------------------------------------------------------------------------
// get_max_preds (heatmaps_process.py:16-44): flat argmax per (b,j), first maximum wins
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(1024) argmax_hw_kernel(const T* __restrict__ hm, int pitch, int32_t* idx_out,
                                                         float* maxval_out, int HW, int J) {
  extern __shared__ float sm[];
  const int RG = blockDim.x / J;
  float* s_v = sm;
  int* s_i = reinterpret_cast<int*>(sm + RG * J);
  const int tid = threadIdx.x;
  const int j = tid % J, rg = tid / J;
  const T* hb = hm + (int64_t)blockIdx.x * HW * pitch;
  if (rg < RG) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int l = rg; l < HW; l += RG) {
      float v = to_f<T>(hb[(int64_t)l * pitch + j]);
      if (v > best || bi == 0x7fffffff) { best = v; bi = l; }
    }
    s_v[rg * J + j] = best;
    s_i[rg * J + j] = bi;
  }
  __syncthreads();
  if (tid < J) {
    float best = s_v[tid];
    int bi = s_i[tid];
    for (int r = 1; r < RG; ++r) {
      float v = s_v[r * J + tid];
      int ii = s_i[r * J + tid];
      if (ii != 0x7fffffff && (v > best || (v == best && ii < bi))) { best = v; bi = ii; }
    }
    idx_out[blockIdx.x * J + tid] = bi;
    maxval_out[blockIdx.x * J + tid] = best;
  }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
#define DISPATCH_T(dtype, ...)                                                 \
  if ((dtype) == FAMI_F32) { using T = float; __VA_ARGS__ }                    \
  else if ((dtype) == FAMI_F16) { using T = __half; __VA_ARGS__ }              \
  else { using T = __nv_bfloat16; __VA_ARGS__ }

int nchw_to_nhwc_launch(const float* src, int64_t sns, void* dst, int dt, int N, int C, int H, int W, int pitch,
                        cudaStream_t st) {
  int64_t tot = (int64_t)N * H * W;
  if (C >= 8) {
    DISPATCH_T(dt, nchw_to_nhwc_tiled_kernel<T><<<cdiv(tot, 64), 256, 0, st>>>(src, sns, (T*)dst, N, C, H * W, pitch);)
  } else {
    DISPATCH_T(dt, nchw_to_nhwc_kernel<T><<<cdiv(tot, 256), 256, 0, st>>>(src, sns, (T*)dst, N, C, H * W, pitch);)
  }
  FAMI_CHECK_LAUNCH("nchw_to_nhwc");
  return 0;
}
int nhwc_to_nchw_launch(const void* src, int dt, int pitch, float* dst, int N, int C, int H, int W, cudaStream_t st) {
  int64_t tot = (int64_t)N * H * W;
  DISPATCH_T(dt, nhwc_to_nchw_kernel<T><<<cdiv(tot, 256), 256, 0, st>>>((const T*)src, pitch, dst, N, C, H * W);)
  FAMI_CHECK_LAUNCH("nhwc_to_nchw");
  return 0;
}
int warp_translate_fwd_launch(const void* src, int sp, const float* txy, void* out, int op, int dt, int B, int H,
                              int W, int C, cudaStream_t st) {
  int64_t tot = (int64_t)B * H * W * (C / 4);
  DISPATCH_T(dt, warp_translate_fwd_kernel<T><<<cdiv(tot, 256), 256, 0, st>>>((const T*)src, sp, txy, (T*)out, op, B,
                                                                               H, W, C / 4);)
  FAMI_CHECK_LAUNCH("warp_translate_fwd");
  return 0;
}
int sub_bcast_launch(const void* a, const void* b, void* out, int dt, int64_t n, int rep, cudaStream_t st) {
  int64_t n4 = n / 4;
  DISPATCH_T(dt, sub_bcast_kernel<T><<<cdiv(n4 * rep, 256), 256, 0, st>>>((const T*)a, (const T*)b, (T*)out, n4, rep);)
  FAMI_CHECK_LAUNCH("sub_bcast");
  return 0;
}
int copy2d_launch(const void* src, int sp, void* dst, int dp, int dt, int64_t rows, int cols, cudaStream_t st) {
  int c4 = cols / 4;
  DISPATCH_T(dt, copy2d_kernel<T><<<cdiv(rows * c4, 256), 256, 0, st>>>((const T*)src, sp, (T*)dst, dp, rows, c4);)
  FAMI_CHECK_LAUNCH("copy2d");
  return 0;
}
int linear_fwd_launch(const float* x, const float* w, const float* b, float* y, int M, int K, int N, cudaStream_t st) {
  int64_t threads = (int64_t)M * N * 32;
  linear_fwd_kernel<<<cdiv(threads, 256), 256, 0, st>>>(x, w, b, y, M, K, N);
  FAMI_CHECK_LAUNCH("linear_fwd");
  return 0;
}
int bn_finalize_launch(const double* stats, const float* gamma, const float* beta, float* rm, float* rv, float* scale,
                       float* shift, float* save_mean, float* save_invstd, int C, int64_t count, float eps,
                       float momentum, cudaStream_t st) {
  bn_finalize_kernel<<<cdiv(C, 128), 128, 0, st>>>(stats, gamma, beta, rm, rv, scale, shift, save_mean, save_invstd, C,
                                                   (double)count, eps, momentum);
  FAMI_CHECK_LAUNCH("bn_finalize");
  return 0;
}
// plain storage cast fp32 -> 16 bit of a dense or pitched NHWC activation (fami_bn_apply_act with scale == shift == null):
// 8 channels per thread and iteration (two 16-byte loads in flight, one 16-byte store), grid-stride
template <typename T>
__global__ void __launch_bounds__(256) cast_f32_to_half_kernel(const float* __restrict__ x, int xp, T* __restrict__ y, int yp,
                                                               int64_t rows, int C) {
  const int c8 = C >> 3;
  const int64_t tot = rows * c8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / c8;
    const int c = (int)(i - r * c8) << 3;
    const float4 a = __ldg(reinterpret_cast<const float4*>(x + r * xp + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(x + r * xp + c + 4));
    T o[8] = {from_f<T>(a.x), from_f<T>(a.y), from_f<T>(a.z), from_f<T>(a.w), from_f<T>(b.x), from_f<T>(b.y), from_f<T>(b.z), from_f<T>(b.w)};
    *reinterpret_cast<uint4*>(y + r * yp + c) = *reinterpret_cast<const uint4*>(o);
  }
}

int bn_apply_act_launch(const void* x, int xdt, int xp, const float* scale, const float* shift, const void* res, int rp,
                        void* y, int yp, int dt, int N, int Ho, int Wo, int C, int up, int relu, cudaStream_t st) {
  int64_t tot = (int64_t)N * Ho * Wo * C;
  const int esz = dt == FAMI_F32 ? 4 : 2;
  if (!scale && !shift) {
    // identity affine = storage cast; the fast path covers what the deformable convolutions of the tf32 arm need
    FAMI_CHECK_ARG(!res && up == 1 && !relu, "fami_bn_apply_act: null scale / shift means a plain cast (no residual, up, relu)");
    FAMI_CHECK_ARG(xdt == FAMI_F32 && dt != FAMI_F32 && C % 8 == 0 && xp % 4 == 0 && yp % 8 == 0 &&
                       (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0,
                   "fami_bn_apply_act: the cast path takes fp32 -> 16 bit, C %% 8 == 0, 16-byte aligned rows");
    const int64_t rows = (int64_t)N * Ho * Wo;
    const int64_t work = rows * (C >> 3);
    int grid = (int)((work + 255) / 256);
    const int cap = num_sms() * 16;
    if (grid > cap) grid = cap;
    if (dt == FAMI_F16) cast_f32_to_half_kernel<__half><<<grid, 256, 0, st>>>((const float*)x, xp, (__half*)y, yp, rows, C);
    else cast_f32_to_half_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const float*)x, xp, (__nv_bfloat16*)y, yp, rows, C);
    FAMI_CHECK_LAUNCH("cast_f32_to_half");
    return 0;
  }
  const bool vec = xdt == FAMI_F32 && C % 4 == 0 && xp % 4 == 0 && yp % 4 == 0 && (!res || rp % 4 == 0) &&
                   (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & (4 * esz - 1)) == 0 &&
                   (!res || (reinterpret_cast<uintptr_t>(res) & (4 * esz - 1)) == 0) &&
                   (reinterpret_cast<uintptr_t>(scale) & 15) == 0 && (reinterpret_cast<uintptr_t>(shift) & 15) == 0;
  if (vec && up == 1) {
    DISPATCH_T(dt, bn_apply_act_vec4x4_kernel<float, T><<<cdiv(tot / 4, 1024), 256, 0, st>>>(
                       (const float*)x, xp, scale, shift, (const T*)res, rp, (T*)y, yp, tot / 4, C, relu);)
  } else if (vec) {
    DISPATCH_T(dt, bn_apply_act_vec4_kernel<float, T><<<cdiv(tot / 4, 256), 256, 0, st>>>(
                       (const float*)x, xp, scale, shift, (const T*)res, rp, (T*)y, yp, N, Ho, Wo, C, up, relu);)
  } else if (xdt == FAMI_F32) {
    DISPATCH_T(dt, bn_apply_act_kernel<float, T><<<cdiv(tot, 256), 256, 0, st>>>(
                       (const float*)x, xp, scale, shift, (const T*)res, rp, (T*)y, yp, N, Ho, Wo, C, up, relu);)
  } else if (xdt == FAMI_F16) {
    DISPATCH_T(dt, bn_apply_act_kernel<__half, T><<<cdiv(tot, 256), 256, 0, st>>>(
                       (const __half*)x, xp, scale, shift, (const T*)res, rp, (T*)y, yp, N, Ho, Wo, C, up, relu);)
  } else {
    DISPATCH_T(dt, bn_apply_act_kernel<__nv_bfloat16, T><<<cdiv(tot, 256), 256, 0, st>>>(
                       (const __nv_bfloat16*)x, xp, scale, shift, (const T*)res, rp, (T*)y, yp, N, Ho, Wo, C, up, relu);)
  }
  FAMI_CHECK_LAUNCH("bn_apply_act");
  return 0;
}
int bn_stats_launch(const void* x, int dt, int pitch, int64_t rows, int C, double* stats, cudaStream_t st) {
  if (dt == FAMI_F32 && C % 4 == 0 && pitch % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && C <= 1024) {
    const int cq = C / 4;
    int threads = 512;
    if (cq > threads) threads = ((cq + 31) / 32) * 32;
    const int RG = threads / cq;
    // at least two blocks per SM: the low-resolution branches (12x9x384 at N = 160: 17 K rows) ran on 17 blocks of 1024 rows
    int rows_per_block = (int)(rows / (2 * num_sms()));
    rows_per_block = rows_per_block > 1024 ? 1024 : (rows_per_block < 4 * RG ? 4 * RG : rows_per_block);
    const size_t smem = (size_t)2 * RG * C * sizeof(double);
    bn_stats_vec4_kernel<<<cdiv(rows, rows_per_block), threads, smem, st>>>((const float*)x, pitch, rows, C, stats,
                                                                           rows_per_block);
    FAMI_CHECK_LAUNCH("bn_stats_vec4");
    return 0;
  }
  int threads = 256;
  if (C > threads) threads = ((C + 31) / 32) * 32;
  int RG = threads / C;
  int rows_per_block = 2048;
  size_t smem = (size_t)2 * RG * C * sizeof(float);
  DISPATCH_T(dt, bn_stats_kernel<T><<<cdiv(rows, rows_per_block), threads, smem, st>>>((const T*)x, pitch, rows, C, stats,
                                                                                       rows_per_block);)
  FAMI_CHECK_LAUNCH("bn_stats");
  return 0;
}
int joint_mse_launch(const void* pred, int dt, int pitch, const float* target, const float* weight, float* loss,
                     float* grad, float gscale, int B, int J, int H, int W, cudaStream_t st) {
  int64_t tot = (int64_t)B * H * W;
  float inv = 1.f / ((float)J * (float)B * (float)(H * W));
  DISPATCH_T(dt, joint_mse_kernel<T><<<cdiv(tot, 256), 256, 0, st>>>((const T*)pred, pitch, target, weight, loss, grad,
                                                                      gscale, B, J, H * W, inv);)
  FAMI_CHECK_LAUNCH("joint_mse");
  return 0;
}
int softmax_pkl_launch(const void* a, int ap, const void* b, int bp, int dt, float* out, int B, int HW, int C,
                       float temperature, cudaStream_t st) {
  int threads = 1024;
  int RG = threads / C;
  size_t smem = (size_t)4 * RG * C * sizeof(float);
  if (smem < 32 * sizeof(float)) smem = 32 * sizeof(float);
  float inv_count = 1.f / ((float)B * (float)C * (float)HW);
  DISPATCH_T(dt, softmax_pkl_kernel<T><<<B, threads, smem, st>>>((const T*)a, ap, (const T*)b, bp, out, HW, C,
                                                                 1.f / temperature, inv_count);)
  FAMI_CHECK_LAUNCH("softmax_pkl");
  return 0;
}
int argmax_hw_launch(const void* hm, int dt, int pitch, int32_t* idx, float* maxv, int B, int HW, int J,
                     cudaStream_t st) {
  int threads = 1024;
  int RG = threads / J;
  size_t smem = (size_t)2 * RG * J * sizeof(float);
  DISPATCH_T(dt, argmax_hw_kernel<T><<<B, threads, smem, st>>>((const T*)hm, pitch, idx, maxv, HW, J);)
  FAMI_CHECK_LAUNCH("argmax_hw");
  return 0;
}

}  // namespace fami
