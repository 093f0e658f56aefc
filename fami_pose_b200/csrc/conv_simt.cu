// fp32 implicit-GEMM convolution with fused (folded-BN scale/shift | bias) + residual + ReLU
// + nearest-upsample-on-write epilogue.  Exact fp32 FMA arithmetic: this is the "fp32 parity"
// arm of fami_conv2d_bn_act_fwd and the fallback for shapes the tcgen05 path does not take.
//
// Replaces (reference): nn.Conv2d/BatchNorm2d/ReLU chains in posetimation/layers/basic_model.py:44-63,
// :83-113, basic_layer.py:55-73 and posetimation/backbones/hrnet.py:89-172,651-680.
//
// GEMM view: M = N*Ho*Wo output pixels, N = Cout, K = kh*kw*Cin.  NHWC activations make each
// (pixel, tap) a contiguous run of Cin floats, so the A tile is gathered with 16-byte cp.async
// (zero-filled at the padding halo) straight into shared memory; weights are pre-packed
// [K][CoutPad] so the B tile is a plain 2-D copy.  3-stage cp.async pipeline, 8x4 register tile.
#include <stdlib.h>

#include "common.cuh"

namespace fami {

struct ConvParams {
  const void* x;      // float (conv modes) or TX (DCN mode)
  const float* w;
  const float* scale;
  const float* shift;
  const void* res;    // TO
  void* y;            // TO
  double* stats;
  int N, H, W, Cin, Cout, CoutPad, kh, kw, stride, pad, dil, Ho, Wo, up, relu;
  int in_pitch, out_pitch, res_pitch;
  int M;       // N*Ho*Wo
  int Ktot;    // kh*kw*Cin
  int ksteps;  // ceil(Ktot/16)
  int vec_store;  // y / residual pointers and pitches allow 16B accesses
  // DCN mode (MODE_DCN): x is sampled bilinearly at offset positions and modulated by mask
  const float* off;
  const float* mask;
  int off_pitch, mask_pitch, cpg;
};

enum { MODE_VEC = 0, MODE_SCALAR = 1, MODE_DCN = 2 };

template <int WM, int WN, int MODE, typename TX, typename TO>
__global__ void __launch_bounds__(WM * WN * 32) conv_f32_kernel(const ConvParams p) {
  const TX* __restrict__ xptr = reinterpret_cast<const TX*>(p.x);
  const float* __restrict__ xptrf = reinterpret_cast<const float*>(p.x);
  const TO* __restrict__ rptr = reinterpret_cast<const TO*>(p.res);
  TO* __restrict__ yptr = reinterpret_cast<TO*>(p.y);
  constexpr int BM = WM * 64, BN = WN * 16, KC = 16, NT = WM * WN * 32, AS = KC + 4, STAGES = 3;
  constexpr int A_CH = (BM * 4 + NT - 1) / NT;        // 16B chunks of the A tile per thread
  constexpr int B_CH = (KC * BN / 4 + NT - 1) / NT;   // 16B chunks of the B tile per thread
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                         // [STAGES][BM][AS]
  float* Bs = smem + STAGES * BM * AS;      // [STAGES][KC][BN]

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int wm = warp / WN, wn = warp % WN;
  const int lm = lane >> 2, ln = lane & 3;
  const int m_blk = blockIdx.x * BM;
  const int n_blk = blockIdx.y * BN;
  const int HoWo = p.Ho * p.Wo;

  // ---- per-thread gather metadata for its A chunks (fixed over the K loop) ------------------
  int a_iy0[A_CH], a_ix0[A_CH];
  int64_t a_img[A_CH];
  bool a_ok[A_CH];
#pragma unroll
  for (int i = 0; i < A_CH; ++i) {
    int c = tid + i * NT;
    int m = m_blk + (c >> 2);
    bool ok = (c < BM * 4) && (m < p.M);
    int mm = ok ? m : 0;
    int n = mm / HoWo;
    int r = mm - n * HoWo;
    int yo = r / p.Wo, xo = r - yo * p.Wo;
    a_iy0[i] = yo * p.stride - p.pad;
    a_ix0[i] = xo * p.stride - p.pad;
    a_img[i] = (int64_t)n * p.H * p.W;
    a_ok[i] = ok;
  }
  const int cpt = p.Cin >> 4;  // 16-channel chunks per tap (VEC path)

  auto load_stage = [&](int s, int ks) {
    float* as = As + s * BM * AS;
    float* bs = Bs + s * KC * BN;
    if (MODE == MODE_DCN) {
      // modulated deformable gather (torchvision deform_conv2d semantics):
      // 16 channels (4 quads) of one tap per k-step; stride 1 so the output pixel index is m.
      int tap = ks / cpt;
      int c0 = (ks - tap * cpt) << 4;
      int r = tap / p.kw, sx = tap - r * p.kw;
      int dy = r * p.dil, dx = sx * p.dil;
#pragma unroll
      for (int i = 0; i < A_CH; ++i) {
        int c = tid + i * NT;
        if (c < BM * 4) {
          float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
          if (a_ok[i]) {
            int ch = c0 + ((c & 3) << 2);
            int g = ch / p.cpg;
            int64_t pix = (int64_t)m_blk + (c >> 2);
            const float* op = p.off + pix * p.off_pitch + g * 2 * p.kh * p.kw + 2 * tap;
            float ody = __ldg(op), odx = __ldg(op + 1);
            float mk = __ldg(p.mask + pix * p.mask_pitch + g * p.kh * p.kw + tap);
            float py = (float)(a_iy0[i] + dy) + ody;
            float px = (float)(a_ix0[i] + dx) + odx;
            if (py > -1.f && py < (float)p.H && px > -1.f && px < (float)p.W) {
              int y0 = (int)floorf(py), x0 = (int)floorf(px);
              float ly = py - (float)y0, lx = px - (float)x0;
              float hy = 1.f - ly, hx = 1.f - lx;
              const TX* xb = xptr + a_img[i] * p.in_pitch + ch;
              const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
              bool y0ok = y0 >= 0, y1ok = y0 + 1 <= p.H - 1, x0ok = x0 >= 0, x1ok = x0 + 1 <= p.W - 1;
              float4 v1 = (y0ok && x0ok) ? ld4<TX>(xb + ((int64_t)y0 * p.W + x0) * p.in_pitch) : z;
              float4 v2 = (y0ok && x1ok) ? ld4<TX>(xb + ((int64_t)y0 * p.W + x0 + 1) * p.in_pitch) : z;
              float4 v3 = (y1ok && x0ok) ? ld4<TX>(xb + ((int64_t)(y0 + 1) * p.W + x0) * p.in_pitch) : z;
              float4 v4 = (y1ok && x1ok) ? ld4<TX>(xb + ((int64_t)(y0 + 1) * p.W + x0 + 1) * p.in_pitch) : z;
              float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
              val.x = mk * (w1 * v1.x + w2 * v2.x + w3 * v3.x + w4 * v4.x);
              val.y = mk * (w1 * v1.y + w2 * v2.y + w3 * v3.y + w4 * v4.y);
              val.z = mk * (w1 * v1.z + w2 * v2.z + w3 * v3.z + w4 * v4.z);
              val.w = mk * (w1 * v1.w + w2 * v2.w + w3 * v3.w + w4 * v4.w);
            }
          }
          *reinterpret_cast<float4*>(as + (c >> 2) * AS + ((c & 3) << 2)) = val;
        }
      }
    } else if (MODE == MODE_VEC) {
      int tap = ks / cpt;
      int c0 = (ks - tap * cpt) << 4;
      int r = tap / p.kw, sx = tap - r * p.kw;
      int dy = r * p.dil, dx = sx * p.dil;
#pragma unroll
      for (int i = 0; i < A_CH; ++i) {
        int c = tid + i * NT;
        if (c < BM * 4) {
          int iy = a_iy0[i] + dy, ix = a_ix0[i] + dx;
          bool ok = a_ok[i] && (unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W;
          const float* src = ok ? xptrf + (a_img[i] + (int64_t)iy * p.W + ix) * p.in_pitch + c0 + ((c & 3) << 2) : xptrf;
          cp_async16(as + (c >> 2) * AS + ((c & 3) << 2), src, ok);
        }
      }
    } else {
      // generic scalar gather over the flattened K axis (stem conv, Cin = 3)
      for (int e = tid; e < BM * KC; e += NT) {
        int ml = e >> 4, kk = e & 15;
        int m = m_blk + ml;
        int k = ks * KC + kk;
        float v = 0.f;
        if (m < p.M && k < p.Ktot) {
          int tap = k / p.Cin, ci = k - tap * p.Cin;
          int r = tap / p.kw, sx = tap - r * p.kw;
          int n = m / HoWo, rr = m - n * HoWo;
          int yo = rr / p.Wo, xo = rr - yo * p.Wo;
          int iy = yo * p.stride - p.pad + r * p.dil, ix = xo * p.stride - p.pad + sx * p.dil;
          if ((unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W)
            v = __ldg(xptrf + ((int64_t)(n * p.H + iy) * p.W + ix) * p.in_pitch + ci);
        }
        as[ml * AS + kk] = v;
      }
    }
#pragma unroll
    for (int i = 0; i < B_CH; ++i) {
      int c = tid + i * NT;
      if (c < KC * BN / 4) {
        int row = c / (BN / 4), col = (c - row * (BN / 4)) << 2;
        const float* src = p.w + (int64_t)(ks * KC + row) * p.CoutPad + n_blk + col;
        cp_async16(bs + row * BN + col, src, true);
      }
    }
  };

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // prologue
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < p.ksteps) load_stage(s, s);
    cp_async_commit();
  }

  const int a_row0 = wm * 64 + lm * 4;  // rows a_row0..+3 and a_row0+32..+35
  const int b_col = wn * 16 + ln * 4;

  for (int ks = 0; ks < p.ksteps; ++ks) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      int nk = ks + STAGES - 1;
      if (nk < p.ksteps) load_stage(nk % STAGES, nk);
      cp_async_commit();
    }
    const float* as = As + (ks % STAGES) * BM * AS;
    const float* bs = Bs + (ks % STAGES) * KC * BN;
#pragma unroll
    for (int kq = 0; kq < 4; ++kq) {
      float4 a[8], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = *reinterpret_cast<const float4*>(as + (a_row0 + i) * AS + kq * 4);
        a[i + 4] = *reinterpret_cast<const float4*>(as + (a_row0 + 32 + i) * AS + kq * 4);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(bs + (kq * 4 + j) * BN + b_col);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[i][0] = fmaf(a[i].x, b[0].x, acc[i][0]); acc[i][1] = fmaf(a[i].x, b[0].y, acc[i][1]);
        acc[i][2] = fmaf(a[i].x, b[0].z, acc[i][2]); acc[i][3] = fmaf(a[i].x, b[0].w, acc[i][3]);
        acc[i][0] = fmaf(a[i].y, b[1].x, acc[i][0]); acc[i][1] = fmaf(a[i].y, b[1].y, acc[i][1]);
        acc[i][2] = fmaf(a[i].y, b[1].z, acc[i][2]); acc[i][3] = fmaf(a[i].y, b[1].w, acc[i][3]);
        acc[i][0] = fmaf(a[i].z, b[2].x, acc[i][0]); acc[i][1] = fmaf(a[i].z, b[2].y, acc[i][1]);
        acc[i][2] = fmaf(a[i].z, b[2].z, acc[i][2]); acc[i][3] = fmaf(a[i].z, b[2].w, acc[i][3]);
        acc[i][0] = fmaf(a[i].w, b[3].x, acc[i][0]); acc[i][1] = fmaf(a[i].w, b[3].y, acc[i][1]);
        acc[i][2] = fmaf(a[i].w, b[3].z, acc[i][2]); acc[i][3] = fmaf(a[i].w, b[3].w, acc[i][3]);
      }
    }
  }
  cp_async_wait<0>();

  // ---- epilogue ------------------------------------------------------------------------------
  const int n0 = n_blk + b_col;
  float sc[4], sh[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int o = n0 + j;
    sc[j] = (p.scale && o < p.Cout) ? __ldg(p.scale + o) : 1.f;
    sh[j] = (p.shift && o < p.Cout) ? __ldg(p.shift + o) : 0.f;
  }
  const bool vec_out = (n0 + 3 < p.Cout) && p.vec_store;
  const int Hout = p.Ho * p.up, Wout = p.Wo * p.up;
  float ssum[4] = {0, 0, 0, 0}, ssq[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m_blk + a_row0 + (i & 3) + ((i >> 2) << 5);
    if (m >= p.M) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = fmaf(acc[i][j], sc[j], sh[j]);
    if (p.stats) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { ssum[j] += v[j]; ssq[j] += v[j] * v[j]; }
    }
    int n = m / HoWo, r = m - n * HoWo;
    int yo = r / p.Wo, xo = r - yo * p.Wo;
    for (int dy = 0; dy < p.up; ++dy)
      for (int dx = 0; dx < p.up; ++dx) {
        int64_t pix = ((int64_t)n * Hout + yo * p.up + dy) * Wout + xo * p.up + dx;
        TO* yp = yptr + pix * p.out_pitch + n0;
        if (vec_out) {
          float4 o = make_float4(v[0], v[1], v[2], v[3]);
          if (rptr) {
            float4 rr = ld4<TO>(rptr + pix * p.res_pitch + n0);
            o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
          }
          if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
          st4<TO>(yp, o);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (n0 + j < p.Cout) {
              float o = v[j];
              if (rptr) o += to_f<TO>(rptr[pix * p.res_pitch + n0 + j]);
              if (p.relu) o = fmaxf(o, 0.f);
              yp[j] = from_f<TO>(o);
            }
          }
        }
      }
  }
  if (p.stats) {
    // reduce over the 8 lanes sharing ln (lane bits 2..4), then one double atomic per channel per warp
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        ssum[j] += __shfl_xor_sync(0xffffffffu, ssum[j], o);
        ssq[j] += __shfl_xor_sync(0xffffffffu, ssq[j], o);
      }
    }
    if (lm == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n0 + j < p.Cout) {
          atomicAdd(p.stats + n0 + j, (double)ssum[j]);
          atomicAdd(p.stats + p.Cout + n0 + j, (double)ssq[j]);
        }
    }
  }
}

template <int WM, int WN, int MODE, typename TX, typename TO>
static int launch_cfg(const ConvParams& p, cudaStream_t st) {
  constexpr int BM = WM * 64, BN = WN * 16;
  constexpr size_t smem = (size_t)3 * (BM * 20 + 16 * BN) * sizeof(float);
  static std::atomic<uint64_t> attr_mask{0};
  set_max_smem_once(attr_mask, conv_f32_kernel<WM, WN, MODE, TX, TO>, (int)smem);
  dim3 grid(cdiv(p.M, BM), p.CoutPad / BN);
  conv_f32_kernel<WM, WN, MODE, TX, TO><<<grid, WM * WN * 32, smem, st>>>(p);
  FAMI_CHECK_LAUNCH("conv_f32_kernel");
  return 0;
}

// ---- stem convolution (HRNet conv1: 3 -> 64, 3x3 stride 2; hrnet.py:651) ------------------------------
// K = 27 is too short for the implicit-GEMM machinery.  One thread computes 16 output channels of 4
// horizontally adjacent output pixels (64 accumulators): the 27 x 64 weights are read from shared
// memory once per 4 pixels, which keeps the kernel FMA-bound instead of LDS-bound; 4 threads cover the 64
// channels of a pixel so a warp writes 128-byte runs.
template <typename TO>
__global__ void __launch_bounds__(256) stem_conv_kernel(const ConvParams p) {
  __shared__ __align__(16) float sw[27 * 64];
  __shared__ float ssc[64], ssh[64];
  const float* __restrict__ x = reinterpret_cast<const float*>(p.x);
  TO* __restrict__ y = reinterpret_cast<TO*>(p.y);
  for (int i = threadIdx.x; i < 27 * 64; i += 256) sw[i] = p.w[i];
  if (threadIdx.x < 64) {
    ssc[threadIdx.x] = p.scale ? p.scale[threadIdx.x] : 1.f;
    ssh[threadIdx.x] = p.shift ? p.shift[threadIdx.x] : 0.f;
  }
  __syncthreads();
  const int q = threadIdx.x & 3;
  const int m0 = (blockIdx.x * 64 + (threadIdx.x >> 2)) * 4;   // first of 4 output pixels (same row: Wo % 4 == 0)
  if (m0 >= p.M) return;
  const int HoWo = p.Ho * p.Wo;
  const int n = m0 / HoWo, r = m0 - n * HoWo;
  const int yo = r / p.Wo, xo = r - yo * p.Wo;
  float acc[4][16];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
#pragma unroll
  for (int fr = 0; fr < 3; ++fr) {
    const int iy = yo * 2 - 1 + fr;
    const bool yok = (unsigned)iy < (unsigned)p.H;
    const float* row = x + (int64_t)(n * p.H + (yok ? iy : 0)) * p.W * p.in_pitch;
#pragma unroll
    for (int fs = 0; fs < 3; ++fs) {
      float in[4][3];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int ix = (xo + i) * 2 - 1 + fs;
        const bool ok = yok && (unsigned)ix < (unsigned)p.W;
        const float* px = row + (int64_t)(ok ? ix : 0) * p.in_pitch;
#pragma unroll
        for (int c = 0; c < 3; ++c) in[i][c] = ok ? __ldg(px + c) : 0.f;
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int k = (fr * 3 + fs) * 3 + c;
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 w = *reinterpret_cast<const float4*>(sw + k * 64 + q * 16 + j4 * 4);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float a = in[i][c];
            acc[i][j4 * 4 + 0] = fmaf(a, w.x, acc[i][j4 * 4 + 0]);
            acc[i][j4 * 4 + 1] = fmaf(a, w.y, acc[i][j4 * 4 + 1]);
            acc[i][j4 * 4 + 2] = fmaf(a, w.z, acc[i][j4 * 4 + 2]);
            acc[i][j4 * 4 + 3] = fmaf(a, w.w, acc[i][j4 * 4 + 3]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    TO* yp = y + (int64_t)(m0 + i) * p.out_pitch + q * 16;
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) {
      float4 o;
      o.x = fmaf(acc[i][j4 * 4 + 0], ssc[q * 16 + j4 * 4 + 0], ssh[q * 16 + j4 * 4 + 0]);
      o.y = fmaf(acc[i][j4 * 4 + 1], ssc[q * 16 + j4 * 4 + 1], ssh[q * 16 + j4 * 4 + 1]);
      o.z = fmaf(acc[i][j4 * 4 + 2], ssc[q * 16 + j4 * 4 + 2], ssh[q * 16 + j4 * 4 + 2]);
      o.w = fmaf(acc[i][j4 * 4 + 3], ssc[q * 16 + j4 * 4 + 3], ssh[q * 16 + j4 * 4 + 3]);
      if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
      st4<TO>(yp + j4 * 4, o);
    }
  }
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int MODE, typename TX, typename TO>
static int dispatch_tile(const ConvParams& p, cudaStream_t st) {
  const int cp = p.CoutPad;
  if (cp % 64 == 0) return launch_cfg<2, 4, MODE, TX, TO>(p, st);
  if (cp % 48 == 0) return launch_cfg<2, 3, MODE, TX, TO>(p, st);
  if (cp % 32 == 0) return launch_cfg<4, 2, MODE, TX, TO>(p, st);
  return launch_cfg<4, 1, MODE, TX, TO>(p, st);
}

// x: float NHWC.  y/residual: float (out_bf16 = 0) or bf16 (out_bf16 = 1, scalar-gather stem path only).
int stem_tc_supported(const fami_conv_desc* d, const void* y);
int stem_tc_launch(const fami_conv_desc* d, const float* x, const float* w, const float* scale, const float* shift, void* y,
                   cudaStream_t st);

int conv_f32_launch(const fami_conv_desc* d, const float* x, const float* w, const float* scale,
                    const float* shift, const void* res, void* y, double* stats, cudaStream_t st) {
  // HRNet stem (fp32 pixels -> 16-bit activations): tensor-core kernel with an in-CTA im2col (csrc/stem_tc.cu)
  static const bool stem_tc_off = getenv("FAMI_DISABLE_STEM_TC") != nullptr;
  // (fp32 output only on request -- the tf32 arm, see fami_conv2d_bn_act_fwd: a plain FAMI_F32 descriptor is the exact-fp32 arm)
  if (!stem_tc_off && !res && is_half_dtype(d->out_dtype) && stem_tc_supported(d, y)) return stem_tc_launch(d, x, w, scale, shift, y, st);
  ConvParams p;
  memset(&p, 0, sizeof(p));
  const bool out_half = is_half_dtype(d->out_dtype);
  const size_t osz = out_half ? 2 : 4;
  p.x = x; p.w = w; p.scale = scale; p.shift = shift; p.res = res; p.y = y; p.stats = d->stats ? stats : nullptr;
  p.N = d->N; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.Cout = d->Cout;
  p.CoutPad = fami_conv_cout_pad(d->Cout);
  p.kh = d->kh; p.kw = d->kw; p.stride = d->stride; p.pad = d->pad; p.dil = d->dil;
  p.Ho = d->Ho; p.Wo = d->Wo; p.up = d->up; p.relu = d->relu;
  p.in_pitch = d->in_pitch; p.out_pitch = d->out_pitch; p.res_pitch = d->res_pitch;
  p.M = d->N * d->Ho * d->Wo;
  p.Ktot = d->kh * d->kw * d->Cin;
  p.ksteps = (p.Ktot + 15) / 16;
  const size_t va = 4 * osz;  // bytes of a 4-channel vector access
  p.vec_store = (reinterpret_cast<uintptr_t>(y) % va == 0) && (d->out_pitch % 4 == 0) &&
                (!res || ((reinterpret_cast<uintptr_t>(res) % va == 0) && d->res_pitch % 4 == 0));
  const bool vec = (d->Cin % 16 == 0) && (d->in_pitch % 4 == 0) && al16(x);
  // HRNet stem: 3 -> 64, 3x3 stride 2 pad 1, no residual / statistics / upsample
  if (d->Cin == 3 && d->Cout == 64 && d->kh == 3 && d->stride == 2 && d->pad == 1 && d->dil == 1 && !res && !p.stats &&
      d->up == 1 && p.vec_store && d->out_pitch % 16 == 0 && d->Wo % 4 == 0) {
    const int grid = cdiv(p.M, 256);
    if (d->out_dtype == FAMI_F16) stem_conv_kernel<__half><<<grid, 256, 0, st>>>(p);
    else if (d->out_dtype == FAMI_BF16) stem_conv_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(p);
    else stem_conv_kernel<float><<<grid, 256, 0, st>>>(p);
    FAMI_CHECK_LAUNCH("stem_conv_kernel");
    return 0;
  }
  if (d->out_dtype == FAMI_BF16) return dispatch_tile<MODE_SCALAR, float, __nv_bfloat16>(p, st);
  if (d->out_dtype == FAMI_F16) return dispatch_tile<MODE_SCALAR, float, __half>(p, st);
  return vec ? dispatch_tile<MODE_VEC, float, float>(p, st) : dispatch_tile<MODE_SCALAR, float, float>(p, st);
}

// v1 DCN forward: same mainloop/epilogue as the conv, A tile produced by the deformable gather.
// x/out: float or bf16 (d->dtype); offset/mask/weights/bias: float.
int dcn_simt_launch(const fami_dcn_desc* d, const void* x, const float* off, const float* mask,
                    const float* w, const float* bias, void* out, cudaStream_t st) {
  ConvParams p;
  memset(&p, 0, sizeof(p));
  const bool bf = is_half_dtype(d->dtype);
  p.x = x; p.w = w; p.scale = nullptr; p.shift = bias; p.res = nullptr; p.y = out; p.stats = nullptr;
  p.N = d->B; p.H = d->H; p.W = d->W; p.Cin = d->C; p.Cout = d->Cout;
  p.CoutPad = fami_conv_cout_pad(d->Cout);
  p.kh = d->kh; p.kw = d->kw; p.stride = 1; p.pad = d->pad; p.dil = d->dil;
  p.Ho = d->H + 2 * d->pad - d->dil * (d->kh - 1); p.Wo = d->W + 2 * d->pad - d->dil * (d->kw - 1);
  p.up = 1; p.relu = 0;
  p.in_pitch = d->x_pitch; p.out_pitch = d->out_pitch; p.res_pitch = 0;
  p.M = d->B * p.Ho * p.Wo;
  p.Ktot = d->kh * d->kw * d->C;
  p.ksteps = p.Ktot / 16;
  p.vec_store = (reinterpret_cast<uintptr_t>(out) % (bf ? 8 : 16) == 0) && (d->out_pitch % 4 == 0);
  p.off = off; p.mask = mask; p.off_pitch = d->off_pitch; p.mask_pitch = d->mask_pitch;
  p.cpg = d->C / d->G;
  if (d->dtype == FAMI_BF16) return dispatch_tile<MODE_DCN, __nv_bfloat16, __nv_bfloat16>(p, st);
  if (d->dtype == FAMI_F16) return dispatch_tile<MODE_DCN, __half, __half>(p, st);
  return dispatch_tile<MODE_DCN, float, float>(p, st);
}

// ---- weight packing: OIHW float -> [kh*kw*Cin (padded to 16)][CoutPad] float ------------------
__global__ void pack_w_f32_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin,
                                  int taps, int Kpad, int CoutPad) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t tot = (int64_t)Kpad * CoutPad;
  if (i >= tot) return;
  int k = (int)(i / CoutPad), o = (int)(i - (int64_t)k * CoutPad);
  float v = 0.f;
  if (k < taps * Cin && o < Cout) {
    int tap = k / Cin, c = k - tap * Cin;
    v = w[((int64_t)o * Cin + c) * taps + tap];
  }
  out[i] = v;
}

int pack_w_f32_launch(const float* w, float* out, int Cout, int Cin, int kh, int kw, cudaStream_t st) {
  int taps = kh * kw, Kpad = ((taps * Cin + 15) / 16) * 16, CoutPad = fami_conv_cout_pad(Cout);
  int64_t tot = (int64_t)Kpad * CoutPad;
  pack_w_f32_kernel<<<cdiv(tot, 256), 256, 0, st>>>(w, out, Cout, Cin, taps, Kpad, CoutPad);
  FAMI_CHECK_LAUNCH("pack_w_f32_kernel");
  return 0;
}

}  // namespace fami
