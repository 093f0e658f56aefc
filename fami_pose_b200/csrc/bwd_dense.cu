// Backward kernels of the dense pieces of the trainable head (fp32, NHWC with pitch):
//   conv dgrad / wgrad      autograd of nn.Conv2d as used by BasicBlock (basic_model.py:44-63),
//                           conv_bn_relu (basic_layer.py:55-73) and the offset / mask / heatmap convs
//                           (Alignment_V15.py:79-106)
//   train/eval BatchNorm bwd (+ the ReLU mask of the fused activation, + the residual's gradient)
//   MI pseudo-KL backward   (Alignment_V15.py:250-277)
//   Linear backward         (feat_global_offset_layers[7..9], Alignment_V15.py:69-71)
// Stride-1 dgrad is the forward convolution of grad_out with the flipped, transposed filter (pad' =
// dil*(k-1) - pad): it runs on the forward kernels.  Everything else here is a straightforward SIMT kernel.
#include <stdlib.h>

#include "common.cuh"

namespace fami {

namespace {

// ---------------------------------------------------------------------------------------------
// flipped + transposed filter in OIHW order: wt[ci][co][kh-1-r][kw-1-s] = w[co][ci][r][s]
// ---------------------------------------------------------------------------------------------
__global__ void flip_transpose_kernel(const float* __restrict__ w, float* __restrict__ wt, int Cout, int Cin, int kh, int kw) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t tot = (int64_t)Cout * Cin * kh * kw;
  if (i >= tot) return;
  int s = (int)(i % kw);
  int64_t t = i / kw;
  int r = (int)(t % kh);
  t /= kh;
  int ci = (int)(t % Cin), co = (int)(t / Cin);
  wt[(((int64_t)ci * Cout + co) * kh + (kh - 1 - r)) * kw + (kw - 1 - s)] = w[i];
}

// ---------------------------------------------------------------------------------------------
// generic dgrad (any stride): thread = (input pixel, 4 input channels); wt is the fp32 packing of
// the flipped/transposed filter: [(tap' * Cout + co)][CinPad]
// ---------------------------------------------------------------------------------------------
__global__ void conv_dgrad_gather_kernel(const float* __restrict__ gy, int gy_pitch, const float* __restrict__ wt,
                                         float* __restrict__ gx, int gx_pitch, int N, int H, int W, int Cin, int Cout,
                                         int kh, int kw, int stride, int pad, int dil, int Ho, int Wo, int CinPad) {
  const int cq = (Cin + 3) / 4;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t tot = (int64_t)N * H * W * cq;
  if (i >= tot) return;
  const int c0 = (int)(i % cq) * 4;
  int64_t pix = i / cq;
  const int xi = (int)(pix % W);
  int64_t t = pix / W;
  const int yi = (int)(t % H), n = (int)(t / H);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int r = 0; r < kh; ++r) {
    const int yy = yi + pad - r * dil;
    if (yy < 0 || yy % stride) continue;
    const int yo = yy / stride;
    if (yo >= Ho) continue;
    for (int s = 0; s < kw; ++s) {
      const int xx = xi + pad - s * dil;
      if (xx < 0 || xx % stride) continue;
      const int xo = xx / stride;
      if (xo >= Wo) continue;
      const float* g = gy + ((int64_t)(n * Ho + yo) * Wo + xo) * gy_pitch;
      const int tapf = (kh - 1 - r) * kw + (kw - 1 - s);
      const float* wrow = wt + (int64_t)tapf * Cout * CinPad + c0;
      for (int co = 0; co < Cout; ++co) {
        const float gv = __ldg(g + co);
        const float4 wv = *reinterpret_cast<const float4*>(wrow + (int64_t)co * CinPad);   // CinPad % 4 == 0, zero padded
        acc[0] = fmaf(gv, wv.x, acc[0]); acc[1] = fmaf(gv, wv.y, acc[1]);
        acc[2] = fmaf(gv, wv.z, acc[2]); acc[3] = fmaf(gv, wv.w, acc[3]);
      }
    }
  }
  float* o = gx + pix * gx_pitch + c0;
  for (int j = 0; j < 4; ++j)
    if (c0 + j < Cin) o[j] = acc[j];
}

// ---------------------------------------------------------------------------------------------
// wgrad: dw[co][ci][r][s] += sum_pixels gy[p][co] * x[p @ tap][ci].  Block = one (tap, 64 ci, 64 co) tile
// over a chunk of output pixels; 256 threads, 4x4 accumulators each; fp32 atomics into dw (caller zeroes).
// ---------------------------------------------------------------------------------------------
constexpr int kWgPix = 32;
template <bool kVec>
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const float* __restrict__ x, int x_pitch,
                                                         const float* __restrict__ gy, int gy_pitch,
                                                         float* __restrict__ dw, int N, int H, int W, int Cin, int Cout,
                                                         int kh, int kw, int stride, int pad, int dil, int Ho, int Wo,
                                                         int ci_tiles, int co_tiles, int pix_per_block) {
  __shared__ __align__(16) float xs[kWgPix][64];
  __shared__ __align__(16) float gs[kWgPix][64];
  int tile = blockIdx.x;
  const int cot = tile % co_tiles; tile /= co_tiles;
  const int cit = tile % ci_tiles; tile /= ci_tiles;
  const int tap = tile;
  const int r = tap / kw, s = tap - r * kw;
  const int tid = threadIdx.x;
  const int tci = tid & 15, tco = tid >> 4;
  const int64_t npix = (int64_t)N * Ho * Wo;
  const int64_t p0 = (int64_t)blockIdx.y * pix_per_block;
  const int64_t p1 = p0 + pix_per_block < npix ? p0 + pix_per_block : npix;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  for (int64_t pb = p0; pb < p1; pb += kWgPix) {
    // stage kWgPix pixels x 64 channels of x (at this tap) and of gy
    if (kVec) {
      // 16-byte loads: thread -> (pixel, 4 channels); Cin, Cout and both pitches are multiples of 4
      for (int e = tid; e < kWgPix * 16; e += 256) {
        const int pl = e >> 4, c = (e & 15) << 2;
        const int64_t p = pb + pl;
        float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), gv = xv;
        if (p < p1) {
          const int xo = (int)(p % Wo);
          const int64_t t = p / Wo;
          const int yo = (int)(t % Ho), n = (int)(t / Ho);
          const int yi = yo * stride - pad + r * dil, xi = xo * stride - pad + s * dil;
          const int ci = cit * 64 + c, co = cot * 64 + c;
          if (ci < Cin && yi >= 0 && yi < H && xi >= 0 && xi < W)
            xv = __ldg(reinterpret_cast<const float4*>(x + ((int64_t)(n * H + yi) * W + xi) * x_pitch + ci));
          if (co < Cout) gv = __ldg(reinterpret_cast<const float4*>(gy + p * gy_pitch + co));
        }
        *reinterpret_cast<float4*>(&xs[pl][c]) = xv;
        *reinterpret_cast<float4*>(&gs[pl][c]) = gv;
      }
    } else {
      for (int e = tid; e < kWgPix * 64; e += 256) {
        const int pl = e >> 6, c = e & 63;
        const int64_t p = pb + pl;
        float xv = 0.f, gv = 0.f;
        if (p < p1) {
          const int xo = (int)(p % Wo);
          const int64_t t = p / Wo;
          const int yo = (int)(t % Ho), n = (int)(t / Ho);
          const int yi = yo * stride - pad + r * dil, xi = xo * stride - pad + s * dil;
          const int ci = cit * 64 + c, co = cot * 64 + c;
          if (ci < Cin && yi >= 0 && yi < H && xi >= 0 && xi < W) xv = __ldg(x + ((int64_t)(n * H + yi) * W + xi) * x_pitch + ci);
          if (co < Cout) gv = __ldg(gy + p * gy_pitch + co);
        }
        xs[pl][c] = xv;
        gs[pl][c] = gv;
      }
    }
    __syncthreads();
#pragma unroll
    for (int pl = 0; pl < kWgPix; ++pl) {
      const float4 xv = *reinterpret_cast<const float4*>(&xs[pl][tci * 4]);
      const float4 gv = *reinterpret_cast<const float4*>(&gs[pl][tco * 4]);
      const float xa[4] = {xv.x, xv.y, xv.z, xv.w}, ga[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(ga[a], xa[b], acc[a][b]);
    }
    __syncthreads();
  }
  const int taps = kh * kw;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int co = cot * 64 + tco * 4 + a;
    if (co >= Cout) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int ci = cit * 64 + tci * 4 + b;
      if (ci < Cin) atomicAdd(dw + ((int64_t)co * Cin + ci) * taps + tap, acc[a][b]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// wgrad on the tensor cores ('tf32' arm): the same (tap, 64 ci, 64 co) tile over a slab of pixels, but the products run on
// mma.sync.m16n8k8 TF32 with fp32 accumulation: D[ci][co] += sum_pixels x[pixel @ tap][ci] * gy[pixel][co], i.e. M = ci,
// N = co, K = pixels.  Both operands are staged pixel-major (as they lie in HBM) with a 72-float pitch: lane (gid, t) of a
// fragment reads [k = t (+4)][m or n = gid (+8)], i.e. bank 8 t + gid -- conflict-free scalar loads, no transposition.
// Operands are rounded to nearest TF32 on their way into shared memory (the tensor core would truncate).  8 warps: four
// 32 x 32 warp tiles x two K groups (pixels 0-15 / 16-31 of a 32-pixel chunk), register double buffering of the global
// loads, the K groups reduced through shared memory before the fp32 atomics.
// ---------------------------------------------------------------------------------------------
constexpr int kWmPix = 32, kWmPitch = 72;
__device__ __forceinline__ float tf32_rna(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
__device__ __forceinline__ void mma_tf32_1688(float (&c)[4], float a0, float a1, float a2, float a3, float b0, float b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(__float_as_uint(a0)), "r"(__float_as_uint(a1)), "r"(__float_as_uint(a2)), "r"(__float_as_uint(a3)),
        "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}
template <bool kVec>
__global__ void __launch_bounds__(256) conv_wgrad_tf32_kernel(const float* __restrict__ x, int x_pitch,
                                                              const float* __restrict__ gy, int gy_pitch,
                                                              float* __restrict__ dw, int N, int H, int W, int Cin, int Cout,
                                                              int kh, int kw, int stride, int pad, int dil, int Ho, int Wo,
                                                              int ci_tiles, int co_tiles, int pix_per_block) {
  __shared__ __align__(16) float xs[2][kWmPix][kWmPitch];
  __shared__ __align__(16) float gs[2][kWmPix][kWmPitch];
  int tile = blockIdx.x;
  const int cot = tile % co_tiles; tile /= co_tiles;
  const int cit = tile % ci_tiles; tile /= ci_tiles;
  const int tap = tile;
  const int r = tap / kw, s = tap - r * kw;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gid = lane >> 2, t = lane & 3;
  const int kg = warp >> 2;                     // K group: pixels 16 kg .. 16 kg + 15 of a chunk
  const int m0 = (warp & 1) * 32, n0 = ((warp >> 1) & 1) * 32;
  const int npix = N * Ho * Wo;                 // (< 2^31: checked by the entry point)
  const int p0 = blockIdx.y * pix_per_block;
  const int p1 = p0 + pix_per_block < npix ? p0 + pix_per_block : npix;
  // this thread stages pixels (tid >> 4) and (tid >> 4) + 16 of a chunk, channels 4 * (tid & 15) .. + 3
  const int c = (tid & 15) << 2;
  const int ci = cit * 64 + c, co = cot * 64 + c;
  float4 xr[2], gr[2];
  // (n, yo, xo) of this thread's two pixels of the chunk being fetched, advanced by 32 pixels per chunk without divisions
  int fn[2], fy[2], fx[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int p = p0 + (tid >> 4) + 16 * h;
    fx[h] = p % Wo;
    const int q = p / Wo;
    fy[h] = q % Ho;
    fn[h] = q / Ho;
  }
  auto fetch = [&](int pb) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int p = pb + (tid >> 4) + 16 * h;
      xr[h] = make_float4(0.f, 0.f, 0.f, 0.f);
      gr[h] = xr[h];
      if (p < p1) {
        const int yi = fy[h] * stride - pad + r * dil, xi = fx[h] * stride - pad + s * dil;
        if (kVec) {       // 16-byte loads: Cin, Cout and both pitches are multiples of 4
          if (ci < Cin && yi >= 0 && yi < H && xi >= 0 && xi < W)
            xr[h] = __ldg(reinterpret_cast<const float4*>(x + ((int64_t)(fn[h] * H + yi) * W + xi) * x_pitch + ci));
          if (co < Cout) gr[h] = __ldg(reinterpret_cast<const float4*>(gy + (int64_t)p * gy_pitch + co));
        } else {          // ragged channel counts (the 17-joint heatmap convolution): element-wise
          if (yi >= 0 && yi < H && xi >= 0 && xi < W) {
            const float* sx = x + ((int64_t)(fn[h] * H + yi) * W + xi) * x_pitch + ci;
            if (ci < Cin) xr[h].x = __ldg(sx);
            if (ci + 1 < Cin) xr[h].y = __ldg(sx + 1);
            if (ci + 2 < Cin) xr[h].z = __ldg(sx + 2);
            if (ci + 3 < Cin) xr[h].w = __ldg(sx + 3);
          }
          const float* sg = gy + (int64_t)p * gy_pitch + co;
          if (co < Cout) gr[h].x = __ldg(sg);
          if (co + 1 < Cout) gr[h].y = __ldg(sg + 1);
          if (co + 2 < Cout) gr[h].z = __ldg(sg + 2);
          if (co + 3 < Cout) gr[h].w = __ldg(sg + 3);
        }
      }
      fx[h] += kWmPix;
      while (fx[h] >= Wo) {
        fx[h] -= Wo;
        if (++fy[h] == Ho) { fy[h] = 0; ++fn[h]; }
      }
    }
  };
  auto store = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int pl = (tid >> 4) + 16 * h;
      *reinterpret_cast<float4*>(&xs[buf][pl][c]) = make_float4(tf32_rna(xr[h].x), tf32_rna(xr[h].y), tf32_rna(xr[h].z), tf32_rna(xr[h].w));
      *reinterpret_cast<float4*>(&gs[buf][pl][c]) = make_float4(tf32_rna(gr[h].x), tf32_rna(gr[h].y), tf32_rna(gr[h].z), tf32_rna(gr[h].w));
    }
  };
  float acc[2][4][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[a][b][e] = 0.f;
  int buf = 0;
  if (p0 < p1) {
    fetch(p0);
    store(0);
  }
  __syncthreads();
  for (int pb = p0; pb < p1; pb += kWmPix) {
    const bool more = pb + kWmPix < p1;
    if (more) fetch(pb + kWmPix);               // in flight while this chunk is multiplied
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int kb = 16 * kg + 8 * ks;
      float a[2][4], b[4][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        a[mt][0] = xs[buf][kb + t][m0 + 16 * mt + gid];
        a[mt][1] = xs[buf][kb + t][m0 + 16 * mt + gid + 8];
        a[mt][2] = xs[buf][kb + t + 4][m0 + 16 * mt + gid];
        a[mt][3] = xs[buf][kb + t + 4][m0 + 16 * mt + gid + 8];
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        b[nt][0] = gs[buf][kb + t][n0 + 8 * nt + gid];
        b[nt][1] = gs[buf][kb + t + 4][n0 + 8 * nt + gid];
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_tf32_1688(acc[mt][nt], a[mt][0], a[mt][1], a[mt][2], a[mt][3], b[nt][0], b[nt][1]);
    }
    if (more) store(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
  // K group 1 hands its partial tile to K group 0 through shared memory (the staging buffers are free now)
  float* red = &xs[0][0][0];                    // 4 warps x 32 values x 32 lanes = 16 KB <= sizeof(xs)
  if (kg == 1) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) red[(((warp & 3) * 32) + (mt * 4 + nt) * 4 + e) * 32 + lane] = acc[mt][nt][e];
  }
  __syncthreads();
  if (kg == 0) {
    const int taps = kh * kw;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float v = acc[mt][nt][e] + red[((warp * 32) + (mt * 4 + nt) * 4 + e) * 32 + lane];
          const int oci = cit * 64 + m0 + 16 * mt + gid + 8 * (e >> 1);
          const int oco = cot * 64 + n0 + 8 * nt + 2 * t + (e & 1);
          if (oci < Cin && oco < Cout) atomicAdd(dw + ((int64_t)oco * Cin + oci) * taps + tap, v);
        }
  }
}

// Same tile decomposition, double buffered: the (pixels x 64 ci) slab of x at this tap and the (pixels x 64 co) slab of gy
// of chunk k+1 are in flight (cp.async, 16 bytes per request, zero-filled outside the image / beyond the channel count)
// while chunk k is multiplied.  The single-buffered kernel above exposed a global round trip per 32 pixels.
__global__ void __launch_bounds__(256) conv_wgrad_db_kernel(const float* __restrict__ x, int x_pitch,
                                                            const float* __restrict__ gy, int gy_pitch,
                                                            float* __restrict__ dw, int N, int H, int W, int Cin, int Cout,
                                                            int kh, int kw, int stride, int pad, int dil, int Ho, int Wo,
                                                            int ci_tiles, int co_tiles, int pix_per_block) {
  __shared__ __align__(16) float xs[2][kWgPix][64];
  __shared__ __align__(16) float gs[2][kWgPix][64];
  int tile = blockIdx.x;
  const int cot = tile % co_tiles; tile /= co_tiles;
  const int cit = tile % ci_tiles; tile /= ci_tiles;
  const int tap = tile;
  const int r = tap / kw, s = tap - r * kw;
  const int tid = threadIdx.x;
  const int tci = tid & 15, tco = tid >> 4;
  const int64_t npix = (int64_t)N * Ho * Wo;
  const int64_t p0 = (int64_t)blockIdx.y * pix_per_block;
  const int64_t p1 = p0 + pix_per_block < npix ? p0 + pix_per_block : npix;
  // this thread stages pixels (tid >> 4) and (tid >> 4) + 16 of a chunk, channels 4 * (tid & 15) .. + 3
  const int c = (tid & 15) << 2;
  const int ci = cit * 64 + c, co = cot * 64 + c;
  auto stage = [&](int buf, int64_t pb) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int pl = (tid >> 4) + 16 * h;
      const int64_t p = pb + pl;
      bool okx = false, okg = false;
      const float* sx = x;
      const float* sg = gy;
      if (p < p1) {
        const int xo = (int)(p % Wo);
        const int64_t t = p / Wo;
        const int yo = (int)(t % Ho), n = (int)(t / Ho);
        const int yi = yo * stride - pad + r * dil, xi = xo * stride - pad + s * dil;
        if (ci < Cin && yi >= 0 && yi < H && xi >= 0 && xi < W) {
          okx = true;
          sx = x + ((int64_t)(n * H + yi) * W + xi) * x_pitch + ci;
        }
        if (co < Cout) {
          okg = true;
          sg = gy + p * gy_pitch + co;
        }
      }
      cp_async16(&xs[buf][pl][c], sx, okx);
      cp_async16(&gs[buf][pl][c], sg, okg);
    }
    cp_async_commit();
  };
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  int buf = 0;
  if (p0 < p1) stage(0, p0);
  for (int64_t pb = p0; pb < p1; pb += kWgPix) {
    if (pb + kWgPix < p1) {
      stage(buf ^ 1, pb + kWgPix);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
#pragma unroll
    for (int pl = 0; pl < kWgPix; ++pl) {
      const float4 xv = *reinterpret_cast<const float4*>(&xs[buf][pl][tci * 4]);
      const float4 gv = *reinterpret_cast<const float4*>(&gs[buf][pl][tco * 4]);
      const float xa[4] = {xv.x, xv.y, xv.z, xv.w}, ga[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(ga[a], xa[b], acc[a][b]);
    }
    __syncthreads();
    buf ^= 1;
  }
  const int taps = kh * kw;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int oc = cot * 64 + tco * 4 + a;
    if (oc >= Cout) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int ic = cit * 64 + tci * 4 + b;
      if (ic < Cin) atomicAdd(dw + ((int64_t)oc * Cin + ic) * taps + tap, acc[a][b]);
    }
  }
}

// per-channel sum over rows (bias gradient); float atomics into out (caller zeroes)
__global__ void __launch_bounds__(1024) col_sum_kernel(const float* __restrict__ x, int pitch, int64_t rows, int C,
                                                       float* __restrict__ out, int rows_per_block) {
  extern __shared__ float sm[];
  const int RG = blockDim.x / C;
  const int tid = threadIdx.x, c = tid % C, rg = tid / C;
  int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float s = 0.f;
  if (rg < RG)
    for (int64_t r = r0 + rg; r < r1; r += RG) s += x[r * pitch + c];
  if (rg < RG) sm[rg * C + c] = s;
  __syncthreads();
  if (tid < C) {
    float t = 0.f;
    for (int r = 0; r < RG; ++r) t += sm[r * C + tid];
    atomicAdd(out + tid, t);
  }
}

// ---------------------------------------------------------------------------------------------
// BatchNorm backward.  g' = gy * [y > 0] (fused ReLU), xhat = (x - mean) * invstd.
//   pass 1: sums[c] = sum g', sums[C + c] = sum g' * xhat   (double atomics, caller zeroes)
//   pass 2: training: dx = gamma*invstd * (g' - sums[c]/M - xhat * sums[C+c]/M)
//           eval    : dx = gamma*invstd * g'
//           g_res (optional) = g'  -- the gradient of the residual operand of the fused epilogue
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) bn_bwd_reduce_kernel(const float* __restrict__ x, int x_pitch,
                                                             const float* __restrict__ gy, int gy_pitch,
                                                             const float* __restrict__ y, int y_pitch,
                                                             const float* __restrict__ mean, const float* __restrict__ invstd,
                                                             int64_t rows, int C, double* __restrict__ sums,
                                                             int rows_per_block) {
  extern __shared__ float sm[];
  const int RG = blockDim.x / C;
  const int tid = threadIdx.x, c = tid % C, rg = tid / C;
  int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float s = 0.f, q = 0.f;
  if (rg < RG) {
    const float mu = mean[c], is = invstd[c];
    for (int64_t r = r0 + rg; r < r1; r += RG) {
      float g = gy[r * gy_pitch + c];
      if (y && !(y[r * y_pitch + c] > 0.f)) g = 0.f;
      s += g;
      q = fmaf(g, (x[r * x_pitch + c] - mu) * is, q);
    }
    sm[rg * C + c] = s;
    sm[(RG + rg) * C + c] = q;
  }
  __syncthreads();
  if (tid < C) {
    double ds = 0, dq = 0;
    for (int r = 0; r < RG; ++r) { ds += sm[r * C + tid]; dq += sm[(RG + r) * C + tid]; }
    atomicAdd(sums + tid, ds);
    atomicAdd(sums + C + tid, dq);
  }
}

__global__ void bn_bwd_apply_kernel(const float* __restrict__ x, int x_pitch, const float* __restrict__ gy, int gy_pitch,
                                    const float* __restrict__ y, int y_pitch, const float* __restrict__ mean,
                                    const float* __restrict__ invstd, const float* __restrict__ gamma,
                                    const double* __restrict__ sums, int64_t rows, int C, int training,
                                    float* __restrict__ dx, int dx_pitch, float* __restrict__ g_res, int gres_pitch,
                                    float* __restrict__ dgamma, float* __restrict__ dbeta) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) {
    if (dgamma) dgamma[i] = (float)sums[C + i];
    if (dbeta) dbeta[i] = (float)sums[i];
  }
  if (i >= rows * C) return;
  const int c = (int)(i % C);
  const int64_t r = i / C;
  float g = gy[r * gy_pitch + c];
  if (y && !(y[r * y_pitch + c] > 0.f)) g = 0.f;
  if (g_res) g_res[r * gres_pitch + c] = g;
  if (!dx) return;
  const float is = invstd[c];
  const float k = (gamma ? gamma[c] : 1.f) * is;
  float v = g;
  if (training) {
    const double inv_m = 1.0 / (double)rows;
    const float xhat = (x[r * x_pitch + c] - mean[c]) * is;
    v = g - (float)(sums[c] * inv_m) - xhat * (float)(sums[C + c] * inv_m);
  }
  dx[r * dx_pitch + c] = k * v;
}

// 16-byte form of bn_bwd_apply_kernel (C, pitches multiples of 4, 16-byte aligned tensors): two pieces per thread, the loads of
// both issued before the first use (the element-per-thread kernel keeps 4 bytes per tensor in flight per thread)
__global__ void __launch_bounds__(256) bn_bwd_apply_vec4_kernel(const float* __restrict__ x, int x_pitch, const float* __restrict__ gy,
                                                                int gy_pitch, const float* __restrict__ y, int y_pitch,
                                                                const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                const float* __restrict__ gamma, const double* __restrict__ sums,
                                                                int64_t rows, int C, int training, float* __restrict__ dx,
                                                                int dx_pitch, float* __restrict__ g_res, int gres_pitch,
                                                                float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int cq = C >> 2;
  const int64_t base = (int64_t)blockIdx.x * 512 + threadIdx.x;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int64_t ch = base + 256 * k;
    if (ch < C) {
      if (dgamma) dgamma[ch] = (float)sums[C + ch];
      if (dbeta) dbeta[ch] = (float)sums[ch];
    }
  }
  const int64_t tot = rows * cq;
  float4 g[2], yv[2], xv[2];
  int c[2];
  int64_t r[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int64_t i = base + 256 * k;
    r[k] = i / cq;
    c[k] = (int)(i - r[k] * cq) << 2;
    if (i < tot) {
      g[k] = __ldg(reinterpret_cast<const float4*>(gy + r[k] * gy_pitch + c[k]));
      if (y) yv[k] = __ldg(reinterpret_cast<const float4*>(y + r[k] * y_pitch + c[k]));
      if (dx && training) xv[k] = __ldg(reinterpret_cast<const float4*>(x + r[k] * x_pitch + c[k]));
    }
  }
  const double inv_m = 1.0 / (double)rows;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    if (base + 256 * k >= tot) continue;
    float ga[4] = {g[k].x, g[k].y, g[k].z, g[k].w};
    if (y) {
      const float ya[4] = {yv[k].x, yv[k].y, yv[k].z, yv[k].w};
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (!(ya[e] > 0.f)) ga[e] = 0.f;
    }
    if (g_res) *reinterpret_cast<float4*>(g_res + r[k] * gres_pitch + c[k]) = make_float4(ga[0], ga[1], ga[2], ga[3]);
    if (!dx) continue;
    const float xa[4] = {xv[k].x, xv[k].y, xv[k].z, xv[k].w};
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int ch = c[k] + e;
      const float is = invstd[ch];
      const float kk = (gamma ? gamma[ch] : 1.f) * is;
      float v = ga[e];
      if (training) {
        const float xhat = (xa[e] - mean[ch]) * is;
        v = ga[e] - (float)(sums[ch] * inv_m) - xhat * (float)(sums[C + ch] * inv_m);
      }
      o[e] = kk * v;
    }
    *reinterpret_cast<float4*>(dx + r[k] * dx_pitch + c[k]) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// ---------------------------------------------------------------------------------------------
// MI pseudo-KL backward.  Per row (sample, channel) over l: p = softmax(a/T), t = softmax(b/T),
// L = sum t (log t - p);  dL/da_j = -(1/T) p_j (t_j - S),  S = sum t p;
// dL/db_j = (1/T) t_j ((log t_j - p_j) - L).  Both scaled by gout / (B*C*HW).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) softmax_pkl_bwd_kernel(const float* __restrict__ a, int a_pitch,
                                                               const float* __restrict__ bsrc, int b_pitch,
                                                               const float* __restrict__ gout, float* __restrict__ ga,
                                                               int ga_pitch, float* __restrict__ gb, int gb_pitch, int HW,
                                                               int C, float inv_temp, float inv_count) {
  extern __shared__ float sm[];  // [4][RG*C]
  const int RG = blockDim.x / C;
  const int tid = threadIdx.x;
  const int c = tid % C, rg = tid / C;
  const bool active = rg < RG;
  const float* ab = a + (int64_t)blockIdx.x * HW * a_pitch;
  const float* bb = bsrc + (int64_t)blockIdx.x * HW * b_pitch;
  float ma = -INFINITY, mb = -INFINITY;
  if (active)
    for (int l = rg; l < HW; l += RG) {
      ma = fmaxf(ma, ab[(int64_t)l * a_pitch + c] * inv_temp);
      mb = fmaxf(mb, bb[(int64_t)l * b_pitch + c] * inv_temp);
    }
  float* s0 = sm; float* s1 = sm + RG * C; float* s2 = sm + 2 * RG * C; float* s3 = sm + 3 * RG * C;
  if (active) { s0[rg * C + c] = ma; s1[rg * C + c] = mb; }
  __syncthreads();
  if (active)
    for (int r = 0; r < RG; ++r) { ma = fmaxf(ma, s0[r * C + c]); mb = fmaxf(mb, s1[r * C + c]); }
  __syncthreads();
  float za = 0.f, zb = 0.f, sb = 0.f, xab = 0.f;
  if (active)
    for (int l = rg; l < HW; l += RG) {
      const float va = ab[(int64_t)l * a_pitch + c] * inv_temp - ma;
      const float vb = bb[(int64_t)l * b_pitch + c] * inv_temp - mb;
      const float ea = expf(va), eb = expf(vb);
      za += ea; zb += eb; sb = fmaf(eb, vb, sb); xab = fmaf(ea, eb, xab);
    }
  if (active) { s0[rg * C + c] = za; s1[rg * C + c] = zb; s2[rg * C + c] = sb; s3[rg * C + c] = xab; }
  __syncthreads();
  if (active) {
    za = zb = sb = xab = 0.f;
    for (int r = 0; r < RG; ++r) { za += s0[r * C + c]; zb += s1[r * C + c]; sb += s2[r * C + c]; xab += s3[r * C + c]; }
    const float lzb = logf(zb);
    const float S = xab / (za * zb);
    const float Lrow = sb / zb - lzb - S;
    const float gs = __ldg(gout) * inv_count * inv_temp;
    const float iza = 1.f / za, izb = 1.f / zb;
    for (int l = rg; l < HW; l += RG) {
      const float va = ab[(int64_t)l * a_pitch + c] * inv_temp - ma;
      const float vb = bb[(int64_t)l * b_pitch + c] * inv_temp - mb;
      const float p = expf(va) * iza, t = expf(vb) * izb;
      if (ga) ga[((int64_t)blockIdx.x * HW + l) * ga_pitch + c] = -gs * p * (t - S);
      if (gb) gb[((int64_t)blockIdx.x * HW + l) * gb_pitch + c] = gs * t * ((vb - lzb - p) - Lrow);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Linear backward: gx = gy W, gW = gy^T x, gb = sum_m gy  (tiny: M = 4B, K <= 144, N <= 64)
// ---------------------------------------------------------------------------------------------
__global__ void linear_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ gy,
                                  float* __restrict__ gx, float* __restrict__ gw, float* __restrict__ gb, int M, int K, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_gx = gx ? M * K : 0, n_gw = gw ? N * K : 0, n_gb = gb ? N : 0;
  if (i < n_gx) {
    const int m = i / K, k = i - m * K;
    float s = 0.f;
    for (int n = 0; n < N; ++n) s = fmaf(gy[m * N + n], w[n * K + k], s);
    gx[i] = s;
  } else if (i < n_gx + n_gw) {
    const int j = i - n_gx, n = j / K, k = j - n * K;
    float s = 0.f;
    for (int m = 0; m < M; ++m) s = fmaf(gy[m * N + n], x[m * K + k], s);
    gw[j] = s;
  } else if (i < n_gx + n_gw + n_gb) {
    const int n = i - n_gx - n_gw;
    float s = 0.f;
    for (int m = 0; m < M; ++m) s += gy[m * N + n];
    gb[n] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// Adam over one flat parameter bucket (torch.optim.Adam semantics: no weight decay, no amsgrad), as built by
// posetimation/optimizer/optimizer.py:66-72.  bias corrections are passed in (host computes 1 - beta^t).
// ---------------------------------------------------------------------------------------------
__global__ void adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, int64_t n, float lr, float beta1, float beta2, float eps,
                                 float bc1, float bc2_sqrt) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  const float mi = beta1 * m[i] + (1.f - beta1) * gi;
  const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] -= (lr / bc1) * (mi / denom);
}

// CUDA-graph form: learning rate and bias corrections come from device memory (hyper = {lr, 1 - beta1^t, sqrt(1 - beta2^t)}),
// so a captured training step replays with the current step count / MultiStepLR value.
__global__ void adam_step_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                     float* __restrict__ v, int64_t n, const float* __restrict__ hyper, float beta1, float beta2,
                                     float eps) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float lr = hyper[0], bc1 = hyper[1], bc2_sqrt = hyper[2];
  const float gi = g[i];
  const float mi = beta1 * m[i] + (1.f - beta1) * gi;
  const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] -= (lr / bc1) * (mi / denom);
}

// sum of the up x up replicas of a nearest-upsampled gradient (+ optional ReLU mask from the block output y):
// backward of Interpolate(scale_factor=up, mode='nearest') fused with the `+ residual -> ReLU` of the HRNet fuse layers
// (hrnet.py:99-112,151-172).  g' = grad_y * [y > 0]; grad_small[n,yo,xo,c] = sum_{dy,dx} g'[n,yo*up+dy,xo*up+dx,c];
// grad_res = g' (may be null).  One thread per (low-resolution pixel, 4 channels).
__global__ void upsample_add_bwd_kernel(const float* __restrict__ gy, int gp, const float* __restrict__ y, int yp,
                                        float* __restrict__ gsmall, int sp, float* __restrict__ gres, int rp, int N, int Ho,
                                        int Wo, int C, int up) {
  const int c4n = C >> 2;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)N * Ho * Wo * c4n) return;
  const int c = (int)(i % c4n) * 4;
  int64_t pix = i / c4n;
  const int xo = (int)(pix % Wo);
  const int yo = (int)((pix / Wo) % Ho);
  const int n = (int)(pix / ((int64_t)Wo * Ho));
  const int Wb = Wo * up, Hb = Ho * up;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int dy = 0; dy < up; ++dy)
    for (int dx = 0; dx < up; ++dx) {
      const int64_t q = ((int64_t)n * Hb + yo * up + dy) * Wb + xo * up + dx;
      float4 g = *reinterpret_cast<const float4*>(gy + q * gp + c);
      if (y) {
        const float4 o = *reinterpret_cast<const float4*>(y + q * yp + c);
        g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f; g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
      }
      if (gres) *reinterpret_cast<float4*>(gres + q * rp + c) = g;
      acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
    }
  *reinterpret_cast<float4*>(gsmall + pix * sp + c) = acc;
}

}  // namespace

int adam_step_dev_launch(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, float beta1, float beta2,
                         float eps, cudaStream_t st) {
  adam_step_dev_kernel<<<cdiv(n, 256), 256, 0, st>>>(p, g, m, v, n, hyper, beta1, beta2, eps);
  FAMI_CHECK_LAUNCH("adam_step_dev_kernel");
  return 0;
}

int upsample_add_bwd_launch(const float* gy, int gp, const float* y, int yp, float* gsmall, int sp, float* gres, int rp, int N,
                            int Ho, int Wo, int C, int up, cudaStream_t st) {
  const int64_t tot = (int64_t)N * Ho * Wo * (C / 4);
  upsample_add_bwd_kernel<<<cdiv(tot, 256), 256, 0, st>>>(gy, gp, y, yp, gsmall, sp, gres, rp, N, Ho, Wo, C, up);
  FAMI_CHECK_LAUNCH("upsample_add_bwd_kernel");
  return 0;
}

int adam_step_launch(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                     float bc1, float bc2_sqrt, cudaStream_t st) {
  adam_step_kernel<<<cdiv(n, 256), 256, 0, st>>>(p, g, m, v, n, lr, beta1, beta2, eps, bc1, bc2_sqrt);
  FAMI_CHECK_LAUNCH("adam_step_kernel");
  return 0;
}

int flip_transpose_launch(const float* w, float* wt, int Cout, int Cin, int kh, int kw, cudaStream_t st) {
  int64_t tot = (int64_t)Cout * Cin * kh * kw;
  flip_transpose_kernel<<<cdiv(tot, 256), 256, 0, st>>>(w, wt, Cout, Cin, kh, kw);
  FAMI_CHECK_LAUNCH("flip_transpose_kernel");
  return 0;
}

int conv_dgrad_launch(const fami_conv_desc* d, const float* gy, const float* wt_packed, float* gx, cudaStream_t st) {
  if (d->stride == 1) {
    // forward convolution of grad_out [N,Ho,Wo,Cout] with the flipped/transposed filter
    fami_conv_desc f = *d;
    f.H = d->Ho; f.W = d->Wo; f.Cin = d->Cout; f.Cout = d->Cin;
    f.pad = d->dil * (d->kh - 1) - d->pad;
    f.Ho = d->H; f.Wo = d->W;
    f.up = 1; f.relu = 0; f.stats = 0; f.om_groups = 0;
    f.in_pitch = d->out_pitch; f.out_pitch = d->in_pitch; f.res_pitch = 0;
    f.out_dtype = FAMI_F32;      // f.dtype stays FAMI_F32 (exact SIMT) or FAMI_TF32 (tcgen05 kind::tf32, same kernels as forward)
    return fami_conv2d_bn_act_fwd(&f, gy, wt_packed, nullptr, nullptr, nullptr, gx, nullptr, (void*)st);
  }
  const int CinPad = fami_conv_cout_pad(d->Cin);
  const int64_t tot = (int64_t)d->N * d->H * d->W * ((d->Cin + 3) / 4);
  conv_dgrad_gather_kernel<<<cdiv(tot, 128), 128, 0, st>>>(gy, d->out_pitch, wt_packed, gx, d->in_pitch, d->N, d->H, d->W,
                                                           d->Cin, d->Cout, d->kh, d->kw, d->stride, d->pad, d->dil, d->Ho,
                                                           d->Wo, CinPad);
  FAMI_CHECK_LAUNCH("conv_dgrad_gather_kernel");
  return 0;
}

int conv_wgrad_launch(const fami_conv_desc* d, const float* x, const float* gy, float* dw, float* dbias, cudaStream_t st) {
  const int ci_tiles = (d->Cin + 63) / 64, co_tiles = (d->Cout + 63) / 64;
  const int tiles = d->kh * d->kw * ci_tiles * co_tiles;
  const int64_t npix = (int64_t)d->N * d->Ho * d->Wo;
  int splits = (8 * num_sms() + tiles - 1) / tiles;
  int64_t per = (npix + splits - 1) / splits;
  per = ((per + kWgPix - 1) / kWgPix) * kWgPix;
  if (per < kWgPix) per = kWgPix;
  splits = (int)((npix + per - 1) / per);
  dim3 grid(tiles, splits);
  const bool vec = d->Cin % 4 == 0 && d->Cout % 4 == 0 && d->in_pitch % 4 == 0 && d->out_pitch % 4 == 0 &&
                   (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(gy) & 15) == 0;
  FAMI_CHECK_ARG(npix < (1ll << 31) - 2 * per, "conv wgrad: too many output pixels");
  static const bool single = getenv("FAMI_WGRAD_SINGLE") != nullptr;     // A/B: the single-buffered kernel
  static const bool no_tc = getenv("FAMI_WGRAD_SIMT") != nullptr;        // A/B: the 'tf32' arm's wgrad on the fp32 FMA kernel
  if (d->dtype == FAMI_TF32 && !no_tc && vec)
    conv_wgrad_tf32_kernel<true><<<grid, 256, 0, st>>>(x, d->in_pitch, gy, d->out_pitch, dw, d->N, d->H, d->W, d->Cin, d->Cout,
                                                       d->kh, d->kw, d->stride, d->pad, d->dil, d->Ho, d->Wo, ci_tiles, co_tiles, (int)per);
  else if (d->dtype == FAMI_TF32 && !no_tc)
    conv_wgrad_tf32_kernel<false><<<grid, 256, 0, st>>>(x, d->in_pitch, gy, d->out_pitch, dw, d->N, d->H, d->W, d->Cin, d->Cout,
                                                        d->kh, d->kw, d->stride, d->pad, d->dil, d->Ho, d->Wo, ci_tiles, co_tiles, (int)per);
  else if (vec && !single)
    conv_wgrad_db_kernel<<<grid, 256, 0, st>>>(x, d->in_pitch, gy, d->out_pitch, dw, d->N, d->H, d->W, d->Cin, d->Cout,
                                               d->kh, d->kw, d->stride, d->pad, d->dil, d->Ho, d->Wo, ci_tiles, co_tiles, (int)per);
  else if (vec)
    conv_wgrad_kernel<true><<<grid, 256, 0, st>>>(x, d->in_pitch, gy, d->out_pitch, dw, d->N, d->H, d->W, d->Cin, d->Cout,
                                                  d->kh, d->kw, d->stride, d->pad, d->dil, d->Ho, d->Wo, ci_tiles, co_tiles,
                                                  (int)per);
  else
    conv_wgrad_kernel<false><<<grid, 256, 0, st>>>(x, d->in_pitch, gy, d->out_pitch, dw, d->N, d->H, d->W, d->Cin, d->Cout,
                                                   d->kh, d->kw, d->stride, d->pad, d->dil, d->Ho, d->Wo, ci_tiles, co_tiles,
                                                   (int)per);
  FAMI_CHECK_LAUNCH("conv_wgrad_kernel");
  if (dbias) {
    FAMI_CHECK_ARG(d->Cout <= 1024, "conv wgrad: bias gradient supports Cout <= 1024");
    int threads = 1024, RG = threads / d->Cout;
    int rows_per_block = 2048;
    col_sum_kernel<<<cdiv(npix, rows_per_block), threads, (size_t)RG * d->Cout * sizeof(float), st>>>(gy, d->out_pitch, npix,
                                                                                                   d->Cout, dbias, rows_per_block);
    FAMI_CHECK_LAUNCH("col_sum_kernel");
  }
  return 0;
}

int bn_bwd_launch(const float* x, int xp, const float* gy, int gp, const float* y, int yp, const float* mean,
                  const float* invstd, const float* gamma, int64_t rows, int C, int training, double* sums, float* dx,
                  int dxp, float* g_res, int grp, float* dgamma, float* dbeta, cudaStream_t st) {
  FAMI_CHECK_ARG(C <= 1024, "bn bwd: C <= 1024");
  int threads = 1024, RG = threads / C;
  // at least two blocks per SM (the head's maps at N = 32 are 221 K rows: 108 blocks of 2048 rows left a quarter of the SMs idle)
  int rows_per_block = (int)(rows / (2 * num_sms()));
  rows_per_block = rows_per_block > 2048 ? 2048 : (rows_per_block < 4 * RG ? 4 * RG : rows_per_block);
  bn_bwd_reduce_kernel<<<cdiv(rows, rows_per_block), threads, (size_t)2 * RG * C * sizeof(float), st>>>(
      x, xp, gy, gp, y, yp, mean, invstd, rows, C, sums, rows_per_block);
  FAMI_CHECK_LAUNCH("bn_bwd_reduce_kernel");
  int64_t tot = rows * C;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const bool vec = C % 4 == 0 && C >= 4 && xp % 4 == 0 && gp % 4 == 0 && (!y || yp % 4 == 0) && (!dx || dxp % 4 == 0) &&
                   (!g_res || grp % 4 == 0) && al16(x) && al16(gy) && al16(y) && al16(dx) && al16(g_res);
  if (vec)
    bn_bwd_apply_vec4_kernel<<<cdiv(tot / 4 > C ? tot / 4 : C, 512), 256, 0, st>>>(x, xp, gy, gp, y, yp, mean, invstd, gamma, sums, rows, C, training,
                                                                 dx, dxp, g_res, grp, dgamma, dbeta);
  else
    bn_bwd_apply_kernel<<<cdiv(tot, 256), 256, 0, st>>>(x, xp, gy, gp, y, yp, mean, invstd, gamma, sums, rows, C, training, dx,
                                                        dxp, g_res, grp, dgamma, dbeta);
  FAMI_CHECK_LAUNCH("bn_bwd_apply_kernel");
  return 0;
}

int softmax_pkl_bwd_launch(const float* a, int ap, const float* b, int bp, const float* gout, float* ga, int gap, float* gb,
                           int gbp, int B, int HW, int C, float temperature, cudaStream_t st) {
  int threads = 1024, RG = threads / C;
  size_t smem = (size_t)4 * RG * C * sizeof(float);
  float inv_count = 1.f / ((float)B * (float)C * (float)HW);
  softmax_pkl_bwd_kernel<<<B, threads, smem, st>>>(a, ap, b, bp, gout, ga, gap, gb, gbp, HW, C, 1.f / temperature, inv_count);
  FAMI_CHECK_LAUNCH("softmax_pkl_bwd_kernel");
  return 0;
}

int linear_bwd_launch(const float* x, const float* w, const float* gy, float* gx, float* gw, float* gb, int M, int K, int N,
                      cudaStream_t st) {
  int tot = (gx ? M * K : 0) + (gw ? N * K : 0) + (gb ? N : 0);
  if (tot == 0) return 0;
  linear_bwd_kernel<<<cdiv(tot, 128), 128, 0, st>>>(x, w, gy, gx, gw, gb, M, K, N);
  FAMI_CHECK_LAUNCH("linear_bwd_kernel");
  return 0;
}

}  // namespace fami
