// Backward kernels of the two alignment ops (fp32 storage, exact-fp32 arithmetic):
//   fami_dcn_bwd            -- torchvision's deformable_col2im / deformable_col2im_coord + weight GEMM
//                              (analytic form: SURVEY.md Appendix B backward)
//   fami_warp_translate_bwd -- backward of kornia warp_affine for pure translations (Alignment_V15.py:133-135)
// First correct versions: straightforward SIMT with atomics for the scatters; they complete the ABI so the
// reference's trainable head (1.06 M parameters with the frozen HRNet default) can be differentiated.
#include <stdlib.h>

#include "common.cuh"

namespace fami {

namespace {

struct BwdP {
  int B, H, W, C, Cout, CoutPad, G, cpg, d;
  int xp, offp, mp, gop;
  int tf32;            // weight gradient products on mma.sync TF32 (FAMI_TF32 descriptors), else exact fp32 FMAs
  int tap_minor;       // weight-gradient grid: (9 taps, chunks) instead of (chunks, 9 taps)
  int colp, gosp;      // shared-memory row pitches of the weight-gradient kernel's column / grad_out slabs
  const float* x;
  const float* off;
  const float* mask;
  const float* w;      // packed [9*C][CoutPad]
  const float* go;     // [B,H,W,Cout] pitch gop
  float* gx;           // [B,H,W,C] dense (pitch C), zeroed by the launcher
  float* goff;         // [B,H,W,18G] dense
  float* gmask;        // [B,H,W,9G] dense
  float* gw;           // packed [9*C][CoutPad], zeroed by the launcher
  float* gb;           // [Cout], zeroed by the launcher
};

__device__ __forceinline__ float tf32_rna_b(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  // 16-byte vector reduction (sm_90+): one L2 atomic transaction for the 4 channels of a corner
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// one thread per (pixel, tap, 4-channel quad): grad wrt input (scatter), offset and mask.
// Persistent CTAs; kSmemW: the filter is staged once per CTA in shared memory, transposed to
// [tap][o][C] so that the quads of a pixel read consecutive 16-byte words (g_col = W^T g_out).
template <bool kSmemW>
__global__ void __launch_bounds__(512, 2) dcn_bwd_data_kernel(const BwdP p) {
  extern __shared__ float s_wt[];
  const int quads = p.C >> 2;
  const int64_t total = (int64_t)p.B * p.H * p.W * 9 * quads;
  if (kSmemW) {
    const int n = 9 * p.C * p.Cout;
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
      const int c = e % p.C;
      const int t = e / p.C;
      const int o = t % p.Cout, tap = t / p.Cout;
      s_wt[e] = __ldg(p.w + (int64_t)(tap * p.C + c) * p.CoutPad + o);
    }
    __syncthreads();
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(i % quads);
    const int tap = (int)((i / quads) % 9);
    const int64_t pix = i / (9 * quads);
    const int x = (int)(pix % p.W);
    const int y = (int)((pix / p.W) % p.H);
    const int b = (int)(pix / ((int64_t)p.W * p.H));
    const int ch = q << 2, g = ch / p.cpg;
    const int fr = tap / 3, fs = tap - fr * 3;

    // g_col[c] = sum_o W[o,c,tap] * g_out[pix,o]
    float gc[4] = {0.f, 0.f, 0.f, 0.f};
    const float* gop = p.go + pix * p.gop;
    if (kSmemW) {
      const float* wr = s_wt + (int64_t)tap * p.Cout * p.C + ch;
      for (int o = 0; o < p.Cout; ++o) {
        const float gv = __ldg(gop + o);
        const float4 wv = *reinterpret_cast<const float4*>(wr + o * p.C);
        gc[0] = fmaf(wv.x, gv, gc[0]); gc[1] = fmaf(wv.y, gv, gc[1]);
        gc[2] = fmaf(wv.z, gv, gc[2]); gc[3] = fmaf(wv.w, gv, gc[3]);
      }
    } else {
      const float* wr = p.w + (int64_t)(tap * p.C + ch) * p.CoutPad;
      for (int o = 0; o < p.Cout; ++o) {
        const float gv = __ldg(gop + o);
#pragma unroll
        for (int c = 0; c < 4; ++c) gc[c] = fmaf(__ldg(wr + (int64_t)c * p.CoutPad + o), gv, gc[c]);
      }
    }
    const float ody = __ldg(p.off + pix * p.offp + g * 18 + 2 * tap);
    const float odx = __ldg(p.off + pix * p.offp + g * 18 + 2 * tap + 1);
    const float mk = __ldg(p.mask + pix * p.mp + g * 9 + tap);
    const float py = (float)(y - p.d + fr * p.d) + ody;
    const float px = (float)(x - p.d + fs * p.d) + odx;
    float g_m = 0.f, g_dy = 0.f, g_dx = 0.f;
    if (py > -1.f && py < (float)p.H && px > -1.f && px < (float)p.W) {
      const int y0 = (int)floorf(py), x0 = (int)floorf(px);
      const float ly = py - (float)y0, lx = px - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
      const bool y0ok = y0 >= 0, y1ok = y0 + 1 <= p.H - 1, x0ok = x0 >= 0, x1ok = x0 + 1 <= p.W - 1;
      const int64_t img = (int64_t)b * p.H * p.W;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* xb = p.x + ch;
      const float4 v1 = (y0ok && x0ok) ? ld4<float>(xb + (img + (int64_t)y0 * p.W + x0) * p.xp) : z;
      const float4 v2 = (y0ok && x1ok) ? ld4<float>(xb + (img + (int64_t)y0 * p.W + x0 + 1) * p.xp) : z;
      const float4 v3 = (y1ok && x0ok) ? ld4<float>(xb + (img + (int64_t)(y0 + 1) * p.W + x0) * p.xp) : z;
      const float4 v4 = (y1ok && x1ok) ? ld4<float>(xb + (img + (int64_t)(y0 + 1) * p.W + x0 + 1) * p.xp) : z;
      const float a1[4] = {v1.x, v1.y, v1.z, v1.w}, a2[4] = {v2.x, v2.y, v2.z, v2.w};
      const float a3[4] = {v3.x, v3.y, v3.z, v3.w}, a4[4] = {v4.x, v4.y, v4.z, v4.w};
      const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
      float gm[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float val = w1 * a1[c] + w2 * a2[c] + w3 * a3[c] + w4 * a4[c];
        g_m = fmaf(gc[c], val, g_m);
        g_dy = fmaf(gc[c], hx * (a3[c] - a1[c]) + lx * (a4[c] - a2[c]), g_dy);
        g_dx = fmaf(gc[c], hy * (a2[c] - a1[c]) + ly * (a4[c] - a3[c]), g_dx);
        gm[c] = gc[c] * mk;
      }
      float* gxb = p.gx + ch;      // gx is dense (pitch C, C % 4 == 0): 16-byte aligned quads
      if (y0ok && x0ok) red_add_v4(gxb + (img + (int64_t)y0 * p.W + x0) * p.C, gm[0] * w1, gm[1] * w1, gm[2] * w1, gm[3] * w1);
      if (y0ok && x1ok) red_add_v4(gxb + (img + (int64_t)y0 * p.W + x0 + 1) * p.C, gm[0] * w2, gm[1] * w2, gm[2] * w2, gm[3] * w2);
      if (y1ok && x0ok) red_add_v4(gxb + (img + (int64_t)(y0 + 1) * p.W + x0) * p.C, gm[0] * w3, gm[1] * w3, gm[2] * w3, gm[3] * w3);
      if (y1ok && x1ok) red_add_v4(gxb + (img + (int64_t)(y0 + 1) * p.W + x0 + 1) * p.C, gm[0] * w4, gm[1] * w4, gm[2] * w4, gm[3] * w4);
      g_dy *= mk;
      g_dx *= mk;
    }
    // the cpg/4 quads of one offset group contribute to the same (offset, mask) entries
    if (p.cpg == 4) {
      p.goff[pix * (18 * p.G) + g * 18 + 2 * tap] = g_dy;
      p.goff[pix * (18 * p.G) + g * 18 + 2 * tap + 1] = g_dx;
      p.gmask[pix * (9 * p.G) + g * 9 + tap] = g_m;
    } else {
      atomicAdd(p.goff + pix * (18 * p.G) + g * 18 + 2 * tap, g_dy);
      atomicAdd(p.goff + pix * (18 * p.G) + g * 18 + 2 * tap + 1, g_dx);
      atomicAdd(p.gmask + pix * (9 * p.G) + g * 9 + tap, g_m);
    }
  }
}

// grad wrt weight / bias: block = (tap, chunk of 64 pixels); columns recomputed into shared memory
constexpr int kWChunk = 64;
__global__ void __launch_bounds__(256) dcn_bwd_weight_kernel(const BwdP p) {
  extern __shared__ float sm[];
  float* s_col = sm;                        // [kWChunk][colp]
  float* s_go = sm + kWChunk * p.colp;      // [kWChunk][gosp]  (pitches == 8 | 24 mod 32: conflict-free fragment loads)
  // the nine taps of a pixel chunk are adjacent blocks (blockIdx.x = tap): the chunk's offsets, masks and grad_out are
  // read from DRAM once and from the L2 eight times (tap-major block order streamed the 287 MB of offsets nine times)
  const int tap = p.tap_minor ? blockIdx.x : blockIdx.y;
  const int64_t pix0 = (int64_t)(p.tap_minor ? blockIdx.y : blockIdx.x) * kWChunk;
  const int64_t npix = (int64_t)p.B * p.H * p.W;
  const int fr = tap / 3, fs = tap - fr * 3;
  const int quads = p.C >> 2;
  for (int e = threadIdx.x; e < kWChunk * quads; e += blockDim.x) {
    const int pl = e / quads, q = e - pl * quads;
    const int64_t pix = pix0 + pl;
    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pix < npix) {
      const int x = (int)(pix % p.W);
      const int y = (int)((pix / p.W) % p.H);
      const int b = (int)(pix / ((int64_t)p.W * p.H));
      const int ch = q << 2, g = ch / p.cpg;
      const float ody = __ldg(p.off + pix * p.offp + g * 18 + 2 * tap);
      const float odx = __ldg(p.off + pix * p.offp + g * 18 + 2 * tap + 1);
      const float mk = __ldg(p.mask + pix * p.mp + g * 9 + tap);
      const float py = (float)(y - p.d + fr * p.d) + ody;
      const float px = (float)(x - p.d + fs * p.d) + odx;
      if (py > -1.f && py < (float)p.H && px > -1.f && px < (float)p.W) {
        const int y0 = (int)floorf(py), x0 = (int)floorf(px);
        const float ly = py - (float)y0, lx = px - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
        const bool y0ok = y0 >= 0, y1ok = y0 + 1 <= p.H - 1, x0ok = x0 >= 0, x1ok = x0 + 1 <= p.W - 1;
        const int64_t img = (int64_t)b * p.H * p.W;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* xb = p.x + ch;
        const float4 v1 = (y0ok && x0ok) ? ld4<float>(xb + (img + (int64_t)y0 * p.W + x0) * p.xp) : z;
        const float4 v2 = (y0ok && x1ok) ? ld4<float>(xb + (img + (int64_t)y0 * p.W + x0 + 1) * p.xp) : z;
        const float4 v3 = (y1ok && x0ok) ? ld4<float>(xb + (img + (int64_t)(y0 + 1) * p.W + x0) * p.xp) : z;
        const float4 v4 = (y1ok && x1ok) ? ld4<float>(xb + (img + (int64_t)(y0 + 1) * p.W + x0 + 1) * p.xp) : z;
        const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
        val.x = mk * (w1 * v1.x + w2 * v2.x + w3 * v3.x + w4 * v4.x);
        val.y = mk * (w1 * v1.y + w2 * v2.y + w3 * v3.y + w4 * v4.y);
        val.z = mk * (w1 * v1.z + w2 * v2.z + w3 * v3.z + w4 * v4.z);
        val.w = mk * (w1 * v1.w + w2 * v2.w + w3 * v3.w + w4 * v4.w);
      }
    }
    if (p.tf32) { val.x = tf32_rna_b(val.x); val.y = tf32_rna_b(val.y); val.z = tf32_rna_b(val.z); val.w = tf32_rna_b(val.w); }
    *reinterpret_cast<float4*>(s_col + pl * p.colp + (q << 2)) = val;
  }
  for (int e = threadIdx.x; e < kWChunk * p.Cout; e += blockDim.x) {
    const int pl = e / p.Cout, o = e - pl * p.Cout;
    const int64_t pix = pix0 + pl;
    const float gv = pix < npix ? __ldg(p.go + pix * p.gop + o) : 0.f;
    s_go[pl * p.gosp + o] = gv;       // (kept unrounded: the bias gradient below sums it exactly; rounded at fragment load)
  }
  __syncthreads();
  if (p.tf32) {
    // D[c][o] += sum_pixels col[pixel][c] * g_out[pixel][o] on mma.sync.m16n8k8 TF32 (M = c, N = o, K = the 64 pixels of
    // the chunk): both slabs are pixel-major, lane (gid, t) of a fragment reads [k = t (+4)][m | n = gid (+8)]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, t = lane & 3;
    const int mt_n = (p.C + 15) >> 4, nt_n = (p.Cout + 7) >> 3;
    for (int tile = warp; tile < mt_n * nt_n; tile += 8) {
      const int mt = tile / nt_n, nt = tile - mt * nt_n;
      const int c0 = 16 * mt + gid, o0 = 8 * nt + gid;
      const bool c0ok = c0 < p.C, c1ok = c0 + 8 < p.C, ook = o0 < p.Cout;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int kb = 0; kb < kWChunk; kb += 8) {
        const float* ca = s_col + (kb + t) * p.colp;
        const float* cb = ca + 4 * p.colp;
        const float* ga = s_go + (kb + t) * p.gosp;
        const float a0 = c0ok ? ca[c0] : 0.f, a1 = c1ok ? ca[c0 + 8] : 0.f, a2 = c0ok ? cb[c0] : 0.f, a3 = c1ok ? cb[c0 + 8] : 0.f;
        const float b0 = ook ? tf32_rna_b(ga[o0]) : 0.f, b1 = ook ? tf32_rna_b(ga[4 * p.gosp + o0]) : 0.f;
        asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
            : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3])
            : "r"(__float_as_uint(a0)), "r"(__float_as_uint(a1)), "r"(__float_as_uint(a2)), "r"(__float_as_uint(a3)),
              "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {      // acc = {(c0, 2t), (c0, 2t+1), (c0+8, 2t), (c0+8, 2t+1)} of the (mt, nt) tile
        const int c = 16 * mt + gid + 8 * (e >> 1), o = 8 * nt + 2 * t + (e & 1);
        if (c < p.C && o < p.Cout) atomicAdd(p.gw + (int64_t)(tap * p.C + c) * p.CoutPad + o, acc[e]);
      }
    }
  } else {
    for (int e = threadIdx.x; e < p.C * p.Cout; e += blockDim.x) {
      const int c = e / p.Cout, o = e - c * p.Cout;
      float s = 0.f;
      for (int pl = 0; pl < kWChunk; ++pl) s = fmaf(s_col[pl * p.colp + c], s_go[pl * p.gosp + o], s);
      atomicAdd(p.gw + (int64_t)(tap * p.C + c) * p.CoutPad + o, s);
    }
  }
  if (tap == 0 && p.gb) {
    for (int o = threadIdx.x; o < p.Cout; o += blockDim.x) {
      float s = 0.f;
      for (int pl = 0; pl < kWChunk; ++pl) s += s_go[pl * p.gosp + o];
      atomicAdd(p.gb + o, s);
    }
  }
}

// ---- translation warp backward --------------------------------------------------------------------
// out[b,y,x,c] = sum_k w_k * src[b, y0+ky, x0+kx, c], (py,px) = (y - ty, x - tx): d/dtx = -d/dpx, d/dty = -d/dpy
__global__ void __launch_bounds__(256) warp_bwd_kernel(const float* __restrict__ src, int sp, const float* __restrict__ txy,
                                                       const float* __restrict__ go, int gop, float* gs, int gsp,
                                                       float* gtxy, int B, int H, int W, int C) {
  const int C4 = C >> 2;
  const int64_t total = (int64_t)B * H * W * C4;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float gtx = 0.f, gty = 0.f;
  int b = 0;
  if (i < total) {
    const int q = (int)(i % C4);
    const int64_t pix = i / C4;
    const int x = (int)(pix % W);
    const int y = (int)((pix / W) % H);
    b = (int)(pix / ((int64_t)W * H));
    const float tx = __ldg(txy + 2 * b), ty = __ldg(txy + 2 * b + 1);
    const float px = (float)x - tx, py = (float)y - ty;
    float fx = floorf(px), fy = floorf(py);
    const float lx = px - fx, ly = py - fy;
    fx = fminf(fmaxf(fx, -2.f), (float)W);
    fy = fminf(fmaxf(fy, -2.f), (float)H);
    const int x0 = (int)fx, y0 = (int)fy;
    const float4 g = ld4<float>(go + pix * gop + q * 4);
    const float gv[4] = {g.x, g.y, g.z, g.w};
    float v[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int yy = y0 + (k >> 1), xx = x0 + (k & 1);
      const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
      const float w = ((k >> 1) ? ly : 1.f - ly) * ((k & 1) ? lx : 1.f - lx);
      float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) {
        const int64_t sp_pix = ((int64_t)b * H + yy) * W + xx;
        s4 = ld4<float>(src + sp_pix * sp + q * 4);
        if (gs) {
          float* gp = gs + sp_pix * gsp + q * 4;
#pragma unroll
          for (int c = 0; c < 4; ++c) atomicAdd(gp + c, gv[c] * w);
        }
      }
      v[k][0] = s4.x; v[k][1] = s4.y; v[k][2] = s4.z; v[k][3] = s4.w;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float ddx = (1.f - ly) * (v[1][c] - v[0][c]) + ly * (v[3][c] - v[2][c]);
      const float ddy = (1.f - lx) * (v[2][c] - v[0][c]) + lx * (v[3][c] - v[1][c]);
      gtx -= gv[c] * ddx;
      gty -= gv[c] * ddy;
    }
  }
  if (gtxy) {
    // threads of a warp may straddle two samples only at sample boundaries: reduce per warp when uniform
    const unsigned full = __activemask();
    const int b0 = __shfl_sync(full, b, 0);
    const bool uniform = __all_sync(full, b == b0 || i >= total);
    if (uniform) {
      gtx = warp_sum(gtx);
      gty = warp_sum(gty);
      if ((threadIdx.x & 31) == 0 && (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31) < total) {
        atomicAdd(gtxy + 2 * b0, gtx);
        atomicAdd(gtxy + 2 * b0 + 1, gty);
      }
    } else if (i < total) {
      atomicAdd(gtxy + 2 * b, gtx);
      atomicAdd(gtxy + 2 * b + 1, gty);
    }
  }
}

}  // namespace

int dcn_bwd_launch(const fami_dcn_desc* d, const float* x, const float* off, const float* mask, const float* w,
                   const float* go, float* gx, float* goff, float* gmask, float* gw, float* gb, cudaStream_t st) {
  FAMI_CHECK_ARG(d->om_layout == 0, "fami_dcn_bwd: torchvision offset/mask layout only");
  FAMI_CHECK_ARG(gx && goff && gmask && gw, "fami_dcn_bwd: gradient buffers must be provided");
  BwdP p;
  memset(&p, 0, sizeof(p));
  p.B = d->B; p.H = d->H; p.W = d->W; p.C = d->C; p.Cout = d->Cout; p.CoutPad = fami_conv_cout_pad(d->Cout);
  p.G = d->G; p.cpg = d->C / d->G; p.d = d->dil;
  p.xp = d->x_pitch; p.offp = d->off_pitch; p.mp = d->mask_pitch; p.gop = d->out_pitch;
  p.x = x; p.off = off; p.mask = mask; p.w = w; p.go = go;
  p.gx = gx; p.goff = goff; p.gmask = gmask; p.gw = gw; p.gb = gb;
  const int64_t npix = (int64_t)d->B * d->H * d->W;
  cudaMemsetAsync(gx, 0, sizeof(float) * npix * d->C, st);
  cudaMemsetAsync(gw, 0, sizeof(float) * (size_t)9 * d->C * p.CoutPad, st);
  if (gb) cudaMemsetAsync(gb, 0, sizeof(float) * d->Cout, st);
  if (p.cpg != 4) {
    cudaMemsetAsync(goff, 0, sizeof(float) * npix * 18 * d->G, st);
    cudaMemsetAsync(gmask, 0, sizeof(float) * npix * 9 * d->G, st);
  }
  const int64_t items = npix * 9 * (d->C / 4);
  FAMI_CHECK_ARG((reinterpret_cast<uintptr_t>(gx) & 15) == 0, "fami_dcn_bwd: grad_x must be 16-byte aligned");
  const size_t wsm = (size_t)9 * d->C * d->Cout * sizeof(float);
  int64_t blocks = cdiv(items, 512);
  if (wsm <= 200 * 1024) {
    const int64_t cap = (int64_t)num_sms() * 2;        // persistent: the transposed filter is staged once per CTA
    if (blocks > cap) blocks = cap;
    cudaFuncSetAttribute(dcn_bwd_data_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsm);
    dcn_bwd_data_kernel<true><<<(unsigned)blocks, 512, wsm, st>>>(p);
  } else {
    dcn_bwd_data_kernel<false><<<(unsigned)blocks, 512, 0, st>>>(p);
  }
  FAMI_CHECK_LAUNCH("dcn_bwd_data_kernel");
  // slab pitches: exact arm dense; tf32 arm == 8 | 24 (mod 32) floats so that the four k rows of a fragment load hit
  // different banks (and a multiple of 4 for the float4 column stores)
  auto frag_pitch = [](int n) { int q = (n + 3) & ~3; while (q % 32 != 8 && q % 32 != 24) q += 4; return q; };
  p.tf32 = d->dtype == FAMI_TF32 && getenv("FAMI_DCN_BWD_SIMT") == nullptr;
  p.colp = p.tf32 ? frag_pitch(d->C) : d->C;
  p.gosp = p.tf32 ? frag_pitch(d->Cout) : d->Cout;
  const size_t smem = (size_t)kWChunk * (p.colp + p.gosp) * sizeof(float);
  FAMI_CHECK_ARG(smem <= 200 * 1024, "fami_dcn_bwd: C + Cout too large for the weight-gradient kernel");
  cudaFuncSetAttribute(dcn_bwd_weight_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int chunks = cdiv(npix, kWChunk);
  p.tap_minor = chunks <= 65535 && getenv("FAMI_DCN_BWD_TAPMAJOR") == nullptr;
  dim3 grid(p.tap_minor ? 9 : chunks, p.tap_minor ? chunks : 9);
  dcn_bwd_weight_kernel<<<grid, 256, smem, st>>>(p);
  FAMI_CHECK_LAUNCH("dcn_bwd_weight_kernel");
  return 0;
}

int warp_translate_bwd_launch(const float* src, int sp, const float* txy, const float* go, int gop, float* gs, int gsp,
                              float* gtxy, int B, int H, int W, int C, cudaStream_t st) {
  FAMI_CHECK_ARG(C % 4 == 0 && sp % 4 == 0 && gop % 4 == 0, "fami_warp_translate_bwd: C and pitches must be multiples of 4");
  FAMI_CHECK_ARG(!gs || gsp == C, "fami_warp_translate_bwd: grad_src must be dense (pitch == C): it is zero-filled here");
  if (gs) cudaMemsetAsync(gs, 0, sizeof(float) * (size_t)B * H * W * C, st);
  if (gtxy) cudaMemsetAsync(gtxy, 0, sizeof(float) * 2 * B, st);
  const int64_t total = (int64_t)B * H * W * (C / 4);
  warp_bwd_kernel<<<cdiv(total, 256), 256, 0, st>>>(src, sp, txy, go, gop, gs, gsp, gtxy, B, H, W, C);
  FAMI_CHECK_LAUNCH("warp_bwd_kernel");
  return 0;
}

}  // namespace fami
