// Modulated deformable convolution v2 on B200 -- the north-star kernel (fami_dcn_fwd, 16-bit arm).
// Replaces torchvision.ops.deform_conv2d as called at posetimation/zoo/Alignment/Alignment_V15.py:146-158.
//
// Structure (persistent CTAs, one 16x8-pixel tile at a time, 544 threads):
//   * the x neighbourhood of the tile ((16+2R) x (8+2R) pixels x 64 channel slots) is brought ONCE into
//     shared memory by a TMA tiled box load (128B-swizzled rows; out-of-image pixels zero-filled, which
//     is exactly torchvision's "corner outside the image contributes 0" rule);
//   * 12 gather warps in three groups; group q owns UMMA A stage q and taps q, q+3, q+6. A thread is one
//     output pixel: per tap it reads the pixel's 3*G (dy | dx | mask) floats as 16-byte loads -- streamed
//     from HBM exactly once, in the tap-major layout the fused offset|mask conv writes -- and for each
//     offset group forms the four bilinear corners from shared memory (8-byte loads), blends them in
//     packed 16-bit arithmetic and stores the modulated columns into the A tile (128B-swizzled K-major,
//     16-byte stores). Samples outside the staged window are collected in a bit mask and resolved
//     afterwards through a bounds-checked global path;
//   * one warp issues tcgen05.mma (M=128, N=Cout, K=C per tap) against the weights resident in shared
//     memory, accumulating the nine taps in TMEM (double buffered across tiles);
//   * 4 epilogue warps add the bias and store the tile (coalesced, through shared-memory staging).
// The [C*9, B*H*W] column buffer of the reference never exists; HBM traffic is the algorithmic minimum:
// offsets+masks (4*27*G B/pixel) + x (~once, via L2) + out.
#include <stdlib.h>

#include "tc_common.cuh"

namespace fami {

__device__ unsigned long long g_dcn_trace[4096];   // [tile it < 16][16 events] of CTA 0, FAMI_DCN_TRACE=1

namespace {

__device__ __forceinline__ void dtrace(int on, int it, int ev) {
  if (on && blockIdx.x == 0 && it < 16) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_dcn_trace[it * 16 + ev] = t;
  }
}

constexpr int kTH = 16, kTW = 8;            // output tile (pixels): 128 = one UMMA M tile
constexpr int kGatherSets = 2;            // two sets of 8 warps; set s gathers taps s, s+2, s+4, ... of every tile
constexpr int kSetWarps = 8;              // 256 threads = 128 pixels x 2 offset-group parities
constexpr int kGatherWarps = kGatherSets * kSetWarps;
constexpr int kGatherThreads = 32 * kGatherWarps;
constexpr int kDcnEpiWarps = 4;
constexpr int kDcnThreads = kGatherThreads + 32 + 32 * kDcnEpiWarps;   // 672
constexpr int kAStages = 3;               // tap t -> A stage t % 3
constexpr int kATile = 128 * 128;          // bytes per A stage

struct DcnTcParams {
  int B, H, W, C, Cout, G, d, R;
  int WH, WW;                       // staged window (pixels)
  int tiles_x, tiles_y, total_tiles;
  int nk, BN;
  int om_pitch, x_pitch, out_pitch, vec_ok, out_f32;
  int om_blocked;                   // 1: offsets|masks in the warp-blocked layout (om_layout 2)
  int64_t om_tap_stride;            // floats between taps in the blocked layout = nblk * (3G/4) * 128
  uint32_t win_bytes, w_tile_bytes, ab_format;
  int trace;
  const float* om;                  // [B*H*W][om_pitch], per pixel [9 taps][dy(G) | dx(G) | mask(G)]
  const void* x;                    // TH NHWC (global fallback path)
  const float* bias;
  void* out;
};

template <typename TH> __device__ __forceinline__ float4 ld4h(const TH* p);   // 4 consecutive 16-bit values (8 B)
template <> __device__ __forceinline__ float4 ld4h<__half>(const __half* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
template <> __device__ __forceinline__ float4 ld4h<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

// packed 16-bit arithmetic for the bilinear blend: acc += w * v on two channels at once
template <typename TH> struct H2;
template <> struct H2<__half> {
  typedef __half2 t;
  static __device__ __forceinline__ t bcast(float w) { return __float2half2_rn(w); }
  static __device__ __forceinline__ t pack(float a, float b) { return __floats2half2_rn(a, b); }
  static __device__ __forceinline__ t lo(t v) { return __low2half2(v); }
  static __device__ __forceinline__ t hi(t v) { return __high2half2(v); }
  static __device__ __forceinline__ t fma(t a, t b, t c) { return __hfma2(a, b, c); }
  static __device__ __forceinline__ t mul(t a, t b) { return __hmul2(a, b); }
};
template <> struct H2<__nv_bfloat16> {
  typedef __nv_bfloat162 t;
  static __device__ __forceinline__ t bcast(float w) { return __float2bfloat162_rn(w); }
  static __device__ __forceinline__ t pack(float a, float b) { return __floats2bfloat162_rn(a, b); }
  static __device__ __forceinline__ t lo(t v) { return __low2bfloat162(v); }
  static __device__ __forceinline__ t hi(t v) { return __high2bfloat162(v); }
  static __device__ __forceinline__ t fma(t a, t b, t c) { return __hfma2(a, b, c); }
  static __device__ __forceinline__ t mul(t a, t b) { return __hmul2(a, b); }
};

template <typename TH, int kG>
__global__ void __launch_bounds__(kDcnThreads, 1)
dcn_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const DcnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* s_win = smem;                                   // WH*WW rows x 128 B
  uint8_t* s_a = s_win + p.win_bytes;                      // kAStages x 16 KB
  uint8_t* s_w = s_a + kAStages * kATile;                  // 9 x BN x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_w + 9 * p.w_tile_bytes);
  const uint32_t bar0 = smem_u32(bars);
  // barriers: win_full, win_free, w_full, a_full[3], a_empty[3], tfull[2], tempty[2]
  const uint32_t win_full = bar0, win_free = bar0 + 8, w_full = bar0 + 16;
  auto a_full = [&](int s) { return bar0 + 8u * (3 + s); };
  auto a_empty = [&](int s) { return bar0 + 8u * (3 + kAStages + s); };
  auto tfull = [&](int a) { return bar0 + 8u * (3 + 2 * kAStages + a); };
  auto tempty = [&](int a) { return bar0 + 8u * (5 + 2 * kAStages + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7 + 2 * kAStages);
  float* s_scale = reinterpret_cast<float*>(bars + 10 + 2 * kAStages);   // 16-byte aligned (float4 reads)
  float* s_shift = s_scale + p.BN;
  uint8_t* stage_base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(s_shift + p.BN) + 15) & ~(uintptr_t)15);
  fill_scale_shift(s_scale, s_shift, nullptr, p.bias, p.Cout, p.BN);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(win_full, 1);
    mbar_init(win_free, kGatherWarps);
    mbar_init(w_full, 1);
    for (int s = 0; s < kAStages; ++s) { mbar_init(a_full(s), kSetWarps); mbar_init(a_empty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), kDcnEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  if (warp == kGatherWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto tile_origin = [&](int tile, int& b, int& y0, int& x0) {
    const int per_img = p.tiles_x * p.tiles_y;
    b = tile / per_img;
    const int t = tile - b * per_img;
    const int ty = t / p.tiles_x;
    y0 = ty * kTH;
    x0 = (t - ty * p.tiles_x) * kTW;
  };

  if (warp < kGatherWarps) {
    // ===================== gather warps =====================
    // Two sets of eight warps; set s owns the taps s, s+2, s+4, ... of every tile (tap t lands in A stage t % 3, the MMA
    // warp consumes the taps in order).  A thread is (output pixel, offset-group parity): it walks the kG/2 groups
    // g = 2j + parity of its pixel for one tap.  Lane layout: 16 lanes = 8 consecutive pixels of an image row x 2 parities.
    // Bank picture of a corner load (8 bytes at window row r, 16-byte chunk (g>>1) ^ (r&7), half g&1): eight consecutive
    // pixels hit eight different chunks and the two parities the two halves, so a half-warp covers all 32 banks when the
    // offsets are smooth (one wavefront instead of the 2-4 of a one-thread-per-pixel mapping, where g&1 was warp-uniform
    // and only half the banks were reachable), and random offsets spread over 16 bank pairs instead of 8.
    constexpr int kQ = kG / 4;
    constexpr int kJ = kG / 2;             // groups per thread
    typedef typename H2<TH>::t h2;
    const TH* xg = reinterpret_cast<const TH*>(p.x);
    const uint32_t win_u32 = smem_u32(s_win);
    const int set = warp / kSetWarps;
    const int wset = warp - set * kSetWarps;                      // warp within the set: image rows 2*wset, 2*wset+1 of the tile
    const int par = lane & 1;
    const int ry = 2 * wset + (lane >> 4), rx = (lane >> 1) & 7;
    const int r = ry * kTW + rx;                                   // A-tile row = pixel of the tile
    const uint32_t rsw = (uint32_t)(r & 7) << 4;
    const uint32_t a_row0 = smem_u32(s_a) + (uint32_t)(r * 128) + ((uint32_t)par << 3);
    const int WW = p.WW;
    const unsigned ylim = (unsigned)(p.WH - 1), xlim = (unsigned)(WW - 1);
    uint32_t wph = 0, fph = 0;
    int git = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++git) {
      int b, y0, x0;
      tile_origin(tile, b, y0, x0);
      if (threadIdx.x == 0) dtrace(p.trace, git, 0);
      const int y = y0 + ry, x = x0 + rx;
      const bool valid = y < p.H && x < p.W;
      // tap-major NHWC (om_layout 1): the 3*kG floats of a (pixel, tap) are one contiguous run.  Warp-blocked (om_layout 2):
      // [tap][tile*8 + warp block][q < 3kG/4][16 pixels][e0 e2 | e1 e3]: this lane's float2 = its two groups of quad q, the
      // warp's 32 lanes read 256 contiguous bytes per instruction.
      const float* po = p.om_blocked
                            ? p.om + ((int64_t)tile * 8 + wset) * (3 * kQ * 64) + (lane & 31) * 2
                            : p.om + ((int64_t)(b * p.H + y) * p.W + x) * p.om_pitch;
      float2 vdy[kQ], vdx[kQ], vmk[kQ];      // groups g = 4i + par (x) and 4i + 2 + par (y)
      auto load_unit = [&](int tap) {
        if (p.om_blocked) {
          const float* o = po + tap * p.om_tap_stride;
#pragma unroll
          for (int i = 0; i < kQ; ++i) {
            vdy[i] = __ldg(reinterpret_cast<const float2*>(o + i * 64));
            vdx[i] = __ldg(reinterpret_cast<const float2*>(o + (kQ + i) * 64));
            vmk[i] = __ldg(reinterpret_cast<const float2*>(o + (2 * kQ + i) * 64));
          }
        } else {
          const float* o = po + tap * 3 * kG + par;
#pragma unroll
          for (int i = 0; i < kQ; ++i) {
            vdy[i] = make_float2(__ldg(o + 4 * i), __ldg(o + 4 * i + 2));
            vdx[i] = make_float2(__ldg(o + kG + 4 * i), __ldg(o + kG + 4 * i + 2));
            vmk[i] = make_float2(__ldg(o + 2 * kG + 4 * i), __ldg(o + 2 * kG + 4 * i + 2));
          }
        }
      };
      if (valid) load_unit(set);
      mbar_wait(win_full, wph);
      wph ^= 1u;
      if (threadIdx.x == 0) dtrace(p.trace, git, 1);
#pragma unroll 1
      for (int tap = set; tap < 9; tap += kGatherSets) {
        const int kr = tap / 3, kc = tap - kr * 3;      // kernel row / column
        const int stage = tap - kr * 3;                 // tap % 3
        const uint32_t u = (uint32_t)(git * 3 + kr);    // use index of this stage
        mbar_wait(a_empty(stage), (u & 1u) ^ 1u);
        const uint32_t a_row = a_row0 + (uint32_t)(stage * kATile);
        if (!valid) {
#pragma unroll
          for (int j = 0; j < kJ; ++j) sts64(a_row + (((uint32_t)j << 4) ^ rsw), make_uint2(0u, 0u));
        } else {
          const float base_y = (float)(ry + kr * p.d + p.R - p.d);
          const float base_x = (float)(rx + kc * p.d + p.R - p.d);
          uint32_t slow = 0;
#pragma unroll
          for (int j = 0; j < kJ; ++j) {
            // group g = 2j + par: quad j >> 1, element (j & 1) of this lane's float2
            const float dy = (j & 1) ? vdy[j >> 1].y : vdy[j >> 1].x;
            const float dx = (j & 1) ? vdx[j >> 1].y : vdx[j >> 1].x;
            const float mk = (j & 1) ? vmk[j >> 1].y : vmk[j >> 1].x;
            // window coordinates: integer shifts of the image-space sample position, so the fractional
            // parts are exactly those of py / px.  floor() without the conversion unit (FRND / F2I run at 16 lanes
            // per clock and were 2/3 busy): adding 1.5 * 2^23 with round-down leaves floor(w) in the low mantissa bits
            // for |w| < 2^22; a sample further out than that is far outside the window either way.
            const float wy = base_y + dy, wx = base_x + dx;
            const float ty = __fadd_rd(wy, 12582912.f), tx = __fadd_rd(wx, 12582912.f);
            const float fy = ty - 12582912.f, fx = tx - 12582912.f;
            const int iy = __float_as_int(ty) - 0x4B400000, ix = __float_as_int(tx) - 0x4B400000;
            uint2 pk = make_uint2(0u, 0u);
            if ((unsigned)iy < ylim && (unsigned)ix < xlim) {
              // all four corners inside the staged (zero-padded) window: 8-byte shared-memory loads,
              // blend in packed 16-bit arithmetic (the column is rounded to 16 bit for the MMA anyway)
              const float ly = wy - fy, lx = wx - fx;
              const float wb = mk * ly, wt = mk - wb;          // mask * (ly | 1 - ly)
              const float w4f = wb * lx, w3f = wb - w4f, w2f = wt * lx, w1f = wt - w2f;
              const uint32_t row00 = (uint32_t)(iy * WW + ix);
              const uint32_t row10 = row00 + (uint32_t)WW;
              const uint32_t gs = (uint32_t)j << 4, gofs = (uint32_t)par << 3;
              const uint2 u1 = lds64(win_u32 + row00 * 128u + ((gs ^ (row00 << 4)) & 0x70u) + gofs);
              const uint2 u2 = lds64(win_u32 + (row00 + 1u) * 128u + ((gs ^ ((row00 + 1u) << 4)) & 0x70u) + gofs);
              const uint2 u3 = lds64(win_u32 + row10 * 128u + ((gs ^ (row10 << 4)) & 0x70u) + gofs);
              const uint2 u4 = lds64(win_u32 + (row10 + 1u) * 128u + ((gs ^ ((row10 + 1u) << 4)) & 0x70u) + gofs);
              // two packed conversions + four lane broadcasts (PRMT) instead of four conversions
              const h2 w12 = H2<TH>::pack(w1f, w2f), w34 = H2<TH>::pack(w3f, w4f);
              const h2 w1 = H2<TH>::lo(w12), w2 = H2<TH>::hi(w12), w3 = H2<TH>::lo(w34), w4 = H2<TH>::hi(w34);
              h2 lo = H2<TH>::mul(w1, *reinterpret_cast<const h2*>(&u1.x));
              h2 hi = H2<TH>::mul(w1, *reinterpret_cast<const h2*>(&u1.y));
              lo = H2<TH>::fma(w2, *reinterpret_cast<const h2*>(&u2.x), lo);
              hi = H2<TH>::fma(w2, *reinterpret_cast<const h2*>(&u2.y), hi);
              lo = H2<TH>::fma(w3, *reinterpret_cast<const h2*>(&u3.x), lo);
              hi = H2<TH>::fma(w3, *reinterpret_cast<const h2*>(&u3.y), hi);
              lo = H2<TH>::fma(w4, *reinterpret_cast<const h2*>(&u4.x), lo);
              hi = H2<TH>::fma(w4, *reinterpret_cast<const h2*>(&u4.y), hi);
              pk.x = *reinterpret_cast<const uint32_t*>(&lo);
              pk.y = *reinterpret_cast<const uint32_t*>(&hi);
            } else {
              slow |= 1u << j;          // outside the staged window: resolved below from global memory
            }
            sts64(a_row + (((uint32_t)j << 4) ^ rsw), pk);
          }
          // large offsets: bounds-checked global corners, fp32 blend (rare; kept out of the main loop so a
          // single far sample does not serialise the warp through this path once per group)
          while (slow) {
            const int j = __ffs((int)slow) - 1;
            slow &= slow - 1u;
            const int g = 2 * j + par;
            float dy, dx, mk;
            if (p.om_blocked) {
              const float* o = po + tap * p.om_tap_stride;
              dy = __ldg(o + (j >> 1) * 64 + (j & 1));
              dx = __ldg(o + (kQ + (j >> 1)) * 64 + (j & 1));
              mk = __ldg(o + (2 * kQ + (j >> 1)) * 64 + (j & 1));
            } else {
              const float* o = po + tap * 3 * kG + g;
              dy = __ldg(o); dx = __ldg(o + kG); mk = __ldg(o + 2 * kG);
            }
            const float py = (float)(y - p.d + kr * p.d) + dy;
            const float px = (float)(x - p.d + kc * p.d) + dx;
            uint2 pk2 = make_uint2(0u, 0u);
            if (py > -1.f && py < (float)p.H && px > -1.f && px < (float)p.W) {
              const int iy0 = (int)floorf(py), ix0 = (int)floorf(px);
              const float ly = py - (float)iy0, lx = px - (float)ix0, hy = 1.f - ly, hx = 1.f - lx;
              const TH* xb = xg + (int64_t)b * p.H * p.W * p.x_pitch + g * 4;
              const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
              const bool y0ok = iy0 >= 0, y1ok = iy0 + 1 <= p.H - 1, x0ok = ix0 >= 0, x1ok = ix0 + 1 <= p.W - 1;
              const float4 v1 = (y0ok && x0ok) ? ld4h<TH>(xb + ((int64_t)iy0 * p.W + ix0) * p.x_pitch) : z;
              const float4 v2 = (y0ok && x1ok) ? ld4h<TH>(xb + ((int64_t)iy0 * p.W + ix0 + 1) * p.x_pitch) : z;
              const float4 v3 = (y1ok && x0ok) ? ld4h<TH>(xb + ((int64_t)(iy0 + 1) * p.W + ix0) * p.x_pitch) : z;
              const float4 v4 = (y1ok && x1ok) ? ld4h<TH>(xb + ((int64_t)(iy0 + 1) * p.W + ix0 + 1) * p.x_pitch) : z;
              const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
              pk2.x = f2_to_h2<TH>(mk * (w1 * v1.x + w2 * v2.x + w3 * v3.x + w4 * v4.x),
                                   mk * (w1 * v1.y + w2 * v2.y + w3 * v3.y + w4 * v4.y));
              pk2.y = f2_to_h2<TH>(mk * (w1 * v1.z + w2 * v2.z + w3 * v3.z + w4 * v4.z),
                                   mk * (w1 * v1.w + w2 * v2.w + w3 * v3.w + w4 * v4.w));
            }
            sts64(a_row + (((uint32_t)j << 4) ^ rsw), pk2);
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the MMA (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full(stage));   // one arrival per warp
        if (tap + kGatherSets < 9 && valid) load_unit(tap + kGatherSets);   // in flight while the MMA drains this stage
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(win_free);   // this warp no longer reads the window of this tile
      if (warp == 0) {
        // refill the window for the next tile as soon as the last gather warp has left it (issued from here,
        // not from the MMA warp, which is still draining the last taps)
        mbar_wait(win_free, fph);
        fph ^= 1u;
        const int next = tile + gridDim.x;
        if (lane == 0 && next < p.total_tiles) {
          int nb, ny0, nx0;
          tile_origin(next, nb, ny0, nx0);
          mbar_arrive_expect_tx(win_full, p.win_bytes);
          tma_tiled_4d(smem_u32(s_win), &tmX, win_full, 0, nx0 - p.R, ny0 - p.R, nb);
        }
        __syncwarp();
      }
    }
  } else if (warp == kGatherWarps) {
    // ===================== TMA + MMA issuer (warp-uniform control flow, elected lane issues) =====
    const bool leader = elect_one();
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t idesc = (1u << 4) | (p.ab_format << 7) | (p.ab_format << 10) | ((uint32_t)(p.BN >> 3) << 17) |
                           ((uint32_t)(128 >> 4) << 24);
    // pull a tile's offsets|masks (16 row segments of 8 pixels x 27G floats, the dominant HBM stream) into
    // L2 one tile ahead, so the gather warps' dependent loads see L2 rather than HBM latency
    auto prefetch_om = [&](int b, int y0, int x0) {
      if (p.om_blocked) {
        // 9 taps x (8 blocks x 3G/4 x 256 B) contiguous per tile
        const int tile_id = (b * p.tiles_y + y0 / kTH) * p.tiles_x + x0 / kTW;
        const uint32_t bytes = (uint32_t)(4 * 3 * (p.G / 4) * 512);
        if (lane < 9) {
          const float* ptr = p.om + (int64_t)lane * p.om_tap_stride + (int64_t)tile_id * (bytes / 4);
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr), "r"(bytes) : "memory");
        }
        return;
      }
      const int row = lane >> 1, half = lane & 1;          // 32 lanes: 16 rows x 2 halves of the 8-pixel segment
      const int y = y0 + row, x = x0 + half * 4;
      if (y < p.H && x < p.W) {
        const int npx = (p.W - x < 4) ? (p.W - x) : 4;
        const float* ptr = p.om + ((int64_t)(b * p.H + y) * p.W + x) * p.om_pitch;
        const uint32_t bytes = (uint32_t)(npx * p.om_pitch * 4) & ~15u;
        if (bytes >= 16 && ((reinterpret_cast<uintptr_t>(ptr) & 15) == 0))
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr), "r"(bytes) : "memory");
      }
    };
    if (leader) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW)) : "memory");
      mbar_arrive_expect_tx(w_full, 9u * p.w_tile_bytes);
      for (int t = 0; t < 9; ++t) tma_tiled_2d(smem_u32(s_w + t * p.w_tile_bytes), &tmW, w_full, t * 64, 0);
    }
    {
      int b, y0, x0;
      if ((int)blockIdx.x < p.total_tiles) {
        tile_origin(blockIdx.x, b, y0, x0);
        if (leader) {
          mbar_arrive_expect_tx(win_full, p.win_bytes);
          tma_tiled_4d(smem_u32(s_win), &tmX, win_full, 0, x0 - p.R, y0 - p.R, b);
        }
        prefetch_om(b, y0, x0);
      }
      if ((int)(blockIdx.x + gridDim.x) < p.total_tiles) {
        tile_origin(blockIdx.x + gridDim.x, b, y0, x0);
        prefetch_om(b, y0, x0);
      }
    }
    mbar_wait(w_full, 0);
    tc_fence_after();
    const uint32_t w_lo0 = sw128_desc_lo(smem_u32(s_w));
    const uint32_t w_step = p.w_tile_bytes >> 4;
    int stage = 0;
    uint32_t aph = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      mbar_wait(tempty(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_u + (uint32_t)(acc * p.BN);
      uint32_t w_lo = w_lo0;
      for (int tap = 0; tap < 9; ++tap, w_lo += w_step) {
        mbar_wait(a_full(stage), aph);
        tc_fence_after();
        const uint32_t a_lo = sw128_desc_lo(smem_u32(s_a + stage * kATile));
        umma_ksteps_n(p.nk, leader, d_tmem, a_lo, w_lo, idesc, tap != 0);
        if (leader) umma_commit(a_empty(stage));
        if (++stage == kAStages) { stage = 0; aph ^= 1u; }
      }
      if (leader) umma_commit(tfull(acc));
      // pull the offsets|masks of the tile after next towards L2 (the window refill is issued by gather warp 0)
      const int next = tile + 2 * gridDim.x;
      if (next < p.total_tiles) {
        int b, y0, x0;
        tile_origin(next, b, y0, x0);
        prefetch_om(b, y0, x0);
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    EpiArgs ea;
    ea.s_scale = smem_u32(s_scale); ea.s_shift = smem_u32(s_shift); ea.res = nullptr; ea.y = p.out;
    ea.Cout = p.Cout; ea.BN = p.BN; ea.ch_base = 0; ea.out_pitch = p.out_pitch; ea.res_pitch = 0;
    ea.out_f32 = p.out_f32; ea.relu = 0; ea.vec_ok = p.vec_ok; ea.up = 1; ea.Wout = p.W;
    ea.spitch = 128 + 16;
    const uint32_t stage = smem_u32(stage_base) + (uint32_t)((warp - kGatherWarps - 1) * 32 * ea.spitch);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      int b, y0, x0;
      tile_origin(tile, b, y0, x0);
      const int y = y0 + (row >> 3), x = x0 + (row & 7);
      const bool valid = y < p.H && x < p.W;
      const int pix = valid ? (b * p.H + y) * p.W + x : 0;
      mbar_wait(tfull(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t)(acc * p.BN) + ((uint32_t)(quarter * 32) << 16);
      epilogue_rows<TH>(ea, t_addr, 0, p.BN, valid, pix, stage, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(acc));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kGatherWarps) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace

// shared-memory footprint for a window radius R; the launcher takes the largest R <= dil + 5 that fits
static size_t dcn_tc_smem(const fami_dcn_desc* d, int R) {
  const int BN = ((d->Cout + 15) / 16) * 16;
  const size_t win = (size_t)(kTH + 2 * R) * (kTW + 2 * R) * 128;
  return win + kAStages * kATile + 9 * (size_t)BN * 128 + 1024 + 256 + (size_t)BN * 8 +
         (size_t)kDcnEpiWarps * 32 * (128 + 16);
}
static int dcn_tc_radius(const fami_dcn_desc* d) {
  for (int R = d->dil + 5; R >= d->dil + 2; --R)
    if (dcn_tc_smem(d, R) <= 227 * 1024) return R;
  return -1;
}

int dcn_tc_supported(const fami_dcn_desc* d) {
  if (!is_half_dtype(d->dtype) || (d->om_layout != 1 && d->om_layout != 2)) return 0;
  if (d->C % 16 != 0 || d->C > 64 || d->G * 4 != d->C) return 0;      // G in {4, 8, 12, 16}
  if (d->om_layout == 1 && d->off_pitch % 4 != 0) return 0;             // 16-byte loads of the (dy|dx|mask) runs
  if (d->Cout > 256 || d->x_pitch % 8 != 0) return 0;
  if (d->kh != 3 || d->kw != 3 || d->pad != d->dil || d->dil > 4) return 0;
  if (dcn_tc_radius(d) < 0) return 0;                                   // filter + window do not fit in shared memory
  return 1;
}

int dcn_tc_launch(const fami_dcn_desc* d, const void* x, const float* om, const void* w, const float* bias, void* out,
                  cudaStream_t st) {
  FAMI_CHECK_ARG(load_driver_fns(), "cuTensorMapEncode* driver entry points unavailable");
  FAMI_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0,
                 "dcn_tc: x / w must be 16-byte aligned");
  FAMI_CHECK_ARG((reinterpret_cast<uintptr_t>(om) & 15) == 0, "dcn_tc: offsets|masks must be 16-byte aligned");
  DcnTcParams p;
  memset(&p, 0, sizeof(p));
  p.B = d->B; p.H = d->H; p.W = d->W; p.C = d->C; p.Cout = d->Cout; p.G = d->G; p.d = d->dil;
  p.R = dcn_tc_radius(d);   // dilation reach + up to 5 px of offset (2.5 sigma of the sigma = 2 px regime); beyond -> global path
  FAMI_CHECK_ARG(p.R > 0, "dcn_tc: filter and window do not fit in shared memory");
  p.WH = kTH + 2 * p.R; p.WW = kTW + 2 * p.R;
  p.tiles_x = (d->W + kTW - 1) / kTW; p.tiles_y = (d->H + kTH - 1) / kTH;
  p.total_tiles = d->B * p.tiles_x * p.tiles_y;
  p.nk = d->C / 16;
  p.BN = ((d->Cout + 15) / 16) * 16;
  p.om_pitch = d->off_pitch; p.x_pitch = d->x_pitch; p.out_pitch = d->out_pitch;
  p.out_f32 = d->out_f32 ? 1 : 0;
  p.vec_ok = ((reinterpret_cast<uintptr_t>(out) & 15) == 0) && (d->out_pitch % (d->out_f32 ? 4 : 8) == 0);
  p.win_bytes = (uint32_t)(p.WH * p.WW) * 128u;
  p.w_tile_bytes = (uint32_t)p.BN * 128u;
  p.ab_format = d->dtype == FAMI_F16 ? 0u : 1u;
  p.om = om; p.x = x; p.bias = bias; p.out = out;
  p.om_blocked = d->om_layout == 2;
  p.om_tap_stride = (int64_t)p.total_tiles * 4 * (3 * (d->G / 4)) * 128;
  static const bool trace_on = getenv("FAMI_DCN_TRACE") != nullptr;
  p.trace = trace_on ? atoi(getenv("FAMI_DCN_TRACE")) : 0;

  const CUtensorMapDataType tm_dtype = d->dtype == FAMI_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap tmX, tmW;
  {
    // a pitch of >= 64 slots lets the box read whole 128-byte rows without out-of-bounds fill on the channel axis
    // (slots C..63 are never consumed: the gather only reads channels < C)
    cuuint64_t dims[4] = {(cuuint64_t)(d->x_pitch >= 64 ? 64 : d->C), (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)d->x_pitch * 2, (cuuint64_t)d->W * d->x_pitch * 2,
                             (cuuint64_t)d->H * d->W * d->x_pitch * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)p.WW, (cuuint32_t)p.WH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode_tiled(&tmX, tm_dtype, 4, const_cast<void*>(x), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FAMI_CHECK_ARG(r == CUDA_SUCCESS, "dcn_tc: cuTensorMapEncodeTiled(x) failed (%d)", (int)r);
  }
  {
    // weights packed by fami_pack_conv_weight(half): [CoutPad][9 taps][64] (Cin <= 64 -> one chunk per tap)
    cuuint64_t dims[2] = {(cuuint64_t)9 * 64, (cuuint64_t)p.BN};
    cuuint64_t strides[1] = {(cuuint64_t)9 * 64 * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)p.BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(&tmW, tm_dtype, 2, const_cast<void*>(w), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FAMI_CHECK_ARG(r == CUDA_SUCCESS, "dcn_tc: cuTensorMapEncodeTiled(w) failed (%d)", (int)r);
  }
  const size_t smem = dcn_tc_smem(d, p.R);
  FAMI_CHECK_ARG(smem <= 227 * 1024, "dcn_tc: shared memory budget exceeded (%zu B)", smem);
  int grid = p.total_tiles;
  const int sms = num_sms();
  if (grid > sms) grid = sms;
#define FAMI_DCN_LAUNCH(TH_, G_)                                                                              \
  {                                                                                                            \
    static bool attr_done = false;                                                                             \
    if (!attr_done) {                                                                                          \
      cudaFuncSetAttribute(dcn_tc_kernel<TH_, G_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);   \
      attr_done = true;                                                                                        \
    }                                                                                                          \
    dcn_tc_kernel<TH_, G_><<<grid, kDcnThreads, smem, st>>>(tmX, tmW, p);                                      \
  }
#define FAMI_DCN_LAUNCH_G(TH_)                                                  \
  switch (d->G) {                                                                \
    case 4: FAMI_DCN_LAUNCH(TH_, 4) break;                                       \
    case 8: FAMI_DCN_LAUNCH(TH_, 8) break;                                       \
    case 12: FAMI_DCN_LAUNCH(TH_, 12) break;                                     \
    default: FAMI_DCN_LAUNCH(TH_, 16) break;                                     \
  }
  if (d->dtype == FAMI_F16) {
    FAMI_DCN_LAUNCH_G(__half)
  } else {
    FAMI_DCN_LAUNCH_G(__nv_bfloat16)
  }
#undef FAMI_DCN_LAUNCH_G
#undef FAMI_DCN_LAUNCH
  FAMI_CHECK_LAUNCH("dcn_tc_kernel");
  return 0;
}

int debug_read_dcn_trace(unsigned long long* host_out, int n) {
  if (n > 4096) n = 4096;
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(host_out, g_dcn_trace, (size_t)n * sizeof(unsigned long long)) == cudaSuccess ? 0 : 1;
}

}  // namespace fami
