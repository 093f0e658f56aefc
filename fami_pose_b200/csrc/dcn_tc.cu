// Modulated deformable convolution v2 on B200 -- the north-star kernel (fami_dcn_fwd, 16-bit arm).
// Replaces torchvision.ops.deform_conv2d as called at posetimation/zoo/Alignment/Alignment_V15.py:146-158.
//
// Structure (persistent CTAs, one 16x8-pixel tile at a time, 640 threads = 20 warps x 96 registers):
//   * the x neighbourhood of the tile ((16+2R) x (8+2R) pixels x 64 channel slots) is brought ONCE into
//     shared memory by a TMA tiled box load, UNSWIZZLED with a 128-byte pixel pitch (out-of-image pixels
//     zero-filled, which is exactly torchvision's "corner outside the image contributes 0" rule);
//   * 16 gather warps, warp w = image row w of the tile.  A LANE is (pixel, offset group): the 4-channel
//     (8-byte) corner of group g sits at byte 8g of its pixel's 128-byte row, i.e. in bank pair g WHATEVER the
//     sampling position is -- lanes with different groups can never conflict, for any offsets (the
//     one-lane-per-pixel mappings before were 2.9x over the ideal wavefront count at sigma = 2 px).  The 8*G
//     samples of a (row, tap) are walked 32 at a time (LaneMap).  Per sample a lane reads its (dy, dx, mask)
//     -- streamed from HBM exactly once, in the row-blocked layout the fused offset|mask convolution writes
//     (128 contiguous bytes per load instruction), pulled into L2 two tiles ahead by bulk prefetches and into
//     registers two taps ahead --, forms the four bilinear corners from shared memory, blends them in packed
//     16-bit arithmetic and stores the modulated 4 channels into the tap's A tile (128B-swizzled K-major).
//     Samples outside the staged window are resolved per tap through a bounds-checked global path;
//   * one warp (also an epilogue warp) issues tcgen05.mma (M=128, N=Cout, K=C per tap) against the weights resident in shared
//     memory, accumulating the nine taps in TMEM (double buffered across tiles);
//   * 4 epilogue warps add the bias and store the tile (a lane owns a pixel row: whole 32-byte sectors per store).
// The [C*9, B*H*W] column buffer of the reference never exists; HBM traffic is the algorithmic minimum:
// offsets+masks (4*27*G B/pixel) + x (~once, via L2) + out.
#include <stdlib.h>

#include "tc_common.cuh"

namespace fami {

__device__ unsigned long long g_dcn_trace[4096];   // [tile it < 16][16 events] of CTA 0, FAMI_DCN_TRACE=1

namespace {

__device__ __forceinline__ void wtrace(int on, uint32_t tcount, int warp, int ev) {
  // one steady-state pair of taps (taps 3 and 4 of the CTA's 4th tile) seen by every warp of CTA 0; compiled into the
  // probes library only (csrc/build.py --probes), the product kernel carries no per-tap instrumentation
#ifdef FAMI_DEBUG_PROBES
  if (on && blockIdx.x == 0 && (tcount == 30u || tcount == 31u) && (threadIdx.x & 31) == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_dcn_trace[2048 + (tcount - 30u) * 256 + warp * 8 + ev] = t;
  }
#endif
}
__device__ __forceinline__ void dtrace(int on, int it, int ev) {
  if (on && blockIdx.x == 0 && it < 16) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_dcn_trace[it * 16 + ev] = t;
  }
}

constexpr int kTH = 16, kTW = 8;            // output tile (pixels): 128 = one UMMA M tile
#ifndef FAMI_DCN_RPW
#define FAMI_DCN_RPW 1
#endif
constexpr int kRPW = FAMI_DCN_RPW;        // tile rows per gather warp
constexpr int kGatherWarps = kTH / kRPW;  // gather warp w owns image rows kRPW*w .. kRPW*w + kRPW-1 of the tile
constexpr int kGatherThreads = 32 * kGatherWarps;
constexpr int kDcnEpiWarps = 4;
constexpr int kDcnThreads = kGatherThreads + 32 * kDcnEpiWarps;   // 640: 20 warps x 96 registers fill the register file
                                                                     // (the first epilogue warp is also the TMA + MMA issuer)
#ifndef FAMI_DCN_ASTAGES
#define FAMI_DCN_ASTAGES 2
#endif
constexpr int kAStages = FAMI_DCN_ASTAGES; // A stages (one tap each), walked round-robin by a running tap counter
constexpr int kATile = 128 * 128;          // bytes per A stage
constexpr int kBarSlots = 13 + 2 * kAStages; // mbarriers of the kernel (8 B each); the TMEM slot and the bias table follow
constexpr uint32_t kMagicBits = 0x4B400000u;   // 1.5 * 2^23: adding it with round-down leaves floor(v) in the low mantissa bits
constexpr float kMagic = 12582912.f;

// lane <-> (pixel, offset group) map of a gather warp: the 8*G samples of a (tile row, tap) are numbered s = pixel*G + g and
// taken 32 at a time, lane = s % 32: NIT = G/4 iterations, every lane busy in every iteration.  For G = 4, 8, 16 a half-warp
// holds whole pixels (all its groups distinct: one wavefront per 8-byte corner load); for G = 12 sixteen consecutive samples
// wrap around the groups once, so four lanes share a bank pair with four others (two wavefronts) -- cheaper than leaving a
// quarter of the lanes idle in every instruction of the sample loop.
template <int kG> struct LaneMap {
  static constexpr int NIT = kG / 4;
};

struct DcnTcParams {
  int B, H, W, C, Cout, G, d, R;
  int WH, WW;                       // staged window (pixels)
  int tiles_x, tiles_y, total_tiles;
  int nk, BN;
  int npass;                        // channel passes of 64 channels (C > 64), else 1
  int w_stream;                     // 0: the 9 weight tiles are resident; NW > 0: streamed through NW stages per (pass, tap)
  int om_pitch, x_pitch, out_pitch, vec_ok, out_f32;
  int om_blocked;                   // 1: offsets|masks in the row-blocked layout (om_layout 2)
  int64_t om_tap_stride;            // floats between taps in the blocked layout = tiles * 128 * 3G
  uint32_t win_bytes, w_tile_bytes, ab_format;
  int trace;
  int ablate;                       // FAMI_DCN_ABLATE (timing experiments, results wrong): 1 no epilogue stores, 2 no window
                                    // reload, 4 no corner loads / blends, 8 no far path, 16 no offset loads; 32 (results right): issuer's epilogue not deferred; 64 no fence.proxy.async
  const float* om;                  // layout 1: [B*H*W][om_pitch], per pixel [9 taps][dy(G) | dx(G) | mask(G)]
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
  static __device__ __forceinline__ t pack(float a, float b) { return __floats2half2_rn(a, b); }
  static __device__ __forceinline__ t lo(t v) { return __low2half2(v); }
  static __device__ __forceinline__ t hi(t v) { return __high2half2(v); }
  static __device__ __forceinline__ t fma(t a, t b, t c) { return __hfma2(a, b, c); }
  static __device__ __forceinline__ t mul(t a, t b) { return __hmul2(a, b); }
};
template <> struct H2<__nv_bfloat16> {
  typedef __nv_bfloat162 t;
  static __device__ __forceinline__ t pack(float a, float b) { return __floats2bfloat162_rn(a, b); }
  static __device__ __forceinline__ t lo(t v) { return __low2bfloat162(v); }
  static __device__ __forceinline__ t hi(t v) { return __high2bfloat162(v); }
  static __device__ __forceinline__ t fma(t a, t b, t c) { return __hfma2(a, b, c); }
  static __device__ __forceinline__ t mul(t a, t b) { return __hmul2(a, b); }
};

// streamed once: no L1 allocation (the three loads of an iteration touch disjoint 32-byte sectors)
__device__ __forceinline__ float ldg_stream(const float* p) {
  float v;
  asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// (dy, dx, mask) of one tap for the NIT iterations of a lane
template <int NIT> struct OmRegs { float dy[NIT], dx[NIT], mk[NIT]; };

template <typename TH, int kG>
__global__ void __launch_bounds__(kDcnThreads, 1)
dcn_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const DcnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // every address below is a 32-bit shared-space address derived ONCE from the aligned base (generic pointers only for the
  // few plain C accesses): the compiler otherwise rematerialises the generic -> shared conversion in the inner loops
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t sbase = (raw_u32 + 1023u) & ~1023u;
  const uint32_t win_u32 = sbase;                                          // WH*WW rows x 128 B
  const uint32_t a_u32 = win_u32 + ((p.win_bytes + 1023u) & ~1023u);       // kAStages x 16 KB (1 KB-aligned swizzle atoms)
  const uint32_t w_u32 = a_u32 + kAStages * kATile;                        // 9 (resident) or NW (streamed) x BN x 128 B
  const uint32_t bar0 = w_u32 + (uint32_t)(p.w_stream ? p.w_stream : 9) * p.w_tile_bytes;
  // barriers: win_full, win_free, w_full, a_full[3], a_empty[3], tfull[2], tempty[2], ws_full[3], ws_empty[3]
  const uint32_t win_full = bar0, win_free = bar0 + 8, w_full = bar0 + 16;
  auto a_full = [&](int s) { return bar0 + 8u * (3 + s); };
  auto a_empty = [&](int s) { return bar0 + 8u * (3 + kAStages + s); };
  auto tfull = [&](int a) { return bar0 + 8u * (3 + 2 * kAStages + a); };
  auto tempty = [&](int a) { return bar0 + 8u * (5 + 2 * kAStages + a); };
  auto ws_full = [&](int s) { return bar0 + 8u * (7 + 2 * kAStages + s); };
  auto ws_empty = [&](int s) { return bar0 + 8u * (10 + 2 * kAStages + s); };
  uint8_t* gen0 = smem_raw + (bar0 - raw_u32);                             // generic pointer to the barrier block
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen0 + 8 * kBarSlots);
  float* s_scale = reinterpret_cast<float*>(gen0 + 8 * (kBarSlots + 3));   // 16-byte aligned (float4 reads)
  float* s_shift = s_scale + p.BN;
  fill_scale_shift(s_scale, s_shift, nullptr, p.bias, p.Cout, p.BN);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(win_full, 1);
    mbar_init(win_free, kGatherWarps);
    mbar_init(w_full, 1);
    for (int s = 0; s < kAStages; ++s) { mbar_init(a_full(s), kGatherWarps); mbar_init(a_empty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), kDcnEpiWarps); }
    for (int w = 0; w < 3; ++w) { mbar_init(ws_full(w), 1); mbar_init(ws_empty(w), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  if (warp == kGatherWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bar0 + 8u * kBarSlots), "r"(512u)
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
    constexpr int NIR = LaneMap<kG>::NIT;                  // iterations per tile row
    constexpr int NIT = kRPW * NIR;                        // iterations per tap of this warp: j = row * NIR + jj
    typedef typename H2<TH>::t h2;
    typedef OmRegs<NIT> Om;
    const TH* xg = reinterpret_cast<const TH*>(p.x);
    const int ry = warp * kRPW;                            // first image row of the tile owned by this warp
    const uint32_t rowpitch = (uint32_t)p.WW * 128u;
    const uint32_t ylim = (uint32_t)(p.WH - 1), xlim = (uint32_t)(p.WW - 1);
    const uint32_t win_safe = win_u32;
    // per iteration j of this lane: sample s = 32 j + lane = pixel * G + g.
    //   kaddr[j]: window address of the sample = iy_bits * rowpitch + ix_bits * 128 + kaddr[j], the bit patterns of the
    //             magic-number floors used directly (32-bit wrap-around arithmetic; ix is relative to pixel 0 of the row)
    //   xbias[j]: ix_bits - xbias[j] = window column of the sample (bounds test)
    //   a_st[j] : A-tile byte address of (row r = ry*8 + pixel, group g), 128B-swizzled K-major: chunk (g>>1) ^ (r&7)
    uint32_t kaddr[NIT], xbias[NIT], a_st[NIT];
    int pixj[NIT], gj[NIT];
#pragma unroll
    for (int j = 0; j < NIT; ++j) {
      const int rj = j / NIR, sidx = 32 * (j - rj * NIR) + lane;      // row within the warp, sample within the row
      pixj[j] = sidx / kG;
      gj[j] = sidx - pixj[j] * kG;
      kaddr[j] = win_u32 + (uint32_t)(gj[j] * 8 + pixj[j] * 128) + (uint32_t)rj * rowpitch - kMagicBits * rowpitch - kMagicBits * 128u;
      xbias[j] = kMagicBits - (uint32_t)pixj[j];
      a_st[j] = a_u32 + (uint32_t)(((ry + rj) * kTW + pixj[j]) * 128) + ((uint32_t)(((gj[j] >> 1) ^ pixj[j]) << 4) | ((uint32_t)(gj[j] & 1) << 3));
    }
    const float fd = (float)p.d;
    const float my0 = kMagic + (float)(ry + p.R - p.d);     // + kr * d per kernel row
    const float mx0 = kMagic + (float)(p.R - p.d);          // + kc * d per kernel column (pixel 0 of the row)
    const float mx1 = mx0 + fd, mx2 = mx1 + fd;

    // A CTA walks a stream of units = (tile, channel pass of 64 channels); C <= 64 has one pass per tile.
    // per-unit state of this lane: pointer to its (dy) float of tap 0 / iteration 0, and the valid-iteration mask
    const int NP = p.npass, Gt = p.G;
    struct TileRef { const float* po; uint32_t vmask; int tile; int pass; };
    auto unit_ref = [&](int useq) {
      TileRef t;
      t.po = p.om; t.vmask = 0;
      const int ti = useq / NP;
      t.pass = useq - ti * NP;
      t.tile = blockIdx.x + ti * gridDim.x;
      if (t.tile >= p.total_tiles) return t;
      int b, y0, x0;
      tile_origin(t.tile, b, y0, x0);
      const int y = y0 + ry;
#pragma unroll
      for (int j = 0; j < NIT; ++j)
        if (y + j / NIR < p.H && x0 + pixj[j] < p.W) t.vmask |= 1u << j;
      if (p.om_blocked) {
        // [tap][tile][row 16][dy | dx | mask][pixel 8][group G]; one pass: sample s of iteration j at float 32 j + lane of
        // its run; passes (16 groups each): lanes 0-15 / 16-31 are pixels 2j / 2j+1, groups 16 pass .. 16 pass + 15
        const int lane_part = (kG == 16) ? (lane >> 4) * Gt + (lane & 15) + 16 * t.pass : lane;
        t.po = p.om + (int64_t)t.tile * (128 * 3 * Gt) + ry * (3 * 8 * Gt) + lane_part;
      } else {
        t.po = p.om + ((int64_t)(b * p.H + (y < p.H ? y : 0)) * p.W + x0) * p.om_pitch + 16 * t.pass;
      }
      return t;
    };
    auto load_tap = [&](const TileRef& t, int tap, Om& o) {
      if (p.ablate & 16) {
#pragma unroll
        for (int j = 0; j < NIT; ++j) { o.dy[j] = 0.25f; o.dx[j] = 0.25f; o.mk[j] = 1.f; }
        return;
      }
      if (p.om_blocked) {
        const float* q = t.po + tap * p.om_tap_stride;
#pragma unroll
        for (int j = 0; j < NIT; ++j) {
          const bool v = (t.vmask >> j) & 1u;
          const int rj = j / NIR, jj = j - rj * NIR;
          const int jo = rj * (3 * 8 * Gt) + ((kG == 16) ? 2 * jj * Gt : 32 * jj);
          o.dy[j] = v ? ldg_stream(q + jo) : 0.f;
          o.dx[j] = v ? ldg_stream(q + 8 * Gt + jo) : 0.f;
          o.mk[j] = v ? ldg_stream(q + 16 * Gt + jo) : 0.f;
        }
      } else {
        const float* q = t.po + tap * 3 * Gt;
#pragma unroll
        for (int j = 0; j < NIT; ++j) {
          const bool v = (t.vmask >> j) & 1u;
          const float* qj = q + ((int64_t)(j / NIR) * p.W + pixj[j]) * p.om_pitch + gj[j];
          o.dy[j] = v ? __ldg(qj) : 0.f;
          o.dx[j] = v ? __ldg(qj + Gt) : 0.f;
          o.mk[j] = v ? __ldg(qj + 2 * Gt) : 0.f;
        }
      }
    };
    // load of the tap two positions ahead in the (unit, tap) stream
    auto load_ahead = [&](const TileRef& cur, const TileRef& nxt, int tap, Om& o) {
      if (tap + 2 < 9) load_tap(cur, tap + 2, o);
      else load_tap(nxt, tap + 2 - 9, o);
    };

    // one tap of one unit: NIT samples of this lane into A stage kc
    auto do_tap = [&](const TileRef& t, const Om& o, int kr, float my, float mx, int kc, uint32_t tcount) {
      const uint32_t stage = tcount % kAStages, u = tcount / kAStages;   // A stage of this tap and its use index
      const uint32_t st_off = stage * kATile;
      wtrace(p.trace, tcount, warp, 0);                       // tap starts (offsets of the tap already in registers)
      bool far = false;                          // any sample of this lane outside the staged window
      constexpr int NB = (NIR % 2) ? NIR : 2;    // samples per batch (8 or 12 corner loads in flight per lane); NB divides NIT
      static_assert(NIT % NB == 0, "batch size must divide the iteration count");
      uint2 pk[NIT];
#pragma unroll
      for (int j = 0; j < NIT; ++j) pk[j] = make_uint2(0u, 0u);
      if (!(p.ablate & 4))
#pragma unroll
      for (int j0 = 0; j0 < NIT; j0 += NB) {
        uint32_t a00[NB], a10[NB];
        h2 w12[NB], w34[NB];
#pragma unroll
        for (int jb = 0; jb < NB; ++jb) {
          const int j = j0 + jb;
          const float dy = o.dy[j], dx = o.dx[j], mk = o.mk[j];
          // floor() through the magic-number add (round-down): exact for |v| < 2^22, the integer lands in the low mantissa bits
          const float ty = __fadd_rd(dy, my), tx = __fadd_rd(dx, mx);
          const float ly = dy - (ty - my), lx = dx - (tx - mx);
          const uint32_t iyb = __float_as_uint(ty), ixb = __float_as_uint(tx);
          const bool in = (iyb - (kMagicBits - (uint32_t)(j / NIR))) < ylim && (ixb - xbias[j]) < xlim;
          const float wb = mk * ly, wt = mk - wb;          // mask * (ly | 1 - ly)
          const float w4f = wb * lx, w3f = wb - w4f, w2f = wt * lx, w1f = wt - w2f;
          w12[jb] = H2<TH>::pack(w1f, w2f);
          w34[jb] = H2<TH>::pack(w3f, w4f);
          const uint32_t a = iyb * rowpitch + (ixb * 128u + kaddr[j]);
          a00[jb] = in ? a : win_safe;                      // outside the staged window: harmless address, fixed up below
          a10[jb] = a00[jb] + rowpitch;
          far |= !in;      // (pixels outside the image carry zero offsets: always inside)
        }
        uint2 u1[NB], u2[NB], u3[NB], u4[NB];
#pragma unroll
        for (int jb = 0; jb < NB; ++jb) {
          const int j = j0 + jb;
          u1[jb] = lds64(a00[jb]);
          u2[jb] = lds64(a00[jb] + 128u);
          u3[jb] = lds64(a10[jb]);
          u4[jb] = lds64(a10[jb] + 128u);
        }
#pragma unroll
        for (int jb = 0; jb < NB; ++jb) {
          const int j = j0 + jb;
          const h2 w1 = H2<TH>::lo(w12[jb]), w2 = H2<TH>::hi(w12[jb]), w3 = H2<TH>::lo(w34[jb]), w4 = H2<TH>::hi(w34[jb]);
          h2 lo = H2<TH>::mul(w1, *reinterpret_cast<const h2*>(&u1[jb].x));
          h2 hi = H2<TH>::mul(w1, *reinterpret_cast<const h2*>(&u1[jb].y));
          lo = H2<TH>::fma(w2, *reinterpret_cast<const h2*>(&u2[jb].x), lo);
          hi = H2<TH>::fma(w2, *reinterpret_cast<const h2*>(&u2[jb].y), hi);
          lo = H2<TH>::fma(w3, *reinterpret_cast<const h2*>(&u3[jb].x), lo);
          hi = H2<TH>::fma(w3, *reinterpret_cast<const h2*>(&u3[jb].y), hi);
          lo = H2<TH>::fma(w4, *reinterpret_cast<const h2*>(&u4[jb].x), lo);
          hi = H2<TH>::fma(w4, *reinterpret_cast<const h2*>(&u4[jb].y), hi);
          // (a pixel outside the image carries mask 0: all four weights are 0 and the column is 0)
          pk[j].x = *reinterpret_cast<const uint32_t*>(&lo);
          pk[j].y = *reinterpret_cast<const uint32_t*>(&hi);
        }
      }
      // the A stage is needed only now: the wait for the MMA that last read it hides behind the loads and blends above
      wtrace(p.trace, tcount, warp, 1);                       // loads and blends done
      mbar_wait(a_empty(stage), (u & 1u) ^ 1u);
      wtrace(p.trace, tcount, warp, 2);                       // A stage free
#pragma unroll
      for (int j = 0; j < NIT; ++j) sts64(a_st[j] + st_off, pk[j]);
      // large offsets: bounds-checked global corners, fp32 blend.  Kept out of the sample loop and entered once per tap
      // by the whole warp, so the dependent global loads of all far samples of the tap are in flight together.
      if (__any_sync(0xffffffffu, far) && !(p.ablate & 8)) {
        uint32_t slow = 0;
#pragma unroll
        for (int j = 0; j < NIT; ++j) {
          const uint32_t iyb = __float_as_uint(__fadd_rd(o.dy[j], my)), ixb = __float_as_uint(__fadd_rd(o.dx[j], mx));
          if (!((iyb - (kMagicBits - (uint32_t)(j / NIR))) < ylim && (ixb - xbias[j]) < xlim)) slow |= 1u << j;
        }
        while (slow) {
          const int j = __ffs((int)slow) - 1;
          slow &= slow - 1u;
          float sdy = 0.f, sdx = 0.f, smk = 0.f;
          int spix = 0, g = 0;
          uint32_t sa = 0;
#pragma unroll
          for (int jj = 0; jj < NIT; ++jj)
            if (jj == j) { sdy = o.dy[jj]; sdx = o.dx[jj]; smk = o.mk[jj]; spix = pixj[jj]; g = gj[jj]; sa = a_st[jj]; }
          int tb, ty0, tx0;
          tile_origin(t.tile, tb, ty0, tx0);
          const int y = ty0 + ry + j / NIR, x = tx0 + spix;
          const float py = (float)(y - p.d + kr * p.d) + sdy;
          const float px = (float)(x - p.d + kc * p.d) + sdx;
          uint2 pk2 = make_uint2(0u, 0u);
          if (py > -1.f && py < (float)p.H && px > -1.f && px < (float)p.W) {
            const int iy0 = (int)floorf(py), ix0 = (int)floorf(px);
            const float ly = py - (float)iy0, lx = px - (float)ix0, hy = 1.f - ly, hx = 1.f - lx;
            const TH* xb = xg + (int64_t)tb * p.H * p.W * p.x_pitch + (16 * t.pass + g) * 4;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const bool y0ok = iy0 >= 0, y1ok = iy0 + 1 <= p.H - 1, x0ok = ix0 >= 0, x1ok = ix0 + 1 <= p.W - 1;
            const float4 v1 = (y0ok && x0ok) ? ld4h<TH>(xb + ((int64_t)iy0 * p.W + ix0) * p.x_pitch) : z;
            const float4 v2 = (y0ok && x1ok) ? ld4h<TH>(xb + ((int64_t)iy0 * p.W + ix0 + 1) * p.x_pitch) : z;
            const float4 v3 = (y1ok && x0ok) ? ld4h<TH>(xb + ((int64_t)(iy0 + 1) * p.W + ix0) * p.x_pitch) : z;
            const float4 v4 = (y1ok && x1ok) ? ld4h<TH>(xb + ((int64_t)(iy0 + 1) * p.W + ix0 + 1) * p.x_pitch) : z;
            const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
            pk2.x = f2_to_h2<TH>(smk * (w1 * v1.x + w2 * v2.x + w3 * v3.x + w4 * v4.x),
                                 smk * (w1 * v1.y + w2 * v2.y + w3 * v3.y + w4 * v4.y));
            pk2.y = f2_to_h2<TH>(smk * (w1 * v1.z + w2 * v2.z + w3 * v3.z + w4 * v4.z),
                                 smk * (w1 * v1.w + w2 * v2.w + w3 * v3.w + w4 * v4.w));
          }
          sts64(sa + st_off, pk2);
        }
        __syncwarp();
      }
      wtrace(p.trace, tcount, warp, 3);                       // stores (and far path) done
      if (!(p.ablate & 64))
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the MMA (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full(stage));   // one arrival per warp
      wtrace(p.trace, tcount, warp, 4);                       // fenced and arrived
    };

    uint32_t wph = 0, fph = 0, tcount = 0;       // tcount: taps gathered so far (all units)
    const int my_tiles = ((int)blockIdx.x < p.total_tiles) ? (p.total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int n_units = my_tiles * NP;
    TileRef cur = unit_ref(0), nxt = unit_ref(1);
    Om o0, o1, o2;                       // taps 3i, 3i+1, 3i+2 of the stream: rotating, two taps in flight
    load_tap(cur, 0, o0);
    load_tap(cur, 1, o1);
    for (int useq = 0; useq < n_units; ++useq) {
      if (threadIdx.x == 0) dtrace(p.trace, useq, 0);
      if (!(p.ablate & 2) || useq == 0) {
        mbar_wait(win_full, wph);
        wph ^= 1u;
      }
      if (threadIdx.x == 0) dtrace(p.trace, useq, 1);
      float my = my0;
#pragma unroll 1
      for (int kr = 0; kr < 3; ++kr, my += fd) {
        load_ahead(cur, nxt, kr * 3 + 0, o2);
        do_tap(cur, o0, kr, my, mx0, 0, tcount);
        load_ahead(cur, nxt, kr * 3 + 1, o0);
        do_tap(cur, o1, kr, my, mx1, 1, tcount + 1);
        load_ahead(cur, nxt, kr * 3 + 2, o1);
        do_tap(cur, o2, kr, my, mx2, 2, tcount + 2);
        tcount += 3;
        if (threadIdx.x == 0) dtrace(p.trace, useq, 2 + kr);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(win_free);   // this warp no longer reads the window of this unit
      if (warp == 0) {
        // refill the window for the next unit as soon as the last gather warp has left it (issued from here,
        // not from the MMA warp, which is still draining the last taps)
        mbar_wait(win_free, fph);
        fph ^= 1u;
        if (lane == 0 && useq + 1 < n_units && !(p.ablate & 2)) {
          int nb, ny0, nx0;
          tile_origin(nxt.tile, nb, ny0, nx0);
          mbar_arrive_expect_tx(win_full, p.win_bytes);
          tma_tiled_4d(win_u32, &tmX, win_full, 64 * nxt.pass, nx0 - p.R, ny0 - p.R, nb);
        }
        __syncwarp();
      }
      cur = nxt;
      nxt = unit_ref(useq + 2);
    }
  } else {
    // ===================== epilogue warps (the first one is also the TMA + MMA issuer) =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    // Direct epilogue: a lane owns one accumulator row (pixel) and stores its Cout channels itself, 16 at a time (32 / 64
    // contiguous bytes per lane and chunk: whole sectors).  The four epilogue warps have a whole tile time to move 12 KB, so
    // the staged routine of the conv kernels buys nothing here, and its 18 KB of staging pay for a window with 6 px of reach.
    const uint32_t s_shift_u32 = bar0 + 8u * (kBarSlots + 3) + (uint32_t)p.BN * 4u;
    auto epi_tile = [&](int tile, int it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      int b, y0, x0;
      tile_origin(tile, b, y0, x0);
      const int y = y0 + (row >> 3), x = x0 + (row & 7);
      const bool valid = y < p.H && x < p.W && !(p.ablate & 1);
      const int64_t pix = valid ? ((int64_t)b * p.H + y) * p.W + x : 0;
      mbar_wait(tfull(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t)(acc * p.BN) + ((uint32_t)(quarter * 32) << 16);
      for (int c0 = 0; c0 < p.Cout; c0 += 16) {          // warp-uniform
        uint32_t v[16];
        tmem_ld16(t_addr + (uint32_t)c0, v);
        tmem_ld_wait();
        float o[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 sh = lds128f(s_shift_u32 + (uint32_t)(c0 + 4 * q) * 4u);     // bias (0 beyond Cout)
          o[4 * q + 0] = __uint_as_float(v[4 * q + 0]) + sh.x;
          o[4 * q + 1] = __uint_as_float(v[4 * q + 1]) + sh.y;
          o[4 * q + 2] = __uint_as_float(v[4 * q + 2]) + sh.z;
          o[4 * q + 3] = __uint_as_float(v[4 * q + 3]) + sh.w;
        }
        const int nc = p.Cout - c0 < 16 ? p.Cout - c0 : 16;
        if (valid) {          // (no early exit: the TMEM loads above are warp-collective)
          if (p.out_f32) {
            float* dst = reinterpret_cast<float*>(p.out) + pix * p.out_pitch + c0;
            if (p.vec_ok && nc == 16) {
#pragma unroll
              for (int q = 0; q < 4; ++q) reinterpret_cast<float4*>(dst)[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
            } else {
#pragma unroll
              for (int c = 0; c < 16; ++c)
                if (c < nc) dst[c] = o[c];
            }
          } else {
            TH* dst = reinterpret_cast<TH*>(p.out) + pix * p.out_pitch + c0;
            if (p.vec_ok && nc == 16) {
              uint4 w0, w1;
              w0.x = f2_to_h2<TH>(o[0], o[1]); w0.y = f2_to_h2<TH>(o[2], o[3]); w0.z = f2_to_h2<TH>(o[4], o[5]); w0.w = f2_to_h2<TH>(o[6], o[7]);
              w1.x = f2_to_h2<TH>(o[8], o[9]); w1.y = f2_to_h2<TH>(o[10], o[11]); w1.z = f2_to_h2<TH>(o[12], o[13]); w1.w = f2_to_h2<TH>(o[14], o[15]);
              reinterpret_cast<uint4*>(dst)[0] = w0;
              reinterpret_cast<uint4*>(dst)[1] = w1;
            } else {
#pragma unroll
              for (int c = 0; c < 16; ++c)
                if (c < nc) dst[c] = from_f<TH>(o[c]);
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(acc));
    };
    if (warp == kGatherWarps) {
      // ---- TMA + MMA issuer (warp-uniform control flow, elected lane issues), then its quarter of the epilogue ----
      const bool leader = elect_one();
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t idesc = (1u << 4) | (p.ab_format << 7) | (p.ab_format << 10) | ((uint32_t)(p.BN >> 3) << 17) |
                             ((uint32_t)(128 >> 4) << 24);
      // pull a tile's offsets|masks (16 row segments of 8 pixels x 27G floats, the dominant HBM stream) into
      // L2 one tile ahead, so the gather warps' dependent loads see L2 rather than HBM latency
      auto prefetch_om = [&](int b, int y0, int x0) {
        if (p.om_blocked) {
          // 9 taps x (128 pixels x 3G floats) contiguous per tile
          const int tile_id = (b * p.tiles_y + y0 / kTH) * p.tiles_x + x0 / kTW;
          const uint32_t bytes = (uint32_t)(128 * 3 * p.G * 4);
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
        if (!p.w_stream) {
          mbar_arrive_expect_tx(w_full, 9u * p.w_tile_bytes);
          for (int t = 0; t < 9; ++t) tma_tiled_2d(w_u32 + t * p.w_tile_bytes, &tmW, w_full, t * 64, 0);
        }
      }
      {
        int b, y0, x0;
        if ((int)blockIdx.x < p.total_tiles) {
          tile_origin(blockIdx.x, b, y0, x0);
          if (leader) {
            mbar_arrive_expect_tx(win_full, p.win_bytes);
            tma_tiled_4d(win_u32, &tmX, win_full, 0, x0 - p.R, y0 - p.R, b);
          }
          prefetch_om(b, y0, x0);
        }
        if ((int)(blockIdx.x + gridDim.x) < p.total_tiles) {
          tile_origin(blockIdx.x + gridDim.x, b, y0, x0);
          prefetch_om(b, y0, x0);
        }
      }
      // Weights: resident (C <= 64: nine [BN][64] tiles loaded once) or streamed (C > 64: one [BN][64] tile per (pass, tap)
      // through NW stages; the slot of the MMA issued one step earlier is refilled, so the issuer never waits on its own MMA).
      const int NP = p.npass, NW = p.w_stream;
      const int my_tiles = ((int)blockIdx.x < p.total_tiles) ? (p.total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
      const int w_total = my_tiles * NP * 9;               // streamed tiles this CTA consumes
      auto w_issue = [&](int id) {                         // tile id of the stream -> (pass, tap) of its unit
        const int k = id % (NP * 9);
        const int pass = k / 9, tap = k - pass * 9;
        const int slot = id % NW;
        if (leader) {
          mbar_arrive_expect_tx(ws_full(slot), p.w_tile_bytes);
          tma_tiled_2d(w_u32 + slot * p.w_tile_bytes, &tmW, ws_full(slot), (tap * NP + pass) * 64, 0);
        }
      };
      if (NW) {
        for (int i = 0; i < NW && i < w_total; ++i) w_issue(i);
      } else {
        mbar_wait(w_full, 0);
      }
      tc_fence_after();
      const uint32_t w_lo0 = sw128_desc_lo(w_u32);
      const uint32_t w_step = p.w_tile_bytes >> 4;
      int stage = 0;
      uint32_t aph = 0;
      int it = 0, wid = 0, last_tile = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; last_tile = tile, tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(tempty(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_u + (uint32_t)(acc * p.BN);
        for (int pass = 0; pass < NP; ++pass) {
          uint32_t w_lo = w_lo0;
          for (int tap = 0; tap < 9; ++tap, w_lo += w_step) {
            // this warp's quarter of the PREVIOUS tile's epilogue, once the first two taps of this tile are issued: the wait
            // for the previous tile's last MMAs and the drain then overlap the gather instead of holding up taps 0 and 1
            if (pass == 0 && tap == 2 && it > 0 && !(p.ablate & 32)) epi_tile(last_tile, it - 1);
            mbar_wait(a_full(stage), aph);
            wtrace(p.trace, (uint32_t)((it * NP + pass) * 9 + tap), 16, 0);    // all sixteen arrivals seen
            int slot = 0;
            if (NW) {
              slot = wid % NW;
              mbar_wait(ws_full(slot), (uint32_t)(wid / NW) & 1u);
            }
            tc_fence_after();
            const uint32_t a_lo = sw128_desc_lo(a_u32 + stage * kATile);
            umma_ksteps_n(p.nk, leader, d_tmem, a_lo, NW ? w_lo0 + (uint32_t)slot * w_step : w_lo, idesc, (pass | tap) != 0);
            if (leader) umma_commit(a_empty(stage));
            wtrace(p.trace, (uint32_t)((it * NP + pass) * 9 + tap), 16, 1);    // MMAs issued, commit issued
            if (NW) {
              if (leader) umma_commit(ws_empty(slot));
              if (wid >= 1 && wid - 1 + NW < w_total) {
                const int prev = wid - 1;
                mbar_wait(ws_empty(prev % NW), (uint32_t)(prev / NW) & 1u);
                w_issue(prev + NW);
              }
              ++wid;
            }
            if (++stage == kAStages) { stage = 0; aph ^= 1u; }
          }
        }
        if (leader) umma_commit(tfull(acc));
        // pull the offsets|masks of the tile after next towards L2 (the window refill is issued by gather warp 0)
        const int next = tile + 2 * gridDim.x;
        if (next < p.total_tiles) {
          int b, y0, x0;
          tile_origin(next, b, y0, x0);
          prefetch_om(b, y0, x0);
        }
        if (p.ablate & 32) epi_tile(tile, it);          // A/B: epilogue quarter at the end of its own tile
      }
      // (this warp's quarter of a tile's epilogue runs after the first taps of the NEXT tile have been issued -- see above)
      if (it > 0 && !(p.ablate & 32)) epi_tile(last_tile, it - 1);
    } else {
      int it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) epi_tile(tile, it);
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

// weight stages: 0 = the nine tiles of the single pass are resident (C <= 64), else streamed per (pass, tap)
static int dcn_tc_wstream(const fami_dcn_desc* d) {
  if (d->C <= 64) return 0;
  const int BN = ((d->Cout + 15) / 16) * 16;
  return BN > 128 ? 2 : 3;
}
// shared-memory footprint for a window radius R; the launcher takes the largest R <= dil + 7 that fits
static size_t dcn_tc_smem(const fami_dcn_desc* d, int R) {
  const int BN = ((d->Cout + 15) / 16) * 16;
  const int NW = dcn_tc_wstream(d);
  const size_t win = ((size_t)(kTH + 2 * R) * (kTW + 2 * R) * 128 + 1023) & ~(size_t)1023;
  return win + kAStages * kATile + (size_t)(NW ? NW : 9) * BN * 128 + 1024 + 8 * (kBarSlots + 3) + 16 + (size_t)BN * 8;
}
static int dcn_tc_radius(const fami_dcn_desc* d) {
  static const int reach = getenv("FAMI_DCN_REACH") ? atoi(getenv("FAMI_DCN_REACH")) : 7;   // A/B switch, default 7 px
  for (int R = d->dil + reach; R >= d->dil + 2; --R)
    if (dcn_tc_smem(d, R) <= 227 * 1024) return R;
  return -1;
}

int dcn_tc_supported(const fami_dcn_desc* d) {
  if (!is_half_dtype(d->dtype) || (d->om_layout != 1 && d->om_layout != 2)) return 0;
  // 4 channels per offset group; one pass of C <= 64 channels (G in {4, 8, 12, 16}) or C / 64 passes of 64 channels
  if (d->C % 16 != 0 || d->G * 4 != d->C || (d->C > 64 && d->C % 64 != 0)) return 0;
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
  p.R = dcn_tc_radius(d);   // dilation reach + up to 7 px of offset (3.5 sigma of the sigma = 2 px regime); beyond -> global path
  FAMI_CHECK_ARG(p.R > 0, "dcn_tc: filter and window do not fit in shared memory");
  p.WH = kTH + 2 * p.R; p.WW = kTW + 2 * p.R;
  p.tiles_x = (d->W + kTW - 1) / kTW; p.tiles_y = (d->H + kTH - 1) / kTH;
  p.total_tiles = d->B * p.tiles_x * p.tiles_y;
  p.npass = d->C > 64 ? d->C / 64 : 1;
  p.w_stream = dcn_tc_wstream(d);
  p.nk = d->C > 64 ? 4 : d->C / 16;
  p.BN = ((d->Cout + 15) / 16) * 16;
  p.om_pitch = d->off_pitch; p.x_pitch = d->x_pitch; p.out_pitch = d->out_pitch;
  p.out_f32 = d->out_f32 ? 1 : 0;
  p.vec_ok = ((reinterpret_cast<uintptr_t>(out) & 15) == 0) && (d->out_pitch % (d->out_f32 ? 4 : 8) == 0);
  p.win_bytes = (uint32_t)(p.WH * p.WW) * 128u;
  p.w_tile_bytes = (uint32_t)p.BN * 128u;
  p.ab_format = d->dtype == FAMI_F16 ? 0u : 1u;
  p.om = om; p.x = x; p.bias = bias; p.out = out;
  p.om_blocked = d->om_layout == 2;
  p.om_tap_stride = (int64_t)p.total_tiles * 128 * 3 * d->G;
  static const bool trace_on = getenv("FAMI_DCN_TRACE") != nullptr;
  p.trace = trace_on ? atoi(getenv("FAMI_DCN_TRACE")) : 0;
  p.ablate = getenv("FAMI_DCN_ABLATE") ? atoi(getenv("FAMI_DCN_ABLATE")) : 0;

  const CUtensorMapDataType tm_dtype = d->dtype == FAMI_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap tmX, tmW;
  {
    // a pitch of >= 64 slots lets the box read whole 128-byte rows without out-of-bounds fill on the channel axis
    // (slots C..63 are never consumed: the gather only reads channels < C)
    cuuint64_t dims[4] = {(cuuint64_t)(d->C > 64 ? d->C : (d->x_pitch >= 64 ? 64 : d->C)), (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)d->x_pitch * 2, (cuuint64_t)d->W * d->x_pitch * 2,
                             (cuuint64_t)d->H * d->W * d->x_pitch * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)p.WW, (cuuint32_t)p.WH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    // unswizzled: pixel pitch 128 B, group g of every pixel in bank pair g
    CUresult r = g_encode_tiled(&tmX, tm_dtype, 4, const_cast<void*>(x), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FAMI_CHECK_ARG(r == CUDA_SUCCESS, "dcn_tc: cuTensorMapEncodeTiled(x) failed (%d)", (int)r);
  }
  {
    // weights packed by fami_pack_conv_weight(half): [CoutPad][9 taps][ceil(Cin/64) chunks][64]
    const int chunks = (d->C + 63) / 64;
    cuuint64_t dims[2] = {(cuuint64_t)9 * chunks * 64, (cuuint64_t)p.BN};
    cuuint64_t strides[1] = {(cuuint64_t)9 * chunks * 64 * 2};
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
    static std::atomic<uint64_t> attr_mask{0};                                                                 \
    set_max_smem_once(attr_mask, dcn_tc_kernel<TH_, G_>, 227 * 1024);                                          \
    dcn_tc_kernel<TH_, G_><<<grid, kDcnThreads, smem, st>>>(tmX, tmW, p);                                      \
  }
#define FAMI_DCN_LAUNCH_G(TH_)                                                  \
  switch (d->C > 64 ? 16 : d->G) {   /* groups per pass */                      \
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
