// Modulated deformable convolution v2 on B200, warp-private form (fami_dcn_fwd, 16-bit arms, C == Cout in {32, 48, 64}).
// Replaces torchvision.ops.deform_conv2d as called at posetimation/zoo/Alignment/Alignment_V15.py:146-158.
//
// dcn_tc.cu hands every tap of a 128-pixel tile from sixteen gather warps to one tcgen05 issuer: its time is the hand-over
// chain (A-tile stores, proxy fence, sixteen-warp rendezvous, commit), not bytes or math.  Here nothing is
// handed over:
//   * a compute warp owns 16 pixels (tile rows 2w, 2w+1) and gathers its samples STRAIGHT INTO THE A FRAGMENT of
//     mma.sync.m16n8k16: lane (gid, t) holds rows gid / gid+8 and k-slots {2t, 2t+1, 2t+8, 2t+9}, which are mapped to the four
//     channels of offset group 4*kstep + t (the k order of a contraction is free; the filter is repacked to match), so a
//     (pixel, group) sample -- one (dy, dx, mask), four 8-byte corners, a packed 16-bit blend -- IS two fragment registers.
//     The warp multiplies them with the filter fragments resident in shared memory and keeps the 16 x Cout fp32 accumulator in
//     registers over the nine taps: no A tile, no fence, no barrier, no TMEM; warps drift freely;
//   * C = 48 (the reference's shape): the window keeps the DENSE 96-byte pixel pitch.  The bank pair of (pixel column px, group
//     4 ks + t) is 4 * ((ks - px) mod 4) + t, so the four pixels of a half-warp reading the same k-step sit in four different
//     bank blocks when their sample columns are consecutive (smooth offsets: one wavefront per half-warp) and collide like any
//     12-group layout must when they are random (about two).  C = 32 / 64 (64 / 128-byte pitch) take their k-steps in a
//     rotated order per pixel instead (rot = gid % KS) and put the samples back in k-step order with selects;
//   * the x window is a RING over a vertical strip of tiles (8 slots of 8 image rows, 64 rows: a power of two, the wrap is one
//     AND) filled with TMA by lane 0 of warp 0, which polls the slot barriers between its own taps; a tile reads five chunks
//     while the next tile's three land, so the window refill that dcn_tc.cu exposes per tile is hidden and an image row is
//     fetched once per strip; zero fill outside the image = torchvision's rule for corners outside;
//   * offsets|masks are streamed from HBM exactly once: the producer convolution writes them k-step-blocked (om_layout 3:
//     [tile][tap][row 16][dy | dx | mask][group / 4][pixel 8][group % 4], so the 32 lanes (pixel, group % 4) of a load read one
//     128-byte line; tap-major NHWC is accepted too), loaded into registers three taps ahead (2.4 taps of lead; bulk L2 prefetches a tile or a tap ahead measured 12-17 %
//     SLOWER and are gone);
//   * samples outside the window (|dy| > 5 px or |dx| > 7 px beyond the dilation) take a bounds-checked global path that
//     costs the one warp that meets them.
#include <stdlib.h>

#include "tc_common.cuh"

namespace fami {

namespace {

constexpr int kLH = 16, kTW = 8;                        // tile of the row-blocked offset layout (fami_dcn_desc.om_layout 2)
#ifndef FAMI_WP_TILEH
#define FAMI_WP_TILEH 24
#endif
constexpr int kTH = FAMI_WP_TILEH;                      // compute tile: kTH rows x 8 columns, one m16 MMA tile per warp
constexpr int kWpWarps = kTH / 2;                       // compute warps: warp w owns tile rows 2w, 2w+1
constexpr int kWpThreads = 32 * kWpWarps;               // (warp 0's lane 0 is also the TMA producer, by polling: a thirteenth
                                                        // warp would round the register allocation up to sixteen warps)
constexpr int kRy = 8;                                  // vertical halo of the window
constexpr int kChunk = (kTH == 16) ? 16 : 8;            // image rows per ring chunk
constexpr int kNCH = (kTH + 2 * kRy) / kChunk;          // chunks a tile reads
constexpr int kADV = kTH / kChunk;                      // new chunks per tile
constexpr int kSlots = (kTH == 16) ? 3 : 8;             // ring slots: the tile's chunks + the next tile's new ones
constexpr int kRingRows = kChunk * kSlots;              // 48 | 64
constexpr bool kRingPow2 = (kRingRows & (kRingRows - 1)) == 0;
static_assert(kTH == 16 || kTH == 24, "tile height 16 or 24");
static_assert(kSlots >= kNCH + kADV, "ring too small for the prefetch of the next tile");
constexpr uint32_t kMagicBits = 0x4B400000u;            // 1.5 * 2^23: adding it with round-down leaves floor(v) in the mantissa
constexpr float kMagic = 12582912.f;

struct DcnWpParams {
  int B, H, W, C, Cout, G, d, Rx, WW;
  int tiles_x, tiles_y, ltiles_y;         // compute tiles (kTH x 8) per image; 16-row layout tiles per image column
  int seg_tiles, segs_per_strip, n_segs;  // a segment = up to seg_tiles vertically consecutive tiles of one strip
  int om_pitch, om_hstride;               // tap-major layout: floats per pixel / between image rows
  int x_pitch, out_pitch, vec_ok, out_f32;
  uint32_t chunk_bytes, rowpitch;
  int ablate;                             // FAMI_DCN_ABLATE (timing experiments, results wrong): 1 no stores, 8 no far path,
                                          // 16 no offset loads
  const float* om;
  const void* x;
  const void* w;                          // packed [CoutPad][9 taps][64] 16-bit (fami_pack_conv_weight)
  const float* bias;
  void* out;
};

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

template <typename TH> __device__ __forceinline__ float4 ld4g(const TH* p) {   // 4 consecutive 16-bit values (8 B), global
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = h2_to_f2<TH>(u.x), b = h2_to_f2<TH>(u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}

// A sample whose 2x2 footprint leaves the staged window: bounds-checked global corners, fp32 blend (torchvision's rule: a
// corner outside the image contributes 0).  Out of line: the nine taps' hot code stays small.
template <typename TH>
__device__ __noinline__ uint2 dcn_far_sample(const TH* xb, int H, int W, int x_pitch, float py, float px, float mk) {
  uint2 pk = make_uint2(0u, 0u);
  if (py > -1.f && py < (float)H && px > -1.f && px < (float)W) {
    const int iy0 = (int)floorf(py), ix0 = (int)floorf(px);
    const float ly = py - (float)iy0, lx = px - (float)ix0, hy = 1.f - ly, hx = 1.f - lx;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool y0ok = iy0 >= 0, y1ok = iy0 + 1 <= H - 1, x0ok = ix0 >= 0, x1ok = ix0 + 1 <= W - 1;
    const float4 v1 = (y0ok && x0ok) ? ld4g<TH>(xb + ((int64_t)iy0 * W + ix0) * x_pitch) : z;
    const float4 v2 = (y0ok && x1ok) ? ld4g<TH>(xb + ((int64_t)iy0 * W + ix0 + 1) * x_pitch) : z;
    const float4 v3 = (y1ok && x0ok) ? ld4g<TH>(xb + ((int64_t)(iy0 + 1) * W + ix0) * x_pitch) : z;
    const float4 v4 = (y1ok && x1ok) ? ld4g<TH>(xb + ((int64_t)(iy0 + 1) * W + ix0 + 1) * x_pitch) : z;
    const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
    pk.x = f2_to_h2<TH>(mk * (w1 * v1.x + w2 * v2.x + w3 * v3.x + w4 * v4.x), mk * (w1 * v1.y + w2 * v2.y + w3 * v3.y + w4 * v4.y));
    pk.y = f2_to_h2<TH>(mk * (w1 * v1.z + w2 * v2.z + w3 * v3.z + w4 * v4.z), mk * (w1 * v1.w + w2 * v2.w + w3 * v3.w + w4 * v4.w));
  }
  return pk;
}

template <typename TH>
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1);
template <>
__device__ __forceinline__ void mma16816<__half>(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                                 uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma16816<__nv_bfloat16>(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                        uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// streamed once, whole lines per instruction: no L1 allocation
__device__ __forceinline__ float ldg_stream(const float* p) {
  float v;
  asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {   // non-blocking
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}

// (dy, dx, mask) of one tap for the 2*KS samples of a lane: j = h * KS + i (h: row gid / gid+8 of the m16 tile, i: iteration)
template <int NS> struct OmRegs { float dy[NS], dx[NS], mk[NS]; };

struct SegIt { int seg, k, nt; uint32_t pos0; };       // tile k of segment seg (nt tiles); pos0: ring position of its chunk 0
struct TileRef { const float* q; uint32_t vmask; int b, y0, x0, wr; };   // wr: the row pair (rows 2 wr, 2 wr + 1) of this warp in the tile

// window pixel pitch (bytes): dense for C = 48, 64 channel slots otherwise
__host__ __device__ constexpr uint32_t wp_pix_bytes(int KS) { return KS == 3 ? 96u : 128u; }

// segment -> (image, strip, first tile row, tiles); out of line: the integer divisions stay out of the tap loop's code
__device__ __noinline__ void wp_seg_decode(int seg, int segs_per_img, int segs_per_strip, int seg_tiles, int tiles_y, int& b, int& tx,
                                           int& ty0, int& nt) {
  b = seg / segs_per_img;
  const int r = seg - b * segs_per_img;
  tx = r / segs_per_strip;
  ty0 = (r - tx * segs_per_strip) * seg_tiles;
  nt = tiles_y - ty0 < seg_tiles ? tiles_y - ty0 : seg_tiles;
}

template <typename TH, int KS, int NT, bool BLK>
__global__ void __launch_bounds__(kWpThreads, 1)
dcn_wp_kernel(const __grid_constant__ CUtensorMap tmX, const DcnWpParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int NS = 2 * KS, G = 4 * KS;
  constexpr bool kRot = KS != 3;
  constexpr uint32_t kPixB = wp_pix_bytes(KS);
  constexpr uint32_t kWfBytes = 9u * KS * (NT / 2) * 32u * 16u;
  constexpr uint32_t kWinRows = kTH + 2 * kRy;                          // window rows of a tile
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t win_u32 = (raw_u32 + 127u) & ~127u;                  // ring: kRingRows x WW pixels x kPixB
  const uint32_t wf_u32 = win_u32 + kSlots * p.chunk_bytes;           // filter fragments [tap][kstep][n-tile pair][lane][16 B]
  const uint32_t bar0 = wf_u32 + kWfBytes;                            // full[kSlots], empty[kSlots]
  const uint32_t bias_u32 = bar0 + 16u * kSlots;
  auto full = [&](uint32_t s) { return bar0 + 8u * s; };
  auto empty = [&](uint32_t s) { return bar0 + 8u * kSlots + 8u * s; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < (uint32_t)kSlots; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), kWpWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
  }

  const int segs_per_img = p.tiles_x * p.segs_per_strip;
  auto seg_decode = [&](int seg, int& b, int& tx, int& ty0, int& nt) {
    wp_seg_decode(seg, segs_per_img, p.segs_per_strip, p.seg_tiles, p.tiles_y, b, tx, ty0, nt);
  };

  // ---- producer state (meaningful in lane 0 of warp 0): next chunk of the CTA's chunk stream ----
  int pr_seg = blockIdx.x, pr_c = 0, pr_nch = 0, pr_b = 0, pr_tx = 0, pr_row0 = 0;
  uint32_t pr_pos = 0;
  auto pr_load_seg = [&]() {
    if (pr_seg < p.n_segs) {
      int ty0, nt;
      seg_decode(pr_seg, pr_b, pr_tx, ty0, nt);
      pr_row0 = ty0 * kTH;
      pr_nch = kADV * nt + (kNCH - kADV);
    }
  };
  // issues the next chunk if its slot has been released (never blocks)
  auto pr_poll = [&]() {
    if (pr_seg >= p.n_segs) return;
    const uint32_t slot = pr_pos % kSlots, use = pr_pos / kSlots;
    if (!mbar_test(empty(slot), (use & 1u) ^ 1u)) return;
    mbar_arrive_expect_tx(full(slot), p.chunk_bytes);
    tma_tiled_4d(win_u32 + slot * p.chunk_bytes, &tmX, full(slot), 0, pr_tx * kTW - p.Rx, pr_row0 - kRy + pr_c * kChunk, pr_b);
    ++pr_pos;
    if (++pr_c == pr_nch) { pr_seg += gridDim.x; pr_c = 0; pr_load_seg(); }
  };
  const bool is_producer = threadIdx.x == 0;
  if (is_producer) {
    pr_load_seg();
    for (int s = 0; s < kSlots; ++s) pr_poll();
  }
  __syncwarp();
  // chunk q of the stream has landed (warp 0 keeps the producer going while it waits)
  auto wait_full = [&](uint32_t q) {
    const uint32_t bar = full(q % kSlots), par = (q / kSlots) & 1u;
    if (warp == 0) {
      for (;;) {
        uint32_t ok = 0;
        if (lane == 0) { ok = mbar_test(bar, par); if (!ok) pr_poll(); }
        ok = __shfl_sync(0xffffffffu, ok, 0);
        if (ok) break;
      }
    }
    mbar_wait(bar, par);
  };

  // ===================== compute warps =====================
  typedef typename H2<TH>::t h2;
  typedef OmRegs<NS> Om;
  const TH* xg = reinterpret_cast<const TH*>(p.x);
  const int gid = lane >> 2, t = lane & 3;
  const int rot = kRot ? gid % KS : 0;
  const uint32_t rowpitch = p.rowpitch;
  const uint32_t xlim = (uint32_t)(p.WW - 1);
  int gofs[KS];                               // 4 * k-step of iteration i of this lane ((i + rot) % KS; the group is gofs + t)
#pragma unroll
  for (int i = 0; i < KS; ++i) gofs[i] = kRot ? 4 * ((i + rot) % KS) : 4 * i;
  const int om_hs = BLK ? 3 * kTW * G : p.om_hstride;   // floats between the two rows of the warp
  constexpr int om_cs = BLK ? kTW * G : G;              // floats between dy, dx, mask
  constexpr int om_ts = BLK ? kLH * kTW * 3 * G : 3 * G;     // floats between taps (layout 3 is tile-major: [tile][tap])
  const float fd = (float)p.d;
  const float mxb = kMagic + (float)(gid + p.Rx - p.d);        // window column of the tap's grid point (+ kc * d)
  // window column address of iteration i = min(column, WW - 2) * pixel pitch + colk[i] (a far sample reads a harmless clamped
  // address and is fixed up afterwards; ring rows are always valid)
  uint32_t colk[KS];
#pragma unroll
  for (int i = 0; i < KS; ++i) colk[i] = win_u32 + (uint32_t)(gofs[i] + t) * 8u;
  const uint32_t ucl = (uint32_t)(p.WW - 2);

  auto it_tile = [&](const SegIt& it) {
    TileRef tr;
    tr.q = p.om; tr.vmask = 0; tr.b = 0; tr.y0 = 0; tr.x0 = 0; tr.wr = warp;
    if (it.seg >= p.n_segs) return tr;
    int b, tx, ty0, nt;
    seg_decode(it.seg, b, tx, ty0, nt);
    const int ty = ty0 + it.k;
    tr.b = b; tr.y0 = ty * kTH; tr.x0 = tx * kTW;
    // the row pairs rotate over the warps from tile to tile: the border rows of a tile meet the window edge (far samples) far
    // more often than the inner ones, and a warp that is always slower holds up the ring for the other eleven
    tr.wr = (warp + it.k + it.seg) % kWpWarps;
    const int y = tr.y0 + 2 * tr.wr, x = tr.x0 + gid;
    if (x < p.W) tr.vmask = (y < p.H ? 1u : 0u) | (y + 1 < p.H ? 2u : 0u);
    if (BLK)
      tr.q = p.om + (((int64_t)b * p.ltiles_y + (y >> 4)) * p.tiles_x + tx) * (9 * kLH * kTW * 3 * G) + (y & (kLH - 1)) * (3 * kTW * G) + gid * 4 + t;
    else
      tr.q = p.om + (((int64_t)b * p.H + (y < p.H ? y : 0)) * p.W + (x < p.W ? x : 0)) * p.om_pitch + t;
    if (!tr.vmask) tr.q = p.om;
    return tr;
  };
  auto it_next = [&](const SegIt& it) {
    SegIt n = it;
    if (it.seg >= p.n_segs) return n;
    if (it.k + 1 < it.nt) { n.k = it.k + 1; return n; }
    n.seg = it.seg + gridDim.x; n.k = 0; n.pos0 = it.pos0 + kADV * it.nt + (kNCH - kADV);
    n.nt = 0;
    if (n.seg < p.n_segs) { int b, tx, ty0; seg_decode(n.seg, b, tx, ty0, n.nt); }
    return n;
  };
  auto load_tap = [&](const TileRef& tr, int tap, Om& o) {
    if (p.ablate & 16) {
#pragma unroll
      for (int j = 0; j < NS; ++j) { o.dy[j] = 0.25f; o.dx[j] = 0.25f; o.mk[j] = 1.f; }
      return;
    }
    const float* q0 = tr.q + tap * om_ts;
    const float* q1 = q0 + om_hs;
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      const int h = j / KS, i = j - h * KS;
      const bool v = (tr.vmask >> h) & 1u;
      // k-step-blocked runs [group / 4][pixel 8][group % 4]: the 32 lanes of a load read one contiguous 128-byte line
      const float* qj = (h ? q1 : q0) + (BLK ? 8 * gofs[i] : gofs[i]);
      if (BLK) {
        o.dy[j] = v ? ldg_stream(qj) : 0.f;
        o.dx[j] = v ? ldg_stream(qj + om_cs) : 0.f;
        o.mk[j] = v ? ldg_stream(qj + 2 * om_cs) : 0.f;
      } else {
        o.dy[j] = v ? __ldg(qj) : 0.f;
        o.dx[j] = v ? __ldg(qj + om_cs) : 0.f;
        o.mk[j] = v ? __ldg(qj + 2 * om_cs) : 0.f;
      }
    }
  };
  // the tap three positions ahead in the (tile, tap) stream, into the registers of the tap just gathered
  auto load_ahead = [&](const TileRef& cur, const TileRef& nxt, int tap, Om& o) {
    if (tap + 3 < 9) load_tap(cur, tap + 3, o);
    else load_tap(nxt, tap + 3 - 9, o);
  };

  float acc[NT][4];
  SegIt cur, nxt;

  // one tap: 2*KS samples of this lane -> A fragments -> KS x NT MMAs
  auto do_tap = [&](const TileRef& tr, const TileRef& trn, Om& o, int kr, int kc, float my0, float mx, uint32_t base_row) {
    if (kc == 0 && is_producer) pr_poll();      // once per kernel row: three polls per tile, three chunks to issue per tile
    const float my1 = my0 + 1.f;
    uint2 rs[NS];
    bool far = false;
    {
      uint32_t a00[NS], a10[NS];
      h2 w12[NS], w34[NS];
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        const int h = j / KS, i = j - h * KS;
        const float dy = o.dy[j], dx = o.dx[j], mk = o.mk[j];
        const float my = h ? my1 : my0;
        // floor() through the magic-number add (round-down): exact for |v| < 2^22, the integer lands in the low mantissa bits
        const float ty = __fadd_rd(dy, my), tx = __fadd_rd(dx, mx);
        const float ly = dy - (ty - my), lx = dx - (tx - mx);
        const uint32_t v = __float_as_uint(ty) - kMagicBits, u = __float_as_uint(tx) - kMagicBits;   // window row / column
        const bool in = v < kWinRows - 1u && u < xlim;
        const float wb = mk * ly, wt = mk - wb;          // mask * (ly | 1 - ly)
        const float w4f = wb * lx, w3f = wb - w4f, w2f = wt * lx, w1f = wt - w2f;
        w12[j] = H2<TH>::pack(w1f, w2f);
        w34[j] = H2<TH>::pack(w3f, w4f);
        uint32_t r0, r1;
        if (kRingPow2) {     // (the magic constant's low bits are 0: the float's bit pattern can be masked directly)
          r0 = (__float_as_uint(ty) + base_row) & (uint32_t)(kRingRows - 1);
          r1 = (__float_as_uint(ty) + base_row + 1u) & (uint32_t)(kRingRows - 1);
        } else {
          const uint32_t vc = v < kWinRows - 1u ? v : 0u;
          r0 = vc + base_row; r1 = r0 + 1u;
          r0 = r0 >= (uint32_t)kRingRows ? r0 - (uint32_t)kRingRows : r0;
          r1 = r1 >= (uint32_t)kRingRows ? r1 - (uint32_t)kRingRows : r1;
        }
        const uint32_t col = min(u, ucl) * kPixB + colk[i];
        a00[j] = col + r0 * rowpitch;
        a10[j] = col + r1 * rowpitch;
        far |= !in;      // (pixels outside the image carry zero offsets: always inside)
      }
      uint2 u1[NS], u2[NS], u3[NS], u4[NS];
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        u1[j] = lds64(a00[j]);
        u2[j] = lds64(a00[j] + kPixB);
        u3[j] = lds64(a10[j]);
        u4[j] = lds64(a10[j] + kPixB);
      }
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        const h2 w1 = H2<TH>::lo(w12[j]), w2 = H2<TH>::hi(w12[j]), w3 = H2<TH>::lo(w34[j]), w4 = H2<TH>::hi(w34[j]);
        h2 lo = H2<TH>::mul(w1, *reinterpret_cast<const h2*>(&u1[j].x));
        h2 hi = H2<TH>::mul(w1, *reinterpret_cast<const h2*>(&u1[j].y));
        lo = H2<TH>::fma(w2, *reinterpret_cast<const h2*>(&u2[j].x), lo);
        hi = H2<TH>::fma(w2, *reinterpret_cast<const h2*>(&u2[j].y), hi);
        lo = H2<TH>::fma(w3, *reinterpret_cast<const h2*>(&u3[j].x), lo);
        hi = H2<TH>::fma(w3, *reinterpret_cast<const h2*>(&u3[j].y), hi);
        lo = H2<TH>::fma(w4, *reinterpret_cast<const h2*>(&u4[j].x), lo);
        hi = H2<TH>::fma(w4, *reinterpret_cast<const h2*>(&u4[j].y), hi);
        rs[j].x = *reinterpret_cast<const uint32_t*>(&lo);
        rs[j].y = *reinterpret_cast<const uint32_t*>(&hi);
      }
    }
    // large offsets: entered by the whole warp so that the dependent global loads of all far samples of the tap overlap
    if (__any_sync(0xffffffffu, far) && !(p.ablate & 8)) {
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        const int h = j / KS, i = j - h * KS;
        const float my = h ? my1 : my0;
        const uint32_t v = __float_as_uint(__fadd_rd(o.dy[j], my)) - kMagicBits, u = __float_as_uint(__fadd_rd(o.dx[j], mx)) - kMagicBits;
        if (!(v < kWinRows - 1u && u < xlim)) {
          const int y = tr.y0 + 2 * tr.wr + h, x = tr.x0 + gid;
          rs[j] = dcn_far_sample<TH>(xg + (int64_t)tr.b * p.H * p.W * p.x_pitch + (gofs[i] + t) * 4, p.H, p.W, p.x_pitch,
                                     (float)(y - p.d + kr * p.d) + o.dy[j], (float)(x - p.d + kc * p.d) + o.dx[j], o.mk[j]);
        }
      }
      __syncwarp();
    }
    // (dy, dx, mask) of the tap THREE ahead, into the registers this tap has just finished with (the far path above was their
    // last reader): three buffers give 2.4 taps of lead instead of the 1.5 of a load into the previous tap's buffer; issued
    // behind the gather, so that a wait for this tap's registers never covers the fresh loads' scoreboard
    load_ahead(tr, trn, kr * 3 + kc, o);
    const uint32_t wtap = wf_u32 + (uint32_t)((kr * 3 + kc) * KS) * (uint32_t)(NT / 2) * 512u + (uint32_t)lane * 16u;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint2 f0 = rs[ks], f1 = rs[KS + ks];
      if (kRot) {      // iteration i of this lane was k-step (i + rot) % KS: put the samples back in k-step order
#pragma unroll
        for (int r = 1; r < KS; ++r) {
          const int src = (ks - r + KS) % KS;
          if (rot == r) { f0 = rs[src]; f1 = rs[KS + src]; }
        }
      }
#pragma unroll
      for (int ntp = 0; ntp < NT / 2; ++ntp) {
        const uint4 bw = lds128(wtap + (uint32_t)(ks * (NT / 2) + ntp) * 512u);
        mma16816<TH>(acc[2 * ntp], f0.x, f1.x, f0.y, f1.y, bw.x, bw.y);
        mma16816<TH>(acc[2 * ntp + 1], f0.x, f1.x, f0.y, f1.y, bw.z, bw.w);
      }
    }
  };

  cur.seg = blockIdx.x; cur.k = 0; cur.pos0 = 0; cur.nt = 0;
  if (cur.seg < p.n_segs) { int b, tx, ty0; seg_decode(cur.seg, b, tx, ty0, cur.nt); }
  nxt = it_next(cur);
  TileRef tc = it_tile(cur), tn = it_tile(nxt);
  Om o0, o1, o2;                       // taps 3i, 3i+1, 3i+2 of the stream: rotating, three taps in flight
  load_tap(tc, 0, o0);
  load_tap(tc, 1, o1);
  load_tap(tc, 2, o2);
  // (the first window chunks and the first two taps' offsets are already in flight: the repack below hides behind them)
  {
    // filter -> fragment order: lane (gid, t) of (tap, kstep, n-tile) reads W[cout = 8 nt + gid][tap][cin = 16 ks + 4 t .. +3]:
    // b0 = k-slots 2t, 2t+1, b1 = k-slots 2t+8, 2t+9 of the k order the gather produces
    const TH* wg = reinterpret_cast<const TH*>(p.w);
    for (int idx = threadIdx.x; idx < 9 * KS * (NT / 2) * 64; idx += kWpThreads) {
      const int q = idx & 1, ln = (idx >> 1) & 31, rest = idx >> 6;
      const int ntp = rest % (NT / 2), tk = rest / (NT / 2);
      const int ks = tk % KS, tap = tk / KS;
      const int cout = 8 * (2 * ntp + q) + (ln >> 2);
      const uint2 v = __ldg(reinterpret_cast<const uint2*>(wg + ((size_t)cout * 9 + tap) * 64 + 16 * ks + 4 * (ln & 3)));
      sts64(wf_u32 + (uint32_t)idx * 8u, v);
    }
    for (int c = threadIdx.x; c < NT * 8; c += kWpThreads) {
      const float b = (p.bias && c < p.Cout) ? p.bias[c] : 0.f;
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_u32 + 4u * c), "f"(b) : "memory");
    }
  }
  __syncthreads();
  while (cur.seg < p.n_segs) {
    const uint32_t P = cur.pos0 + (uint32_t)(kADV * cur.k);       // first chunk of this tile in the stream
    for (uint32_t c = cur.k == 0 ? 0u : (uint32_t)(kNCH - kADV); c < (uint32_t)kNCH; ++c) wait_full(P + c);
    const uint32_t base_row = (P * kChunk) % (uint32_t)kRingRows;
#pragma unroll
    for (int n = 0; n < NT; ++n) { acc[n][0] = 0.f; acc[n][1] = 0.f; acc[n][2] = 0.f; acc[n][3] = 0.f; }
    float my = kMagic + (float)(2 * tc.wr + kRy - p.d);   // window row of the tap's grid point (row h = 0; + kr * d)
#pragma unroll 1
    for (int kr = 0; kr < 3; ++kr, my += fd) {
      do_tap(tc, tn, o0, kr, 0, my, mxb, base_row);
      do_tap(tc, tn, o1, kr, 1, my, mxb + fd, base_row);
      do_tap(tc, tn, o2, kr, 2, my, mxb + 2.f * fd, base_row);
    }
    __syncwarp();
    if (lane == 0) {             // this warp no longer reads the tile's first kADV chunks (all of them at the end of a segment)
      const uint32_t nrel = cur.k + 1 == cur.nt ? (uint32_t)kNCH : (uint32_t)kADV;
      for (uint32_t c = 0; c < nrel; ++c) mbar_arrive(empty((P + c) % kSlots));
    }
    // ---- epilogue: acc[nt] = {(row gid, couts 8nt+2t, +1), (row gid+8, same couts)} ----
    if (!(p.ablate & 1)) {
      const int x = tc.x0 + gid;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int y = tc.y0 + 2 * tc.wr + h;
        const bool valid = y < p.H && x < p.W;                    // (shuffles below are executed by every lane)
        const int64_t pix = valid ? ((int64_t)tc.b * p.H + y) * p.W + x : 0;
        if (p.out_f32) {
          float* dst = reinterpret_cast<float*>(p.out) + pix * p.out_pitch;
#pragma unroll
          for (int n = 0; n < NT; ++n) {
            const int c = 8 * n + 2 * t;
            float b0, b1;
            asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(b0), "=f"(b1) : "r"(bias_u32 + 4u * c));
            const float v0 = acc[n][2 * h] + b0, v1 = acc[n][2 * h + 1] + b1;
            if (valid) {
              if (p.vec_ok && c + 1 < p.Cout) *reinterpret_cast<float2*>(dst + c) = make_float2(v0, v1);
              else { if (c < p.Cout) dst[c] = v0; if (c + 1 < p.Cout) dst[c + 1] = v1; }
            }
          }
        } else {
          TH* dst = reinterpret_cast<TH*>(p.out) + pix * p.out_pitch;
#pragma unroll
          for (int np = 0; np < NT / 2; ++np) {
            float b0, b1, b2, b3;
            asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(b0), "=f"(b1) : "r"(bias_u32 + 4u * (16 * np + 2 * t)));
            asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(b2), "=f"(b3) : "r"(bias_u32 + 4u * (16 * np + 8 + 2 * t)));
            const uint32_t lo = f2_to_h2<TH>(acc[2 * np][2 * h] + b0, acc[2 * np][2 * h + 1] + b1);          // couts 16np + 2t, +1
            const uint32_t hi = f2_to_h2<TH>(acc[2 * np + 1][2 * h] + b2, acc[2 * np + 1][2 * h + 1] + b3);  // couts 16np + 8 + 2t, +1
            // even t keeps the low n-tile and takes its right neighbour's, odd t keeps the high one and takes its left
            // neighbour's: every lane then stores 8 contiguous bytes, a quad a whole 32-byte sector
            const uint32_t got = __shfl_xor_sync(0xffffffffu, (t & 1) ? lo : hi, 1);
            const int c = 16 * np + ((t & 1) ? 8 + 2 * (t - 1) : 2 * t);
            const uint2 val = (t & 1) ? make_uint2(got, hi) : make_uint2(lo, got);
            if (valid) {
              if (p.vec_ok && c + 3 < p.Cout) *reinterpret_cast<uint2*>(dst + c) = val;
              else {
                const TH* hv = reinterpret_cast<const TH*>(&val);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  if (c + e < p.Cout) dst[c + e] = hv[e];
              }
            }
          }
        }
      }
    }
    cur = nxt; tc = tn;
    nxt = it_next(cur);
    tn = it_tile(nxt);
  }
  // the producer has nothing left to issue here: every chunk of the stream was waited for by warp 0 itself
}

int wp_env(const char* name, int dflt) {
  const char* s = getenv(name);
  return s ? atoi(s) : dflt;
}

size_t dcn_wp_smem(int KS, int NT, int Rx) {
  const size_t ring = (size_t)kRingRows * (kTW + 2 * Rx) * wp_pix_bytes(KS);
  return 128 + ring + (size_t)9 * KS * (NT / 2) * 512 + 16 * kSlots + (size_t)NT * 8 * 4;
}
int dcn_wp_rx(const fami_dcn_desc* d) {
  const int KS = d->C / 16, NT = d->Cout / 8;
  for (int Rx = d->dil + 7; Rx >= d->dil + 4; --Rx)
    if (dcn_wp_smem(KS, NT, Rx) <= 227 * 1024) return Rx;
  return -1;
}

}  // namespace

int dcn_wp_supported(const fami_dcn_desc* d) {
  static const int on = wp_env("FAMI_DCN_WP", 1);
  if (!on) return 0;
  if (!is_half_dtype(d->dtype) || (d->om_layout != 1 && d->om_layout != 3)) return 0;
  if (d->C != d->Cout || (d->C != 32 && d->C != 48) || d->G * 4 != d->C) return 0;
  if (d->x_pitch % 8 != 0) return 0;
  if (d->kh != 3 || d->kw != 3 || d->pad != d->dil || d->dil < 1 || d->dil > 4) return 0;
  return dcn_wp_rx(d) > 0;
}

int dcn_wp_launch(const fami_dcn_desc* d, const void* x, const float* om, const void* w, const float* bias, void* out,
                  cudaStream_t st) {
  FAMI_CHECK_ARG(load_driver_fns(), "cuTensorMapEncode* driver entry points unavailable");
  FAMI_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0,
                 "dcn_wp: x / w must be 16-byte aligned");
  FAMI_CHECK_ARG((reinterpret_cast<uintptr_t>(om) & 15) == 0, "dcn_wp: offsets|masks must be 16-byte aligned");
  const int KS = d->C / 16, NT = d->Cout / 8;
  const uint32_t pixb = wp_pix_bytes(KS);
  DcnWpParams p;
  memset(&p, 0, sizeof(p));
  p.B = d->B; p.H = d->H; p.W = d->W; p.C = d->C; p.Cout = d->Cout; p.G = d->G; p.d = d->dil;
  p.Rx = dcn_wp_rx(d);
  FAMI_CHECK_ARG(p.Rx > 0, "dcn_wp: filter and window do not fit in shared memory");
  p.WW = kTW + 2 * p.Rx;
  p.tiles_x = (d->W + kTW - 1) / kTW; p.tiles_y = (d->H + kTH - 1) / kTH; p.ltiles_y = (d->H + kLH - 1) / kLH;
  const int sms = num_sms();
  {
    // segment length: the longest strips that still balance over the SMs (a segment of s tiles loads kNCH - kADV extra chunks)
    int best = 1;
    double best_cost = 1e30;
    for (int s = 1; s <= p.tiles_y; ++s) {
      const int64_t units = (int64_t)d->B * p.tiles_x * ((p.tiles_y + s - 1) / s);
      const double cost = (double)((units + sms - 1) / sms) * (s + 0.35);
      if (cost < best_cost - 1e-9 || (cost < best_cost + 1e-9 && s > best)) { best_cost = cost; best = s; }
    }
    p.seg_tiles = wp_env("FAMI_DCN_WP_SEG", best);
    if (p.seg_tiles < 1) p.seg_tiles = 1;
    if (p.seg_tiles > p.tiles_y) p.seg_tiles = p.tiles_y;
  }
  p.segs_per_strip = (p.tiles_y + p.seg_tiles - 1) / p.seg_tiles;
  p.n_segs = d->B * p.tiles_x * p.segs_per_strip;
  const bool blocked = d->om_layout == 3;
  p.om_pitch = d->off_pitch;
  p.om_hstride = d->W * d->off_pitch;
  p.x_pitch = d->x_pitch; p.out_pitch = d->out_pitch;
  p.out_f32 = d->out_f32 ? 1 : 0;
  p.vec_ok = ((reinterpret_cast<uintptr_t>(out) & 7) == 0) && (d->out_pitch % (d->out_f32 ? 2 : 4) == 0);
  p.rowpitch = (uint32_t)p.WW * pixb;
  p.chunk_bytes = (uint32_t)kChunk * p.rowpitch;
  p.ablate = wp_env("FAMI_DCN_ABLATE", 0);
  p.om = om; p.x = x; p.w = w; p.bias = bias; p.out = out;

  const CUtensorMapDataType tm_dtype = d->dtype == FAMI_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap tmX;
  {
    // box = kChunk rows x WW pixels x (pixel pitch / 2) channel slots; slots >= C are never consumed (zero-filled when the
    // tensor's channel extent ends there, the next pixel's first channels when its pitch is wider)
    const cuuint32_t slots = pixb / 2;
    cuuint64_t dims[4] = {(cuuint64_t)((cuuint32_t)d->x_pitch >= slots ? slots : d->C), (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)d->x_pitch * 2, (cuuint64_t)d->W * d->x_pitch * 2,
                             (cuuint64_t)d->H * d->W * d->x_pitch * 2};
    cuuint32_t box[4] = {slots, (cuuint32_t)p.WW, (cuuint32_t)kChunk, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode_tiled(&tmX, tm_dtype, 4, const_cast<void*>(x), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FAMI_CHECK_ARG(r == CUDA_SUCCESS, "dcn_wp: cuTensorMapEncodeTiled(x) failed (%d)", (int)r);
  }
  const size_t smem = dcn_wp_smem(KS, NT, p.Rx);
  int grid = p.n_segs < sms ? p.n_segs : sms;
#define FAMI_WP_LAUNCH(TH_, KS_, NT_, BLK_)                                                                    \
  {                                                                                                            \
    static std::atomic<uint64_t> attr_mask{0};                                                                 \
    set_max_smem_once(attr_mask, dcn_wp_kernel<TH_, KS_, NT_, BLK_>, 227 * 1024);                              \
    dcn_wp_kernel<TH_, KS_, NT_, BLK_><<<grid, kWpThreads, smem, st>>>(tmX, p);                                \
  }
#define FAMI_WP_LAUNCH_C(TH_, BLK_)                                             \
  if (KS == 2) FAMI_WP_LAUNCH(TH_, 2, 4, BLK_) else FAMI_WP_LAUNCH(TH_, 3, 6, BLK_)
  if (d->dtype == FAMI_F16) {
    if (blocked) { FAMI_WP_LAUNCH_C(__half, true) } else { FAMI_WP_LAUNCH_C(__half, false) }
  } else {
    if (blocked) { FAMI_WP_LAUNCH_C(__nv_bfloat16, true) } else { FAMI_WP_LAUNCH_C(__nv_bfloat16, false) }
  }
#undef FAMI_WP_LAUNCH_C
#undef FAMI_WP_LAUNCH
  FAMI_CHECK_LAUNCH("dcn_wp_kernel");
  return 0;
}

}  // namespace fami
