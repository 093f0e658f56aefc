// 3x3 stride-1 "same" convolution (pad == dil) on tcgen05 with the input tile staged ONCE in shared
// memory (halo form).  This is the kernel for the HRNet BasicBlock convs (84 % of the model's FLOPs)
// and the dilated offset/mask convs of the alignment head.
//
// Why: the TMA-im2col form (conv_tc.cu) fetches every activation byte 9x from L2 and re-streams the
// weights for every 128-pixel tile; measured, it saturates the L2->SM fabric (~25 B/clk/SM) at 12-32 %
// tensor utilisation.  Here one TMA *tiled* box load brings (BH+2d) x (W+2d) pixels x 64 channels
// (zero-filled outside the image / beyond Cin) into a 128B-swizzled tile whose rows are the
// flattened, width-padded pixel positions.  With that layout filter tap (r,s) of output position q is
// simply row  q + (r*Wp + s)*d : the nine taps are nine UMMA A-descriptors with shifted start
// addresses into the SAME tile (the hardware applies the 128B swizzle on absolute address bits, so
// any 128-byte row shift is legal -- established by tools/probe_umma.py).  Outputs are computed for
// all Wp = W+2d columns of BH rows (NM = ceil(BH*Wp/128) UMMA M-tiles sharing every B tile); the 2d
// padded columns per row are discarded by the epilogue.  Weights are either resident in shared
// memory for the whole kernel (narrow convs) or streamed once per CTA tile and shared by the NM
// M-tiles.
//
// Warp roles (512 threads, persistent): warp 0 TMA producer; warps 1-3 MMA issuers, issuer j owning the
// M-tiles m = j (mod n_iss) -- the per-MMA issue cost of a single thread (~150 clk measured) exceeds the
// tensor-pipe time of an N=48 MMA (~44 clk), so the M-tiles of a CTA tile are issued from separate warps
// into separate accumulators; warps 4-15 epilogue (3 column groups x 4 TMEM lane quarters, coalesced
// through shared-memory staging).
#include <stdlib.h>

#include "tc_common.cuh"

namespace fami {

// optional per-role timeline of CTA 0 (p.trace != 0; read back with fami_debug_read_trace)
__device__ unsigned long long g_trace[8192];   // [4 slots][64 tiles][8 events]

namespace {

__device__ __forceinline__ void trace(int on, int slot, int it, int ev) {
  if (on && blockIdx.x == 0 && it < 64) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_trace[(slot * 64 + it) * 8 + ev] = t;
  }
}

constexpr int kMmaWarps = 3;                       // MMA issuer warps (one per M-tile, see below)
constexpr int kEpiWarp0 = 1 + kMmaWarps;           // first epilogue warp
constexpr int kHThreads = 32 * (1 + kMmaWarps + kEpiWarps);
constexpr int kEpiThreads = 32 * kEpiWarps;

struct HaloParams {
  int N, H, W, Wp, d;        // image count/size, padded width W+2d, dilation (= padding)
  int BH, NM, HR;            // rows per CTA tile, M-tiles per CTA tile, smem rows per A stage
  int tiles_per_img, n_tiles, total_tiles;
  int cchunks, last_kk;
  int Cout, BN;
  int relu, out_f32, vec_ok;
  int out_pitch, res_pitch;
  int sA, sB, b_resident, acc_bufs, n_iss, acc_stride;
  uint32_t a_stage_bytes, b_tile_bytes, a_box_bytes;
  int trace;
  int om_groups, om_tiles_x, om_tiles_y;   // > 0: y is the row-blocked DCN offset|mask buffer
  int64_t om_tap_stride;
  int om_kblocked;                         // fami_conv_desc.om_layout == 3
  const float* scale;
  const float* shift;
  const void* res;
  void* y;
  const float* res32;   // fp32 residual stream mode (16-bit arms), may be null
  float* y32;
  int res32_pitch, y32_pitch;
};

// The nine taps of one resident-weights channel chunk, issued by one warp for its M-tiles m = issuer, issuer + n_iss, ...
// NK = 16-channel MMA steps per tap.  Everything but the descriptor increments is hoisted: at N <= 96 the issuing
// thread's instruction stream, not the tensor pipe, sets the MMA rate.
// one (A rows, B tile) pair of NK K-steps, single CTA or CTA pair (cta_group::2)
template <int NK, bool kTf32, bool kPair>
__device__ __forceinline__ void halo_mma(bool leader, uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, bool acc) {
  if constexpr (kPair) umma_pair_ksteps<NK, kTf32>(leader, d, a_lo, b_lo, idesc, acc);
  else umma_ksteps<NK, kTf32>(leader, d, a_lo, b_lo, idesc, acc);
}
template <bool kTf32, bool kPair>
__device__ __forceinline__ void halo_mma_n(int nk, bool leader, uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, bool acc) {
  if constexpr (kPair) umma_pair_ksteps_n<kTf32>(nk, leader, d, a_lo, b_lo, idesc, acc);
  else umma_ksteps_n<kTf32>(nk, leader, d, a_lo, b_lo, idesc, acc);
}
template <bool kPair> __device__ __forceinline__ void halo_commit(uint32_t bar) {
  if constexpr (kPair) umma_commit_pair(bar);
  else umma_commit(bar);
}

template <int NK, bool kTf32, bool kPair>
__device__ __forceinline__ void halo_issue_taps(bool leader, uint32_t a_base, uint32_t b_lo, uint32_t b_step, uint32_t wp8,
                                                uint32_t d8, uint32_t a_inc, uint32_t d_inc, uint32_t d0, uint32_t idesc,
                                                int issuer, int NM, int n_iss) {
  uint32_t a_row = a_base;
#pragma unroll 1
  for (int fr = 0; fr < 3; ++fr, a_row += wp8) {
    uint32_t a_tap = a_row;
#pragma unroll
    for (int fs = 0; fs < 3; ++fs, a_tap += d8, b_lo += b_step) {
      uint32_t a_lo = a_tap, dcol = d0;
      for (int m = issuer; m < NM; m += n_iss, a_lo += a_inc, dcol += d_inc)
        halo_mma<NK, kTf32, kPair>(leader, dcol, a_lo, b_lo, idesc, (fr | fs) != 0);
    }
  }
}

// kPair: CTA pairs (cluster of 2, tcgen05 cta_group::2): the two CTAs take two consecutive CTA tiles; every MMA is M = 256
// (M-tile m of BOTH tiles), issued by the leader CTA only -- half the MMA instructions per pixel for the issue-bound narrow
// classes -- and each CTA keeps only HALF of the rows of every weight tile, so resident weights cost half the shared memory
// (what lets the tf32 48->48 conv keep its weights resident AND double-buffer its A tile).
template <typename TH, bool kPair>
__global__ void __launch_bounds__(kHThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int kKC = TcTraits<TH>::kKC;       // channels per 128-byte row: 64 (16-bit) / 32 (tf32)
  constexpr bool kTf32 = TcTraits<TH>::kTf32;
  const uint32_t cta_rank = kPair ? cluster_ctarank() : 0u;
  const bool is_leader_cta = cta_rank == 0u;
  const int tile0 = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;          // work units: CTA tiles, or pairs of them
  const int tile_step = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int n_units = kPair ? (p.total_tiles + 1) >> 1 : p.total_tiles;
  auto unit_tile = [&](int unit, bool& dup) -> int {
    if (!kPair) { dup = false; return unit; }
    const int t = 2 * unit + (int)cta_rank;
    dup = t >= p.total_tiles;
    return dup ? p.total_tiles - 1 : t;
  };
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int ksteps = 9 * p.cchunks;
  const int nB = p.b_resident ? ksteps : p.sB;
  uint8_t* smemA = smem;
  uint8_t* smemB = smem + (size_t)p.sA * p.a_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smemB + (size_t)nB * p.b_tile_bytes);
  const uint32_t bar0 = smem_u32(bars);
  // barriers: fullA[sA] emptyA[sA] fullB[nBbar] emptyB[nBbar] tfull[2] tempty[2]
  const int nBbar = p.b_resident ? 1 : p.sB;
  auto fullA = [&](int s) { return bar0 + 8u * s; };
  auto emptyA = [&](int s) { return bar0 + 8u * (p.sA + s); };
  auto fullB = [&](int s) { return bar0 + 8u * (2 * p.sA + s); };
  auto emptyB = [&](int s) { return bar0 + 8u * (2 * p.sA + nBbar + s); };
  auto tfull = [&](int a) { return bar0 + 8u * (2 * p.sA + 2 * nBbar + a); };
  auto tempty = [&](int a) { return bar0 + 8u * (2 * p.sA + 2 * nBbar + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.sA + 2 * nBbar + 4);
  float* s_scale = reinterpret_cast<float*>(bars + 2 * p.sA + 2 * nBbar + 6);
  float* s_shift = s_scale + p.n_tiles * p.BN;
  uint8_t* stage_base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(s_shift + p.n_tiles * p.BN) + 15) & ~(uintptr_t)15);
  fill_scale_shift(s_scale, s_shift, p.scale, p.shift, p.Cout, p.n_tiles * p.BN);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.sA; ++s) { mbar_init(fullA(s), 1); mbar_init(emptyA(s), p.n_iss); }
    for (int s = 0; s < nBbar; ++s) { mbar_init(fullB(s), 1); mbar_init(emptyB(s), p.n_iss); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), p.n_iss); mbar_init(tempty(a), kPair ? 2 * kEpiWarps : kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (kPair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if constexpr (kPair) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int acc_cols = p.NM * p.acc_stride;     // TMEM columns of one accumulator set

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform control flow, one elected lane issues) =====
    {
      const bool leader = elect_one();
      if (leader) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
      }
      // pair mode: b_tile_bytes is this CTA's HALF of a weight tile (rows [rank * BN/2, +BN/2)); every load of either CTA
      // completes on the LEADER's barrier, which the leader arms for the bytes of both
      const int b_row0 = kPair ? (int)cta_rank * (p.BN >> 1) : 0;
      if (p.b_resident) {
        // all weights of this conv (single N tile) stay in shared memory for the whole kernel
        if (leader && is_leader_cta) mbar_arrive_expect_tx(fullB(0), (kPair ? 2u : 1u) * (uint32_t)ksteps * p.b_tile_bytes);
        for (int ks = 0; ks < ksteps; ++ks)
          if (leader) {
            if constexpr (kPair) tma_tiled_2d_2sm(smem_u32(smemB + (size_t)ks * p.b_tile_bytes), &tmB, fullB(0), ks * kKC, b_row0);
            else tma_tiled_2d(smem_u32(smemB + (size_t)ks * p.b_tile_bytes), &tmB, fullB(0), ks * kKC, 0);
          }
      }
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      for (int unit = tile0; unit < n_units; unit += tile_step) {
        bool dup;
        const int tile = unit_tile(unit, dup);
        const int nt = tile % p.n_tiles;
        const int t2 = tile / p.n_tiles;
        const int ty = t2 % p.tiles_per_img, img = t2 / p.tiles_per_img;
        const int y0 = ty * p.BH;
        for (int cc = 0; cc < p.cchunks; ++cc) {
          mbar_wait(emptyA(sa), pa ^ 1u);
          if (leader) {
            trace(p.trace, 0, (unit - tile0) / tile_step, 0);
            if constexpr (kPair) {
              if (is_leader_cta) mbar_arrive_expect_tx(fullA(sa), 2u * p.a_box_bytes);
              tma_tiled_4d_2sm(smem_u32(smemA + (size_t)sa * p.a_stage_bytes), &tmA, fullA(sa), cc * kKC, -p.d, y0 - p.d, img);
            } else {
              mbar_arrive_expect_tx(fullA(sa), p.a_box_bytes);
              tma_tiled_4d(smem_u32(smemA + (size_t)sa * p.a_stage_bytes), &tmA, fullA(sa), cc * kKC, -p.d, y0 - p.d, img);
            }
          }
          if (++sa == p.sA) { sa = 0; pa ^= 1u; }
          if (!p.b_resident) {
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(emptyB(sb), pb ^ 1u);
              if (leader) {
                if constexpr (kPair) {
                  if (is_leader_cta) mbar_arrive_expect_tx(fullB(sb), 2u * p.b_tile_bytes);
                  tma_tiled_2d_2sm(smem_u32(smemB + (size_t)sb * p.b_tile_bytes), &tmB, fullB(sb), (tap * p.cchunks + cc) * kKC,
                                   nt * p.BN + b_row0);
                } else {
                  mbar_arrive_expect_tx(fullB(sb), p.b_tile_bytes);
                  tma_tiled_2d(smem_u32(smemB + (size_t)sb * p.b_tile_bytes), &tmB, fullB(sb), (tap * p.cchunks + cc) * kKC,
                               nt * p.BN);
                }
              }
              if (++sb == p.sB) { sb = 0; pb ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp <= kMmaWarps) {
   if (warp - 1 < p.n_iss && is_leader_cta) {
    const int issuer = warp - 1;
    // ===================== MMA issuers (warp-uniform control flow, one elected lane issues) ======
    // Operands of tcgen05.mma live in uniform registers: keeping the whole warp converged lets the
    // compiler keep descriptors there; issuing from divergent code (if lane == 0) costs 2-3x per MMA.
    {
      const bool leader = elect_one();
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t idesc = kPair ? umma_idesc_pair<TH>(p.BN) : umma_idesc<TH>(p.BN);
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int it = 0;
      if (p.b_resident) { mbar_wait(fullB(0), 0); tc_fence_after(); }
      for (int unit = tile0; unit < n_units; unit += tile_step, ++it) {
        const int acc = (p.acc_bufs == 2) ? (it & 1) : 0;
        const uint32_t acc_phase = (p.acc_bufs == 2) ? ((uint32_t)(it >> 1) & 1u) : ((uint32_t)it & 1u);
        if (leader && issuer == 0) trace(p.trace, 1, it, 3);
        mbar_wait(tempty(acc), acc_phase ^ 1u);
        tc_fence_after();
        if (leader && issuer == 0) trace(p.trace, 1, it, 0);
        const uint32_t d_tmem = tmem_u + (uint32_t)(acc * acc_cols);
        if (p.b_resident && p.cchunks == 1 && !p.trace) {
          // Hot path (Cin <= 64, weights resident): the issuing thread's own instruction stream bounds the MMA rate
          // at small N, so this loop carries nothing but the descriptor increments.
          mbar_wait(fullA(sa), pa);
          tc_fence_after();
          const uint32_t a_base = sw128_desc_lo(smem_u32(smemA + (size_t)sa * p.a_stage_bytes)) + (uint32_t)issuer * 1024u;
          const uint32_t b_step = p.b_tile_bytes >> 4;
          const uint32_t wp8 = (uint32_t)(p.Wp * p.d) * 8u, d8 = (uint32_t)p.d * 8u;
          const uint32_t a_inc = (uint32_t)p.n_iss * 1024u, d_inc = (uint32_t)(p.n_iss * p.acc_stride);
          const uint32_t d0 = d_tmem + (uint32_t)(issuer * p.acc_stride);
          const uint32_t b_lo0 = sw128_desc_lo(smem_u32(smemB));
          switch (p.last_kk) {
            case 1: halo_issue_taps<1, kTf32, kPair>(leader, a_base, b_lo0, b_step, wp8, d8, a_inc, d_inc, d0, idesc, issuer, p.NM, p.n_iss); break;
            case 2: halo_issue_taps<2, kTf32, kPair>(leader, a_base, b_lo0, b_step, wp8, d8, a_inc, d_inc, d0, idesc, issuer, p.NM, p.n_iss); break;
            case 3: halo_issue_taps<3, kTf32, kPair>(leader, a_base, b_lo0, b_step, wp8, d8, a_inc, d_inc, d0, idesc, issuer, p.NM, p.n_iss); break;
            default: halo_issue_taps<4, kTf32, kPair>(leader, a_base, b_lo0, b_step, wp8, d8, a_inc, d_inc, d0, idesc, issuer, p.NM, p.n_iss); break;
          }
          if (leader) halo_commit<kPair>(emptyA(sa));
          if (++sa == p.sA) { sa = 0; pa ^= 1u; }
        } else if (!p.b_resident && !p.trace) {
          // Streamed weights (Cin > 64 or several N tiles): one B tile per (tap, channel chunk) through the ring, each
          // serving this issuer's M-tiles.  Same discipline as the hot path: nothing in the loop but increments.
          const uint32_t wp8 = (uint32_t)(p.Wp * p.d) * 8u, d8 = (uint32_t)p.d * 8u;
          const uint32_t a_inc = (uint32_t)p.n_iss * 1024u, d_inc = (uint32_t)(p.n_iss * p.acc_stride);
          const uint32_t d0 = d_tmem + (uint32_t)(issuer * p.acc_stride);
          const uint32_t b_ring0 = sw128_desc_lo(smem_u32(smemB)), b_step = p.b_tile_bytes >> 4;
          bool accum = false;
          for (int cc = 0; cc < p.cchunks; ++cc) {
            mbar_wait(fullA(sa), pa);
            tc_fence_after();
            const int nk = (cc == p.cchunks - 1) ? p.last_kk : 4;
            uint32_t a_row = sw128_desc_lo(smem_u32(smemA + (size_t)sa * p.a_stage_bytes)) + (uint32_t)issuer * 1024u;
            for (int fr = 0; fr < 3; ++fr, a_row += wp8) {
              uint32_t a_tap = a_row;
              for (int fs = 0; fs < 3; ++fs, a_tap += d8) {
                mbar_wait(fullB(sb), pb);
                tc_fence_after();
                const uint32_t b_lo = b_ring0 + (uint32_t)sb * b_step;
                uint32_t a_lo = a_tap, dcol = d0;
                for (int m = issuer; m < p.NM; m += p.n_iss, a_lo += a_inc, dcol += d_inc)
                  halo_mma_n<kTf32, kPair>(nk, leader, dcol, a_lo, b_lo, idesc, accum);
                accum = true;
                if (leader) halo_commit<kPair>(emptyB(sb));
                if (++sb == p.sB) { sb = 0; pb ^= 1u; }
              }
            }
            if (leader) halo_commit<kPair>(emptyA(sa));
            if (++sa == p.sA) { sa = 0; pa ^= 1u; }
          }
        } else
        for (int cc = 0; cc < p.cchunks; ++cc) {
          mbar_wait(fullA(sa), pa);
          tc_fence_after();
          if (cc == 0 && leader && issuer == 0) trace(p.trace, 1, it, 1);
          const uint32_t a_lo0 = sw128_desc_lo(smem_u32(smemA + (size_t)sa * p.a_stage_bytes));
          const int nk = (cc == p.cchunks - 1) ? p.last_kk : 4;
          const uint32_t b_step = p.b_tile_bytes >> 4;                      // one B tile, in 16-byte units
          uint32_t b_res_lo = sw128_desc_lo(smem_u32(smemB)) + (uint32_t)cc * b_step;   // resident: tile (tap*cchunks + cc)
          const uint32_t b_res_inc = (uint32_t)p.cchunks * b_step;
          const uint32_t wp8 = (uint32_t)(p.Wp * p.d) * 8u, d8 = (uint32_t)p.d * 8u;    // row shifts in 16-byte units
          int tap = 0;
          for (int fr = 0; fr < 3; ++fr) {
            uint32_t a_tap = a_lo0 + (uint32_t)fr * wp8;
            for (int fs = 0; fs < 3; ++fs, ++tap, a_tap += d8) {
              uint32_t b_lo;
              if (p.b_resident) {
                b_lo = b_res_lo;
                b_res_lo += b_res_inc;
              } else {
                mbar_wait(fullB(sb), pb);
                tc_fence_after();
                b_lo = sw128_desc_lo(smem_u32(smemB + (size_t)sb * p.b_tile_bytes));
              }
              const bool acc_first = (cc | tap) != 0;
              uint32_t a_lo = ((p.trace & 4) ? a_lo0 : a_tap) + (uint32_t)issuer * 1024u, dcol = d_tmem + (uint32_t)(issuer * p.acc_stride);
              const uint32_t a_inc = (uint32_t)p.n_iss * 1024u, d_inc = (uint32_t)(p.n_iss * p.acc_stride);
              if (p.trace & 8) b_lo = sw128_desc_lo(smem_u32(smemB));
              for (int m = issuer; m < p.NM; m += p.n_iss, a_lo += a_inc, dcol += d_inc)
                halo_mma_n<kTf32, kPair>(nk, leader, dcol, a_lo, b_lo, idesc, acc_first);
              if (leader && tap < 8 && issuer == 0) trace(p.trace, 3, it, tap);
              if (!p.b_resident) {
                if (leader) halo_commit<kPair>(emptyB(sb));
                if (++sb == p.sB) { sb = 0; pb ^= 1u; }
              }
            }
          }
          if (leader) halo_commit<kPair>(emptyA(sa));
          if (++sa == p.sA) { sa = 0; pa ^= 1u; }
        }
        if (leader) halo_commit<kPair>(tfull(acc));
        if (leader && issuer == 0) trace(p.trace, 1, it, 2);
        if (leader && issuer == 0 && p.trace && blockIdx.x == 0 && it < 64) g_trace[(1 * 64 + it) * 8 + 5] = (unsigned long long)clock64();
      }
    }
   }
  } else {
    // ===================== epilogue (warps 4..15) =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    int col_begin, col_end;
    epi_col_range(p.BN, (warp - kEpiWarp0) >> 2, col_begin, col_end);
    EpiArgs ea;
    // Lean pipelined epilogue (16-column groups, pitch 48 B: staging + two residual buffers per warp fill exactly the
    // 12 x 32 x 144 B allocation) whenever the output is a plain 16-bit NHWC tile; the generic routine otherwise.
    const bool lean = epi_pipe_ok(p.BN, p.Cout, p.vec_ok, p.out_f32, 1, p.res != nullptr, sizeof(TH) == 4) &&
                      p.Cout == p.BN * p.n_tiles && p.om_groups == 0 && !(p.trace & 16);
    const bool pf_on = lean && p.res != nullptr;
    const bool lean_nores = lean && p.res == nullptr;
    ea.spitch = lean ? epi_pipe_pitch(16) : epi_stage_pitch(p.BN, p.out_f32);
    const uint32_t stage = smem_u32(stage_base) + (uint32_t)((warp - kEpiWarp0) * 32 * ea.spitch);
    const uint32_t rb0 = smem_u32(stage_base) + (uint32_t)(kEpiWarps * 32 * ea.spitch);
    const uint32_t rbuf[2] = {rb0 + (uint32_t)((2 * (warp - kEpiWarp0)) * 32 * ea.spitch),
                              rb0 + (uint32_t)((2 * (warp - kEpiWarp0) + 1) * 32 * ea.spitch)};
    int pf_have = 0, pf_sel = 0;
    ea.s_scale = smem_u32(s_scale); ea.s_shift = smem_u32(s_shift); ea.res = p.res; ea.y = p.y;
    ea.Cout = p.Cout; ea.BN = p.BN; ea.out_pitch = p.out_pitch; ea.res_pitch = p.res_pitch;
    ea.out_f32 = p.out_f32; ea.relu = p.relu; ea.vec_ok = p.vec_ok; ea.up = 1; ea.Wout = p.W;
    ea.res32 = p.res32; ea.y32 = p.y32; ea.res32_pitch = p.res32_pitch; ea.y32_pitch = p.y32_pitch;
    const bool stream_mode = p.res32 != nullptr || p.y32 != nullptr;
    // stream mode on a plain tile: the pipelined twin (same 48-byte staging / residual-buffer layout as the lean routines)
    const bool stream_lean = stream_mode && sizeof(TH) == 2 && p.Cout == p.BN * p.n_tiles && p.om_groups == 0 && p.res == nullptr &&
                             epi_stream_pipe_ok(p.BN, p.Cout, 1, p.out_f32, p.out_pitch, p.y, p.res32, p.res32_pitch, p.y32, p.y32_pitch);
    uint32_t sstage = 0, srb0 = 0, srb1 = 0;
    if (stream_lean) {
      const uint32_t sp = (uint32_t)epi_pipe_pitch(16);      // 48
      sstage = smem_u32(stage_base) + (uint32_t)(warp - kEpiWarp0) * 32u * sp;
      const uint32_t b0 = smem_u32(stage_base) + (uint32_t)kEpiWarps * 32u * sp;
      srb0 = b0 + (uint32_t)(2 * (warp - kEpiWarp0)) * 32u * sp;
      srb1 = b0 + (uint32_t)(2 * (warp - kEpiWarp0) + 1) * 32u * sp;
    }
    int it = 0;
    for (int unit = tile0; unit < n_units; unit += tile_step, ++it) {
      bool dup;
      const int tile = unit_tile(unit, dup);
      const int acc = (p.acc_bufs == 2) ? (it & 1) : 0;
      const uint32_t acc_phase = (p.acc_bufs == 2) ? ((uint32_t)(it >> 1) & 1u) : ((uint32_t)it & 1u);
      const int nt = tile % p.n_tiles;
      const int t2 = tile / p.n_tiles;
      const int ty = t2 % p.tiles_per_img, img = t2 / p.tiles_per_img;
      const int y0 = ty * p.BH;
      ea.ch_base = nt * p.BN;
      if (warp == kEpiWarp0 && lane == 0) trace(p.trace, 2, it, 2);
      // row -> pixel for M-tile m of this CTA tile
      auto row_pix = [&](int m, bool& valid) -> int {
        const int q = m * 128 + row;
        const int yy = q / p.Wp, xx = q - yy * p.Wp;
        valid = (yy < p.BH) && (xx < p.W) && (y0 + yy < p.H) && !dup;
        return valid ? (img * p.H + (y0 + yy)) * p.W + xx : 0;
      };
      // The residual of the NEXT M-tile (or of the first M-tile of this CTA's next tile) is copied into one of
      // two per-warp shared-memory buffers with cp.async while the current M-tile is drained: its global-load
      // latency, which used to be exposed once per M-tile, is off the critical path and costs no registers.
      // (A register prefetch was tried first and measured slower: 207 vs 173 us on the 48->48 conv.)
      mbar_wait(tfull(acc), acc_phase);
      tc_fence_after();
      if (warp == kEpiWarp0 && lane == 0) trace(p.trace, 2, it, 0);
      for (int m = 0; m < p.NM; ++m) {
        bool valid;
        const int pix = row_pix(m, valid);
        const uint32_t t_addr = tmem_base + (uint32_t)(acc * acc_cols + m * p.acc_stride) + ((uint32_t)(quarter * 32) << 16);
        if (stream_lean) {
         if constexpr (sizeof(TH) == 2) {
          // next unit of this warp (residual prefetch): the next M-tile, or the first M-tile of this CTA's next tile
          bool nvalid = false, have_next = true;
          int npix = 0, nchb = ea.ch_base;
          if (m + 1 < p.NM) {
            npix = row_pix(m + 1, nvalid);
          } else if (unit + tile_step < n_units) {
            bool ndup;
            const int ntile = unit_tile(unit + tile_step, ndup);
            const int nnt = ntile % p.n_tiles;
            const int nt2 = ntile / p.n_tiles;
            const int nty = nt2 % p.tiles_per_img, nimg = nt2 / p.tiles_per_img;
            const int ny0 = nty * p.BH;
            const int yy = row / p.Wp, xx = row - yy * p.Wp;
            nvalid = (yy < p.BH) && (xx < p.W) && (ny0 + yy < p.H) && !ndup;
            npix = nvalid ? (nimg * p.H + (ny0 + yy)) * p.W + xx : 0;
            nchb = nnt * p.BN;
          } else {
            have_next = false;
          }
          if (p.res32)
            epilogue_rows_pipelined_stream<TH, true>(ea, t_addr, col_begin, col_end, valid, pix, sstage, srb0, srb1, lane, pf_sel,
                                                     pf_have, have_next, nvalid, npix, nchb);
          else
            epilogue_rows_pipelined_stream<TH, false>(ea, t_addr, col_begin, col_end, valid, pix, sstage, 0u, 0u, lane, pf_sel,
                                                      pf_have, false, false, 0, 0);
         }
        } else if (stream_mode) {
          epilogue_rows_stream<TH>(ea, t_addr, col_begin, col_end, valid, pix);
        } else if (p.om_groups > 0) {
          const int q = m * 128 + row;
          const int yy = q / p.Wp, xx = q - yy * p.Wp;
          OmBlocked ob;
          ob.base = reinterpret_cast<float*>(p.y); ob.tiles_x = p.om_tiles_x; ob.tiles_y = p.om_tiles_y;
          ob.G3 = 3 * p.om_groups; ob.tap_stride = p.om_tap_stride; ob.kblocked = p.om_kblocked;
          if (!(p.trace & 2)) epilogue_rows_om_blocked(ea, t_addr, col_begin, col_end, valid, img, y0 + yy, xx, ob);
        } else if (pf_on) {
         if constexpr (sizeof(TH) == 2) {
          // next unit of this warp: the next M-tile, or the first M-tile of this CTA's next tile
          bool nvalid = false, have_next = true;
          int npix = 0, nchb = ea.ch_base;
          if (m + 1 < p.NM) {
            npix = row_pix(m + 1, nvalid);
          } else if (unit + tile_step < n_units) {
            bool ndup;
            const int ntile = unit_tile(unit + tile_step, ndup);
            const int nnt = ntile % p.n_tiles;
            const int nt2 = ntile / p.n_tiles;
            const int nty = nt2 % p.tiles_per_img, nimg = nt2 / p.tiles_per_img;
            const int ny0 = nty * p.BH;
            const int yy = row / p.Wp, xx = row - yy * p.Wp;
            nvalid = (yy < p.BH) && (xx < p.W) && (ny0 + yy < p.H) && !ndup;
            npix = nvalid ? (nimg * p.H + (ny0 + yy)) * p.W + xx : 0;
            nchb = nnt * p.BN;
          } else {
            have_next = false;
          }
          if (!(p.trace & 2))
            epilogue_rows_pipelined<TH>(ea, t_addr, col_begin, col_end, valid, pix, stage, rbuf[0], rbuf[1], lane, pf_sel, pf_have,
                                        have_next, nvalid, npix, nchb, 16);
         } else {
          bool nvalid = false, have_next = true;
          int npix = 0, nchb = ea.ch_base;
          if (m + 1 < p.NM) {
            npix = row_pix(m + 1, nvalid);
          } else if (unit + tile_step < n_units) {
            bool ndup;
            const int ntile = unit_tile(unit + tile_step, ndup);
            const int nnt = ntile % p.n_tiles;
            const int nt2 = ntile / p.n_tiles;
            const int nty = nt2 % p.tiles_per_img, nimg = nt2 / p.tiles_per_img;
            const int ny0 = nty * p.BH;
            const int yy = row / p.Wp, xx = row - yy * p.Wp;
            nvalid = (yy < p.BH) && (xx < p.W) && (ny0 + yy < p.H) && !ndup;
            npix = nvalid ? (nimg * p.H + (ny0 + yy)) * p.W + xx : 0;
            nchb = nnt * p.BN;
          } else {
            have_next = false;
          }
          epilogue_rows_pipelined_f32<true>(ea, t_addr, col_begin, col_end, valid, pix, stage, rbuf[0], rbuf[1], lane, pf_sel, pf_have,
                                            have_next, nvalid, npix, nchb);
         }
        } else if (lean_nores) {
         if constexpr (sizeof(TH) == 2) {
          if (!(p.trace & 2))
            epilogue_rows_pipelined<TH, false>(ea, t_addr, col_begin, col_end, valid, pix, stage, 0u, 0u, lane, pf_sel, pf_have, false,
                                               false, 0, 0, 16);
         } else {
          epilogue_rows_pipelined_f32<false>(ea, t_addr, col_begin, col_end, valid, pix, stage, 0u, 0u, lane, pf_sel, pf_have, false,
                                             false, 0, 0);
         }
        } else if (!(p.trace & 2)) {
          epilogue_rows<TH>(ea, t_addr, col_begin, col_end, valid, pix, stage, lane);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {                             // one arrival per warp, on the issuing (leader) CTA's barrier
        if constexpr (kPair) mbar_arrive_leader(tempty(acc));
        else mbar_arrive(tempty(acc));
      }
      if (warp == kEpiWarp0 && lane == 0) trace(p.trace, 2, it, 1);
    }
  }

  tc_fence_before();
  if constexpr (kPair) cluster_sync_all();
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (kPair)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

struct HaloCfg {
  int BN, n_tiles, CoutPad, cchunks;
  int BH, NM, HR, sA, sB, b_resident, acc_bufs, acc_stride;
  size_t smem;
  bool ok;
};

constexpr size_t kSmemBudget = 220 * 1024;

// Chooses the CTA tile: BH image rows (NM = ceil(BH*Wp/128) M-tiles).  Preference: double-buffered
// accumulators, high fraction of useful MMA rows, weights resident if they fit.
HaloCfg halo_cfg(int H, int W, int Cin, int Cout, int d, int kKC, bool pair = false) {
  HaloCfg best;
  best.ok = false;
  const int Wp = W + 2 * d;
  if (Wp > 256) return best;
  int n_tiles = (Cout + 255) / 256;
  int per = (Cout + n_tiles - 1) / n_tiles;
  int BN = ((per + 15) / 16) * 16;
  const int cchunks = (Cin + kKC - 1) / kKC;
  const int ksteps = 9 * cchunks;
  double best_score = -1;
  static const int ast_env = getenv("FAMI_HALO_ACCSTRIDE") ? atoi(getenv("FAMI_HALO_ACCSTRIDE")) : 0;
  const int ast = ast_env > 0 ? ((BN + ast_env - 1) / ast_env) * ast_env : BN;   // TMEM column stride between M-tile accumulators
  static const int nm_max = getenv("FAMI_HALO_NMMAX") ? atoi(getenv("FAMI_HALO_NMMAX")) : 8;   // experiment knob
  for (int NM = 1; NM <= nm_max; ++NM) {
    if (NM * ast > 512) break;
    int BH = (NM * 128) / Wp;
    if (BH < 1) continue;
    if (BH > H) BH = H;
    if (BH + 2 * d > 256) continue;
    const int nm = (BH * Wp + 127) / 128;   // actual M-tiles needed for BH rows
    const int acc_bufs = (2 * nm * ast <= 512) ? 2 : 1;
    int HR = nm * 128 + (2 * Wp + 2) * d;
    const int box_rows = (BH + 2 * d) * Wp;
    if (HR < box_rows) HR = box_rows;
    HR = ((HR + 7) / 8) * 8;
    const size_t a_stage = (size_t)HR * 128;
    const size_t b_tile = (size_t)(pair ? BN / 2 : BN) * 128;      // a CTA of a pair holds half of the rows of a weight tile
    const size_t b_all = b_tile * ksteps;
    static const bool no_resident = getenv("FAMI_HALO_NORES") != nullptr;   // experiment knob: streamed weights only
    if (pair && n_tiles != 1) break;        // the two CTAs of a pair share one N tile
    for (int resident = 1; resident >= 0; --resident) {
      if (resident && (n_tiles != 1 || no_resident)) continue;
      size_t bbytes;
      int sB;
      if (resident) { bbytes = b_all; sB = 0; }
      else { static const int sb_env = getenv("FAMI_HALO_SB") ? atoi(getenv("FAMI_HALO_SB")) : 4; sB = sb_env; bbytes = b_tile * sB; }
      for (int sA = 2; sA >= 1; --sA) {
        size_t smem = a_stage * sA + bbytes + 1024 + 256 + (size_t)kEpiWarps * 32 * (128 + 16) + (size_t)BN * n_tiles * 8;
        if (smem > kSmemBudget) continue;
        const int tiles_per_img = (H + BH - 1) / BH;
        const double eff = (double)H * W / ((double)tiles_per_img * nm * 128);
        if (eff < 0.7) continue;   // small maps waste too many MMA rows here: the im2col kernel takes them
        // feed cost model (bytes per useful output pixel), lower is better
        const double a_bytes = (double)box_rows * 128.0 * cchunks;
        const double b_bytes = resident ? 0.0 : (double)b_all;
        const double feed = (a_bytes + b_bytes) / ((double)BH * W);
        // resident weights: big tiles win even single-buffered (48->48: NM=5, sA=1 measured best).  Streamed weights:
        // the accumulator and the A stage must be double-buffered or the tile loses to the im2col kernel
        // (96->96: NM=2/sA=2/acc=2 72 us, NM=5/sA=1/acc=1 92 us, im2col 81 us).
        // tf32 (32-channel rows): a single A stage with several channel chunks serialises load and MMA per chunk
        // (48->48 tf32: resident NM=2 sA=1 291 us, streamed NM=3 sA=2 160 us)
        const double sa1 = (kKC == 32 && cchunks > 1) ? 0.35 : 0.85;
        double score = resident ? eff / feed * (acc_bufs == 2 ? 1.0 : 0.8) * (sA == 2 ? 1.0 : sa1)
                                : eff / feed * (acc_bufs == 2 ? 1.0 : 0.4) * (sA == 2 ? 1.0 : 0.5);
        if (score > best_score) {
          best_score = score;
          best.ok = true;
          best.BN = BN; best.n_tiles = n_tiles; best.CoutPad = BN * n_tiles; best.cchunks = cchunks;
          best.BH = BH; best.NM = nm; best.HR = HR; best.sA = sA; best.sB = resident ? 1 : sB;
          best.b_resident = resident; best.acc_bufs = acc_bufs; best.smem = smem; best.acc_stride = ast;
        }
        break;  // largest sA that fits for this (NM, resident)
      }
    }
  }
  return best;
}

}  // namespace

// Default policy for the pair form, from tools/time_convs.py on B200 (N = 160, us single -> pair; FAMI_HALO_PAIR=1 forces it):
//   fp16: 48->48 92 -> 100 (+res 98 -> 121), 96->96 69 -> 87, 64->64 112 -> 133, 192->48 68 -> 81, 96->48 66 -> 54
//   tf32: 48->48 159 -> 236 (+res 175 -> 248), 96->96 102 -> 97, 64->64 224 -> 173, 96->48 68.5 -> 67.6
// An M = 256 MMA with N <= 96 is dominated by the cross-SM exchange of the weight halves, and the two CTAs of a pair wait for
// the slower of two A loads at every hand-over: the narrow classes this kernel exists for lose.  The pair form is therefore
// taken only where it measured >= 15 % faster: tf32 64->64 (weights become resident) and 16-bit 96->48.
static bool halo_pair_default(const fami_conv_desc* d, const HaloCfg& single, const HaloCfg& pair) {
  (void)single; (void)pair;
  if (d->dtype == FAMI_TF32) return d->Cin == 64 && d->Cout == 64;
  return d->Cin == 96 && d->Cout == 48;
}

int conv_halo_supported(const fami_conv_desc* d) {
  if (d->kh != 3 || d->kw != 3 || d->stride != 1 || d->pad != d->dil || d->up != 1) return 0;
  const bool tf32 = d->dtype == FAMI_TF32;
  if (tf32 ? (d->Cin % 4 != 0 || d->in_pitch % 4 != 0) : (d->Cin % 16 != 0 || d->in_pitch % 8 != 0)) return 0;
  if (d->dil < 1 || d->dil > 8) return 0;
  // Measured (tools/prof_conv.py, tools/time_conv_shape.py, N=160 fp16): the halo form wins when the weights stay
  // resident in shared memory (Cin <= 64: 48->48 90 us vs 225 us im2col) and when many input channels feed few
  // output channels (256->48: 395 vs 590 us, 192->48: 85 vs 91 us: each streamed 6 KB weight tile serves NM M-tiles);
  // 96->96 wins with double-buffered accumulators and A stages (72 vs 81 us); it loses for the wider square
  // classes whose weights must be re-streamed per CTA tile (192->192: 75 vs 51 us).
  HaloCfg c = halo_cfg(d->H, d->W, d->Cin, d->Cout, d->dil, tf32 ? 32 : 64);
  static const bool force = getenv("FAMI_HALO_FORCE") != nullptr;   // experiment: also take streamed-weights shapes
  return (c.ok && (force || c.b_resident || c.n_tiles > 1 || (c.BN <= 64 && c.cchunks >= 3) ||
                   (c.BN <= 96 && c.acc_bufs == 2 && c.sA == 2))) ? 1 : 0;
}

int conv_halo_launch(const fami_conv_desc* d, const void* x, const void* w, const float* scale, const float* shift,
                     const void* res, void* y, cudaStream_t st, const float* res32, float* y32, int y32_pitch) {
  FAMI_CHECK_ARG(load_driver_fns(), "cuTensorMapEncode* driver entry points unavailable");
  const bool tf32 = d->dtype == FAMI_TF32;
  const int kKC = tf32 ? 32 : 64;
  const cuuint64_t es = tf32 ? 4 : 2;
  FAMI_CHECK_ARG(!tf32 || d->out_dtype == FAMI_F32 || d->out_dtype == FAMI_TF32, "conv_halo: tf32 convolutions write float");
  HaloCfg c = halo_cfg(d->H, d->W, d->Cin, d->Cout, d->dil, kKC);
  FAMI_CHECK_ARG(c.ok, "conv_halo: no tile configuration fits");
  // CTA pairs (cta_group::2): FAMI_HALO_PAIR=1 takes the pair form wherever it is legal (single N tile, no blocked output),
  // =0 never; default: see halo_pair_default()
  static const int pair_env = getenv("FAMI_HALO_PAIR") ? atoi(getenv("FAMI_HALO_PAIR")) : -1;
  bool pair = false;
  if (pair_env != 0 && c.n_tiles == 1 && d->om_groups == 0 && !res32 && !y32) {
    HaloCfg cp = halo_cfg(d->H, d->W, d->Cin, d->Cout, d->dil, kKC, true);
    if (cp.ok && (pair_env == 1 || halo_pair_default(d, c, cp))) { c = cp; pair = true; }
  }
  const CUtensorMapDataType tm_dtype = tm_dtype_of(d->dtype);
  const CUtensorMapDataType tm_wdtype = tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : tm_dtype;   // weights are pre-rounded
  const int dl = d->dil, Wp = d->W + 2 * dl;
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
    cuuint64_t strides[3] = {(cuuint64_t)d->in_pitch * es, (cuuint64_t)d->W * d->in_pitch * es,
                             (cuuint64_t)d->H * d->W * d->in_pitch * es};
    cuuint32_t box[4] = {(cuuint32_t)kKC, (cuuint32_t)Wp, (cuuint32_t)(c.BH + 2 * dl), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode_tiled(&tmA, tm_dtype, 4, const_cast<void*>(x), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FAMI_CHECK_ARG(r == CUDA_SUCCESS, "conv_halo: cuTensorMapEncodeTiled(A) failed (%d)", (int)r);
  }
  const int Kp = 9 * c.cchunks * kKC;
  {
    cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)c.CoutPad};
    cuuint64_t strides[1] = {(cuuint64_t)Kp * es};
    cuuint32_t box[2] = {(cuuint32_t)kKC, (cuuint32_t)(pair ? c.BN / 2 : c.BN)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(&tmB, tm_wdtype, 2, const_cast<void*>(w), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FAMI_CHECK_ARG(r == CUDA_SUCCESS, "conv_halo: cuTensorMapEncodeTiled(B) failed (%d)", (int)r);
  }
  HaloParams p;
  memset(&p, 0, sizeof(p));
  p.N = d->N; p.H = d->H; p.W = d->W; p.Wp = Wp; p.d = dl;
  p.BH = c.BH; p.NM = c.NM; p.HR = c.HR;
  p.tiles_per_img = (d->H + c.BH - 1) / c.BH;
  p.n_tiles = c.n_tiles;
  p.total_tiles = d->N * p.tiles_per_img * c.n_tiles;
  p.cchunks = c.cchunks;
  p.last_kk = (d->Cin - (c.cchunks - 1) * kKC + kKC / 4 - 1) / (kKC / 4);
  p.Cout = d->Cout; p.BN = c.BN;
  p.relu = d->relu; p.out_f32 = d->out_dtype == FAMI_F32 || d->out_dtype == FAMI_TF32;
  const size_t osz = p.out_f32 ? 4 : 2;
  p.vec_ok = ((reinterpret_cast<uintptr_t>(y) & 15) == 0) && ((d->out_pitch * osz) % 16 == 0) &&
             (!res || (((reinterpret_cast<uintptr_t>(res) & 15) == 0) && ((d->res_pitch * es) % 16 == 0)));
  p.out_pitch = d->out_pitch; p.res_pitch = d->res_pitch;
  p.sA = c.sA; p.sB = c.sB; p.b_resident = c.b_resident; p.acc_bufs = c.acc_bufs; p.acc_stride = c.acc_stride;
  p.n_iss = c.NM < kMmaWarps ? c.NM : kMmaWarps;
  if (getenv("FAMI_HALO_ISS")) { int v = atoi(getenv("FAMI_HALO_ISS")); if (v >= 1 && v <= p.n_iss) p.n_iss = v; }
  p.a_stage_bytes = (uint32_t)c.HR * 128u;
  p.b_tile_bytes = (uint32_t)(pair ? c.BN / 2 : c.BN) * 128u;
  p.a_box_bytes = (uint32_t)((c.BH + 2 * dl) * Wp) * 128u;
  p.scale = scale; p.shift = shift; p.res = res; p.y = y;
  p.res32 = res32; p.y32 = y32; p.res32_pitch = d->res_pitch; p.y32_pitch = y32_pitch;
  p.om_groups = d->om_groups;
  p.om_kblocked = d->om_layout == 3;
  p.om_tiles_x = (d->W + 7) / 8; p.om_tiles_y = (d->H + 15) / 16;
  p.om_tap_stride = (int64_t)d->N * p.om_tiles_x * p.om_tiles_y * 4 * (3 * (d->om_groups / 4)) * 128;
  if (p.om_kblocked) p.om_tap_stride = (int64_t)128 * 3 * d->om_groups;   // layout 3 is tile-major: [tile][tap]
  static const bool trace_on = getenv("FAMI_HALO_TRACE") != nullptr;
  p.trace = trace_on ? atoi(getenv("FAMI_HALO_TRACE")) : 0;

  static std::atomic<uint64_t> attr_h{0}, attr_b{0}, attr_t{0}, attr_h2{0}, attr_b2{0}, attr_t2{0};
  const int sms = num_sms();
  if (pair) {
    int clusters = (p.total_tiles + 1) / 2;
    if (clusters > sms / 2) clusters = sms / 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(kHThreads);
    cfg.dynamicSmemBytes = c.smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t e;
    if (tf32) {
      set_max_smem_once(attr_t2, conv_halo_kernel<float, true>, 227 * 1024);
      e = cudaLaunchKernelEx(&cfg, conv_halo_kernel<float, true>, tmA, tmB, p);
    } else if (d->dtype == FAMI_F16) {
      set_max_smem_once(attr_h2, conv_halo_kernel<__half, true>, 227 * 1024);
      e = cudaLaunchKernelEx(&cfg, conv_halo_kernel<__half, true>, tmA, tmB, p);
    } else {
      set_max_smem_once(attr_b2, conv_halo_kernel<__nv_bfloat16, true>, 227 * 1024);
      e = cudaLaunchKernelEx(&cfg, conv_halo_kernel<__nv_bfloat16, true>, tmA, tmB, p);
    }
    FAMI_CHECK_ARG(e == cudaSuccess, "conv_halo_kernel (pair): launch failed: %s", cudaGetErrorString(e));
    FAMI_CHECK_LAUNCH("conv_halo_kernel");
    return 0;
  }
  int grid = p.total_tiles;
  if (grid > sms) grid = sms;
  if (tf32) {
    set_max_smem_once(attr_t, conv_halo_kernel<float, false>, 227 * 1024);
    conv_halo_kernel<float, false><<<grid, kHThreads, c.smem, st>>>(tmA, tmB, p);
  } else if (d->dtype == FAMI_F16) {
    set_max_smem_once(attr_h, conv_halo_kernel<__half, false>, 227 * 1024);
    conv_halo_kernel<__half, false><<<grid, kHThreads, c.smem, st>>>(tmA, tmB, p);
  } else {
    set_max_smem_once(attr_b, conv_halo_kernel<__nv_bfloat16, false>, 227 * 1024);
    conv_halo_kernel<__nv_bfloat16, false><<<grid, kHThreads, c.smem, st>>>(tmA, tmB, p);
  }
  FAMI_CHECK_LAUNCH("conv_halo_kernel");
  return 0;
}

int debug_read_trace(unsigned long long* host_out, int n) {
  if (n > 8192) n = 8192;
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(host_out, g_trace, (size_t)n * sizeof(unsigned long long)) == cudaSuccess ? 0 : 1;
}

}  // namespace fami
