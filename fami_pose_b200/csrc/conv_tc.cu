// bf16 implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 + TMEM), fed by TMA.
// Tensor-core arm of fami_conv2d_bn_act_fwd (reference chains: basic_model.py:44-63,83-113;
// basic_layer.py:55-73; hrnet.py:89-172,651-680).
//
// GEMM view: D[M=128 pixels][N=BN couts] += A[128][64 ch] * B[BN][64 ch]^T per (tap, 64-channel chunk).
//   A: TMA *im2col* load straight from the NHWC bf16 activation: 128 consecutive output pixels
//      (linearised n,yo,xo), 64 channels of filter tap (r,s); the padding halo and the channel tail
//      (Cin % 64) are zero-filled by the TMA unit; stride-2 convs use the traversal stride.
//   B: TMA tiled load from weights pre-packed [CoutPad][taps*chunks*64] (K-major, zero padded).
//   Both land in 128-byte-swizzled K-major shared-memory tiles that tcgen05.mma reads through
//   shared-memory descriptors; the fp32 accumulator lives in TMEM (double buffered, 2 x 256 cols).
// Warp roles (192 threads, one persistent CTA per SM): warp 0 = TMA producer, warp 1 = TMEM owner +
// single-thread MMA issuer, warps 2-5 = epilogue (tcgen05.ld -> scale/shift (+residual) (+ReLU) ->
// bf16/fp32 NHWC stores with optional nearest-upsample replication).
#include "tc_common.cuh"

namespace fami {

namespace {

constexpr int kBM = 128;
constexpr int kABytes = kBM * 128;      // 16 KB: 128 pixels x one 128-byte swizzle row (TcTraits<TH>::kKC channels)
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kAccCols = 256;           // TMEM columns per accumulator buffer

struct TcParams {
  int M, Ho, Wo, HoWo;
  int stride, pad, dil, kw, taps;
  int cchunks, last_kk;
  int Cout, BN, n_tiles, m_tiles;
  int up, relu;
  int out_pitch, res_pitch;
  int out_f32, vec_ok;
  int stages, pipe;   // pipe: pipelined residual epilogue (epilogue_rows_pipelined)
  int om_groups, om_tiles_x, om_tiles_y;   // > 0: y is the row-blocked DCN offset|mask buffer (fami_conv_desc.om_groups)
  int64_t om_tap_stride;
  int om_kblocked;                         // fami_conv_desc.om_layout == 3
  const float* scale;
  const float* shift;
  const void* res;   // TH
  void* y;
  const float* res32;   // fp32 residual stream mode (16-bit arms): float residual / additional float output, may be null
  float* y32;
  int res32_pitch, y32_pitch;
};

// kPair: CTA pairs (cluster of 2, tcgen05 cta_group::2).  A pair processes two adjacent 128-pixel M-tiles as ONE
// M = 256 MMA per K-step: each CTA loads its own A tile and HALF of the B (weight) tile -- half the weight traffic from
// L2 and half the shared memory per stage, so deeper rings -- and every stage hand-over (the 1.4 us barrier round trip that
// bounds the single-CTA form) now covers twice the tensor work.  Only the leader CTA issues MMAs; the
// peer's TMA loads complete on the leader's full barrier; tcgen05.commit multicasts `empty` / `tfull` to both CTAs; the
// peer's epilogue warps release the accumulator on the leader's `tempty`.
template <typename TH, bool kPair>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int kKC = TcTraits<TH>::kKC;
  constexpr bool kTf32 = TcTraits<TH>::kTf32;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int stages = p.stages;
  const uint32_t cta_rank = kPair ? cluster_ctarank() : 0u;
  const bool is_leader_cta = cta_rank == 0u;
  const int bn_cta = kPair ? (p.BN >> 1) : p.BN;              // rows of the B tile held by this CTA
  const uint32_t b_bytes = (uint32_t)bn_cta * 128u;
  const uint32_t stage_bytes = kABytes + b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
  const uint32_t bar0 = smem_u32(bars);
  // barrier i at bar0 + 8*i: full[0..S), empty[S..2S), tfull[2S..2S+2), tempty[2S+2..2S+4)
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (stages + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * stages + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * stages + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * stages + 4);
  float* s_scale = reinterpret_cast<float*>(bars + 2 * stages + 6);
  float* s_shift = s_scale + p.n_tiles * p.BN;
  uint8_t* stage_base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(s_shift + p.n_tiles * p.BN) + 15) & ~(uintptr_t)15);
  fill_scale_shift(s_scale, s_shift, p.scale, p.shift, p.Cout, p.n_tiles * p.BN);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), kPair ? 2 * kEpiWarps : kEpiWarps); }
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
  if constexpr (kPair) cluster_sync_all();    // barrier inits and TMEM allocation of BOTH CTAs visible before any cross-CTA signal
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work units: (M-tile, N-tile); a CTA pair takes two adjacent M-tiles of one N-tile, the CTA of rank r the M-tile 2*u + r
  // (clamped to the last M-tile when the count is odd: computed twice, stored once)
  const int m_units = kPair ? (p.m_tiles + 1) >> 1 : p.m_tiles;
  const int total_tiles = m_units * p.n_tiles;
  const int tile0 = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  auto unit_mt = [&](int unit, bool& dup) -> int {
    if (!kPair) { dup = false; return unit; }
    const int mt = 2 * unit + (int)cta_rank;
    dup = mt >= p.m_tiles;
    return dup ? p.m_tiles - 1 : mt;
  };

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform control flow, one elected lane issues) =====
    {
      const bool leader = elect_one();
      if (leader) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        const int unit = tile / p.n_tiles, nt = tile - unit * p.n_tiles;
        bool dup;
        const int mt = unit_mt(unit, dup);
        const int m0 = mt * kBM;
        const int n = m0 / p.HoWo, r = m0 - n * p.HoWo;
        const int yo = r / p.Wo, xo = r - yo * p.Wo;
        const int cw = xo * p.stride - p.pad, ch = yo * p.stride - p.pad;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int fr = tap / p.kw, fs = tap - fr * p.kw;
          for (int cc = 0; cc < p.cchunks; ++cc) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            if (leader) {
              const uint32_t a_dst = smem_u32(smem + (size_t)stage * stage_bytes);
              if constexpr (kPair) {
                // the leader arms its full barrier for the bytes of BOTH CTAs; the peer's loads complete on it too
                if (is_leader_cta) mbar_arrive_expect_tx(full_bar(stage), 2u * stage_bytes);
                tma_im2col_4d_2sm(a_dst, &tmA, full_bar(stage), cc * kKC, cw, ch, n, (uint16_t)(fs * p.dil),
                                  (uint16_t)(fr * p.dil));
                tma_tiled_2d_2sm(a_dst + kABytes, &tmB, full_bar(stage), (tap * p.cchunks + cc) * kKC,
                                 nt * p.BN + (int)cta_rank * bn_cta);
              } else {
                mbar_arrive_expect_tx(full_bar(stage), stage_bytes);
                tma_im2col_4d(a_dst, &tmA, full_bar(stage), cc * kKC, cw, ch, n, (uint16_t)(fs * p.dil),
                              (uint16_t)(fr * p.dil));
                tma_tiled_2d(a_dst + kABytes, &tmB, full_bar(stage), (tap * p.cchunks + cc) * kKC, nt * p.BN);
              }
            }
            if (++stage == stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) =======
    if (is_leader_cta) {
      const bool leader = elect_one();
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      // instruction descriptor: D=f32, A=B=f16|bf16|tf32, both K-major, M=128 (256 for a CTA pair), N=BN
      const uint32_t idesc = kPair ? umma_idesc_pair<TH>(p.BN) : umma_idesc<TH>(p.BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      const uint32_t a_lo0 = sw128_desc_lo(smem_u32(smem));           // stage 0, A tile (descriptor low word: addr >> 4)
      const uint32_t stage_step = stage_bytes >> 4, b_off = (uint32_t)kABytes >> 4;
      const uint32_t empty_off = 8u * (uint32_t)stages;               // empty_bar(s) = full_bar(s) + empty_off
      uint32_t a_lo = a_lo0, full_cur = bar0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_u + (uint32_t)acc * kAccCols;
        // The issuing thread's own instruction stream is the bottleneck of this kernel when it is long (a k-step of
        // four N=96 MMAs is ~220 clk of tensor pipe): no divisions, descriptors and barrier addresses advanced
        // incrementally, the channel-chunk loop nested inside the tap loop.
        bool accum = false;
        for (int tap = 0; tap < p.taps; ++tap) {
          for (int cc = 0; cc < p.cchunks; ++cc) {
            mbar_wait(full_cur, phase);
            tc_fence_after();
            const int nk = (cc == p.cchunks - 1) ? p.last_kk : 4;
            if constexpr (kPair) {
              umma_pair_ksteps_n<kTf32>(nk, leader, d_tmem, a_lo, a_lo + b_off, idesc, accum);
              if (leader) umma_commit_pair(full_cur + empty_off);   // frees the slot in BOTH CTAs
            } else {
              umma_ksteps_n<kTf32>(nk, leader, d_tmem, a_lo, a_lo + b_off, idesc, accum);
              if (leader) umma_commit(full_cur + empty_off);   // frees the smem slot once the MMAs above have read it
            }
            accum = true;
            a_lo += stage_step;
            full_cur += 8u;
            if (++stage == stages) { stage = 0; phase ^= 1u; a_lo = a_lo0; full_cur = bar0; }
          }
        }
        if (leader) {                                  // accumulator complete -> epilogue (of both CTAs)
          if constexpr (kPair) umma_commit_pair(tfull_bar(acc));
          else umma_commit(tfull_bar(acc));
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    int col_begin, col_end;
    epi_col_range(p.BN, (warp - 2) >> 2, col_begin, col_end);
    const int Hout = p.Ho * p.up, Wout = p.Wo * p.up;
    EpiArgs ea;
    ea.spitch = epi_stage_pitch(p.BN, p.out_f32);
    const uint32_t stage = smem_u32(stage_base) + (uint32_t)((warp - 2) * 32 * ea.spitch);
    // replicate-on-write epilogue with pipelined residual: two extra per-warp buffers behind the staging buffers
    // (they fit in the allocation when the staging pitch is <= 48 B, i.e. 16 columns per warp)
    const uint32_t rbuf = smem_u32(stage_base) + (uint32_t)(kEpiWarps * 32 * ea.spitch) + (uint32_t)((warp - 2) * 64 * ea.spitch);
    const int up_fast = (p.up > 1 && p.res != nullptr && 3 * ea.spitch <= (128 + 16)) ? -1 : 0;
    // pipelined residual epilogue: staging + two residual buffers per warp at the 32-column pitch
    const uint32_t pp = (uint32_t)(sizeof(TH) == 4 ? epi_pipe_pitch(kPipeColsF32 * 2) : epi_pipe_pitch());   // 48 B (fp32, 8 cols) / 80 B
    const uint32_t pstage = smem_u32(stage_base) + (uint32_t)(warp - 2) * (p.res ? 96u : 32u) * pp;
    int psel = 0, pprimed = 0;
    ea.s_scale = smem_u32(s_scale); ea.s_shift = smem_u32(s_shift); ea.res = p.res; ea.y = p.y;
    ea.Cout = p.Cout; ea.BN = p.BN; ea.out_pitch = p.out_pitch; ea.res_pitch = p.res_pitch;
    ea.out_f32 = p.out_f32; ea.relu = p.relu; ea.vec_ok = p.vec_ok; ea.up = p.up; ea.Wout = Wout;
    ea.res32 = p.res32; ea.y32 = p.y32; ea.res32_pitch = p.res32_pitch; ea.y32_pitch = p.y32_pitch;
    const bool stream_mode = p.res32 != nullptr || p.y32 != nullptr;
    // stream mode on a plain tile: the pipelined twin; staging + two residual buffers per warp at the 48-byte pitch fit the
    // generic routine's allocation (kEpiWarps x 32 x 144 B)
    const bool stream_lean = stream_mode && sizeof(TH) == 2 && p.Cout == p.BN * p.n_tiles && p.om_groups == 0 && p.res == nullptr &&
                             epi_stream_pipe_ok(p.BN, p.Cout, p.up, p.out_f32, p.out_pitch, p.y, p.res32, p.res32_pitch, p.y32, p.y32_pitch);
    const uint32_t sps = 48u;
    const uint32_t sstage = smem_u32(stage_base) + (uint32_t)(warp - 2) * 96u * sps;
    int it = 0;
    for (int tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      const int unit = tile / p.n_tiles, nt = tile - unit * p.n_tiles;
      bool dup;
      const int mt = unit_mt(unit, dup);
      const int m = mt * kBM + row;
      const bool valid = m < p.M && !dup;
      int pix0 = 0;
      if (valid) {
        if (p.up == 1) {
          pix0 = m;
        } else {
          const int n = m / p.HoWo, r = m - n * p.HoWo;
          const int yo = r / p.Wo, xo = r - yo * p.Wo;
          pix0 = (n * Hout + yo * p.up) * Wout + xo * p.up;
        }
      }
      ea.ch_base = nt * p.BN;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t)acc * kAccCols + ((uint32_t)(quarter * 32) << 16);
      if (stream_lean) {
       if constexpr (sizeof(TH) == 2) {
        const int ntile = tile + tile_step;
        const bool have_next = ntile < total_tiles;
        const int nunit = ntile / p.n_tiles, nnt = ntile - nunit * p.n_tiles;
        bool ndup;
        const int nmt = unit_mt(nunit, ndup);
        const int nm = nmt * kBM + row;
        const bool nvalid = have_next && nm < p.M && !ndup;
        int npix = 0;                       // first replica of the next unit's row
        if (nvalid) {
          if (p.up == 1) {
            npix = nm;
          } else {
            const int n2 = nm / p.HoWo, r2 = nm - n2 * p.HoWo;
            const int yo2 = r2 / p.Wo, xo2 = r2 - yo2 * p.Wo;
            npix = (n2 * Hout + yo2 * p.up) * Wout + xo2 * p.up;
          }
        }
        if (p.res32)
          epilogue_rows_pipelined_stream<TH, true>(ea, t_addr, col_begin, col_end, valid, pix0, sstage, sstage + 32u * sps,
                                                   sstage + 64u * sps, lane, psel, pprimed, have_next, nvalid, npix, nnt * p.BN);
        else
          epilogue_rows_pipelined_stream<TH, false>(ea, t_addr, col_begin, col_end, valid, pix0, sstage, 0u, 0u, lane, psel, pprimed,
                                                    false, false, 0, 0);
       }
      } else if (stream_mode) {
        epilogue_rows_stream<TH>(ea, t_addr, col_begin, col_end, valid, pix0);
      } else if (p.om_groups > 0) {
        const int n = m / p.HoWo, r = m - n * p.HoWo;
        const int yo = r / p.Wo, xo = r - yo * p.Wo;
        OmBlocked ob;
        ob.base = reinterpret_cast<float*>(p.y); ob.tiles_x = p.om_tiles_x; ob.tiles_y = p.om_tiles_y;
        ob.G3 = 3 * p.om_groups; ob.tap_stride = p.om_tap_stride; ob.kblocked = p.om_kblocked;
        epilogue_rows_om_blocked(ea, t_addr, col_begin, col_end, valid, n, yo, xo, ob);
      } else if (p.pipe) {
        const int ntile = tile + tile_step;
        const bool have_next = ntile < total_tiles;
        const int nunit = ntile / p.n_tiles, nnt = ntile - nunit * p.n_tiles;
        bool ndup;
        const int nmt = unit_mt(nunit, ndup);
        const int nm = nmt * kBM + row;
        const bool nvalid = have_next && nm < p.M && !ndup;
       if constexpr (sizeof(TH) == 4) {
        if (p.res)
          epilogue_rows_pipelined_f32<true>(ea, t_addr, col_begin, col_end, valid, pix0, pstage, pstage + 32u * pp, pstage + 64u * pp,
                                            lane, psel, pprimed, have_next, nvalid, nvalid ? nm : 0, nnt * p.BN);
        else
          epilogue_rows_pipelined_f32<false>(ea, t_addr, col_begin, col_end, valid, pix0, pstage, 0u, 0u, lane, psel, pprimed, false,
                                             false, 0, 0);
       } else {
        if (p.res)
          epilogue_rows_pipelined<TH, true>(ea, t_addr, col_begin, col_end, valid, pix0, pstage, pstage + 32u * pp, pstage + 64u * pp,
                                            lane, psel, pprimed, have_next, nvalid, nvalid ? nm : 0, nnt * p.BN);
        else
          epilogue_rows_pipelined<TH, false>(ea, t_addr, col_begin, col_end, valid, pix0, pstage, 0u, 0u, lane, psel, pprimed, false,
                                             false, 0, 0);
       }
      } else {
        epilogue_rows<TH>(ea, t_addr, col_begin, col_end, valid, pix0, stage, lane, rbuf, up_fast);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {                               // one arrival per warp, on the issuing (leader) CTA's barrier
        if constexpr (kPair) mbar_arrive_leader(tempty_bar(acc));
        else mbar_arrive(tempty_bar(acc));
      }
    }
  }

  tc_fence_before();
  if constexpr (kPair) cluster_sync_all();    // no CTA frees its TMEM / exits while the peer's MMAs or signals may still touch it
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (kPair)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- host side ---------------------------------------------------------------------------------
struct TileCfg {
  int BN, n_tiles, CoutPad, cchunks, Kp;
};

TileCfg tile_cfg(int Cout, int Cin, int kh, int kw, int kKC) {
  TileCfg t;
  t.n_tiles = (Cout + 255) / 256;
  int per = (Cout + t.n_tiles - 1) / t.n_tiles;
  t.BN = ((per + 15) / 16) * 16;
  t.CoutPad = t.BN * t.n_tiles;
  t.cchunks = (Cin + kKC - 1) / kKC;
  t.Kp = kh * kw * t.cchunks * kKC;
  return t;
}

}  // namespace

static inline int kc_of(int dtype) { return dtype == FAMI_TF32 ? 32 : 64; }

int conv_bf16_tc_supported(const fami_conv_desc* d) {
  // K-steps of 32 bytes (16 halves / 8 floats); TMA needs 16-byte pixel strides
  // (tf32: Cin a multiple of 4 is enough -- the last 8-channel K-step reads TMA zero fill against zero-padded weights;
  //  the 324-channel offset|mask gradient of the alignment head needs it)
  if (d->dtype == FAMI_TF32 ? (d->Cin % 4 != 0 || d->in_pitch % 4 != 0) : (d->Cin % 16 != 0 || d->in_pitch % 8 != 0)) return 0;
  if (d->kh != d->kw || (d->kh != 1 && d->kh != 3)) return 0;
  if (d->stride < 1 || d->stride > 8) return 0;
  if (d->pad > 127 || (d->kh - 1) * d->dil - d->pad > 128) return 0;
  return 1;
}

int64_t pack_w_bf16_elems(int Cout, int Cin, int kh, int kw, int dtype) {
  TileCfg t = tile_cfg(Cout, Cin, kh, kw, kc_of(dtype));
  return (int64_t)t.CoutPad * t.Kp;
}

// OIHW float -> [CoutPad][taps][cchunks*64] half (f16 or bf16), zero padded
template <typename TH>
__global__ void pack_w_bf16_kernel(const float* __restrict__ w, TH* __restrict__ out, int Cout, int Cin,
                                   int taps, int cchunks, int CoutPad) {
  constexpr int kKC = TcTraits<TH>::kKC;
  const int Kp = taps * cchunks * kKC;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)CoutPad * Kp) return;
  int o = (int)(i / Kp), k = (int)(i - (int64_t)o * Kp);
  int tap = k / (cchunks * kKC), c = k - tap * cchunks * kKC;
  float v = 0.f;
  if (o < Cout && c < Cin) v = w[((int64_t)o * Cin + c) * taps + tap];
  if constexpr (TcTraits<TH>::kTf32) out[i] = f32_to_tf32_rna(v);   // the tensor core would truncate: round here instead
  else out[i] = from_f<TH>(v);
}

int pack_w_bf16_launch(const float* w, void* out, int Cout, int Cin, int kh, int kw, int dtype, cudaStream_t st) {
  TileCfg t = tile_cfg(Cout, Cin, kh, kw, kc_of(dtype));
  int64_t tot = (int64_t)t.CoutPad * t.Kp;
  if (dtype == FAMI_TF32)
    pack_w_bf16_kernel<float><<<cdiv(tot, 256), 256, 0, st>>>(w, (float*)out, Cout, Cin, kh * kw, t.cchunks, t.CoutPad);
  else if (dtype == FAMI_F16)
    pack_w_bf16_kernel<__half><<<cdiv(tot, 256), 256, 0, st>>>(w, (__half*)out, Cout, Cin, kh * kw, t.cchunks, t.CoutPad);
  else
    pack_w_bf16_kernel<__nv_bfloat16><<<cdiv(tot, 256), 256, 0, st>>>(w, (__nv_bfloat16*)out, Cout, Cin, kh * kw, t.cchunks, t.CoutPad);
  FAMI_CHECK_LAUNCH("pack_w_bf16_kernel");
  return 0;
}

// x, residual: bf16 NHWC; y: bf16 or fp32 (d->out_dtype)
int conv_bf16_tc_launch(const fami_conv_desc* d, const void* x, const void* w, const float* scale, const float* shift,
                        const void* res, void* y, double* stats, cudaStream_t st, const float* res32, float* y32,
                        int y32_pitch) {
  (void)stats;
  FAMI_CHECK_ARG(!d->stats, "bf16 tensor-core conv: fused BN statistics are not supported (use fami_bn_stats)");
  FAMI_CHECK_ARG(load_driver_fns(), "cuTensorMapEncode* driver entry points unavailable");
  FAMI_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0,
                 "bf16 tensor-core conv: x / w must be 16-byte aligned");
  const int out_f32 = d->out_dtype == FAMI_F32 || d->out_dtype == FAMI_TF32;
  const bool tf32 = d->dtype == FAMI_TF32;
  FAMI_CHECK_ARG(!tf32 || out_f32, "tf32 tensor-core conv: the output is float");
  const int kKC = kc_of(d->dtype);
  const cuuint64_t es = tf32 ? 4 : 2;     // element size
  TileCfg t = tile_cfg(d->Cout, d->Cin, d->kh, d->kw, kKC);
  // CTA pairs (cta_group::2).  Measured on B200 (tools/time_convs.py, N=160, fp16 / tf32): the wide 3x3 stride-1 classes gain
  // (192->192 +res 60.2 -> 56.2 us / 97.1 -> 90.9 us, 384->384 +res 56.2 -> 52.1 / 88.9 -> 84.8); 1x1, stride-2 and upsampling
  // convs, whose time is epilogue / HBM bound, LOSE to the extra cross-CTA hand-over (256->64 1x1: 123 -> 172 us), so the pair
  // form is selected for 3x3 stride-1 convs with Cin >= 128 only.  FAMI_TC_PAIR=0 disables it, =1 takes it wherever legal.
  static const int pair_env = getenv("FAMI_TC_PAIR") ? atoi(getenv("FAMI_TC_PAIR")) : -1;
  const int m_tiles_all = (d->N * d->Ho * d->Wo + kBM - 1) / kBM;
  const bool pair_legal = m_tiles_all >= 2 && d->om_groups == 0;
  const bool pair = pair_legal && (pair_env == 1 || (pair_env != 0 && d->kh == 3 && d->stride == 1 && d->up == 1 && d->Cin >= 128));

  CUtensorMap tmA, tmB;
  const CUtensorMapDataType tm_dtype = tm_dtype_of(d->dtype);
  const CUtensorMapDataType tm_wdtype = tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : tm_dtype;   // weights are pre-rounded
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
    cuuint64_t strides[3] = {(cuuint64_t)d->in_pitch * es, (cuuint64_t)d->W * d->in_pitch * es,
                             (cuuint64_t)d->H * d->W * d->in_pitch * es};
    int lower[2] = {-d->pad, -d->pad};
    int upper[2] = {d->pad - (d->kw - 1) * d->dil, d->pad - (d->kh - 1) * d->dil};
    cuuint32_t estr[4] = {1, (cuuint32_t)d->stride, (cuuint32_t)d->stride, 1};
    CUresult r = g_encode_im2col(&tmA, tm_dtype, 4, const_cast<void*>(x), dims, strides, lower,
                                 upper, kKC, kBM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FAMI_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeIm2col failed (%d)", (int)r);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)t.Kp, (cuuint64_t)t.CoutPad};
    cuuint64_t strides[1] = {(cuuint64_t)t.Kp * es};
    cuuint32_t box[2] = {(cuuint32_t)kKC, (cuuint32_t)(pair ? t.BN / 2 : t.BN)};   // a CTA of a pair holds half of the N rows
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(&tmB, tm_wdtype, 2, const_cast<void*>(w), dims, strides, box,
                                estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FAMI_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  }

  TcParams p;
  memset(&p, 0, sizeof(p));
  p.M = d->N * d->Ho * d->Wo;
  p.Ho = d->Ho; p.Wo = d->Wo; p.HoWo = d->Ho * d->Wo;
  p.stride = d->stride; p.pad = d->pad; p.dil = d->dil; p.kw = d->kw; p.taps = d->kh * d->kw;
  p.cchunks = t.cchunks;
  p.last_kk = (d->Cin - (t.cchunks - 1) * kKC + kKC / 4 - 1) / (kKC / 4);   // 32-byte K-steps in the last channel chunk
  p.Cout = d->Cout; p.BN = t.BN; p.n_tiles = t.n_tiles; p.m_tiles = (p.M + kBM - 1) / kBM;
  p.up = d->up; p.relu = d->relu;
  p.out_pitch = d->out_pitch; p.res_pitch = d->res_pitch;
  p.out_f32 = out_f32;
  const size_t osz = out_f32 ? 4 : 2;
  p.vec_ok = ((reinterpret_cast<uintptr_t>(y) & 15) == 0) && ((d->out_pitch * osz) % 16 == 0) &&
             (!res || (((reinterpret_cast<uintptr_t>(res) & 15) == 0) && ((d->res_pitch * es) % 16 == 0)));
  p.scale = scale; p.shift = shift; p.res = res; p.y = y;
  p.res32 = res32; p.y32 = y32; p.res32_pitch = d->res_pitch; p.y32_pitch = y32_pitch;
  p.om_groups = d->om_groups;
  p.om_kblocked = d->om_layout == 3;
  p.om_tiles_x = (d->Wo + 7) / 8; p.om_tiles_y = (d->Ho + 15) / 16;
  p.om_tap_stride = (int64_t)d->N * p.om_tiles_x * p.om_tiles_y * 4 * (3 * (d->om_groups / 4)) * 128;
  if (p.om_kblocked) p.om_tap_stride = (int64_t)128 * 3 * d->om_groups;   // layout 3 is tile-major: [tile][tap]

  const int stage_bytes = kABytes + (pair ? t.BN / 2 : t.BN) * 128;
  p.pipe = (!res32 && !y32 && d->om_groups == 0 && epi_pipe_ok(t.BN, d->Cout, p.vec_ok, out_f32, d->up, res != nullptr, tf32) && d->Cout == t.BN * t.n_tiles &&
            getenv("FAMI_NO_EPI_PIPE") == nullptr) ? 1 : 0;
  const size_t pipe_pitch = tf32 ? epi_pipe_pitch(kPipeColsF32 * 2) : epi_pipe_pitch();
  const size_t epi_bytes = p.pipe ? (size_t)kEpiWarps * 32 * (res ? 3 : 1) * pipe_pitch : (size_t)kEpiWarps * 32 * (128 + 16);
  int stages = (int)((228000 - (size_t)t.CoutPad * 8 - epi_bytes) / stage_bytes);   // all the shared memory there is: the kernel is bound by bytes in flight
  if (stages > 8) stages = 8;
  if (stages < 2) stages = 2;
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024 /*align slack*/ + (2 * stages + 4) * 8 + 64 +
                      (size_t)t.CoutPad * 8 + epi_bytes;
  static std::atomic<uint64_t> attr_h{0}, attr_b{0}, attr_t{0}, attr_h2{0}, attr_b2{0}, attr_t2{0};
  const int sms = num_sms();
  if (pair) {
    int clusters = ((p.m_tiles + 1) / 2) * p.n_tiles;
    if (clusters > sms / 2) clusters = sms / 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t e;
    if (tf32) {
      set_max_smem_once(attr_t2, conv_tc_kernel<float, true>, 227 * 1024);
      e = cudaLaunchKernelEx(&cfg, conv_tc_kernel<float, true>, tmA, tmB, p);
    } else if (d->dtype == FAMI_F16) {
      set_max_smem_once(attr_h2, conv_tc_kernel<__half, true>, 227 * 1024);
      e = cudaLaunchKernelEx(&cfg, conv_tc_kernel<__half, true>, tmA, tmB, p);
    } else {
      set_max_smem_once(attr_b2, conv_tc_kernel<__nv_bfloat16, true>, 227 * 1024);
      e = cudaLaunchKernelEx(&cfg, conv_tc_kernel<__nv_bfloat16, true>, tmA, tmB, p);
    }
    FAMI_CHECK_ARG(e == cudaSuccess, "conv_tc_kernel (pair): launch failed: %s", cudaGetErrorString(e));
    FAMI_CHECK_LAUNCH("conv_tc_kernel");
    return 0;
  }
  int grid = p.m_tiles * p.n_tiles;
  if (grid > sms) grid = sms;
  if (tf32) {
    set_max_smem_once(attr_t, conv_tc_kernel<float, false>, 227 * 1024);
    conv_tc_kernel<float, false><<<grid, kThreads, smem, st>>>(tmA, tmB, p);
  } else if (d->dtype == FAMI_F16) {
    set_max_smem_once(attr_h, conv_tc_kernel<__half, false>, 227 * 1024);
    conv_tc_kernel<__half, false><<<grid, kThreads, smem, st>>>(tmA, tmB, p);
  } else {
    set_max_smem_once(attr_b, conv_tc_kernel<__nv_bfloat16, false>, 227 * 1024);
    conv_tc_kernel<__nv_bfloat16, false><<<grid, kThreads, smem, st>>>(tmA, tmB, p);
  }
  FAMI_CHECK_LAUNCH("conv_tc_kernel");
  return 0;
}

}  // namespace fami
