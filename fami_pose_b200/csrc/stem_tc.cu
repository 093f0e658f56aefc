// HRNet stem convolution on the tensor cores: 3 -> Cout (64), 3x3, stride 2, pad 1, fp32 NHWC pixels in, 16-bit NHWC
// activations out, folded BatchNorm + ReLU in the epilogue (posetimation/backbones/hrnet.py:569-575 conv1 / bn1).
//
// Cin = 3 cannot be a TMA im2col operand (6 bytes per pixel), so the A tile is built by the CTA itself:
//   * 12 gather warps: thread = (output pixel of a 128-pixel tile, kernel row kr).  The three taps of a kernel row are
//     nine contiguous fp32 values of the input row ((2xo-1 .. 2xo+1) x 3 channels); they are converted to 16 bit and
//     stored as 24 bytes (9 values + 3 zeros) at K offset 12*kr of the pixel's 128-byte K-major row (128B swizzle);
//   * one warp issues 3 tcgen05.mma (M=128, N=Cout, K=16 each: K = 36 used of 48) against the filter, which every CTA
//     converts once from the fp32 packing into the same K order;
//   * 8 epilogue warps (2 column groups x 4 TMEM lane quarters) apply scale/shift/ReLU and store coalesced.
// The SIMT kernel this replaces spent 1728 FMAs per output pixel; here a pixel costs ~40 gather instructions.
#include "tc_common.cuh"

namespace fami {

namespace {

constexpr int kSGatherWarps = 12;
constexpr int kSEpiWarps = 8;
constexpr int kSThreads = 32 * (kSGatherWarps + 1 + kSEpiWarps);   // 672
constexpr int kSStages = 3;
constexpr int kSATile = 128 * 128;

struct StemParams {
  int N, H, W, Ho, Wo, M;
  int Cout, BN, CoutPad;
  int in_pitch, out_pitch, relu, vec_ok;
  int out_f32;          // 1: fp32 output (the tf32 arm: fp16 multiplicands -- the TF32 significand --, fp32 accumulate and store)
  int total_tiles;
  uint32_t ab_format;
  const float* x;
  const float* w;       // fp32 packing [27 (tap*3 + c)][CoutPad]
  const float* scale;
  const float* shift;
  void* y;
};

template <typename TH>
__global__ void __launch_bounds__(kSThreads, 1) stem_tc_kernel(const StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* s_a = smem;                                         // kSStages x 16 KB
  uint8_t* s_w = s_a + kSStages * kSATile;                     // BN x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_w + p.BN * 128);
  const uint32_t bar0 = smem_u32(bars);
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (kSStages + s); };
  auto tfull = [&](int a) { return bar0 + 8u * (2 * kSStages + a); };
  auto tempty = [&](int a) { return bar0 + 8u * (2 * kSStages + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kSStages + 4);
  float* s_scale = reinterpret_cast<float*>(bars + 2 * kSStages + 6);
  float* s_shift = s_scale + p.BN;
  uint8_t* stage_base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(s_shift + p.BN) + 15) & ~(uintptr_t)15);
  fill_scale_shift(s_scale, s_shift, p.scale, p.shift, p.Cout, p.BN);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // zero the A stages (K columns the gather never writes must read as 0) and build the filter tile
  for (int i = threadIdx.x; i < (kSStages * kSATile) / 16; i += blockDim.x) reinterpret_cast<uint4*>(s_a)[i] = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < p.BN * 64; i += blockDim.x) {
    const int co = i >> 6, kk = i & 63;              // kk = 12*kr + j, j = 3*ks + c for j < 9
    const int kr = kk / 12, j = kk - kr * 12;
    float v = 0.f;
    if (kr < 3 && j < 9 && co < p.Cout) v = __ldg(p.w + (int64_t)((kr * 3 + j / 3) * 3 + (j % 3)) * p.CoutPad + co);
    const uint32_t off = (uint32_t)co * 128u + ((((uint32_t)kk >> 3) ^ ((uint32_t)co & 7u)) << 4) + (((uint32_t)kk & 7u) << 1);
    *reinterpret_cast<TH*>(s_w + off) = from_f<TH>(v);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < kSStages; ++s) { mbar_init(a_full(s), kSGatherWarps); mbar_init(a_empty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), kSEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic stores above -> visible to the MMA
  if (warp == kSGatherWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int HoWo = p.Ho * p.Wo;

  if (warp < kSGatherWarps) {
    // ===================== gather warps =====================
    const int kr = threadIdx.x >> 7, r = threadIdx.x & 127;
    const uint32_t rsw = (uint32_t)(r & 7);
    // byte offsets 24*kr + {0, 8, 16} of the row, swizzled per 16-byte chunk
    uint32_t dst[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const uint32_t off = (uint32_t)(24 * kr + 8 * i);
      dst[i] = (uint32_t)r * 128u + (((off >> 4) ^ rsw) << 4) + (off & 15u);
    }
    int stage = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int m = tile * 128 + r;
      float v[9];
#pragma unroll
      for (int j = 0; j < 9; ++j) v[j] = 0.f;
      if (m < p.M) {
        const int n = m / HoWo, rr = m - n * HoWo;
        const int yo = rr / p.Wo, xo = rr - yo * p.Wo;
        const int yi = 2 * yo - 1 + kr, xi0 = 2 * xo - 1;
        if (yi >= 0 && yi < p.H) {
          const float* src = p.x + ((int64_t)(n * p.H + yi) * p.W + xi0) * p.in_pitch;
#pragma unroll
          for (int ks = 0; ks < 3; ++ks) {
            const int xi = xi0 + ks;
            if (xi >= 0 && xi < p.W) {
#pragma unroll
              for (int c = 0; c < 3; ++c) v[3 * ks + c] = __ldg(src + ks * p.in_pitch + c);
            }
          }
        }
      }
      mbar_wait(a_empty(stage), ph ^ 1u);
      const uint32_t a_st = smem_u32(s_a) + (uint32_t)(stage * kSATile);
      sts64(a_st + dst[0], make_uint2(f2_to_h2<TH>(v[0], v[1]), f2_to_h2<TH>(v[2], v[3])));
      sts64(a_st + dst[1], make_uint2(f2_to_h2<TH>(v[4], v[5]), f2_to_h2<TH>(v[6], v[7])));
      sts64(a_st + dst[2], make_uint2(f2_to_h2<TH>(v[8], 0.f), 0u));
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full(stage));
      if (++stage == kSStages) { stage = 0; ph ^= 1u; }
    }
  } else if (warp == kSGatherWarps) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t idesc = (1u << 4) | (p.ab_format << 7) | (p.ab_format << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t b_lo = sw128_desc_lo(smem_u32(s_w));
    int stage = 0, it = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      mbar_wait(tempty(acc), acc_phase ^ 1u);
      mbar_wait(a_full(stage), ph);
      tc_fence_after();
      umma_ksteps<3>(leader, tmem_u + (uint32_t)(acc * p.BN), sw128_desc_lo(smem_u32(s_a + stage * kSATile)), b_lo, idesc, false);
      if (leader) umma_commit(a_empty(stage));
      if (leader) umma_commit(tfull(acc));
      if (++stage == kSStages) { stage = 0; ph ^= 1u; }
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - kSGatherWarps - 1;        // 0..7
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int half_cols = ((p.BN / 16 + 1) / 2) * 16;
    const int cg = ew >> 2;
    // warps are numbered so that (warp & 3) is the TMEM lane quarter; the column group alternates every four warps
    const int col_begin = cg == 0 ? 0 : half_cols, col_end = cg == 0 ? half_cols : p.BN;
    EpiArgs ea;
    ea.s_scale = smem_u32(s_scale); ea.s_shift = smem_u32(s_shift); ea.res = nullptr; ea.y = p.y;
    ea.Cout = p.Cout; ea.BN = p.BN; ea.ch_base = 0; ea.out_pitch = p.out_pitch; ea.res_pitch = 0;
    ea.out_f32 = p.out_f32; ea.relu = p.relu; ea.vec_ok = p.vec_ok; ea.up = 1; ea.Wout = p.Wo;
    ea.spitch = epi_pipe_pitch(32);
    const uint32_t pstage = smem_u32(stage_base) + (uint32_t)(ew * 32 * epi_pipe_pitch(32));
    int sel = 0, primed = 0, it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      const int m = tile * 128 + row;
      const bool valid = m < p.M;
      mbar_wait(tfull(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t)(acc * p.BN) + ((uint32_t)(quarter * 32) << 16);
      if (p.out_f32)
        epilogue_rows_pipelined_f32<false>(ea, t_addr, col_begin, col_end, valid, valid ? m : 0, pstage, 0u, 0u, lane, sel, primed, false,
                                           false, 0, 0);
      else
        epilogue_rows_pipelined<TH, false>(ea, t_addr, col_begin, col_end, valid, valid ? m : 0, pstage, 0u, 0u, lane, sel, primed, false,
                                           false, 0, 0, 32);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(acc));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kSGatherWarps) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace

int stem_tc_supported(const fami_conv_desc* d, const void* y) {
  if (d->Cin != 3 || d->kh != 3 || d->kw != 3 || d->stride != 2 || d->pad != 1 || d->dil != 1 || d->up != 1 || d->stats) return 0;
  // fp32 pixels in; 16-bit out (fp16 / bf16 arms) or fp32 out (tf32 arm, selected by the caller: fami_conv2d_bn_act_fwd with
  // dtype FAMI_TF32 and Cin = 3)
  if (d->dtype != FAMI_F32 || !(is_half_dtype(d->out_dtype) || d->out_dtype == FAMI_F32)) return 0;
  if (d->Cout % 16 != 0 || d->Cout > 128 || d->out_pitch % (d->out_dtype == FAMI_F32 ? 4 : 8) != 0 ||
      (reinterpret_cast<uintptr_t>(y) & 15) != 0)
    return 0;
  if ((int64_t)d->N * d->Ho * d->Wo >= (1ll << 31) - 256) return 0;
  return 1;
}

int stem_tc_launch(const fami_conv_desc* d, const float* x, const float* w, const float* scale, const float* shift, void* y,
                   cudaStream_t st) {
  StemParams p;
  memset(&p, 0, sizeof(p));
  p.N = d->N; p.H = d->H; p.W = d->W; p.Ho = d->Ho; p.Wo = d->Wo; p.M = d->N * d->Ho * d->Wo;
  p.Cout = d->Cout; p.BN = d->Cout; p.CoutPad = fami_conv_cout_pad(d->Cout);
  p.in_pitch = d->in_pitch; p.out_pitch = d->out_pitch; p.relu = d->relu; p.vec_ok = 1;
  p.total_tiles = (p.M + 127) / 128;
  p.ab_format = d->out_dtype == FAMI_BF16 ? 1u : 0u;      // fp16 multiplicands unless the arm is bf16
  p.out_f32 = d->out_dtype == FAMI_F32 ? 1 : 0;
  p.x = x; p.w = w; p.scale = scale; p.shift = shift; p.y = y;
  const size_t smem = (size_t)kSStages * kSATile + (size_t)p.BN * 128 + 1024 + 256 + (size_t)p.BN * 8 +
                      (size_t)kSEpiWarps * 32 * epi_pipe_pitch(32);
  int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  if (d->out_dtype != FAMI_BF16) {
    static std::atomic<uint64_t> mask{0};
    set_max_smem_once(mask, stem_tc_kernel<__half>, 227 * 1024);
    stem_tc_kernel<__half><<<grid, kSThreads, smem, st>>>(p);
  } else {
    static std::atomic<uint64_t> mask{0};
    set_max_smem_once(mask, stem_tc_kernel<__nv_bfloat16>, 227 * 1024);
    stem_tc_kernel<__nv_bfloat16><<<grid, kSThreads, smem, st>>>(p);
  }
  FAMI_CHECK_LAUNCH("stem_tc_kernel");
  return 0;
}

}  // namespace fami
