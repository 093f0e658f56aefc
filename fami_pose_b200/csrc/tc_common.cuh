// Shared tcgen05 / TMA / mbarrier helpers for the tensor-core kernels (sm_100a inline PTX).
#pragma once
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace fami {

template <typename TH> __device__ __forceinline__ float2 h2_to_f2(uint32_t w);
template <> __device__ __forceinline__ float2 h2_to_f2<__nv_bfloat16>(uint32_t w) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
}
template <> __device__ __forceinline__ float2 h2_to_f2<__half>(uint32_t w) {
  return __half22float2(*reinterpret_cast<const __half2*>(&w));
}
template <typename TH> __device__ __forceinline__ uint32_t f2_to_h2(float a, float b);
template <> __device__ __forceinline__ uint32_t f2_to_h2<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
template <> __device__ __forceinline__ uint32_t f2_to_h2<__half>(float a, float b) {
  __half2 t = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
// suspend-time hint of mbarrier.try_wait: the waiting warp sleeps in hardware (woken by the completing arrival) instead of
// re-issuing the poll every ~40 clk -- a CTA's spinning epilogue / issuer warps executed 30 % of all warp instructions of the
// deformable kernel and competed with its gather warps for issue slots
constexpr uint32_t kMbarSuspendHint = 0x989680u;
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar), "r"(parity), "r"(kMbarSuspendHint)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_im2col_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c, int w, int h,
                                              int n, uint16_t offw, uint16_t offh) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(offw), "h"(offh)
      : "memory");
}
__device__ __forceinline__ void tma_tiled_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_tiled_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2,
                                             int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// ---- lean MMA issue ---------------------------------------------------------------------------
// Measured on B200 (tools/probe_umma_rate.py): a tcgen05.mma of M=128,K=16 costs max(N/2, ~44) clk in
// the tensor pipe, but the ISSUE side is easily slower: every uniform-datapath instruction between two
// MMAs adds to the per-MMA cost (a loop that rebuilds 64-bit descriptors measured 110-190 clk/MMA).
// So descriptors are kept as (constant high word, 32-bit low word) and only the low word is advanced.
constexpr uint32_t kSw128DescHi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO=1024, version 1, SWIZZLE_128B
__device__ __forceinline__ uint32_t sw128_desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }

// Operand traits of the tensor-core kernels, keyed on the STORAGE type of the activations:
//   __half / __nv_bfloat16 -> tcgen05.mma.kind::f16 (K = 16 per instruction, 64 channels per 128-byte swizzle row)
//   float                  -> tcgen05.mma.kind::tf32 (K = 8 per instruction, 32 channels per 128-byte row; the tensor
//                             core reads the upper 19 bits of each fp32 operand: the "tf32" arm = fp32 storage, TF32
//                             multiplicands, fp32 accumulation -- what cuDNN does for the reference's fp32 convs on a GPU)
// A K-step is 32 bytes of the row in both cases, so descriptor arithmetic is identical.
template <typename TH> struct TcTraits;
template <> struct TcTraits<__half> { static constexpr int kKC = 64; static constexpr uint32_t kFmt = 0u; static constexpr bool kTf32 = false; };
template <> struct TcTraits<__nv_bfloat16> { static constexpr int kKC = 64; static constexpr uint32_t kFmt = 1u; static constexpr bool kTf32 = false; };
template <> struct TcTraits<float> { static constexpr int kKC = 32; static constexpr uint32_t kFmt = 2u; static constexpr bool kTf32 = true; };
// instruction descriptor: D = f32, A / B formats from the traits, both K-major, M = 128, N = BN
template <typename TH> __device__ __forceinline__ uint32_t umma_idesc(int BN) {
  return (1u << 4) | (TcTraits<TH>::kFmt << 7) | (TcTraits<TH>::kFmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// ---- CTA pairs (tcgen05 cta_group::2) -----------------------------------------------------------
// Two CTAs of a cluster (same TPC) execute ONE MMA of M = 256: each CTA's tensor core produces its own 128 accumulator
// rows from its own A tile and reads N/2 rows of B from its own shared memory and N/2 from the peer's.  Only the leader
// (cluster rank 0) issues; tcgen05.commit multicasts the completion to the same barrier in both CTAs; TMA loads of the
// peer signal the LEADER's full barrier (address with the peer bit cleared).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_im2col_4d_2sm(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c, int w, int h, int n,
                                                  uint16_t offw, uint16_t offh) {
  asm volatile(
      "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar & kPeerBitMask), "r"(c), "r"(w), "r"(h), "r"(n), "h"(offw), "h"(offh)
      : "memory");
}
__device__ __forceinline__ void tma_tiled_2d_2sm(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_tiled_4d_2sm(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// arrive on the LEADER CTA's copy of a barrier (from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
}
// completion of all prior MMAs of this thread -> arrive on the same barrier in both CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
template <bool kTf32>
__device__ __forceinline__ void umma_pair_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, bool accumulate) {
  const uint64_t ad = ((uint64_t)kSw128DescHi << 32) | a_lo;
  const uint64_t bd = ((uint64_t)kSw128DescHi << 32) | b_lo;
  if constexpr (kTf32) {
    if (accumulate) {
      asm volatile("{\n.reg .pred p;\nsetp.eq.u32 p, 1, 1;\ntcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                   ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc) : "memory");
    } else {
      asm volatile("{\n.reg .pred p;\nsetp.eq.u32 p, 1, 0;\ntcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                   ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc) : "memory");
    }
  } else {
    if (accumulate) {
      asm volatile("{\n.reg .pred p;\nsetp.eq.u32 p, 1, 1;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                   ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc) : "memory");
    } else {
      asm volatile("{\n.reg .pred p;\nsetp.eq.u32 p, 1, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                   ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc) : "memory");
    }
  }
}
template <int NK, bool kTf32>
__device__ __forceinline__ void umma_pair_ksteps(bool leader, uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                                 bool first_accumulates) {
  if (leader) {
    if (first_accumulates) umma_pair_lo<kTf32>(tmem_d, a_lo, b_lo, idesc, true);
    else umma_pair_lo<kTf32>(tmem_d, a_lo, b_lo, idesc, false);
#pragma unroll
    for (int k = 1; k < NK; ++k) umma_pair_lo<kTf32>(tmem_d, a_lo + 2u * k, b_lo + 2u * k, idesc, true);
  }
}
template <bool kTf32>
__device__ __forceinline__ void umma_pair_ksteps_n(int nk, bool leader, uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo,
                                                   uint32_t idesc, bool first_accumulates) {
  switch (nk) {
    case 1: umma_pair_ksteps<1, kTf32>(leader, tmem_d, a_lo, b_lo, idesc, first_accumulates); break;
    case 2: umma_pair_ksteps<2, kTf32>(leader, tmem_d, a_lo, b_lo, idesc, first_accumulates); break;
    case 3: umma_pair_ksteps<3, kTf32>(leader, tmem_d, a_lo, b_lo, idesc, first_accumulates); break;
    default: umma_pair_ksteps<4, kTf32>(leader, tmem_d, a_lo, b_lo, idesc, first_accumulates); break;
  }
}
// instruction descriptor of a pair MMA: M = 256
template <typename TH> __device__ __forceinline__ uint32_t umma_idesc_pair(int BN) {
  return (1u << 4) | (TcTraits<TH>::kFmt << 7) | (TcTraits<TH>::kFmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

template <bool kTf32 = false>
__device__ __forceinline__ void umma_f16_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, bool accumulate) {
  const uint64_t ad = ((uint64_t)kSw128DescHi << 32) | a_lo;
  const uint64_t bd = ((uint64_t)kSw128DescHi << 32) | b_lo;
  if constexpr (kTf32) {
    if (accumulate) {
      asm volatile("{\n.reg .pred p;\nsetp.eq.u32 p, 1, 1;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                   ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc) : "memory");
    } else {
      asm volatile("{\n.reg .pred p;\nsetp.eq.u32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                   ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc) : "memory");
    }
  } else {
    if (accumulate) {
      asm volatile("{\n.reg .pred p;\nsetp.eq.u32 p, 1, 1;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                   ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc) : "memory");
    } else {
      asm volatile("{\n.reg .pred p;\nsetp.eq.u32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                   ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc) : "memory");
    }
  }
}
// NK K-steps (32 B of the 128-byte row = +2 in the descriptor address field each) of one (A rows, B tile) pair
template <int NK, bool kTf32 = false>
__device__ __forceinline__ void umma_ksteps(bool leader, uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                            bool first_accumulates) {
  if (leader) {
    if (first_accumulates) umma_f16_lo<kTf32>(tmem_d, a_lo, b_lo, idesc, true);
    else umma_f16_lo<kTf32>(tmem_d, a_lo, b_lo, idesc, false);
#pragma unroll
    for (int k = 1; k < NK; ++k) umma_f16_lo<kTf32>(tmem_d, a_lo + 2u * k, b_lo + 2u * k, idesc, true);
  }
}
template <bool kTf32 = false>
__device__ __forceinline__ void umma_ksteps_n(int nk, bool leader, uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo,
                                              uint32_t idesc, bool first_accumulates) {
  switch (nk) {
    case 1: umma_ksteps<1, kTf32>(leader, tmem_d, a_lo, b_lo, idesc, first_accumulates); break;
    case 2: umma_ksteps<2, kTf32>(leader, tmem_d, a_lo, b_lo, idesc, first_accumulates); break;
    case 3: umma_ksteps<3, kTf32>(leader, tmem_d, a_lo, b_lo, idesc, first_accumulates); break;
    default: umma_ksteps<4, kTf32>(leader, tmem_d, a_lo, b_lo, idesc, first_accumulates); break;
  }
}
// round-to-nearest (ties away) fp32 -> tf32, result as an fp32 bit pattern with the low 13 mantissa bits zero
__device__ __forceinline__ float f32_to_tf32_rna(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// explicit shared-space accesses with 32-bit addresses: pointers derived from the manually aligned dynamic
// shared-memory base lose their address space and compile to generic LD/ST (slower than LDS/STS)
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float4 lds128f(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts128f(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t a, uint2 v) {
  asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(a), "r"(v.x), "r"(v.y) : "memory");
}

// whole-warp wait with a single polling lane (32 lanes polling one mbarrier add contention for nothing)
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity, int lane) {
  if (lane == 0) mbar_wait(bar, parity);
  __syncwarp();
}

// true in exactly one lane of a fully converged warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, 128B-swizzled operand tile: rows of 128 B, 8-row swizzle atoms 1024 B apart.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);   // start address  [0,14)
  d |= (uint64_t)1 << 16;                    // LBO (unused for swizzled K-major) [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;          // SBO = 1024 B [32,46)
  d |= (uint64_t)1 << 46;                    // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }


// ------------------------------------------------------------------------------------------------
// Shared epilogue: one warp drains its 32 accumulator rows (TMEM lane quarter) of a 128 x BN tile.
//   y[pix(row) (+ replica), ch] = act(scale*acc + shift + residual)
// Direct per-lane global accesses would touch 32 different cache lines per instruction (each lane
// owns one pixel row), which measured as the bottleneck of both conv kernels.  Instead every column
// group (64 x 16-bit or 32 x fp32 = 128 B per row) goes through a per-warp shared-memory staging
// tile with an odd 16-byte row pitch (conflict-free): residual is fetched and the result is written
// with fully coalesced 16-byte pieces (consecutive lanes -> consecutive addresses).
// ------------------------------------------------------------------------------------------------
constexpr int kEpiGroups = 3;                       // epilogue warp groups (column ranges), 4 warps each
constexpr int kEpiWarps = 4 * kEpiGroups;

// staging row pitch (bytes) for a conv with BN output columns per tile: payload of the widest column
// range handled by one epilogue group (<= 128 B) + 16 (odd multiple of 16 -> conflict-free rows)
__host__ __device__ inline int epi_stage_pitch(int BN, int out_f32) {
  const int chunks = BN >> 4;
  const int widest = ((chunks + kEpiGroups - 1) / kEpiGroups) << 4;   // columns
  int bytes = widest * (out_f32 ? 4 : 2);
  if (bytes > 128) bytes = 128;
  return bytes + 16;
}

// fills the per-CTA scale/shift tables (defaults 1 / 0 beyond Cout or when the pointer is null)
__device__ __forceinline__ void fill_scale_shift(float* s_scale, float* s_shift, const float* scale, const float* shift,
                                                 int Cout, int CoutPad) {
  for (int c = threadIdx.x; c < CoutPad; c += blockDim.x) {
    s_scale[c] = (scale && c < Cout) ? scale[c] : 1.f;
    s_shift[c] = (shift && c < Cout) ? shift[c] : 0.f;
  }
}

// column range [begin,end) of epilogue group g (multiples of 16, as even as possible)
__device__ __forceinline__ void epi_col_range(int BN, int g, int& begin, int& end) {
  const int chunks = BN >> 4;
  const int base = chunks / kEpiGroups, rem = chunks - base * kEpiGroups;
  const int b = g * base + (g < rem ? g : rem);
  begin = b << 4;
  end = (b + base + (g < rem ? 1 : 0)) << 4;
}

struct EpiArgs {
  uint32_t s_scale;       // shared-memory address of float[CoutPad] (1/0 defaults beyond Cout)
  uint32_t s_shift;
  const void* res;        // TH, may be null
  void* y;                // TH or float
  int Cout, BN, ch_base;  // ch_base = first output channel of this N tile
  int out_pitch, res_pitch;
  int out_f32, relu, vec_ok;
  int up, Wout;           // nearest-upsample replication (1 = none); Wout = output row width in pixels
  int spitch;             // staging row pitch in bytes (epi_stage_pitch)
  // fp32 residual stream of the 16-bit arms (fami_conv2d_bn_act_fwd_stream): the residual operand is read from a float
  // tensor and the result is additionally written, before rounding, to a float tensor, so that the ~100 sequential
  // residual additions of the HRNet trunk accumulate in fp32 while every MMA operand stays 16-bit
  const float* res32;     // may be null
  float* y32;             // may be null
  int res32_pitch, y32_pitch;
};

// Direct (unstaged) epilogue of the fp32-residual-stream mode: each lane owns one accumulator row and moves 64-byte
// runs itself.  Slower than the staged routines (32 cache lines per warp instruction) -- this is the parity mode of
// the bf16 arm, not the throughput path.
template <typename TH>
__device__ __forceinline__ void epilogue_rows_stream(const EpiArgs& a, uint32_t t_addr, int col_begin, int col_end, bool valid,
                                                     int pix0) {
  if (col_begin >= col_end) return;   // warp-uniform
  const bool rvec = a.res32 && (a.res32_pitch % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.res32) & 15) == 0);
  const bool y32vec = a.y32 && (a.y32_pitch % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.y32) & 15) == 0);
  for (int c0 = col_begin; c0 < col_end; c0 += 16) {
    const int ch0 = a.ch_base + c0;
    if (ch0 >= a.Cout) break;   // warp-uniform
    uint32_t v[16];
    tmem_ld16(t_addr + (uint32_t)c0, v);
    tmem_ld_wait();
    float o[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 sc = lds128f(a.s_scale + (uint32_t)(ch0 + 4 * j) * 4u);
      const float4 sh = lds128f(a.s_shift + (uint32_t)(ch0 + 4 * j) * 4u);
      o[4 * j + 0] = fmaf(__uint_as_float(v[4 * j + 0]), sc.x, sh.x);
      o[4 * j + 1] = fmaf(__uint_as_float(v[4 * j + 1]), sc.y, sh.y);
      o[4 * j + 2] = fmaf(__uint_as_float(v[4 * j + 2]), sc.z, sh.z);
      o[4 * j + 3] = fmaf(__uint_as_float(v[4 * j + 3]), sc.w, sh.w);
    }
    if (!valid) continue;
    const bool full = ch0 + 16 <= a.Cout;
    for (int dy = 0; dy < a.up; ++dy)
      for (int dx = 0; dx < a.up; ++dx) {
        const int64_t pix = (int64_t)pix0 + dy * a.Wout + dx;
        float t[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) t[j] = o[j];
        if (a.res32) {
          const float* r = a.res32 + pix * a.res32_pitch + ch0;
          if (full && rvec) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 q = __ldg(reinterpret_cast<const float4*>(r) + j);
              t[4 * j] += q.x; t[4 * j + 1] += q.y; t[4 * j + 2] += q.z; t[4 * j + 3] += q.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) if (ch0 + j < a.Cout) t[j] += __ldg(r + j);
          }
        } else if (a.res) {
          const TH* r = reinterpret_cast<const TH*>(a.res) + pix * a.res_pitch + ch0;
#pragma unroll
          for (int j = 0; j < 16; ++j) if (ch0 + j < a.Cout) t[j] += to_f<TH>(r[j]);
        }
        if (a.relu) {
#pragma unroll
          for (int j = 0; j < 16; ++j) t[j] = fmaxf(t[j], 0.f);
        }
        if (a.y32) {
          float* d = a.y32 + pix * a.y32_pitch + ch0;
          if (full && y32vec) {
#pragma unroll
            for (int j = 0; j < 4; ++j) reinterpret_cast<float4*>(d)[j] = make_float4(t[4 * j], t[4 * j + 1], t[4 * j + 2], t[4 * j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) if (ch0 + j < a.Cout) d[j] = t[j];
          }
        }
        if (a.out_f32) {
          float* d = reinterpret_cast<float*>(a.y) + pix * a.out_pitch + ch0;
#pragma unroll
          for (int j = 0; j < 16; ++j) if (ch0 + j < a.Cout) d[j] = t[j];
        } else if constexpr (sizeof(TH) == 2) {
          TH* d = reinterpret_cast<TH*>(a.y) + pix * a.out_pitch + ch0;
          if (full && a.vec_ok) {
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) w[j] = f2_to_h2<TH>(t[2 * j], t[2 * j + 1]);
            reinterpret_cast<uint4*>(d)[0] = make_uint4(w[0], w[1], w[2], w[3]);
            reinterpret_cast<uint4*>(d)[1] = make_uint4(w[4], w[5], w[6], w[7]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) if (ch0 + j < a.Cout) d[j] = from_f<TH>(t[j]);
          }
        }
      }
  }
}

// One warp drains columns [col_begin, col_end) of its 32 accumulator rows.  pix0 = pixel index of this
// lane's row (first replica); valid = row maps to a real output pixel.
// Asynchronous residual prefetch (cp.async, no registers held): the residual of the FIRST column group of a
// FUTURE epilogue_rows call is copied into the per-warp shared-memory buffer `rbuf` (same row pitch as the
// staging buffer) and committed as one cp.async group.  Returns false (warp-uniform) if that call will not use
// the vector path; epilogue_rows then loads the residual itself.
template <typename TH>
__device__ __forceinline__ bool epi_prefetch_async(const EpiArgs& a, int col_begin, int col_end, bool valid, int pix0,
                                                   uint32_t rbuf, int lane) {
  if (col_begin >= col_end || !a.res || a.up != 1 || a.out_f32) return false;
  const int gc = (col_end - col_begin < 64) ? (col_end - col_begin) : 64;
  const int chg = a.ch_base + col_begin;
  if (!(a.vec_ok && (chg + gc <= a.Cout))) return false;
  const unsigned vmask = __ballot_sync(0xffffffffu, valid);
  const int ppr = (gc * 2) >> 4;
  const int lg = ppr <= 2 ? 1 : (ppr <= 4 ? 2 : 3);
  const int lpr = 1 << lg, rpi = 32 >> lg;
  const int sub_r = lane >> lg, sub_c = lane & (lpr - 1);
  const bool lane_on = sub_c < ppr;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    if (it < lpr) {
      const int r = it * rpi + sub_r;
      const int pr = __shfl_sync(0xffffffffu, pix0, r);
      if (lane_on && ((vmask >> r) & 1u)) {
        const void* g = reinterpret_cast<const uint4*>(reinterpret_cast<const TH*>(a.res) + (int64_t)pr * a.res_pitch + chg) + sub_c;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(rbuf + (uint32_t)(r * a.spitch + sub_c * 16)), "l"(g)
                     : "memory");
      }
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  return true;
}

template <typename TH>
__device__ __forceinline__ void epilogue_rows(const EpiArgs& a, uint32_t t_addr, int col_begin, int col_end, bool valid,
                                              int pix0, uint32_t stage, int lane, uint32_t rbuf = 0, int r_ready = 0) {
  // r_ready: 0 = load the residual here; 1 / 2 = the residual of the first column group was prefetched into
  // `rbuf` by epi_prefetch_async and is the last (1) / second-to-last (2) cp.async group this thread committed
  if (col_begin >= col_end) return;   // warp-uniform
  const unsigned vmask = __ballot_sync(0xffffffffu, valid);
  // ---- replicate-on-write fast path (fuse layers: 1x1 conv + BN + nearest x`up` + running sum + ReLU, hrnet.py:99-112)
  // for a single 16-column chunk: the accumulator is read and scaled ONCE, then each of the up*up replicas adds its
  // own residual and is stored; the residual of replica i+1 is in flight (cp.async into the other of two buffers
  // `rbuf`, `rbuf + 32*spitch`) while replica i is processed.  r_ready == -1 selects it (the caller guarantees the
  // two buffers exist).
  if constexpr (sizeof(TH) == 2)
  if (r_ready == -1 && a.up > 1 && a.res && !a.out_f32 && col_end - col_begin == 16 && a.vec_ok &&
      a.ch_base + col_end <= a.Cout) {
    const int chg = a.ch_base + col_begin;
    const uint32_t my_row = stage + (uint32_t)(lane * a.spitch);
    const int sub_r = lane >> 1, sub_c = lane & 1;     // 2 x 16-byte pieces per row, 16 rows per iteration
    auto fetch = [&](int rep, uint32_t buf) {
      const int dy = rep / a.up, dx = rep - dy * a.up;
      const int pix = pix0 + dy * a.Wout + dx;
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int r = it * 16 + sub_r;
        const int pr = __shfl_sync(0xffffffffu, pix, r);
        if ((vmask >> r) & 1u) {
          const void* g = reinterpret_cast<const uint4*>(reinterpret_cast<const TH*>(a.res) + (int64_t)pr * a.res_pitch + chg) + sub_c;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(buf + (uint32_t)(r * a.spitch + sub_c * 16)), "l"(g)
                       : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    fetch(0, rbuf);
    uint32_t v[16];
    tmem_ld16(t_addr + (uint32_t)col_begin, v);
    tmem_ld_wait();
    float o[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 sc = lds128f(a.s_scale + (uint32_t)(chg + 4 * j) * 4u);
      const float4 sh = lds128f(a.s_shift + (uint32_t)(chg + 4 * j) * 4u);
      o[4 * j + 0] = fmaf(__uint_as_float(v[4 * j + 0]), sc.x, sh.x);
      o[4 * j + 1] = fmaf(__uint_as_float(v[4 * j + 1]), sc.y, sh.y);
      o[4 * j + 2] = fmaf(__uint_as_float(v[4 * j + 2]), sc.z, sh.z);
      o[4 * j + 3] = fmaf(__uint_as_float(v[4 * j + 3]), sc.w, sh.w);
    }
    const int nrep = a.up * a.up;
    for (int rep = 0; rep < nrep; ++rep) {
      const uint32_t cur = rbuf + (uint32_t)((rep & 1) * 32 * a.spitch);
      if (rep + 1 < nrep) {
        fetch(rep + 1, rbuf + (uint32_t)(((rep + 1) & 1) * 32 * a.spitch));
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncwarp();
      const uint4 r0 = lds128(cur + (uint32_t)(lane * a.spitch));
      const uint4 r1 = lds128(cur + (uint32_t)(lane * a.spitch + 16));
      const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
      uint32_t w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 t = h2_to_f2<TH>(rw[j]);
        float x0 = o[2 * j] + t.x, x1 = o[2 * j + 1] + t.y;
        if (a.relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
        w[j] = f2_to_h2<TH>(x0, x1);
      }
      sts128(my_row, make_uint4(w[0], w[1], w[2], w[3]));
      sts128(my_row + 16, make_uint4(w[4], w[5], w[6], w[7]));
      __syncwarp();
      const int dy = rep / a.up, dx = rep - dy * a.up;
      const int pix = pix0 + dy * a.Wout + dx;
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int r = it * 16 + sub_r;
        const int pr = __shfl_sync(0xffffffffu, pix, r);
        if ((vmask >> r) & 1u) {
          const uint4 val = lds128(stage + (uint32_t)(r * a.spitch + sub_c * 16));
          uint8_t* dst = reinterpret_cast<uint8_t*>(a.y) + ((int64_t)pr * a.out_pitch + chg) * 2 + sub_c * 16;
          *reinterpret_cast<uint4*>(dst) = val;
        }
      }
      __syncwarp();
    }
    return;
  }
  if (r_ready < 0) r_ready = 0;
  const int esz = a.out_f32 ? 4 : 2;
  const int gmax = a.out_f32 ? 32 : 64;   // columns per staged group (128 B per row)
  const uint32_t my_row = stage + (uint32_t)(lane * a.spitch);
  for (int g0 = col_begin; g0 < col_end; g0 += gmax) {
    const int gc = (col_end - g0 < gmax) ? (col_end - g0) : gmax;
    const int chg = a.ch_base + g0;
    if (chg >= a.Cout) break;   // warp-uniform
    // 16-bit residual with fp32 output has no common staging geometry (heatmap convs never carry a residual);
    // fp32 storage (TH = float, the tf32 arm) stages residual and output with the same 4-byte elements
    const bool grp_vec = a.vec_ok && (chg + gc <= a.Cout) && !(a.res && a.out_f32 && sizeof(TH) == 2);
    // 16-byte pieces per row, mapped with power-of-two lanes per row (no divisions)
    const int ppr = (gc * esz) >> 4;
    const int lg = ppr <= 2 ? 1 : (ppr <= 4 ? 2 : 3);
    const int lpr = 1 << lg;                 // lanes per row
    const int rpi = 32 >> lg;                // rows per iteration
    const int sub_r = lane >> lg, sub_c = lane & (lpr - 1);
    const bool lane_on = sub_c < ppr;
    for (int dy = 0; dy < a.up; ++dy)
      for (int dx = 0; dx < a.up; ++dx) {
        const int pix = pix0 + dy * a.Wout + dx;
        // ---- 1. residual: coalesced global -> staging ------------------------------------------
        uint32_t res_row = my_row;
        if (a.res && grp_vec && r_ready && g0 == col_begin && a.up == 1) {
          if (r_ready == 1) asm volatile("cp.async.wait_group 0;" ::: "memory");
          else asm volatile("cp.async.wait_group 1;" ::: "memory");
          __syncwarp();
          res_row = rbuf + (uint32_t)(lane * a.spitch);
        } else if (a.res && grp_vec) {
          uint4 rreg[8];
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            if (it < lpr) {
              const int r = it * rpi + sub_r;
              const int pr = __shfl_sync(0xffffffffu, pix, r);
              rreg[it] = make_uint4(0, 0, 0, 0);
              if (lane_on && ((vmask >> r) & 1u))
                rreg[it] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const TH*>(a.res) + (int64_t)pr * a.res_pitch + chg) + sub_c);
            }
          }
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            if (it < lpr) {
              const int r = it * rpi + sub_r;
              if (lane_on) sts128(stage + (uint32_t)(r * a.spitch + sub_c * 16), rreg[it]);
            }
          }
          __syncwarp();
        }
        // ---- 2. accumulator chunks: TMEM -> registers -> (+res) -> staging (or direct scalar) -----
        for (int c0 = 0; c0 < gc; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(t_addr + (uint32_t)(g0 + c0), v);
          tmem_ld_wait();
          const int ch0 = chg + c0;
          float o[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 sc = lds128f(a.s_scale + (uint32_t)(ch0 + 4 * j) * 4u);
            const float4 sh = lds128f(a.s_shift + (uint32_t)(ch0 + 4 * j) * 4u);
            o[4 * j + 0] = fmaf(__uint_as_float(v[4 * j + 0]), sc.x, sh.x);
            o[4 * j + 1] = fmaf(__uint_as_float(v[4 * j + 1]), sc.y, sh.y);
            o[4 * j + 2] = fmaf(__uint_as_float(v[4 * j + 2]), sc.z, sh.z);
            o[4 * j + 3] = fmaf(__uint_as_float(v[4 * j + 3]), sc.w, sh.w);
          }
          if (grp_vec) {
            if (a.res) {
              if constexpr (sizeof(TH) == 4) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float4 t = lds128f(res_row + (uint32_t)(c0 * 4 + 16 * j));
                  o[4 * j] += t.x; o[4 * j + 1] += t.y; o[4 * j + 2] += t.z; o[4 * j + 3] += t.w;
                }
              } else {
                const uint4 r0 = lds128(res_row + (uint32_t)(c0 * 2));
                const uint4 r1 = lds128(res_row + (uint32_t)(c0 * 2 + 16));
                const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float2 t = h2_to_f2<TH>(rw[j]);
                  o[2 * j] += t.x;
                  o[2 * j + 1] += t.y;
                }
              }
            }
            if (a.relu) {
#pragma unroll
              for (int j = 0; j < 16; ++j) o[j] = fmaxf(o[j], 0.f);
            }
            if (a.out_f32) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                sts128f(my_row + (uint32_t)(c0 * 4 + 16 * j), make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]));
            } else if constexpr (sizeof(TH) == 2) {
              uint32_t w[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) w[j] = f2_to_h2<TH>(o[2 * j], o[2 * j + 1]);
              sts128(my_row + (uint32_t)(c0 * 2), make_uint4(w[0], w[1], w[2], w[3]));
              sts128(my_row + (uint32_t)(c0 * 2 + 16), make_uint4(w[4], w[5], w[6], w[7]));
            }
          } else if (valid) {
            // ragged tail (Cout not a multiple of the vector width, odd pitches): per-lane scalar path
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int c = ch0 + j;
              if (c < a.Cout) {
                float t = o[j];
                if (a.res) t += to_f<TH>(reinterpret_cast<const TH*>(a.res)[(int64_t)pix * a.res_pitch + c]);
                if (a.relu) t = fmaxf(t, 0.f);
                if (a.out_f32) reinterpret_cast<float*>(a.y)[(int64_t)pix * a.out_pitch + c] = t;
                else reinterpret_cast<TH*>(a.y)[(int64_t)pix * a.out_pitch + c] = from_f<TH>(t);
              }
            }
          }
        }
        // ---- 3. staging -> global, coalesced ----------------------------------------------------
        if (grp_vec) {
          __syncwarp();
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            if (it < lpr) {
              const int r = it * rpi + sub_r;
              const int pr = __shfl_sync(0xffffffffu, pix, r);
              if (lane_on && ((vmask >> r) & 1u)) {
                const uint4 val = lds128(stage + (uint32_t)(r * a.spitch + sub_c * 16));
                uint8_t* dst = reinterpret_cast<uint8_t*>(a.y) + ((int64_t)pr * a.out_pitch + chg) * esz + sub_c * 16;
                *reinterpret_cast<uint4*>(dst) = val;
              }
            }
          }
          __syncwarp();
        }
      }
  }
}

// Pipelined residual epilogue (up == 1, 16-bit output, vector path): the warp's column range is drained in groups of
// 32 columns; while group g is processed, the residual of the next unit -- group g+1 of this tile, or group 0 of the
// warp's next tile -- is already in flight (cp.async into the other of two per-warp buffers).  Only the very first
// unit of a warp pays the global-load latency.  `sel` (buffer parity) and `primed` (the first group of this call has
// been prefetched by the previous call) carry the state across calls.
constexpr int kPipeCols = 32;
__host__ __device__ inline int epi_pipe_pitch(int gcols = kPipeCols) { return gcols * 2 + 16; }
__host__ __device__ inline bool epi_pipe_ok(int BN, int Cout_total, int vec_ok, int out_f32, int up, bool has_res,
                                            bool f32_storage = false) {
  (void)has_res;   // with and without residual (kRes)
  // out_f32 with 16-bit storage (heatmap / offset convs) takes the generic routine; with fp32 storage (tf32 arm) the
  // fp32 twin epilogue_rows_pipelined_f32
  return up == 1 && (f32_storage ? out_f32 != 0 : !out_f32) && vec_ok && (Cout_total % 16 == 0) && BN % 16 == 0;
}
template <typename TH>
__device__ __forceinline__ void epi_pipe_fetch(const EpiArgs& a, int chg, int gc, unsigned vmask, int pix, uint32_t buf, int lane,
                                               int pitch) {
  const int ppr = (gc * (int)sizeof(TH)) >> 4;   // 16-byte pieces per row: 2 or 4
  const int lg = ppr <= 2 ? 1 : 2;
  const int lpr = 1 << lg, rpi = 32 >> lg;
  const int sub_r = lane >> lg, sub_c = lane & (lpr - 1);
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    if (it < lpr) {
      const int r = it * rpi + sub_r;
      const int pr = __shfl_sync(0xffffffffu, pix, r);
      if ((vmask >> r) & 1u) {
        const void* g = reinterpret_cast<const uint4*>(reinterpret_cast<const TH*>(a.res) + (int64_t)pr * a.res_pitch + chg) + sub_c;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(buf + (uint32_t)(r * pitch + sub_c * 16)), "l"(g) : "memory");
      }
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

// fp32-storage twin of epilogue_rows_pipelined (the tf32 arm: float residual, float output): groups of 8 columns =
// 32 bytes per row, row pitch 48 B -- the same per-warp footprint (staging + two residual buffers = 3 x 32 x 48 B) as
// the 16-bit routine at 16 columns, so both share one shared-memory allocation.  While group g is drained the residual
// of the next unit (next group / next M-tile / next CTA tile) is in flight (cp.async).
constexpr int kPipeColsF32 = 8;
template <bool kRes>
__device__ __forceinline__ void epilogue_rows_pipelined_f32(const EpiArgs& a, uint32_t t_addr, int col_begin, int col_end, bool valid,
                                                            int pix0, uint32_t stage, uint32_t rb0, uint32_t rb1, int lane, int& sel,
                                                            int& primed, bool have_next, bool next_valid, int next_pix,
                                                            int next_ch_base) {
  if (col_begin >= col_end) return;   // warp-uniform
  constexpr int gcols = kPipeColsF32;
  constexpr int pitch = gcols * 4 + 16;   // 48
  const unsigned vmask = __ballot_sync(0xffffffffu, valid);
  const unsigned nmask = __ballot_sync(0xffffffffu, next_valid);
  const uint32_t my_row = stage + (uint32_t)(lane * pitch);
  const int sub_r = lane >> 1, sub_c = lane & 1;   // 2 x 16-byte pieces per row, 16 rows per iteration
  if (kRes && !primed) epi_pipe_fetch<float>(a, a.ch_base + col_begin, gcols, vmask, pix0, sel ? rb1 : rb0, lane, pitch);
  for (int g0 = col_begin; g0 < col_end; g0 += gcols) {
    const int chg = a.ch_base + g0;
    const uint32_t cur = sel ? rb1 : rb0, nxt = sel ? rb0 : rb1;
    bool pending = false;
    if (kRes) {
      if (g0 + gcols < col_end) {
        epi_pipe_fetch<float>(a, chg + gcols, gcols, vmask, pix0, nxt, lane, pitch);
        pending = true;
      } else if (have_next) {
        epi_pipe_fetch<float>(a, next_ch_base + col_begin, gcols, nmask, next_pix, nxt, lane, pitch);
        pending = true;
      }
      if (pending) asm volatile("cp.async.wait_group 1;" ::: "memory");
      else asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
    }
    uint32_t v[8];
    tmem_ld8(t_addr + (uint32_t)g0, v);
    tmem_ld_wait();
    const float4 sc0 = lds128f(a.s_scale + (uint32_t)chg * 4u), sc1 = lds128f(a.s_scale + (uint32_t)(chg + 4) * 4u);
    const float4 sh0 = lds128f(a.s_shift + (uint32_t)chg * 4u), sh1 = lds128f(a.s_shift + (uint32_t)(chg + 4) * 4u);
    float4 o0, o1;
    o0.x = fmaf(__uint_as_float(v[0]), sc0.x, sh0.x); o0.y = fmaf(__uint_as_float(v[1]), sc0.y, sh0.y);
    o0.z = fmaf(__uint_as_float(v[2]), sc0.z, sh0.z); o0.w = fmaf(__uint_as_float(v[3]), sc0.w, sh0.w);
    o1.x = fmaf(__uint_as_float(v[4]), sc1.x, sh1.x); o1.y = fmaf(__uint_as_float(v[5]), sc1.y, sh1.y);
    o1.z = fmaf(__uint_as_float(v[6]), sc1.z, sh1.z); o1.w = fmaf(__uint_as_float(v[7]), sc1.w, sh1.w);
    if (kRes) {
      const uint32_t res_row = cur + (uint32_t)(lane * pitch);
      const float4 r0 = lds128f(res_row), r1 = lds128f(res_row + 16);
      o0.x += r0.x; o0.y += r0.y; o0.z += r0.z; o0.w += r0.w;
      o1.x += r1.x; o1.y += r1.y; o1.z += r1.z; o1.w += r1.w;
    }
    if (a.relu) {
      o0.x = fmaxf(o0.x, 0.f); o0.y = fmaxf(o0.y, 0.f); o0.z = fmaxf(o0.z, 0.f); o0.w = fmaxf(o0.w, 0.f);
      o1.x = fmaxf(o1.x, 0.f); o1.y = fmaxf(o1.y, 0.f); o1.z = fmaxf(o1.z, 0.f); o1.w = fmaxf(o1.w, 0.f);
    }
    sts128f(my_row, o0);
    sts128f(my_row + 16, o1);
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int r = it * 16 + sub_r;
      const int pr = __shfl_sync(0xffffffffu, pix0, r);
      if ((vmask >> r) & 1u) {
        const uint4 val = lds128(stage + (uint32_t)(r * pitch + sub_c * 16));
        uint8_t* dst = reinterpret_cast<uint8_t*>(a.y) + ((int64_t)pr * a.out_pitch + chg) * 4 + sub_c * 16;
        *reinterpret_cast<uint4*>(dst) = val;
      }
    }
    __syncwarp();
    sel ^= 1;
    primed = pending ? 1 : 0;
  }
}
// fp32-residual-stream twin of the pipelined routines (16-bit arms with an fp32 residual stream,
// fami_conv2d_bn_act_fwd_stream): 8-column groups as in epilogue_rows_pipelined_f32.  The residual comes from the FLOAT
// tensor a.res32 (cp.async pipeline, kRes), the result goes -- before rounding -- to the float tensor a.y32 when there is
// one (staged, coalesced 16-byte pieces) and, rounded, to the 16-bit tensor a.y (each lane stores the 16 bytes of its
// row's group itself).  Replaces the unstaged epilogue_rows_stream wherever the tile has whole 8-column groups and 16-byte
// aligned rows (incl. the replicate-on-write fuse layers); that routine keeps the ragged cases.
template <typename TH, bool kRes>
__device__ __forceinline__ void epilogue_rows_pipelined_stream(const EpiArgs& a0, uint32_t t_addr, int col_begin, int col_end,
                                                               bool valid, int pix0, uint32_t stage, uint32_t rb0, uint32_t rb1,
                                                               int lane, int& sel, int& primed, bool have_next, bool next_valid,
                                                               int next_pix, int next_ch_base) {
  if (col_begin >= col_end) return;   // warp-uniform
  constexpr int gcols = kPipeColsF32;
  constexpr int pitch = gcols * 4 + 16;   // 48
  EpiArgs a = a0;
  a.res = a0.res32;                        // epi_pipe_fetch<float> reads a.res / a.res_pitch
  a.res_pitch = a0.res32_pitch;
  const unsigned vmask = __ballot_sync(0xffffffffu, valid);
  const unsigned nmask = __ballot_sync(0xffffffffu, next_valid);
  const uint32_t my_row = stage + (uint32_t)(lane * pitch);
  const int sub_r = lane >> 1, sub_c = lane & 1;   // 2 x 16-byte pieces per row, 16 rows per iteration
  // nearest-upsample replication (the HighResolutionModule fuse layers): the accumulator group is read and scaled once,
  // every one of the up*up replicas adds its own residual and is stored; the pipeline unit is (column group, replica)
  const int nrep = a.up * a.up;
  auto rep_pix = [&](int rep) { const int dy = rep / a.up; return pix0 + dy * a.Wout + (rep - dy * a.up); };
  if (kRes && !primed) epi_pipe_fetch<float>(a, a.ch_base + col_begin, gcols, vmask, pix0, sel ? rb1 : rb0, lane, pitch);
  for (int g0 = col_begin; g0 < col_end; g0 += gcols) {
    const int chg = a.ch_base + g0;
    uint32_t v[8];
    tmem_ld8(t_addr + (uint32_t)g0, v);
    tmem_ld_wait();
    const float4 sc0 = lds128f(a.s_scale + (uint32_t)chg * 4u), sc1 = lds128f(a.s_scale + (uint32_t)(chg + 4) * 4u);
    const float4 sh0 = lds128f(a.s_shift + (uint32_t)chg * 4u), sh1 = lds128f(a.s_shift + (uint32_t)(chg + 4) * 4u);
    float4 b0, b1;
    b0.x = fmaf(__uint_as_float(v[0]), sc0.x, sh0.x); b0.y = fmaf(__uint_as_float(v[1]), sc0.y, sh0.y);
    b0.z = fmaf(__uint_as_float(v[2]), sc0.z, sh0.z); b0.w = fmaf(__uint_as_float(v[3]), sc0.w, sh0.w);
    b1.x = fmaf(__uint_as_float(v[4]), sc1.x, sh1.x); b1.y = fmaf(__uint_as_float(v[5]), sc1.y, sh1.y);
    b1.z = fmaf(__uint_as_float(v[6]), sc1.z, sh1.z); b1.w = fmaf(__uint_as_float(v[7]), sc1.w, sh1.w);
    for (int rep = 0; rep < nrep; ++rep) {          // warp-uniform
      const int pix = rep_pix(rep);
      const uint32_t cur = sel ? rb1 : rb0, nxt = sel ? rb0 : rb1;
      bool pending = false;
      if (kRes) {
        if (rep + 1 < nrep) {
          epi_pipe_fetch<float>(a, chg, gcols, vmask, rep_pix(rep + 1), nxt, lane, pitch);
          pending = true;
        } else if (g0 + gcols < col_end) {
          epi_pipe_fetch<float>(a, chg + gcols, gcols, vmask, pix0, nxt, lane, pitch);
          pending = true;
        } else if (have_next) {
          epi_pipe_fetch<float>(a, next_ch_base + col_begin, gcols, nmask, next_pix, nxt, lane, pitch);
          pending = true;
        }
        if (pending) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
      }
      float4 o0 = b0, o1 = b1;
      if (kRes) {
        const uint32_t res_row = cur + (uint32_t)(lane * pitch);
        const float4 r0 = lds128f(res_row), r1 = lds128f(res_row + 16);
        o0.x += r0.x; o0.y += r0.y; o0.z += r0.z; o0.w += r0.w;
        o1.x += r1.x; o1.y += r1.y; o1.z += r1.z; o1.w += r1.w;
      }
      if (a.relu) {
        o0.x = fmaxf(o0.x, 0.f); o0.y = fmaxf(o0.y, 0.f); o0.z = fmaxf(o0.z, 0.f); o0.w = fmaxf(o0.w, 0.f);
        o1.x = fmaxf(o1.x, 0.f); o1.y = fmaxf(o1.y, 0.f); o1.z = fmaxf(o1.z, 0.f); o1.w = fmaxf(o1.w, 0.f);
      }
      // the 16-bit operand twin: this lane's 8 channels = 16 bytes of its pixel row
      if (valid) {
        uint4 h;
        h.x = f2_to_h2<TH>(o0.x, o0.y); h.y = f2_to_h2<TH>(o0.z, o0.w); h.z = f2_to_h2<TH>(o1.x, o1.y); h.w = f2_to_h2<TH>(o1.z, o1.w);
        *reinterpret_cast<uint4*>(reinterpret_cast<TH*>(a.y) + (int64_t)pix * a.out_pitch + chg) = h;
      }
      // the fp32 stream: staged, coalesced 16-byte pieces
      if (a.y32) {
        sts128f(my_row, o0);
        sts128f(my_row + 16, o1);
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          const int r = it * 16 + sub_r;
          const int pr = __shfl_sync(0xffffffffu, pix, r);
          if ((vmask >> r) & 1u) {
            const uint4 val = lds128(stage + (uint32_t)(r * pitch + sub_c * 16));
            uint8_t* dst = reinterpret_cast<uint8_t*>(a.y32) + ((int64_t)pr * a.y32_pitch + chg) * 4 + sub_c * 16;
            *reinterpret_cast<uint4*>(dst) = val;
          }
        }
        __syncwarp();
      }
      if (kRes) {
        sel ^= 1;
        primed = pending ? 1 : 0;
      }
    }
  }
}
// host / device: may a stream-mode tile take epilogue_rows_pipelined_stream?
__host__ __device__ inline bool epi_stream_pipe_ok(int BN, int Cout_total, int up, int out_f32, int out_pitch, const void* y,
                                                   const float* res32, int res32_pitch, const float* y32, int y32_pitch) {
  return (up == 1 || up == 2 || up == 4 || up == 8) && !out_f32 && BN % 8 == 0 && Cout_total % 8 == 0 && out_pitch % 8 == 0 &&
         ((uintptr_t)y & 15) == 0 &&
         (!res32 || (res32_pitch % 4 == 0 && ((uintptr_t)res32 & 15) == 0)) && (!y32 || (y32_pitch % 4 == 0 && ((uintptr_t)y32 & 15) == 0));
}
template <typename TH, bool kRes = true>
__device__ __forceinline__ void epilogue_rows_pipelined(const EpiArgs& a, uint32_t t_addr, int col_begin, int col_end, bool valid,
                                                        int pix0, uint32_t stage, uint32_t rb0, uint32_t rb1, int lane, int& sel,
                                                        int& primed, bool have_next, bool next_valid, int next_pix,
                                                        int next_ch_base, int gcols = kPipeCols) {
  if (col_begin >= col_end) return;   // warp-uniform
  const unsigned vmask = __ballot_sync(0xffffffffu, valid);
  const unsigned nmask = __ballot_sync(0xffffffffu, next_valid);
  const int pitch = epi_pipe_pitch(gcols);
  const uint32_t my_row = stage + (uint32_t)(lane * pitch);
  if (kRes && !primed) {   // very first unit of this warp: nobody prefetched it
    const int gc0 = (col_end - col_begin < gcols) ? (col_end - col_begin) : gcols;
    epi_pipe_fetch<TH>(a, a.ch_base + col_begin, gc0, vmask, pix0, sel ? rb1 : rb0, lane, pitch);
  }
  for (int g0 = col_begin; g0 < col_end; g0 += gcols) {
    const int gc = (col_end - g0 < gcols) ? (col_end - g0) : gcols;
    const int chg = a.ch_base + g0;
    const uint32_t cur = sel ? rb1 : rb0, nxt = sel ? rb0 : rb1;
    // next unit in flight
    bool pending = false;
    if (!kRes) {
      // no residual: nothing to fetch
    } else if (g0 + gcols < col_end) {
      const int ngc = (col_end - (g0 + gcols) < gcols) ? (col_end - (g0 + gcols)) : gcols;
      epi_pipe_fetch<TH>(a, chg + gcols, ngc, vmask, pix0, nxt, lane, pitch);
      pending = true;
    } else if (have_next) {
      const int ngc = (col_end - col_begin < gcols) ? (col_end - col_begin) : gcols;
      epi_pipe_fetch<TH>(a, next_ch_base + col_begin, ngc, nmask, next_pix, nxt, lane, pitch);
      pending = true;
    }
    if (kRes) {
      if (pending) asm volatile("cp.async.wait_group 1;" ::: "memory");
      else asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
    }
    const uint32_t res_row = cur + (uint32_t)(lane * pitch);
    for (int c0 = 0; c0 < gc; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(t_addr + (uint32_t)(g0 + c0), v);
      tmem_ld_wait();
      const int ch0 = chg + c0;
      float o[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 sc = lds128f(a.s_scale + (uint32_t)(ch0 + 4 * j) * 4u);
        const float4 sh = lds128f(a.s_shift + (uint32_t)(ch0 + 4 * j) * 4u);
        o[4 * j + 0] = fmaf(__uint_as_float(v[4 * j + 0]), sc.x, sh.x);
        o[4 * j + 1] = fmaf(__uint_as_float(v[4 * j + 1]), sc.y, sh.y);
        o[4 * j + 2] = fmaf(__uint_as_float(v[4 * j + 2]), sc.z, sh.z);
        o[4 * j + 3] = fmaf(__uint_as_float(v[4 * j + 3]), sc.w, sh.w);
      }
      uint32_t rw[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
      if (kRes) {
        const uint4 r0 = lds128(res_row + (uint32_t)(c0 * 2));
        const uint4 r1 = lds128(res_row + (uint32_t)(c0 * 2 + 16));
        rw[0] = r0.x; rw[1] = r0.y; rw[2] = r0.z; rw[3] = r0.w; rw[4] = r1.x; rw[5] = r1.y; rw[6] = r1.z; rw[7] = r1.w;
      }
      uint32_t w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float x0 = o[2 * j], x1 = o[2 * j + 1];
        if (kRes) {
          const float2 t = h2_to_f2<TH>(rw[j]);
          x0 += t.x; x1 += t.y;
        }
        if (a.relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
        w[j] = f2_to_h2<TH>(x0, x1);
      }
      sts128(my_row + (uint32_t)(c0 * 2), make_uint4(w[0], w[1], w[2], w[3]));
      sts128(my_row + (uint32_t)(c0 * 2 + 16), make_uint4(w[4], w[5], w[6], w[7]));
    }
    __syncwarp();
    {
      const int ppr = (gc * 2) >> 4;
      const int lg = ppr <= 2 ? 1 : 2;
      const int lpr = 1 << lg, rpi = 32 >> lg;
      const int sub_r = lane >> lg, sub_c = lane & (lpr - 1);
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        if (it < lpr) {
          const int r = it * rpi + sub_r;
          const int pr = __shfl_sync(0xffffffffu, pix0, r);
          if ((vmask >> r) & 1u) {
            const uint4 val = lds128(stage + (uint32_t)(r * pitch + sub_c * 16));
            uint8_t* dst = reinterpret_cast<uint8_t*>(a.y) + ((int64_t)pr * a.out_pitch + chg) * 2 + sub_c * 16;
            *reinterpret_cast<uint4*>(dst) = val;
          }
        }
      }
    }
    __syncwarp();
    sel ^= 1;
    primed = pending ? 1 : 0;
  }
}

// Epilogue of the fused offset|mask producer convolution writing the row-blocked layout the deformable kernel reads
// (fami_dcn_desc.om_layout = 2): [tap][tile][row 16][dy | dx | mask][pixel 8][group G] over 16x8-pixel tiles.  A gather warp of
// the deformable kernel owns one tile row; its lanes walk the 8*G samples (pixel, group) of a (row, tap) 32 at a time, so
// every load instruction of the warp reads 128 contiguous bytes and the warp's reads of a (row, tap) are one contiguous run
// of 24*G floats.  Channel n = tap*3G + k*G + g (k = dy, dx, mask), so the four channels of a float4 never straddle a
// (tap, k) run; stored straight from registers, no staging.
struct OmBlocked {
  float* base;
  int tiles_x, tiles_y, G3;     // DCN tiles per image, 3*G
  int64_t tap_stride;           // floats between taps
  int kblocked;                 // 1: k-step-blocked runs [group / 4][pixel 8][group % 4] (fami_dcn_desc.om_layout 3)
};
__device__ __forceinline__ void epilogue_rows_om_blocked(const EpiArgs& a, uint32_t t_addr, int col_begin, int col_end, bool valid,
                                                         int img, int y, int x, const OmBlocked& ob) {
  if (col_begin >= col_end) return;   // warp-uniform
  const int G = ob.G3 / 3;
  const int ry = y & 15, rx = x & 7;
  const int64_t tile = (int64_t)(img * ob.tiles_y + (y >> 4)) * ob.tiles_x + (x >> 3);
  // + k * 8G + g per (dy | dx | mask) run
  // + k * 8G + g per (dy | dx | mask) run; k-step-blocked: + k * 8G + (g / 4) * 32 (g is a multiple of 4 here)
  // (layout 3 is tile-major: the nine taps of a tile are contiguous, ob.tap_stride = 128 * 3G)
  float* lane_base = ob.base + tile * (int64_t)(128 * ob.G3) * (ob.kblocked ? 9 : 1) + (int64_t)(ry * 3) * (8 * G) + (ob.kblocked ? rx * 4 : rx * G);
  const int kstride = 8 * G, gmul = ob.kblocked ? 8 : 1;
  for (int c0 = col_begin; c0 < col_end; c0 += 16) {
    const int ch0 = a.ch_base + c0;
    if (ch0 >= a.Cout) break;   // warp-uniform
    uint32_t v[16];
    tmem_ld16(t_addr + (uint32_t)c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c4 = ch0 + 4 * q;
      if (c4 < a.Cout) {          // warp-uniform
        const float4 sc = lds128f(a.s_scale + (uint32_t)c4 * 4u);
        const float4 sh = lds128f(a.s_shift + (uint32_t)c4 * 4u);
        const int tap = c4 / ob.G3, f = c4 - tap * ob.G3;
        const int k = f / G, g0 = f - k * G;
        if (valid) {
          float4 o;
          o.x = fmaf(__uint_as_float(v[4 * q + 0]), sc.x, sh.x);
          o.y = fmaf(__uint_as_float(v[4 * q + 1]), sc.y, sh.y);
          o.z = fmaf(__uint_as_float(v[4 * q + 2]), sc.z, sh.z);
          o.w = fmaf(__uint_as_float(v[4 * q + 3]), sc.w, sh.w);
          *reinterpret_cast<float4*>(lane_base + (int64_t)tap * ob.tap_stride + k * kstride + g0 * gmul) = o;
        }
      }
    }
  }
}

// ---- driver entry points for tensor-map encoding (resolved at run time, no -lcuda) ------------
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// tensor-map element type of an activation operand.  FAMI_TF32: the TMA unit converts fp32 -> tf32 while it copies
// (CU_TENSOR_MAP_DATA_TYPE_TFLOAT32; round-to-nearest measured by tools/probe_tma_tf32.py), so activations stay
// full fp32 in HBM -- the residual stream keeps all 24 bits -- and only the multiplicand is rounded, as cuDNN's
// TF32 convolutions do.  FAMI_TMA_TF32=0 selects a plain FLOAT32 copy (the tensor core then truncates).
inline CUtensorMapDataType tm_dtype_of(int dtype) {
  if (dtype == FAMI_TF32) {
    static const bool plain = getenv("FAMI_TMA_TF32") != nullptr && atoi(getenv("FAMI_TMA_TF32")) == 0;
    return plain ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32;
  }
  return dtype == FAMI_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
}

inline EncodeIm2colFn g_encode_im2col = nullptr;
inline EncodeTiledFn g_encode_tiled = nullptr;

inline bool load_driver_fns() {
  if (g_encode_im2col && g_encode_tiled) return true;
  void* f1 = nullptr;
  void* f2 = nullptr;
  cudaDriverEntryPointQueryResult q1, q2;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f1, cudaEnableDefault, &q1) != cudaSuccess ||
      q1 != cudaDriverEntryPointSuccess)
    return false;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f2, cudaEnableDefault, &q2) != cudaSuccess ||
      q2 != cudaDriverEntryPointSuccess)
    return false;
  g_encode_im2col = reinterpret_cast<EncodeIm2colFn>(f1);
  g_encode_tiled = reinterpret_cast<EncodeTiledFn>(f2);
  return true;
}


}  // namespace fami
