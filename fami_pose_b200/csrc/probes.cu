// Hardware probes (test / tooling infrastructure, NOT part of the product path): compiled into the library only
// with -DFAMI_DEBUG_PROBES (python fami_pose_b200/csrc/build.py --probes -> libfami_b200_probes.so), together with the
// fami_debug_* exports of api.cu.  tools/probe_*.py drive them.
#ifdef FAMI_DEBUG_PROBES
#include "tc_common.cuh"

namespace fami {

// ------------------------------------------------------------------------------------------------
// Hardware probe (test infrastructure): does a K-major SWIZZLE_128B UMMA descriptor whose start
// address is shifted by `shift` 128-byte rows inside a TMA-written tile address rows
// [shift, shift+128)?  mode 0: base_offset field = 0; mode 1: base_offset = (addr >> 7) & 7.
// x: f16 [R][64], w: f16 [16][64], out: f32 [128][16] = x[shift:shift+128] @ w^T.
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(128, 1)
umma_rowshift_probe(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, float* out,
                    int R, int shift, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sX = smem;                    // R rows x 128 B (R multiple of 8)
  uint8_t* sW = smem + (size_t)R * 128;  // 16 rows x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + 16 * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const uint32_t bar_load = smem_u32(bars), bar_mma = smem_u32(bars + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_load, (uint32_t)(R * 128 + 16 * 128));
    for (int r0 = 0; r0 < R; r0 += 8) tma_tiled_2d(smem_u32(sX + (size_t)r0 * 128), &tmX, bar_load, 0, r0);
    tma_tiled_2d(smem_u32(sW), &tmW, bar_load, 0, 0);
    mbar_wait(bar_load, 0);
    tc_fence_after();
    const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_addr = smem_u32(sX) + (uint32_t)shift * 128u;
    uint64_t adesc = make_sw128_desc(a_addr);
    if (mode == 1) adesc |= (uint64_t)((a_addr >> 7) & 7u) << 49;
    const uint64_t bdesc = make_sw128_desc(smem_u32(sW));
    for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k ? 1u : 0u);
    umma_commit(bar_mma);
  }
  mbar_wait(bar_mma, 0);
  tc_fence_after();
  uint32_t v[16];
  tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  const int row = warp * 32 + lane;
  for (int j = 0; j < 16; ++j) out[row * 16 + j] = __uint_as_float(v[j]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(32u) : "memory");
  }
}
}  // namespace

int debug_umma_rowshift_launch(const void* x, const void* w, float* out, int R, int shift, int mode, cudaStream_t st) {
  FAMI_CHECK_ARG(load_driver_fns(), "driver entry points unavailable");
  FAMI_CHECK_ARG(R % 8 == 0 && R >= 128 + shift + 8 && R <= 1024, "bad R");
  CUtensorMap tmX, tmW;
  {
    cuuint64_t dims[2] = {64, (cuuint64_t)R};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, 8};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(x), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FAMI_CHECK_ARG(r == CUDA_SUCCESS, "encode X failed %d", (int)r);
  }
  {
    cuuint64_t dims[2] = {64, 16};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, 16};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FAMI_CHECK_ARG(r == CUDA_SUCCESS, "encode W failed %d", (int)r);
  }
  size_t smem = (size_t)R * 128 + 16 * 128 + 1024 + 64;
  cudaFuncSetAttribute(umma_rowshift_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  umma_rowshift_probe<<<1, 128, smem, st>>>(tmX, tmW, out, R, shift, mode);
  FAMI_CHECK_LAUNCH("umma_rowshift_probe");
  return 0;
}


// ------------------------------------------------------------------------------------------------
// Hardware probe: back-to-back tcgen05.mma throughput (M=128, K=16, f16) as a function of N and of
// how the issuing thread builds its descriptors.  out[0] = clock cycles for `iters` MMAs (clock64 around
// issue + final commit wait), out[1] = cycles for the issue loop alone.
// variant 0: constant descriptors; 1: descriptors recomputed per MMA from a rotating row shift (as
// the halo conv does); 2: as 0 but 4 different accumulator column offsets round-robin.
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(512, 1) umma_rate_probe(long long* out, int N, int iters, int variant) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                  // 1024 rows x 128 B
  uint8_t* sB = smem + 1024 * 128;     // 256 rows x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 256 * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const uint32_t bar_mma = smem_u32(bars);
  for (int i = threadIdx.x; i < (1024 + 256) * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0h
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) {
    // warp-uniform control flow; a single elected lane issues (variant >= 3) or lane 0 in divergent code
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    const uint64_t adesc0 = make_sw128_desc(a0), bdesc0 = make_sw128_desc(b0);
    long long t0 = clock64(), t1 = 0;
    if (variant >= 3) {
      const bool leader = elect_one();
      for (int i = 0; i < iters; ++i) {
        uint64_t ad = adesc0, bd = bdesc0;
        uint32_t dt = tm;
        if (variant >= 6) {
          // mimic the halo conv issue pattern: tap -> m (3 accumulators) -> k (3 K-steps)
          const int k = i % 3, m = (i / 3) % 3, tap = (i / 9) % 9;
          const int fr = tap / 3, fs = tap - fr * 3;
          ad = make_sw128_desc(a0 + (uint32_t)(m * 128 + fr * 74 + fs) * 128u) + (uint64_t)(2 * k);
          bd = make_sw128_desc(b0 + (uint32_t)(variant == 8 ? 0 : tap) * (uint32_t)N * 128u % (256u * 128u)) + (uint64_t)(2 * k);
          dt = tm + (uint32_t)(m * (variant == 7 ? 64 : N));
        }
        if (variant == 4) {
          const uint32_t shift = (uint32_t)((i * 37) & 511);
          ad = make_sw128_desc(a0 + shift * 128u) + (uint64_t)(2 * (i & 3));
          bd = bdesc0 + (uint64_t)(2 * (i & 3));
        }
        if (variant == 5) dt = tm + (uint32_t)((i & 1) * 256);
        if (leader) umma_bf16(dt, ad, bd, idesc, i > 1 ? 1u : 0u);
      }
      t1 = clock64();
      if (leader) umma_commit(bar_mma);
    } else if (threadIdx.x == 0) {
      for (int i = 0; i < iters; ++i) {
        if (variant == 1) {
          const uint32_t shift = (uint32_t)((i * 37) & 511);
          const uint64_t ad = make_sw128_desc(a0 + shift * 128u) + (uint64_t)(2 * (i & 3));
          umma_bf16(tmem_base, ad, bdesc0 + (uint64_t)(2 * (i & 3)), idesc, i ? 1u : 0u);
        } else if (variant == 2) {
          umma_bf16(tmem_base + (uint32_t)((i & 1) * 256), adesc0, bdesc0, idesc, i > 1 ? 1u : 0u);
        } else {
          umma_bf16(tmem_base, adesc0, bdesc0, idesc, i ? 1u : 0u);
        }
      }
      t1 = clock64();
      umma_commit(bar_mma);
    }
    __syncwarp();
    mbar_wait(bar_mma, 0);
    long long t2 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
      out[0] = t2 - t0;
      out[1] = t1 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}
}  // namespace

int debug_umma_rate_launch(long long* out, int N, int iters, int variant, cudaStream_t st) {
  FAMI_CHECK_ARG(N >= 16 && N <= 256 && N % 16 == 0 && iters > 0, "bad N/iters");
  size_t smem = (size_t)(1024 + 256) * 128 + 1024 + 64;
  cudaFuncSetAttribute(umma_rate_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int grid = (variant % 200) >= 100 ? 148 : 1;   // variant + 100: all SMs run the probe concurrently (CTA 0 reports)
  const int threads = variant >= 200 ? 512 : 128;      // variant + 200: 15 more warps spinning on the completion barrier
  umma_rate_probe<<<grid, threads, smem, st>>>(out, N, iters, variant % 100);
  FAMI_CHECK_LAUNCH("umma_rate_probe");
  return 0;
}


// ------------------------------------------------------------------------------------------------
// Hardware probe: what does a TMA load do to fp32 data when the tensor map's element type is
// CU_TENSOR_MAP_DATA_TYPE_TFLOAT32?  Loads [rows][32] floats (one 128-byte row each, no swizzle) and
// writes the shared-memory bit patterns back.  mode 0: FLOAT32 map, 1: TFLOAT32, 2: TFLOAT32_FTZ.
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(128, 1) tma_tf32_probe(const __grid_constant__ CUtensorMap tm, uint32_t* out, int rows) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)rows * 128);
  const uint32_t b = smem_u32(bar);
  if (threadIdx.x == 0) {
    mbar_init(b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(b, (uint32_t)rows * 128u);
    tma_tiled_2d(smem_u32(smem), &tm, b, 0, 0);
  }
  mbar_wait(b, 0);
  for (int i = threadIdx.x; i < rows * 32; i += blockDim.x) out[i] = reinterpret_cast<const uint32_t*>(smem)[i];
}
}  // namespace

int debug_tma_tf32_launch(const float* x, uint32_t* out, int rows, int mode, cudaStream_t st) {
  FAMI_CHECK_ARG(load_driver_fns(), "driver entry points unavailable");
  FAMI_CHECK_ARG(rows > 0 && rows <= 256, "bad rows");
  CUtensorMap tm;
  cuuint64_t dims[2] = {32, (cuuint64_t)rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {32, (cuuint32_t)rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = mode == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                           : (mode == 1 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32_FTZ);
  CUresult r = g_encode_tiled(&tm, dt, 2, const_cast<float*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FAMI_CHECK_ARG(r == CUDA_SUCCESS, "encode failed %d", (int)r);
  tma_tf32_probe<<<1, 128, (size_t)rows * 128 + 1024 + 64, st>>>(tm, out, rows);
  FAMI_CHECK_LAUNCH("tma_tf32_probe");
  return 0;
}

}  // namespace fami
#endif  // FAMI_DEBUG_PROBES
