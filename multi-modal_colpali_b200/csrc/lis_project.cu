// K3: fused retrieval head for ingestion.  Restates what the reference runs at the end of
// model(**batch) (functions.py:795, 839, 888; 05_experiment02.py:211; body HF
// modeling_colpali.py:148-155):
//     e = Linear(hidden -> 128)(h);  e = e / ||e||_2 (no epsilon);  e = e * attention_mask
// as ONE kernel: TMA streams 128-token x 64-feature blocks of h and W, tcgen05 accumulates the
// [128 tokens x 128 dims] tile in TMEM, and the epilogue (thread = token) adds the bias, takes the
// row norm, scales, masks and writes the 16-bit row that goes straight into the page store.
#include <algorithm>
#include <cstdlib>

#include "lis_common.h"
#include "lis_ptx.cuh"

namespace lis {

constexpr int kPTile = 128;                     // tokens per tile (UMMA M)
constexpr int kPOut = 128;                      // output dims (UMMA N)
constexpr int kPBlockBytes = kPTile * 128;      // one 64-feature block of A or W: 16 KB
constexpr int kPThreads = 192;
// SUB = token sub-tiles (of 128) that share one load of every weight block: 1 -> a stage is 16 KB of h + 16 KB of W,
// 2 -> 32 KB of h + 16 KB of W (the weight stream, which comes from L2 once per tile, is halved per token).
__host__ __device__ constexpr int p_stage_bytes(int sub) { return (sub + 1) * kPBlockBytes; }

struct ProjectArgs {
  const void* bias;     // [128] 16-bit or null
  const void* mask;     // [n_tok] integers of mask_size bytes (1, 4 or 8), or null
  int32_t mask_size;
  void* out;            // [n_tok, 128] 16-bit
  int64_t n_tok;
  int32_t kblocks;      // hidden_dim / 64
  int32_t is_bf16;
  int32_t round_ref;    // 1 = round the Linear output, the norm and the quotient to the 16-bit dtype like the reference model
  const int32_t* dst_row;  // [n_tok] destination row of each token in `out`, < 0 = drop; null = identity
};

__device__ __forceinline__ float load16(const void* p, int i, int is_bf16) {
  return is_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i])
                 : __half2float(reinterpret_cast<const __half*>(p)[i]);
}
__device__ __forceinline__ float round16(float x, int is_bf16) {
  return is_bf16 ? __bfloat162float(__float2bfloat16_rn(x)) : __half2float(__float2half_rn(x));
}
__device__ __forceinline__ uint32_t pack16(float a, float b, int is_bf16) {
  if (is_bf16) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <int SUB>
__global__ void __launch_bounds__(kPThreads, SUB == 1 ? 2 : 1)
project_kernel(const __grid_constant__ CUtensorMap tmap_h, const __grid_constant__ CUtensorMap tmap_w,
               const ProjectArgs args, const int NS) {
  constexpr int kPStageBytes = p_stage_bytes(SUB);
  constexpr int kRows = SUB * kPTile;            // tokens per tile
  constexpr int kAccCols = SUB * kPOut;          // TMEM columns of one accumulator buffer
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tail = smem + (size_t)NS * kPStageBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(tail);   // [NS]
  uint64_t* empty = full + 8;                           // [NS]
  uint64_t* acc_full = empty + 8;                       // [2]
  uint64_t* acc_empty = acc_full + 2;                   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* sbias = reinterpret_cast<float*>(tmem_slot + 2);  // [128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ntiles = (args.n_tok + kRows - 1) / kRows;
  const int kb = args.kblocks;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_h);
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < NS; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(acc_full + a, 1); mbar_init(acc_empty + a, 4); }
    fence_barrier_init();
  }
  if (threadIdx.x < kPOut) sbias[threadIdx.x] = args.bias ? load16(args.bias, threadIdx.x, args.is_bf16) : 0.f;
  if (warp == 1) {
    tmem_alloc(tmem_slot, 2 * kAccCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int b = 0; b < kb; ++b, ++it) {
          const int s = it % NS;
          mbar_wait(empty + s, ((it / NS) & 1u) ^ 1u);
          mbar_arrive_expect_tx(full + s, kPStageBytes);
          uint8_t* dst = smem + (size_t)s * kPStageBytes;
          tma_load_2d(dst, &tmap_h, full + s, b * 64, (int32_t)(tile * kRows), kPolicyEvictFirst);      // SUB x 128 rows
          tma_load_2d(dst + SUB * kPBlockBytes, &tmap_w, full + s, b * 64, 0, kPolicyEvictLast);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(args.is_bf16 ? 1u : 0u, kPTile, kPOut);
      uint32_t it = 0, use = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++use) {
        const uint32_t a = use & 1u;
        mbar_wait(acc_empty + a, ((use >> 1) & 1u) ^ 1u);
        tc_fence_after();
        for (int b = 0; b < kb; ++b, ++it) {
          const int s = it % NS;
          mbar_wait(full + s, (it / NS) & 1u);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)s * kPStageBytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int h = 0; h < SUB; ++h)        // the sub-tiles share the weight block
              umma_f16(tmem_base + a * kAccCols + h * kPOut, make_kmajor_sw128_desc(sa + h * kPBlockBytes + k * 32),
                       make_kmajor_sw128_desc(sa + SUB * kPBlockBytes + k * 32), idesc, (b | k) ? 1u : 0u);
          umma_commit(empty + s);
        }
        umma_commit(acc_full + a);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    uint32_t use = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++use) {
      const uint32_t a = use & 1u;
      mbar_wait(acc_full + a, (use >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int h = 0; h < SUB; ++h) {
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + a * kAccCols + h * kPOut;
      const int64_t tok = tile * kRows + h * kPTile + row;
      const int rr = args.round_ref, bf = args.is_bf16;
      float ss = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(taddr + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float x = __uint_as_float(v[i]) + sbias[c * 32 + i];
          if (rr) x = round16(x, bf);      // the reference's Linear returns the model dtype
          ss = fmaf(x, x, ss);
        }
      }
      float nrm = sqrtf(ss);
      if (rr) nrm = round16(nrm, bf);      // x.norm() of a 16-bit tensor: fp32 accumulation, 16-bit result
      int64_t drow = tok;
      if (args.dst_row != nullptr && tok < args.n_tok) drow = __ldg(args.dst_row + tok);
      float m = 1.f;
      if (args.mask != nullptr && tok < args.n_tok) {
        bool on;
        if (args.mask_size == 1) on = __ldg(static_cast<const uint8_t*>(args.mask) + tok) != 0;
        else if (args.mask_size == 4) on = __ldg(static_cast<const int32_t*>(args.mask) + tok) != 0;
        else on = __ldg(static_cast<const long long*>(args.mask) + tok) != 0;
        m = on ? 1.f : 0.f;
      }
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(taddr + c * 32, v);
        tmem_ld_wait();
        if (tok < args.n_tok && drow >= 0) {
          uint4* dst = reinterpret_cast<uint4*>(static_cast<uint8_t*>(args.out) + drow * 256 + c * 64);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i = j * 8 + e * 2;
              float x0 = __uint_as_float(v[i]) + sbias[c * 32 + i];
              float x1 = __uint_as_float(v[i + 1]) + sbias[c * 32 + i + 1];
              if (rr) { x0 = round16(x0, bf); x1 = round16(x1, bf); }
              w[e] = pack16(x0 / nrm * m, x1 / nrm * m, bf);
            }
            dst[j] = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + a);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * kAccCols);
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static int encode_2d(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int box_rows, int dtype) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || !p) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return LIS_E_CUDA;
  }
  LIS_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor base %p is not 16-byte aligned", base);
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<PFN_encodeTiled>(p)(
      map, dtype == LIS_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
      const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return LIS_E_CUDA;
  }
  return LIS_OK;
}

}  // namespace lis

using namespace lis;

extern "C" int lis_project_normalize(const void* hidden, int64_t n_tok, int64_t hidden_dim, const void* weight,
                                     const void* bias, const void* mask, int mask_itemsize, int dtype,
                                     int round_mode, const int32_t* dst_row, void* out, void* stream) {
  LIS_REQUIRE(round_mode == LIS_ROUND_F32 || round_mode == LIS_ROUND_REFERENCE, "round_mode must be 0 (f32) or 1 (reference)");
  LIS_REQUIRE(mask == nullptr || mask_itemsize == 1 || mask_itemsize == 4 || mask_itemsize == 8,
              "mask_itemsize must be 1, 4 or 8");
  LIS_REQUIRE(hidden && weight && out, "lis_project_normalize: null pointer");
  LIS_REQUIRE(dtype == LIS_BF16 || dtype == LIS_F16, "dtype must be bf16 or f16");
  LIS_REQUIRE(n_tok > 0 && n_tok < (int64_t(1) << 31), "n_tok=%lld out of range", (long long)n_tok);
  LIS_REQUIRE(hidden_dim >= 64 && hidden_dim % 64 == 0 && hidden_dim <= 16384,
              "hidden_dim=%lld must be a multiple of 64 in [64, 16384]", (long long)hidden_dim);
  LIS_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "out is not 16-byte aligned");
  int dev = 0;
  LIS_CUDA_CHECK(cudaGetDevice(&dev));
  const int sms = sm_count(dev);
  LIS_REQUIRE(sms > 0, "no CUDA device");
  // Three launch shapes (LIS_K3_OCC overrides: 1, 2 or 3), measured in profiles/k3_occupancy_r2.txt:
  //   small batches   one CTA per SM, 128-token tiles, 6-stage ring (latency: the ring depth matters most)
  //   from 2 tiles/SM two CTAs per SM, 128-token tiles, 3 stages each (twice the TMA issuers, two epilogues in flight)
  //   (override 3)    one CTA per SM, 256-token tiles that share every weight block (half the L2 -> SM weight stream):
  //                   measured SLOWER than two CTAs per SM at every size (0.216 vs 0.179 ms at 264 k tokens) -- the weight
  //                   stream is not what limits the kernel -- and kept only as an experiment
  const int64_t ntiles128 = (n_tok + kPTile - 1) / kPTile;
  static const int occ_env = [] { const char* e = getenv("LIS_K3_OCC"); return e ? atoi(e) : 0; }();
  const int shape = occ_env >= 1 && occ_env <= 3 ? occ_env
                    : (ntiles128 >= 2 * (int64_t)sms ? 2 : 1);
  const int sub = shape == 3 ? 2 : 1;
  const int occ = shape == 2 ? 2 : 1;
  const int ns = shape == 1 ? 6 : (shape == 2 ? 3 : 4);
  const int smem = 1024 + ns * p_stage_bytes(sub) + 1024;
  CUtensorMap th, tw;
  int rc = encode_2d(&th, hidden, n_tok, hidden_dim, sub * kPTile, dtype);
  if (rc) return rc;
  rc = encode_2d(&tw, weight, kPOut, hidden_dim, kPOut, dtype);
  if (rc) return rc;
  static std::atomic<bool> configured[64];
  if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
    LIS_CUDA_CHECK(cudaFuncSetAttribute(project_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        1024 + 6 * p_stage_bytes(1) + 1024));
    LIS_CUDA_CHECK(cudaFuncSetAttribute(project_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        1024 + 4 * p_stage_bytes(2) + 1024));
    if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_release);
  }
  ProjectArgs a;
  a.bias = bias; a.mask = mask; a.mask_size = mask_itemsize; a.out = out; a.n_tok = n_tok;
  a.kblocks = (int32_t)(hidden_dim / 64);
  a.is_bf16 = dtype == LIS_BF16;
  a.round_ref = round_mode == LIS_ROUND_REFERENCE;
  a.dst_row = dst_row;
  const int64_t ntiles = (n_tok + sub * kPTile - 1) / (sub * kPTile);
  const int grid = (int)std::min<int64_t>((int64_t)sms * occ, ntiles);
  if (sub == 2) project_kernel<2><<<grid, kPThreads, smem, (cudaStream_t)stream>>>(th, tw, a, ns);
  else project_kernel<1><<<grid, kPThreads, smem, (cudaStream_t)stream>>>(th, tw, a, ns);
  count_launch();
  LIS_CUDA_CHECK(cudaGetLastError());
  return LIS_OK;
}
