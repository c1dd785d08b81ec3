// Microbenchmark: steady-state (power-capped) rate of the bare tcgen05.mma main loop with RANDOM operands:
//   cta_group::1  M=128 N=256 (what K1 issues)   vs   cta_group::2  M=256 N=256 (CTA pair shares the B operand)
// No TMA, no epilogue: this is the ceiling the tensor pipe + operand fetch allow under the 1 kW cap.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../multi-modal_colpali_b200/csrc/lis_ptx.cuh"
using namespace lis;

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r;
}

template <int CG>  // cta_group 1 or 2
__global__ void __launch_bounds__(128, 1) mma_loop(int groups, int random, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[2];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  constexpr int kABytes = 96 * 1024;                       // 3 A tiles of 128 rows (per CTA)
  constexpr int kBRows = CG == 2 ? 128 : 256;              // rows of the 256-row B tile held by this CTA
  constexpr int kBStage = kBRows * 256;
  // operands: bf16 values in [-1, 1) from a hash (or zeros)
  for (int i = threadIdx.x; i < (kABytes + 2 * kBStage) / 4; i += blockDim.x) {
    uint32_t h = hash32(i * 2654435761u + blockIdx.x);
    // two bf16: sign + exponent 0x3f (0.5..1) / 0x3e, random mantissa
    uint32_t lo = (h & 0x807f) | 0x3f00, hi = ((h >> 16) & 0x807f) | 0x3e80;
    reinterpret_cast<uint32_t*>(smem)[i] = random ? (lo | (hi << 16)) : 0u;
  }
  if (threadIdx.x == 0) { mbar_init(bars, 1); mbar_init(bars + 1, 1); fence_barrier_init(); }
  if (warp == 0) {
    if (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else { tmem_alloc(&slot, 512); tmem_relinquish(); }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = slot;
  const bool leader = CG == 1 || cluster_rank() == 0;
  if (warp == 1 && leader) {
    const uint32_t idesc = make_idesc_f16(1, CG == 2 ? 256 : 128, 256);
    const uint32_t a_base = smem_u32(smem), b_base = a_base + kABytes;
    const long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      const uint32_t a = a_base + (g % 3) * 32768;
      const uint32_t b = b_base + ((g / 3) & 1) * kBStage;
      const uint32_t d = tmem + (g & 1) * 256;
      if (g >= 2) mbar_wait(bars + (g & 1), ((g - 2) >> 1) & 1);
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t ka = (k >> 2) * 16384 + (k & 3) * 32;
          const uint32_t kb = (k >> 2) * (kBRows * 128) + (k & 3) * 32;
          const uint64_t ad = make_kmajor_sw128_desc(a + ka), bd = make_kmajor_sw128_desc(b + kb);
          const uint32_t acc = k > 0;
          if (CG == 2)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
          else umma_f16(d, ad, bd, idesc, acc);
        }
        if (CG == 2)
          asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bars + (g & 1))) : "memory");
        else umma_commit(bars + (g & 1));
      }
      __syncwarp();
    }
    mbar_wait(bars + ((groups - 1) & 1), ((groups - 1) >> 1) & 1);
    if (groups >= 2) mbar_wait(bars + ((groups - 2) & 1), ((groups - 2) >> 1) & 1);
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    else tmem_dealloc(tmem, 512);
  }
}

template <int CG>
void run(const char* name, int random) {
  long long* d; cudaMalloc(&d, 148 * 8); cudaMemset(d, 0, 148 * 8);
  const int smem = 96 * 1024 + 2 * (CG == 2 ? 128 : 256) * 256;
  cudaFuncSetAttribute(mma_loop<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int groups = 20000, reps = 260, timed = 60;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int r = 0; r < reps; ++r) {
    if (r == reps - timed) cudaEventRecord(e0);
    cudaLaunchKernelEx(&cfg, mma_loop<CG>, groups, random, d);
  }
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= timed;
  long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double sum = 0; int n = 0; for (int i = 0; i < 148; ++i) if (h[i] > 0) { sum += h[i]; ++n; }
  const double avg = n ? sum / n : 0;
  const double flops = (CG == 2 ? 74.0 * 2 : 148.0) * groups * 8 * 2.0 * 128 * 256 * 16;
  printf("%-34s cycles/MMA=%6.1f  %.3f ms/launch -> %7.1f TFLOP/s  eff clock %.0f MHz  (%s)\n", name, avg / (groups * 8.0), ms,
         flops / (ms * 1e-3) / 1e12, avg / (ms * 1e-3) / 1e6, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<1>("cta_group::1 M128 N256 zeros", 0);
  run<1>("cta_group::1 M128 N256 random", 1);
  run<2>("cta_group::2 M256 N256 random", 1);
  run<1>("cta_group::1 M128 N256 random", 1);
  run<2>("cta_group::2 M256 N256 random", 1);
  return 0;
}
