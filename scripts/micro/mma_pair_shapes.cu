// Microbenchmark: rate of the bare tcgen05.mma.cta_group::2 loop for the instruction shapes the CTA-pair kernel
// can choose from (no TMA, no epilogue; random operands; accumulators alternate between TMEM slots).
//   M = 256: one 128-row query tile per CTA;  M = 128: one query tile split 64/64 rows over the pair.
// Answers: does an M = 128 pair instruction run at half rate (as M = 64 does on one CTA), and what does N = 128 cost
// against N = 256 / 192?
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

template <int M, int N>
__global__ void __launch_bounds__(128, 1) mma_loop(int groups, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[4];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  constexpr int kARows = M / 2;                            // per CTA
  constexpr int kATile = kARows * 256;
  constexpr int kABytes = 3 * kATile;
  constexpr int kBRows = N / 2;                            // rows of the B tile held by this CTA
  constexpr int kBStage = kBRows * 256;
  constexpr int kSlots = 512 / N >= 4 ? 4 : 512 / N;       // accumulator ring (N columns each... M=128: N/2 columns)
  for (int i = threadIdx.x; i < (kABytes + 2 * kBStage) / 4; i += blockDim.x) {
    uint32_t h = hash32(i * 2654435761u + blockIdx.x);
    uint32_t lo = (h & 0x807f) | 0x3f00, hi = ((h >> 16) & 0x807f) | 0x3e80;
    reinterpret_cast<uint32_t*>(smem)[i] = lo | (hi << 16);
  }
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(bars + i, 1); fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1 && cluster_rank() == 0) {
    const uint32_t idesc = make_idesc_f16(1, M, N);
    const uint32_t a_base = smem_u32(smem), b_base = a_base + kABytes;
    const long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      const uint32_t a = a_base + (g % 3) * kATile;
      const uint32_t b = b_base + ((g / 3) & 1) * kBStage;
      const uint32_t s = g % kSlots;
      const uint32_t d = tmem + s * (M == 128 ? N / 2 : N);
      if (g >= kSlots) mbar_wait(bars + s, ((g / kSlots) - 1) & 1);
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t ka = (k >> 2) * (kARows * 128) + (k & 3) * 32;
          const uint32_t kb = (k >> 2) * (kBRows * 128) + (k & 3) * 32;
          const uint64_t ad = make_kmajor_sw128_desc(a + ka), bd = make_kmajor_sw128_desc(b + kb);
          const uint32_t acc = k > 0;
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bars + s)) : "memory");
      }
      __syncwarp();
    }
    for (int g = groups - kSlots < 0 ? 0 : groups - kSlots; g < groups; ++g) mbar_wait(bars + g % kSlots, (g / kSlots) & 1);
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

template <int M, int N>
void run() {
  long long* d; cudaMalloc(&d, 148 * 8); cudaMemset(d, 0, 148 * 8);
  const int smem = 3 * (M / 2) * 256 + 2 * (N / 2) * 256;
  cudaFuncSetAttribute(mma_loop<M, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int groups = 20000, reps = 40, timed = 20;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int r = 0; r < reps; ++r) {
    if (r == reps - timed) cudaEventRecord(e0);
    cudaLaunchKernelEx(&cfg, mma_loop<M, N>, groups, d);
  }
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= timed;
  long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double sum = 0; int n = 0; for (int i = 0; i < 148; ++i) if (h[i] > 0) { sum += h[i]; ++n; }
  const double avg = n ? sum / n : 0;
  const double flops = 74.0 * groups * 8 * 2.0 * M * N * 16;
  const double ideal = (double)M * N / 512.0;     // cycles per instruction at 8192 FLOP/clk/SM on both SMs
  printf("cta_group::2 M%3d N%3d  cycles/MMA=%6.1f (full-rate %5.1f, util %.3f)  %.3f ms/launch -> %7.1f TFLOP/s  eff clock %.0f MHz  (%s)\n",
         M, N, avg / (groups * 8.0), ideal, ideal / (avg / (groups * 8.0)), ms, flops / (ms * 1e-3) / 1e12, avg / (ms * 1e-3) / 1e6,
         cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<256, 256>(); run<256, 192>(); run<256, 128>(); run<256, 64>();
  run<128, 256>(); run<128, 192>(); run<128, 128>(); run<128, 64>();
  return 0;
}
