// Microbenchmark: back-to-back tcgen05.mma issue rate (no epilogue, no TMA) on every SM.
// Reports cycles per 128xNx16 MMA and the implied fraction of 8192 FLOP/clk/SM.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../multi-modal_colpali_b200/csrc/lis_ptx.cuh"
using namespace lis;

template <int N, bool TS, bool ELECT>
__global__ void __launch_bounds__(128, 1) mma_rate(int groups, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[4];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  // zero the operands so no NaN/denormal slow paths can matter
  for (int i = threadIdx.x; i < (96 + 128) * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(bars + i, 1);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (ELECT ? (warp == 1) : (threadIdx.x == 32)) {
    const uint32_t idesc = make_idesc_f16(1, 128, N);
    const uint32_t a_base = smem_u32(smem);
    const uint32_t b_base = a_base + 96 * 1024;
    const long long t0 = clock64();
    long long issue_cycles = 0;
    for (int g = 0; g < groups; ++g) {
      const uint32_t a = a_base + (g % 3) * 32768;
      const uint32_t b = b_base + ((g / 3) & 1) * (N * 256);
      const uint32_t d = tmem + (TS ? 192 : 0) + (g & 1) * N;
      if (g >= 2) mbar_wait(bars + (g & 1), ((g - 2) >> 1) & 1);
      if (ELECT) { if (!elect_one_sync()) continue; }
      const long long i0 = clock64();
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t ka = (k >> 2) * 16384 + (k & 3) * 32;
        const uint32_t kb = (k >> 2) * (N * 128) + (k & 3) * 32;
        if (TS) umma_f16_ts(d, tmem + (g % 3) * 64 + k * 8, make_kmajor_sw128_desc(b + kb), idesc, k > 0);
        else umma_f16(d, make_kmajor_sw128_desc(a + ka), make_kmajor_sw128_desc(b + kb), idesc, k > 0);
      }
      umma_commit(bars + (g & 1));
      issue_cycles += clock64() - i0;
    }
    mbar_wait(bars + ((groups - 1) & 1), ((groups - 1) >> 1) & 1);
    const long long t1 = clock64();
    if (ELECT ? elect_one_sync() : true) { cycles[blockIdx.x] = t1 - t0; cycles[148 + blockIdx.x] = issue_cycles; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N, bool TS, bool ELECT>
void run() {
  long long* d;
  cudaMalloc(&d, 2 * 148 * 8);
  const int smem = (96 + 128) * 1024;
  cudaFuncSetAttribute(mma_rate<N, TS, ELECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int groups = 20000;
  mma_rate<N, TS, ELECT><<<148, 128, smem>>>(groups, d);
  mma_rate<N, TS, ELECT><<<148, 128, smem>>>(groups, d);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  mma_rate<N, TS, ELECT><<<148, 128, smem>>>(groups, d);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[296];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double iss = 0; for (int i = 0; i < 148; ++i) iss += h[148 + i]; iss /= 148;
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  const double per_mma = avg / (groups * 8.0);
  const double flops = 148.0 * groups * 8 * 2.0 * 128 * N * 16;
  printf("%s N=%3d cycles/MMA=%6.1f (ideal %d)  util=%.3f  wall %.3f ms -> %.1f TFLOP/s, eff clock %.0f MHz  (%s)\n", ELECT ? (TS ? "TS/elect" : "SS/elect") : (TS ? "TS/lane0" : "SS/lane0"), N, per_mma, N / 2,
         (N / 2) / per_mma, ms, flops / (ms * 1e-3) / 1e12, avg / (ms * 1e-3) / 1e6, cudaGetErrorString(e));
  printf("      issue (8 MMAs + commit) takes %.0f cycles of the %.0f-cycle group period\n", iss / groups, avg / groups);
  cudaFree(d);
}

int main() {
  run<256, false, true>(); run<128, false, true>(); run<64, true, true>();
  return 0;
}
