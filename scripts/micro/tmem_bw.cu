// Microbenchmark: TMEM -> register-file read throughput of tcgen05.ld on one SM (all SMs run it).
// Prints bytes/clk/SM for 4/8/16 reading warps and x32/x64/x128 column widths.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int NCOL>
__device__ __forceinline__ uint32_t ld_cols(uint32_t taddr);

template <>
__device__ __forceinline__ uint32_t ld_cols<32>(uint32_t taddr) {
  uint32_t v[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
        "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
        "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) x ^= v[i];
  return x;
}

// two x32 loads in flight before the wait
template <>
__device__ __forceinline__ uint32_t ld_cols<64>(uint32_t taddr) {
  uint32_t v[64];
#pragma unroll
  for (int h = 0; h < 2; ++h)
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[h * 32 + 0]), "=r"(v[h * 32 + 1]), "=r"(v[h * 32 + 2]), "=r"(v[h * 32 + 3]), "=r"(v[h * 32 + 4]),
          "=r"(v[h * 32 + 5]), "=r"(v[h * 32 + 6]), "=r"(v[h * 32 + 7]), "=r"(v[h * 32 + 8]), "=r"(v[h * 32 + 9]),
          "=r"(v[h * 32 + 10]), "=r"(v[h * 32 + 11]), "=r"(v[h * 32 + 12]), "=r"(v[h * 32 + 13]),
          "=r"(v[h * 32 + 14]), "=r"(v[h * 32 + 15]), "=r"(v[h * 32 + 16]), "=r"(v[h * 32 + 17]),
          "=r"(v[h * 32 + 18]), "=r"(v[h * 32 + 19]), "=r"(v[h * 32 + 20]), "=r"(v[h * 32 + 21]),
          "=r"(v[h * 32 + 22]), "=r"(v[h * 32 + 23]), "=r"(v[h * 32 + 24]), "=r"(v[h * 32 + 25]),
          "=r"(v[h * 32 + 26]), "=r"(v[h * 32 + 27]), "=r"(v[h * 32 + 28]), "=r"(v[h * 32 + 29]),
          "=r"(v[h * 32 + 30]), "=r"(v[h * 32 + 31])
        : "r"(taddr + h * 32)
        : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 64; ++i) x ^= v[i];
  return x;
}

// 16x256b shape: x8 -> 32 registers per thread too, but 16 lanes x 64 columns... (row pairs interleaved)
template <>
__device__ __forceinline__ uint32_t ld_cols<1>(uint32_t taddr) {
  uint32_t v[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
        "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
        "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) x ^= v[i];
  return x;
}

template <int MODE>
__global__ void bw_kernel(int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t x = 0;
  const int span = MODE == 64 ? 64 : 32;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    const uint32_t col = (uint32_t)((i * span + (warp >> 2) * 128) & 511) & ~(uint32_t)(span - 1);
    x ^= ld_cols<MODE>(base + (col & (512 - span)));
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (x == 0x12345678) sink[0] = x;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512));
}

template <int MODE>
void run(const char* name, int warps, int bytes_per_ld) {
  long long* d;
  uint32_t* sink;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaMalloc(&sink, 4);
  const int iters = 4096;
  bw_kernel<MODE><<<148, warps * 32>>>(iters, d, sink);
  bw_kernel<MODE><<<148, warps * 32>>>(iters, d, sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  printf("%-22s warps=%2d  cycles/ld(per warp)=%7.1f  bytes/clk/SM=%7.1f  (%s)\n", name, warps, avg / iters,
         (double)warps * iters * bytes_per_ld / avg, cudaGetErrorString(e));
  cudaFree(d);
  cudaFree(sink);
}

int main() {
  for (int w : {1, 4, 8, 16}) run<32>("32x32b.x32 (1 in flight)", w, 32 * 32 * 4);
  for (int w : {4, 8, 16}) run<64>("32x32b.x32 (2 in flight)", w, 2 * 32 * 32 * 4);
  for (int w : {4, 8}) run<1>("16x256b.x8", w, 32 * 32 * 4);
  return 0;
}
