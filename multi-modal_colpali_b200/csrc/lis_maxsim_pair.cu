// Launcher of the CTA-pair form of K1 (maxsim_pair_kernel.cuh): thread-block clusters of two,
// tcgen05 cta_group::2.  Called from lis_maxsim.cu (tiling policy) -- no C-ABI entry point of its own.
#include <algorithm>
#include <type_traits>

#include "lis_common.h"
#include "maxsim_pair_kernel.cuh"

namespace lis {

constexpr int kPairSmemBudget = 232448;  // 227 KB opt-in maximum per CTA on sm_100

template <int NF, bool ODD, bool DBG>
static int launch_pair(const CUtensorMap& q, const CUtensorMap& p, const MaxSimArgs& a, int grid, cudaStream_t st) {
  constexpr int a_bytes = NF * 32768 + (ODD ? 16384 : 0);
  constexpr int stage = 32768;
  constexpr int kPairTail = pair_tail_bytes(NF);
  const int ns = std::min(8, (kPairSmemBudget - kPairTail - a_bytes) / stage);
  if (ns < 2) {
    set_error("pair kernel: %d query tiles do not fit shared memory", 2 * NF + (ODD ? 1 : 0));
    return LIS_E_INVALID;
  }
  const int smem = a_bytes + ns * stage + kPairTail;
  auto kern = maxsim_pair_kernel<NF, ODD, DBG>;
  static std::atomic<bool> configured[64];
  int dev = 0;
  LIS_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
    LIS_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_release);
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kCtrlThreads + 256);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  LIS_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, q, p, a, ns));
  count_launch();
  LIS_CUDA_CHECK(cudaGetLastError());
  return LIS_OK;
}

// n_mt query M tiles (2..10) in one pass: n_mt / 2 uses with M = 256 and, when n_mt is odd, one with M = 128.
// q, p: tensor maps of the packed query rows and of the token store, both with 64-row boxes.
int dispatch_maxsim_pair(const CUtensorMap& q, const CUtensorMap& p, const MaxSimArgs& a, int grid, cudaStream_t st,
                         bool dbg) {
  if (grid < 2 || (grid & 1)) {
    set_error("pair kernel: grid %d must be even", grid);
    return LIS_E_INVALID;
  }
  if (dbg) {   // raw-similarity dump: one even and one odd shape are enough to pin both accumulator layouts
    if (a.n_mt == 3) return launch_pair<1, true, true>(q, p, a, grid, st);
    if (a.n_mt == 4) return launch_pair<2, false, true>(q, p, a, grid, st);
    set_error("pair kernel: the debug dump exists for 3 and 4 query tiles");
    return LIS_E_INVALID;
  }
#define LIS_PAIR_CASE(N_, NF_, ODD_) \
  if (a.n_mt == N_) return launch_pair<NF_, ODD_, false>(q, p, a, grid, st);
  LIS_PAIR_CASE(2, 1, false)
  LIS_PAIR_CASE(3, 1, true)
  LIS_PAIR_CASE(4, 2, false)
  LIS_PAIR_CASE(5, 2, true)
  LIS_PAIR_CASE(6, 3, false)
  LIS_PAIR_CASE(7, 3, true)
  LIS_PAIR_CASE(8, 4, false)
  LIS_PAIR_CASE(9, 4, true)
  LIS_PAIR_CASE(10, 5, false)
#undef LIS_PAIR_CASE
  set_error("pair kernel: unsupported tile count %d (2..10)", a.n_mt);
  return LIS_E_INVALID;
}

}  // namespace lis
