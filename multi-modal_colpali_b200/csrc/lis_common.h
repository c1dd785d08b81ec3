// Host-side helpers shared by the C-ABI translation units: thread-local error string,
// CUDA error mapping, launch counter, driver entry point for tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/lis.h"

namespace lis {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

struct Tuning {
  int tile_n = 0;
  int group = 0;
  int max_ctas = 0;
  int epi_halves = 0;  // 0 = auto
  int a_operand = 0;   // 0 = auto, 1 = shared memory (SS), 2 = tensor memory (TS)
  int ablate = 0;      // timing experiments only
  // pass planner: steady-state cost of one pass over a reference store (any unit), by resident query tiles;
  // defaults measured on a power-capped B200 (profiles/pass_costs_r2.jsonl), replaceable with lis_set_pass_costs
  float cost_single[4] = {0.f, 2.65f, 3.90f, 5.40f};
  float cost_pair[11] = {0.f, 0.f, 4.10f, 4.95f, 6.10f, 7.40f, 8.55f, 10.00f, 11.40f, 13.10f, 14.15f};
};
extern Tuning g_tuning;

// K2 tournament (lis_topk.cu).  seg_len > 0: row q of the input consists of blocks of seg_len candidates lying
// seg_stride_s floats / seg_stride_i int64 apart (all-gathered per-rank candidate blocks).
int run_tournament(const float* s, int64_t ld_s, const int64_t* ids, int64_t ld_ids, int64_t id_base, int64_t nq,
                   int64_t n, int k, float* out_s, int64_t* out_id, void* ws, int64_t ws_bytes, cudaStream_t st,
                   int seg_len, int64_t seg_stride_s, int64_t seg_stride_i);
// ncclAllGather of `bytes` bytes per rank (lis_comm.cu)
int comm_all_gather(lis_comm* c, const void* send, void* recv, size_t bytes, cudaStream_t st);
Tuning tuning_snapshot();
// Destroy every cached search graph that captured a collective of `c` (lis_index.cu).  NCCL keeps a communicator alive
// while a captured graph refers to it -- ncclCommDestroy would wait for ever -- so lis_comm_destroy calls this first.
void index_release_comm(lis_comm* c);

// Encode a 2-D row-major [rows, 128] 16-bit tensor with a (64 col x box_rows) box, 128-byte swizzle.
int encode_rows_tmap(CUtensorMap* map, const void* base, int64_t rows, int box_rows, int dtype);
int sm_count(int device);

}  // namespace lis

#define LIS_CUDA_CHECK(expr)                                                                  \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::lis::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return LIS_E_CUDA;                                                                      \
    }                                                                                         \
  } while (0)

#define LIS_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      ::lis::set_error(__VA_ARGS__);    \
      return LIS_E_INVALID;             \
    }                                   \
  } while (0)
