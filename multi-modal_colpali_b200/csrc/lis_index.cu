// GPU-resident page index: the in-HBM replacement for the reference's Qdrant multivector
// collection (schema 01_create_context_qdrant.py:208-222: 128-d, COSINE, MAX_SIM; upsert
// functions.py:865; query functions.py:894-926) and for the pickled list of page embeddings that
// score_results re-stacks on every call (05_experiment02.py:213).
//
// HBM layout (all owned by the index, sized once at create time -- no allocation on the add path):
//   tokens   [cap_rows, 128] 16-bit, row-major: every page's token rows back to back (ragged, no padding)
//   offsets  int64 [cap_pages+1]: page p owns rows [offsets[p], offsets[p+1])
//   ids      int64 [cap_pages]:   caller's page id (payload key on the Python side)
//   clamp    uint8 [cap_pages]:   zero-padding semantics flag (see lis.h, p_clamp)
// Search scratch (score rows, top-k tournament buffers) is grown on demand and reused.
#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cstring>
#include <map>
#include <thread>
#include <mutex>
#include <tuple>
#include <vector>

#include "lis_common.h"
#include "lis_ptx.cuh"

namespace lis {
// Everything a replayable search is specialised on; any change of these needs a new CUDA graph.
typedef std::tuple<int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int, int, int, int, int, uint64_t> GraphKey;
}

struct lis_index {
  int device = 0;
  int dtype = LIS_BF16;
  int64_t cap_rows = 0, cap_pages = 0;
  int64_t n_rows = 0, n_pages = 0;
  void* tokens = nullptr;      // F32X2: hi plane rows [0, cap_rows), lo plane rows [cap_rows, 2*cap_rows)
  int64_t* offsets = nullptr;
  int64_t* ids = nullptr;
  uint8_t* clamp = nullptr;
  // scratch
  float* seg_scores = nullptr; int64_t seg_scores_bytes = 0;
  float* q_scores = nullptr;   int64_t q_scores_bytes = 0;
  void* topk_ws = nullptr;     int64_t topk_ws_bytes = 0;
  std::mutex mu;               // one search in flight per index (the scratch above is shared)
  // one-shot search (lis_index_search_sharded): fixed staging so that the whole sequence replays as a CUDA graph
  cudaStream_t sstream = nullptr;       // owned; captured / replayed on
  cudaEvent_t sevent = nullptr;
  uint8_t* h_stage = nullptr; int64_t h_stage_bytes = 0;   // pinned: tables | query rows
  uint8_t* d_stage = nullptr; int64_t d_stage_bytes = 0;   // same layout (+ hi/lo planes of fp32 queries)
  uint8_t* h_out = nullptr;   int64_t h_out_bytes = 0;     // pinned: scores | ids
  uint8_t* d_out = nullptr;   int64_t d_out_bytes = 0;
  uint8_t* sendbuf = nullptr; int64_t sendbuf_bytes = 0;   // this rank's candidates: scores | ids
  uint8_t* recvbuf = nullptr; int64_t recvbuf_bytes = 0;   // world x the same
  uint64_t gen = 0;                     // bumped whenever a buffer above moved or the content changed
  struct Replay { cudaGraphExec_t exec; int kernels; };
  std::map<lis::GraphKey, Replay> graphs;
  lis_comm* graph_comm = nullptr;       // the communicator whose all-gather the cached graphs captured (if any)
  int64_t graph_replays = 0, graph_captures = 0;
};

namespace lis {

__device__ __forceinline__ uint64_t mix64(uint64_t x) {  // splitmix64 finaliser
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// One warp per row: lane l produces elements 4l..4l+3 as Box-Muller normals from a counter hash of
// (seed, global row, pair index), the warp reduces the squared norm, and the row is stored unit-norm.
__global__ void __launch_bounds__(256)
fill_rows_kernel(void* __restrict__ dst, int64_t row0, int64_t n_rows, uint64_t seed, int is_bf16) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp0; r < n_rows; r += nwarps) {
    const uint64_t grow = (uint64_t)(row0 + r);
    float x[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint64_t bits = mix64(seed ^ mix64(grow * 64 + (uint64_t)(lane * 2 + h)));
      const float u1 = ((uint32_t)(bits >> 40) + 1u) * (1.0f / 16777217.0f);  // (0, 1)
      const float u2 = (uint32_t)(bits & 0xFFFFFFu) * (1.0f / 16777216.0f);   // [0, 1)
      const float rad = sqrtf(-2.0f * __logf(u1));
      float sn, cs;
      __sincosf(6.28318530718f * u2, &sn, &cs);
      x[2 * h] = rad * cs;
      x[2 * h + 1] = rad * sn;
    }
    float ss = x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float inv = rsqrtf(fmaxf(ss, 1e-30f));
    uint2 w;
    if (is_bf16) {
      __nv_bfloat162 a = __floats2bfloat162_rn(x[0] * inv, x[1] * inv);
      __nv_bfloat162 b = __floats2bfloat162_rn(x[2] * inv, x[3] * inv);
      w.x = *reinterpret_cast<uint32_t*>(&a);
      w.y = *reinterpret_cast<uint32_t*>(&b);
    } else {
      __half2 a = __floats2half2_rn(x[0] * inv, x[1] * inv);
      __half2 b = __floats2half2_rn(x[2] * inv, x[3] * inv);
      w.x = *reinterpret_cast<uint32_t*>(&a);
      w.y = *reinterpret_cast<uint32_t*>(&b);
    }
    reinterpret_cast<uint2*>(static_cast<uint8_t*>(dst) + r * 256)[lane] = w;
  }
}

__global__ void merge_planes_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                    int64_t n, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __bfloat162float(hi[i]) + __bfloat162float(lo[i]);
}

static inline int elem_dtype(const lis_index* ix) { return ix->dtype == LIS_F32X2 ? LIS_BF16 : ix->dtype; }
static inline uint8_t* lo_plane(const lis_index* ix) {
  return static_cast<uint8_t*>(ix->tokens) + (size_t)ix->cap_rows * 256;
}

static void drop_graphs(lis_index* ix) {
  for (auto& kv : ix->graphs) cudaGraphExecDestroy(kv.second.exec);
  ix->graphs.clear();
  ix->graph_comm = nullptr;
  ++ix->gen;
}

// live indexes, so that a communicator can find the graphs that captured it
static std::mutex g_registry_mu;
static std::vector<lis_index*> g_registry;

void index_release_comm(lis_comm* c) {
  std::lock_guard<std::mutex> reg(g_registry_mu);
  for (lis_index* ix : g_registry) {
    std::lock_guard<std::mutex> lock(ix->mu);
    if (ix->graph_comm == c) {
      cudaSetDevice(ix->device);
      if (ix->sstream) cudaStreamSynchronize(ix->sstream);
      drop_graphs(ix);
    }
  }
}

static int ensure(lis_index* ix, void** p, int64_t* have, int64_t need) {
  if (*have >= need) return LIS_OK;
  if (ix) drop_graphs(ix);   // captured launches hold the old pointer
  if (*p) cudaFree(*p);
  *p = nullptr;
  *have = 0;
  const int64_t bytes = need + need / 4;  // head-room so repeated growth is rare
  cudaError_t e = cudaMalloc(p, (size_t)bytes);
  if (e != cudaSuccess) {
    set_error("cudaMalloc(%lld) failed: %s", (long long)bytes, cudaGetErrorString(e));
    return LIS_E_NOMEM;
  }
  *have = bytes;
  return LIS_OK;
}

}  // namespace lis

using namespace lis;

extern "C" {

int lis_fill_synthetic_rows(void* dst, int64_t row0, int64_t n_rows, uint64_t seed, int dtype, void* stream) {
  LIS_REQUIRE(n_rows >= 0 && row0 >= 0, "negative row range");
  if (n_rows == 0) return LIS_OK;
  LIS_REQUIRE(dst, "null destination");
  LIS_REQUIRE(dtype == LIS_BF16 || dtype == LIS_F16, "dtype must be bf16 or f16");
  int dev = 0;
  LIS_CUDA_CHECK(cudaGetDevice(&dev));
  const int sms = sm_count(dev);
  const int64_t want = (n_rows + 7) / 8;  // 8 warps (rows) per block
  const unsigned grid = (unsigned)std::min<int64_t>(want, (int64_t)sms * 16);
  fill_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dst, row0, n_rows, seed, dtype == LIS_BF16);
  count_launch();
  LIS_CUDA_CHECK(cudaGetLastError());
  return LIS_OK;
}

int lis_index_create(lis_index** out, int device, int dtype, int64_t cap_rows, int64_t cap_pages) {
  LIS_REQUIRE(out, "null out");
  *out = nullptr;
  LIS_REQUIRE(dtype == LIS_BF16 || dtype == LIS_F16 || dtype == LIS_F32X2, "dtype must be bf16, f16 or f32x2");
  LIS_REQUIRE(cap_rows > 0 && cap_pages > 0, "capacities must be positive");
  LIS_REQUIRE(cap_rows < (int64_t(1) << 31), "cap_rows must be < 2^31 per index (shard the corpus)");
  int rc = lis_device_supported(device);
  if (rc) return rc;
  LIS_CUDA_CHECK(cudaSetDevice(device));
  lis_index* ix = new lis_index();
  ix->device = device;
  ix->dtype = dtype;
  ix->cap_rows = cap_rows;
  ix->cap_pages = cap_pages;
  cudaError_t e = cudaMalloc(&ix->tokens, (size_t)cap_rows * 256 * (dtype == LIS_F32X2 ? 2 : 1));
  if (e == cudaSuccess) e = cudaMalloc((void**)&ix->offsets, (size_t)(cap_pages + 1) * 8);
  if (e == cudaSuccess) e = cudaMalloc((void**)&ix->ids, (size_t)cap_pages * 8);
  if (e == cudaSuccess) e = cudaMalloc((void**)&ix->clamp, (size_t)cap_pages);
  if (e == cudaSuccess) e = cudaMemset(ix->offsets, 0, 8);
  if (e != cudaSuccess) {
    set_error("index allocation failed (%lld rows, %lld pages): %s", (long long)cap_rows, (long long)cap_pages,
              cudaGetErrorString(e));
    lis_index_destroy(ix);
    return LIS_E_NOMEM;
  }
  {
    std::lock_guard<std::mutex> reg(g_registry_mu);
    g_registry.push_back(ix);
  }
  *out = ix;
  return LIS_OK;
}

void lis_index_destroy(lis_index* ix) {
  if (!ix) return;
  {
    std::lock_guard<std::mutex> reg(g_registry_mu);
    g_registry.erase(std::remove(g_registry.begin(), g_registry.end(), ix), g_registry.end());
  }
  cudaFree(ix->tokens);
  cudaFree(ix->offsets);
  cudaFree(ix->ids);
  cudaFree(ix->clamp);
  cudaFree(ix->seg_scores);
  cudaFree(ix->q_scores);
  cudaFree(ix->topk_ws);
  drop_graphs(ix);
  if (ix->sstream) cudaStreamDestroy(ix->sstream);
  if (ix->sevent) cudaEventDestroy(ix->sevent);
  if (ix->h_stage) cudaFreeHost(ix->h_stage);
  if (ix->h_out) cudaFreeHost(ix->h_out);
  cudaFree(ix->d_stage);
  cudaFree(ix->d_out);
  cudaFree(ix->sendbuf);
  cudaFree(ix->recvbuf);
  delete ix;
}

int64_t lis_index_num_pages(const lis_index* ix) { return ix ? ix->n_pages : 0; }
int64_t lis_index_num_rows(const lis_index* ix) { return ix ? ix->n_rows : 0; }
const void* lis_index_tokens(const lis_index* ix) { return ix ? ix->tokens : nullptr; }
const void* lis_index_tokens_lo(const lis_index* ix) { return ix && ix->dtype == LIS_F32X2 ? lo_plane(ix) : nullptr; }
const int64_t* lis_index_offsets(const lis_index* ix) { return ix ? ix->offsets : nullptr; }
const int64_t* lis_index_ids(const lis_index* ix) { return ix ? ix->ids : nullptr; }
const uint8_t* lis_index_clamp(const lis_index* ix) { return ix ? ix->clamp : nullptr; }

// Shared by add / fill: validate capacity, upload the page tables for n new pages.
static int append_tables(lis_index* ix, const int32_t* lens, int32_t fixed_len, const int64_t* ids,
                         int64_t id_base, const uint8_t* clamp, int64_t n, int64_t* new_rows, cudaStream_t st) {
  LIS_REQUIRE(ix, "null index");
  LIS_REQUIRE(n > 0, "nothing to add");
  LIS_REQUIRE(ix->n_pages + n <= ix->cap_pages, "page capacity exceeded: %lld + %lld > %lld",
              (long long)ix->n_pages, (long long)n, (long long)ix->cap_pages);
  std::vector<int64_t> off(n + 1), idv(n);
  std::vector<uint8_t> cl(n, 0);
  int64_t row = ix->n_rows;
  off[0] = row;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t len = lens ? lens[i] : fixed_len;
    LIS_REQUIRE(len >= 0, "negative page length at %lld", (long long)i);
    row += len;
    off[i + 1] = row;
    idv[i] = ids ? ids[i] : id_base + i;
    LIS_REQUIRE(idv[i] >= 0, "page ids must be non-negative");
    if (clamp) cl[i] = clamp[i] ? 1 : 0;
  }
  LIS_REQUIRE(row <= ix->cap_rows, "row capacity exceeded: %lld > %lld", (long long)row, (long long)ix->cap_rows);
  LIS_CUDA_CHECK(cudaMemcpyAsync(ix->offsets + ix->n_pages, off.data(), (size_t)(n + 1) * 8,
                                 cudaMemcpyHostToDevice, st));
  LIS_CUDA_CHECK(cudaMemcpyAsync(ix->ids + ix->n_pages, idv.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
  LIS_CUDA_CHECK(cudaMemcpyAsync(ix->clamp + ix->n_pages, cl.data(), (size_t)n, cudaMemcpyHostToDevice, st));
  LIS_CUDA_CHECK(cudaStreamSynchronize(st));  // the staging vectors die with this frame
  *new_rows = row - ix->n_rows;
  return LIS_OK;
}

int lis_index_add(lis_index* ix, const void* tokens, const int32_t* lens, const int64_t* ids,
                  const uint8_t* clamp, int64_t n, void* stream) {
  LIS_REQUIRE(ix && lens, "lis_index_add: null pointer");
  std::lock_guard<std::mutex> lock(ix->mu);      // ingestion and search never overlap on one index
  cudaStream_t st = (cudaStream_t)stream;
  LIS_CUDA_CHECK(cudaSetDevice(ix->device));
  int64_t new_rows = 0;
  int rc = append_tables(ix, lens, 0, ids, ix->n_pages, clamp, n, &new_rows, st);
  if (rc) return rc;
  if (new_rows > 0) {
    LIS_REQUIRE(tokens, "lis_index_add: null tokens");
    uint8_t* hi = static_cast<uint8_t*>(ix->tokens) + ix->n_rows * 256;
    if (ix->dtype == LIS_F32X2) {
      // fp32 rows land in the index's own scratch (the search workspace, grown on demand like for a search: no
      // allocation per call), then get split into the planes
      int rc1 = ensure(ix, &ix->topk_ws, &ix->topk_ws_bytes, new_rows * 512);
      if (rc1) return rc1;
      float* stage = static_cast<float*>(ix->topk_ws);
      cudaError_t e = cudaMemcpyAsync(stage, tokens, (size_t)new_rows * 512, cudaMemcpyDefault, st);
      int rc2 = e == cudaSuccess ? lis_split_f32(stage, new_rows, hi, lo_plane(ix) + ix->n_rows * 256, stream) : LIS_E_CUDA;
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) { set_error("index add (fp32) failed: %s", cudaGetErrorString(e)); return LIS_E_CUDA; }
      if (rc2) return rc2;
    } else {
      LIS_CUDA_CHECK(cudaMemcpyAsync(hi, tokens, (size_t)new_rows * 256, cudaMemcpyDefault, st));
      LIS_CUDA_CHECK(cudaStreamSynchronize(st));
    }
  }
  ix->n_rows += new_rows;
  ix->n_pages += n;
  return LIS_OK;
}

int lis_index_fill_synthetic(lis_index* ix, int64_t n, const int32_t* lens, int32_t fixed_len, uint64_t seed,
                             int64_t id_base, void* stream) {
  LIS_REQUIRE(ix, "null index");
  LIS_REQUIRE(lens || fixed_len > 0, "need lens or a positive fixed_len");
  cudaStream_t st = (cudaStream_t)stream;
  LIS_CUDA_CHECK(cudaSetDevice(ix->device));
  int64_t new_rows = 0;
  int rc = append_tables(ix, lens, fixed_len, nullptr, id_base, nullptr, n, &new_rows, st);
  if (rc) return rc;
  // global row index = id_base-independent position in this index; callers that shard a corpus
  // pass distinct seeds per shard
  rc = lis_fill_synthetic_rows(static_cast<uint8_t*>(ix->tokens) + ix->n_rows * 256, ix->n_rows, new_rows, seed,
                               elem_dtype(ix), stream);
  if (rc) return rc;
  if (ix->dtype == LIS_F32X2)  // synthetic rows are exactly representable in bf16: the lo plane is zero
    LIS_CUDA_CHECK(cudaMemsetAsync(lo_plane(ix) + ix->n_rows * 256, 0, (size_t)new_rows * 256, st));
  LIS_CUDA_CHECK(cudaStreamSynchronize(st));
  ix->n_rows += new_rows;
  ix->n_pages += n;
  return LIS_OK;
}

int lis_index_read_rows(const lis_index* ix, int64_t row0, int64_t n_rows, void* dst, void* stream) {
  LIS_REQUIRE(ix && dst, "null pointer");
  LIS_REQUIRE(row0 >= 0 && n_rows >= 0 && row0 + n_rows <= ix->n_rows, "row range out of bounds");
  if (n_rows == 0) return LIS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (ix->dtype == LIS_F32X2) {
    float* stage = nullptr;
    LIS_CUDA_CHECK(cudaMalloc((void**)&stage, (size_t)n_rows * 512));
    const int64_t n = n_rows * 128;
    merge_planes_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 65535), 256, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(static_cast<const uint8_t*>(ix->tokens) + row0 * 256),
        reinterpret_cast<const __nv_bfloat16*>(lo_plane(ix) + row0 * 256), n, stage);
    count_launch();
    cudaError_t e = cudaMemcpyAsync(dst, stage, (size_t)n_rows * 512, cudaMemcpyDefault, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(stage);
    LIS_CUDA_CHECK(e);
    return LIS_OK;
  }
  LIS_CUDA_CHECK(cudaMemcpyAsync(dst, static_cast<const uint8_t*>(ix->tokens) + row0 * 256, (size_t)n_rows * 256,
                                 cudaMemcpyDefault, st));
  LIS_CUDA_CHECK(cudaStreamSynchronize(st));
  return LIS_OK;
}

int lis_index_dtype(const lis_index* ix) { return ix ? ix->dtype : -1; }

int lis_index_read_plane(const lis_index* ix, int plane, int64_t row0, int64_t n_rows, void* dst, void* stream) {
  LIS_REQUIRE(ix && dst, "null pointer");
  LIS_REQUIRE(plane == 0 || (plane == 1 && ix->dtype == LIS_F32X2), "no such plane");
  LIS_REQUIRE(row0 >= 0 && n_rows >= 0 && row0 + n_rows <= ix->n_rows, "row range out of bounds");
  if (n_rows == 0) return LIS_OK;
  const uint8_t* base = plane ? lo_plane(ix) : static_cast<const uint8_t*>(ix->tokens);
  cudaStream_t st = (cudaStream_t)stream;
  LIS_CUDA_CHECK(cudaMemcpyAsync(dst, base + row0 * 256, (size_t)n_rows * 256, cudaMemcpyDefault, st));
  LIS_CUDA_CHECK(cudaStreamSynchronize(st));
  return LIS_OK;
}

int lis_index_write_rows(lis_index* ix, int plane, int64_t row0, int64_t n_rows, const void* src, void* stream) {
  LIS_REQUIRE(ix && src, "null pointer");
  LIS_REQUIRE(plane == 0 || (plane == 1 && ix->dtype == LIS_F32X2), "no such plane");
  LIS_REQUIRE(row0 >= 0 && n_rows >= 0 && row0 + n_rows <= ix->cap_rows, "row range exceeds capacity");
  if (n_rows == 0) return LIS_OK;
  uint8_t* base = plane ? lo_plane(ix) : static_cast<uint8_t*>(ix->tokens);
  cudaStream_t st = (cudaStream_t)stream;
  LIS_CUDA_CHECK(cudaSetDevice(ix->device));
  LIS_CUDA_CHECK(cudaMemcpyAsync(base + row0 * 256, src, (size_t)n_rows * 256, cudaMemcpyDefault, st));
  LIS_CUDA_CHECK(cudaStreamSynchronize(st));
  return LIS_OK;
}

int lis_index_set_tables(lis_index* ix, const int64_t* offsets, const int64_t* ids, const uint8_t* clamp,
                         int64_t n_pages, void* stream) {
  LIS_REQUIRE(ix && offsets && ids, "null pointer");
  LIS_REQUIRE(n_pages >= 0 && n_pages <= ix->cap_pages, "page count exceeds capacity");
  LIS_REQUIRE(offsets[0] == 0, "offsets must start at 0");
  for (int64_t i = 0; i < n_pages; ++i) {
    LIS_REQUIRE(offsets[i + 1] >= offsets[i], "offsets must be ascending (page %lld)", (long long)i);
    LIS_REQUIRE(ids[i] >= -1, "page ids must be non-negative (or -1: a removed page)");
  }
  LIS_REQUIRE(offsets[n_pages] <= ix->cap_rows, "rows exceed capacity");
  cudaStream_t st = (cudaStream_t)stream;
  LIS_CUDA_CHECK(cudaSetDevice(ix->device));
  LIS_CUDA_CHECK(cudaMemcpyAsync(ix->offsets, offsets, (size_t)(n_pages + 1) * 8, cudaMemcpyHostToDevice, st));
  if (n_pages > 0) {
    LIS_CUDA_CHECK(cudaMemcpyAsync(ix->ids, ids, (size_t)n_pages * 8, cudaMemcpyHostToDevice, st));
    if (clamp) LIS_CUDA_CHECK(cudaMemcpyAsync(ix->clamp, clamp, (size_t)n_pages, cudaMemcpyHostToDevice, st));
    else LIS_CUDA_CHECK(cudaMemsetAsync(ix->clamp, 0, (size_t)n_pages, st));
  }
  LIS_CUDA_CHECK(cudaStreamSynchronize(st));
  ix->n_pages = n_pages;
  ix->n_rows = offsets[n_pages];
  return LIS_OK;
}

}  // extern "C" (reopened below)

namespace lis {

// Size the shared scratch for a search of this shape (may reallocate: never call while capturing).
static int prepare_scratch(lis_index* ix, int64_t n_seg, int64_t nq, int k, bool direct, int world) {
  const int64_t np = std::max<int64_t>(ix->n_pages, 1);
  int rc = ensure(ix, (void**)&ix->seg_scores, &ix->seg_scores_bytes, n_seg * np * 4);
  if (rc) return rc;
  int64_t ws_need = lis_topk_workspace_bytes(nq, np, k);
  if (world > 1) ws_need = std::max(ws_need, lis_topk_workspace_bytes(nq, (int64_t)world * k, k));
  rc = ensure(ix, &ix->topk_ws, &ix->topk_ws_bytes, std::max<int64_t>(ws_need, 256));
  if (rc) return rc;
  if (!direct) {
    rc = ensure(ix, (void**)&ix->q_scores, &ix->q_scores_bytes, nq * np * 4);
    if (rc) return rc;
  }
  return LIS_OK;
}

// K1 -> (segment sums) -> K2 on `stream`; all pointers device; scratch sized by prepare_scratch.
static int enqueue_local_search(lis_index* ix, const void* q, const void* q_lo, int64_t q_rows, const int32_t* seg_lo,
                                const int32_t* seg_hi, const int32_t* mt_seg, int64_t n_seg, int64_t n_mtiles,
                                const int32_t* seg_first, int64_t nq, int round_mode, int k, float* out_scores,
                                int64_t* out_ids, void* stream) {
  const bool direct = seg_first == nullptr;
  const int64_t np = ix->n_pages;
  const int k1_round = direct ? round_mode : (round_mode | LIS_ROUND_DEFER_SUM);
  int rc;
  if (ix->dtype == LIS_F32X2) {
    LIS_REQUIRE(q_lo, "an f32x2 index needs the low plane of the queries");
    round_mode = LIS_ROUND_F32;
    rc = lis_maxsim_scores_f32x2(q, q_lo, q_rows, seg_lo, seg_hi, mt_seg, n_seg, n_mtiles, ix->tokens, lo_plane(ix),
                                 ix->n_rows, ix->offsets, ix->clamp, np, ix->seg_scores, np, stream);
  } else {
    rc = lis_maxsim_scores(q, q_rows, seg_lo, seg_hi, mt_seg, n_seg, n_mtiles, ix->tokens, ix->n_rows, ix->offsets,
                           ix->clamp, np, ix->dtype, k1_round, ix->seg_scores, np, stream);
  }
  if (rc) return rc;
  const float* scores = ix->seg_scores;
  if (!direct) {
    rc = lis_reduce_segments(ix->seg_scores, np, seg_first, nq, np, round_mode, elem_dtype(ix), ix->q_scores, np, stream);
    if (rc) return rc;
    scores = ix->q_scores;
  }
  return lis_topk(scores, np, nq, np, ix->ids, 0, k, out_scores, out_ids, ix->topk_ws, ix->topk_ws_bytes, stream);
}

__global__ void fill_padding_kernel(float* s, int64_t* ids, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    s[i] = -INFINITY;
    ids[i] = -1;
  }
}

// ---- ingestion fusion: where each token of a padded encoder batch lands in the ragged store ----
__device__ __forceinline__ bool mask_on(const void* mask, int itemsize, int64_t i) {
  if (itemsize == 1) return __ldg(static_cast<const uint8_t*>(mask) + i) != 0;
  if (itemsize == 4) return __ldg(static_cast<const int32_t*>(mask) + i) != 0;
  return __ldg(static_cast<const long long*>(mask) + i) != 0;
}

// One block per page: page-local exclusive prefix of the mask -> dst_row (or -1), kept-token count -> lens.
__global__ void __launch_bounds__(256)
page_prefix_kernel(const void* __restrict__ mask, int itemsize, int64_t seq, int32_t* __restrict__ dst_row,
                   int32_t* __restrict__ lens) {
  __shared__ int warp_sums[8];
  __shared__ int carry;
  const int64_t b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t t0 = 0; t0 < seq; t0 += 256) {
    const int64_t t = t0 + threadIdx.x;
    const bool on = t < seq && mask_on(mask, itemsize, b * seq + t);
    const unsigned bal = __ballot_sync(0xffffffffu, on);
    const int before = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) warp_sums[warp] = __popc(bal);
    __syncthreads();
    int wbase = carry;
    for (int w = 0; w < warp; ++w) wbase += warp_sums[w];
    if (t < seq) dst_row[b * seq + t] = on ? wbase + before : -1;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < 8; ++w) tot += warp_sums[w];
      carry += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) lens[b] = carry;
}

// One block: running row offsets of the new pages, their clamp flags, and the row base of every page.
__global__ void __launch_bounds__(256)
page_tables_kernel(const int32_t* __restrict__ lens, int64_t n, int64_t seq, int64_t row0, int64_t* __restrict__ offsets_out,
                   uint8_t* __restrict__ clamp_out, int64_t* __restrict__ base_out) {
  __shared__ int64_t sh[256];
  __shared__ int64_t carry;
  if (threadIdx.x == 0) carry = row0;
  __syncthreads();
  for (int64_t i0 = 0; i0 < n; i0 += 256) {
    const int64_t i = i0 + threadIdx.x;
    const int64_t len = i < n ? lens[i] : 0;
    sh[threadIdx.x] = len;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {      // Hillis-Steele inclusive scan
      const int64_t add = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += add;
      __syncthreads();
    }
    if (i < n) {
      const int64_t end = carry + sh[threadIdx.x];
      base_out[i] = end - len;
      offsets_out[i + 1] = end;               // offsets_out points at offsets[n_pages_old]
      clamp_out[i] = len < seq ? 1 : 0;
    }
    __syncthreads();
    if (threadIdx.x == 255) carry += sh[255];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
page_rebase_kernel(int32_t* __restrict__ dst_row, const int64_t* __restrict__ base, int64_t seq, int64_t n_tok) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_tok; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t d = dst_row[i];
    if (d >= 0) dst_row[i] = (int32_t)(base[i / seq] + d);
  }
}

static inline int64_t al256(int64_t x) { return (x + 255) & ~int64_t(255); }

// Select the index's device for the duration of a call and give the caller's device back afterwards.
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev); else prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

static int ensure_host(uint8_t** p, int64_t* have, int64_t need) {
  if (*have >= need) return LIS_OK;
  if (*p) cudaFreeHost(*p);
  *p = nullptr;
  *have = 0;
  const int64_t bytes = need + need / 2;
  cudaError_t e = cudaHostAlloc((void**)p, (size_t)bytes, cudaHostAllocDefault);
  if (e != cudaSuccess) {
    set_error("cudaHostAlloc(%lld) failed: %s", (long long)bytes, cudaGetErrorString(e));
    return LIS_E_NOMEM;
  }
  *have = bytes;
  return LIS_OK;
}

}  // namespace lis

extern "C" {

int lis_index_search(lis_index* ix, const void* q, const void* q_lo, int64_t q_rows, const int32_t* seg_lo, const int32_t* seg_hi,
                     const int32_t* mt_seg, int64_t n_seg, int64_t n_mtiles, const int32_t* seg_first, int64_t nq,
                     int round_mode, int k, float* out_scores, int64_t* out_ids, void* stream) {
  LIS_REQUIRE(ix, "null index");
  LIS_REQUIRE(ix->n_pages > 0, "index is empty");
  LIS_REQUIRE(nq > 0 && n_seg > 0, "bad query counts nq=%lld n_seg=%lld", (long long)nq, (long long)n_seg);
  // seg_first == NULL is the caller's statement that segment s IS query s; the counts alone cannot prove it
  // (an empty query owns no segment), so anything else must come with the segment table
  LIS_REQUIRE(seg_first != nullptr || n_seg == nq, "split or empty queries need seg_first");
  LIS_REQUIRE(k >= 1 && k <= LIS_MAX_K, "k out of range");
  std::lock_guard<std::mutex> lock(ix->mu);
  int rc = prepare_scratch(ix, n_seg, nq, k, seg_first == nullptr, 1);
  if (rc) return rc;
  return enqueue_local_search(ix, q, q_lo, q_rows, seg_lo, seg_hi, mt_seg, n_seg, n_mtiles, seg_first, nq, round_mode, k,
                              out_scores, out_ids, stream);
}

int lis_index_search_sharded(lis_index* ix, lis_comm* comm, const void* q, int64_t q_rows, const int32_t* seg_lo,
                             const int32_t* seg_hi, const int32_t* mt_seg, int64_t n_seg, int64_t n_mtiles,
                             const int32_t* seg_first, int64_t nq, int round_mode, int k, float* out_scores,
                             int64_t* out_ids, void* stream) {
  LIS_REQUIRE(ix && q && seg_lo && seg_hi && mt_seg && out_scores && out_ids, "lis_index_search_sharded: null pointer");
  LIS_REQUIRE(nq > 0 && n_seg > 0 && n_mtiles > 0 && q_rows > 0, "bad query counts nq=%lld n_seg=%lld", (long long)nq, (long long)n_seg);
  LIS_REQUIRE(seg_first != nullptr || n_seg == nq, "split or empty queries need seg_first");
  LIS_REQUIRE(k >= 1 && k <= LIS_MAX_K, "k out of range");
  LIS_REQUIRE(round_mode == LIS_ROUND_F32 || round_mode == LIS_ROUND_REFERENCE, "bad round_mode");
  const int world = lis_comm_world(comm);
  LIS_REQUIRE(world > 1 || ix->n_pages > 0, "index is empty");
  const bool direct = seg_first == nullptr;
  const bool f32 = ix->dtype == LIS_F32X2;
  std::lock_guard<std::mutex> lock(ix->mu);
  DeviceGuard device_guard(ix->device);
  if (!ix->sstream) {
    LIS_CUDA_CHECK(cudaStreamCreateWithFlags(&ix->sstream, cudaStreamNonBlocking));
    LIS_CUDA_CHECK(cudaEventCreateWithFlags(&ix->sevent, cudaEventDisableTiming));
  }
  cudaStream_t st = ix->sstream;

  // staging layout: seg_lo | seg_hi | mt_seg | seg_first | (pad to 256) | query rows [| hi plane | lo plane]
  const int64_t tab_ints = 2 * n_seg + (n_mtiles + 1) + (nq + 1);
  const int64_t q_off = al256(tab_ints * 4);
  const int64_t q_bytes = q_rows * (f32 ? 512 : 256);
  const int64_t hi_off = al256(q_off + q_bytes), lo_off = hi_off + al256(q_rows * 256);
  const int64_t d_need = f32 ? lo_off + al256(q_rows * 256) : q_off + al256(q_bytes);
  const int64_t s_bytes = al256(nq * (int64_t)k * 4);
  const int64_t cand_bytes = s_bytes + al256(nq * (int64_t)k * 8);      // scores | ids, one rank's block

  cudaPointerAttributes pa;
  bool q_dev = false;
  if (cudaPointerGetAttributes(&pa, q) == cudaSuccess) q_dev = pa.type == cudaMemoryTypeDevice || pa.type == cudaMemoryTypeManaged;
  else cudaGetLastError();

  int rc = ensure_host(&ix->h_stage, &ix->h_stage_bytes, q_off + q_bytes);
  if (rc) return rc;
  const void* old_ptrs[4] = {ix->h_stage, ix->h_out, nullptr, nullptr};
  rc = ensure_host(&ix->h_out, &ix->h_out_bytes, cand_bytes);
  if (rc) return rc;
  if (old_ptrs[0] != ix->h_stage || old_ptrs[1] != ix->h_out) drop_graphs(ix);
  rc = ensure(ix, (void**)&ix->d_stage, &ix->d_stage_bytes, d_need);
  if (rc) return rc;
  rc = ensure(ix, (void**)&ix->d_out, &ix->d_out_bytes, cand_bytes);
  if (rc) return rc;
  if (world > 1) {
    rc = ensure(ix, (void**)&ix->sendbuf, &ix->sendbuf_bytes, cand_bytes);
    if (rc) return rc;
    rc = ensure(ix, (void**)&ix->recvbuf, &ix->recvbuf_bytes, cand_bytes * world);
    if (rc) return rc;
  }
  rc = prepare_scratch(ix, n_seg, nq, k, direct, world);
  if (rc) return rc;

  // host side of the upload
  int32_t* ht = reinterpret_cast<int32_t*>(ix->h_stage);
  memcpy(ht, seg_lo, (size_t)n_seg * 4);
  memcpy(ht + n_seg, seg_hi, (size_t)n_seg * 4);
  memcpy(ht + 2 * n_seg, mt_seg, (size_t)(n_mtiles + 1) * 4);
  if (seg_first) memcpy(ht + 2 * n_seg + n_mtiles + 1, seg_first, (size_t)(nq + 1) * 4);
  if (q_dev) {
    // ordering against the stream that produced q, then a device-to-device copy into the fixed staging
    LIS_CUDA_CHECK(cudaEventRecord(ix->sevent, (cudaStream_t)stream));
    LIS_CUDA_CHECK(cudaStreamWaitEvent(st, ix->sevent, 0));
    LIS_CUDA_CHECK(cudaMemcpyAsync(ix->d_stage + q_off, q, (size_t)q_bytes, cudaMemcpyDeviceToDevice, st));
  } else {
    memcpy(ix->h_stage + q_off, q, (size_t)q_bytes);
  }
  const int64_t up_bytes = q_dev ? tab_ints * 4 : q_off + q_bytes;

  const int32_t* d_tab = reinterpret_cast<const int32_t*>(ix->d_stage);
  const int32_t* d_seg_lo = d_tab;
  const int32_t* d_seg_hi = d_tab + n_seg;
  const int32_t* d_mt_seg = d_tab + 2 * n_seg;
  const int32_t* d_seg_first = direct ? nullptr : d_tab + 2 * n_seg + n_mtiles + 1;
  const void* d_q = ix->d_stage + q_off;
  const void* d_q_lo = nullptr;
  float* fin_s = reinterpret_cast<float*>(ix->d_out);
  int64_t* fin_i = reinterpret_cast<int64_t*>(ix->d_out + s_bytes);

  // the device sequence; runs eagerly the first time a shape is seen, is captured once, and replayed afterwards
  auto enqueue = [&]() -> int {
    LIS_CUDA_CHECK(cudaMemcpyAsync(ix->d_stage, ix->h_stage, (size_t)up_bytes, cudaMemcpyHostToDevice, st));
    if (f32) {
      int r = lis_split_f32(reinterpret_cast<const float*>(ix->d_stage + q_off), q_rows, ix->d_stage + hi_off,
                            ix->d_stage + lo_off, st);
      if (r) return r;
    }
    const void* qq = f32 ? ix->d_stage + hi_off : d_q;
    const void* ql = f32 ? ix->d_stage + lo_off : d_q_lo;
    float* loc_s = world > 1 ? reinterpret_cast<float*>(ix->sendbuf) : fin_s;
    int64_t* loc_i = world > 1 ? reinterpret_cast<int64_t*>(ix->sendbuf + s_bytes) : fin_i;
    int r;
    if (ix->n_pages > 0) {
      r = enqueue_local_search(ix, qq, ql, q_rows, d_seg_lo, d_seg_hi, d_mt_seg, n_seg, n_mtiles, d_seg_first, nq,
                               round_mode, k, loc_s, loc_i, st);
      if (r) return r;
    } else {   // an empty shard contributes padding, and still takes part in the collective
      fill_padding_kernel<<<(unsigned)std::min<int64_t>((nq * k + 255) / 256, 1024), 256, 0, st>>>(loc_s, loc_i, nq * k);
      count_launch();
      LIS_CUDA_CHECK(cudaGetLastError());
    }
    if (world > 1) {
      r = comm_all_gather(comm, ix->sendbuf, ix->recvbuf, (size_t)cand_bytes, st);
      if (r) return r;
      r = run_tournament(reinterpret_cast<const float*>(ix->recvbuf), k,
                         reinterpret_cast<const int64_t*>(ix->recvbuf + s_bytes), k, 0, nq, (int64_t)world * k, k, fin_s,
                         fin_i, ix->topk_ws, ix->topk_ws_bytes, st, k, cand_bytes / 4, cand_bytes / 8);
      if (r) return r;
    }
    LIS_CUDA_CHECK(cudaMemcpyAsync(ix->h_out, ix->d_out, (size_t)cand_bytes, cudaMemcpyDeviceToHost, st));
    return LIS_OK;
  };

  if (world > 1 && ix->graph_comm != nullptr && ix->graph_comm != comm) drop_graphs(ix);   // one communicator per cache
  const Tuning tn = tuning_snapshot();
  const int tune_sig = tn.tile_n * 1000003 + tn.group * 10007 + tn.max_ctas * 101 + tn.epi_halves * 17 + tn.a_operand * 5 + tn.ablate;
  const GraphKey key(q_rows, n_seg, n_mtiles, nq, ix->n_pages, ix->n_rows, k, round_mode, world, (direct ? 1 : 0) | (q_dev ? 2 : 0),
                     tune_sig, ix->gen);
  auto it = ix->graphs.find(key);
  if (it != ix->graphs.end()) {
    LIS_CUDA_CHECK(cudaGraphLaunch(it->second.exec, st));
    count_launch(it->second.kernels);
    ++ix->graph_replays;
  } else {
    rc = enqueue();      // eager: also performs every first-use initialisation (function attributes, NCCL channels)
    if (rc) return rc;
    LIS_CUDA_CHECK(cudaStreamSynchronize(st));
    // capture the same sequence for the next call of this shape (capturing executes nothing)
    if (ix->graphs.size() >= 64) drop_graphs(ix);
    const GraphKey key2(q_rows, n_seg, n_mtiles, nq, ix->n_pages, ix->n_rows, k, round_mode, world,
                        (direct ? 1 : 0) | (q_dev ? 2 : 0), tune_sig, ix->gen);
    cudaGraph_t graph = nullptr;
    LIS_CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    const int64_t launches_before = lis_launch_count();
    rc = enqueue();
    const int kernels = (int)(lis_launch_count() - launches_before);
    g_launches.fetch_add(-(int64_t)kernels, std::memory_order_relaxed);   // captured, not run
    cudaError_t ce = cudaStreamEndCapture(st, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) { set_error("graph capture of the search failed: %s", cudaGetErrorString(ce)); return LIS_E_CUDA; }
    cudaGraphExec_t exec = nullptr;
    ce = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) { set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(ce)); return LIS_E_CUDA; }
    ix->graphs[key2] = lis_index::Replay{exec, kernels};
    if (world > 1) ix->graph_comm = comm;
    ++ix->graph_captures;
    memcpy(out_scores, ix->h_out, (size_t)nq * k * 4);
    memcpy(out_ids, ix->h_out + s_bytes, (size_t)nq * k * 8);
    return LIS_OK;
  }
  LIS_CUDA_CHECK(cudaStreamSynchronize(st));
  memcpy(out_scores, ix->h_out, (size_t)nq * k * 4);
  memcpy(out_ids, ix->h_out + s_bytes, (size_t)nq * k * 8);
  return LIS_OK;
}

// ---- raw row planes <-> files, through a pinned double buffer ---------------------------------------------
// Each chunk is read by `io_threads` pread()s in parallel (page cache / NVMe queue depth) into pinned memory and
// DMA'd while the next chunk is being read; saving is the mirror image.  No intermediate pageable copy.
static int file_rows_io(lis_index* ix, bool load, int plane, int64_t row0, int64_t n_rows, const char* path,
                        int64_t file_offset, int io_threads, cudaStream_t st) {
  LIS_REQUIRE(ix && path, "null pointer");
  LIS_REQUIRE(plane == 0 || (plane == 1 && ix->dtype == LIS_F32X2), "no such plane");
  LIS_REQUIRE(row0 >= 0 && n_rows >= 0 && row0 + n_rows <= (load ? ix->cap_rows : ix->n_rows), "row range out of bounds");
  LIS_REQUIRE(file_offset >= 0, "negative file offset");
  if (n_rows == 0) return LIS_OK;
  if (io_threads <= 0) io_threads = (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
  LIS_CUDA_CHECK(cudaSetDevice(ix->device));
  const int fd = load ? open(path, O_RDONLY) : open(path, O_WRONLY | O_CREAT, 0644);
  if (fd < 0) {
    set_error("cannot open %s: %s", path, strerror(errno));
    return LIS_E_INVALID;
  }
  constexpr int64_t kChunk = int64_t(64) << 20;     // 64 MiB
  uint8_t* pin[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  int rc = LIS_OK;
  for (int i = 0; i < 2 && rc == LIS_OK; ++i) {
    if (cudaHostAlloc((void**)&pin[i], (size_t)kChunk, cudaHostAllocDefault) != cudaSuccess ||
        cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) {
      set_error("file i/o: cannot allocate the pinned staging buffers");
      rc = LIS_E_NOMEM;
    }
  }
  uint8_t* dbase = (plane ? lo_plane(ix) : static_cast<uint8_t*>(ix->tokens)) + row0 * 256;
  const int64_t total = n_rows * 256;
  std::atomic<int> io_err{0};
  auto parallel_io = [&](uint8_t* buf, int64_t off, int64_t bytes) {
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(io_threads, bytes >> 20));
    const int64_t per = ((bytes + nt - 1) / nt + 4095) & ~int64_t(4095);
    auto work = [&](int64_t a, int64_t b) {
      while (a < b) {
        const ssize_t n = load ? pread(fd, buf + a, (size_t)(b - a), (off_t)(file_offset + off + a))
                               : pwrite(fd, buf + a, (size_t)(b - a), (off_t)(file_offset + off + a));
        if (n <= 0) { io_err.store(n == 0 ? EIO : errno); return; }
        a += n;
      }
    };
    if (nt == 1) { work(0, bytes); return; }
    std::vector<std::thread> ts;
    for (int t = 0; t < nt; ++t) {
      const int64_t a = t * per, b = std::min(bytes, a + per);
      if (a < b) ts.emplace_back(work, a, b);
    }
    for (auto& t : ts) t.join();
  };
  bool used[2] = {false, false};
  cudaError_t ce = cudaSuccess;
  if (rc == LIS_OK) {
    if (load) {
      int c = 0;
      for (int64_t off = 0; off < total && ce == cudaSuccess && !io_err.load(); off += kChunk, ++c) {
        const int b = c & 1;
        const int64_t bytes = std::min(kChunk, total - off);
        if (used[b]) ce = cudaEventSynchronize(ev[b]);     // the previous upload from this buffer has finished
        parallel_io(pin[b], off, bytes);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(dbase + off, pin[b], (size_t)bytes, cudaMemcpyHostToDevice, st);
        if (ce == cudaSuccess) ce = cudaEventRecord(ev[b], st);
        used[b] = true;
      }
    } else {
      // download chunk c+1 while chunk c is being written
      const int64_t nchunks = (total + kChunk - 1) / kChunk;
      auto start = [&](int64_t c) {
        const int b = (int)(c & 1);
        const int64_t off = c * kChunk, bytes = std::min(kChunk, total - off);
        ce = cudaMemcpyAsync(pin[b], dbase + off, (size_t)bytes, cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess) ce = cudaEventRecord(ev[b], st);
      };
      start(0);
      for (int64_t c = 0; c < nchunks && ce == cudaSuccess && !io_err.load(); ++c) {
        const int b = (int)(c & 1);
        if (c + 1 < nchunks) start(c + 1);
        if (ce == cudaSuccess) ce = cudaEventSynchronize(ev[b]);
        const int64_t off = c * kChunk;
        if (ce == cudaSuccess) parallel_io(pin[b], off, std::min(kChunk, total - off));
      }
    }
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
  }
  for (int i = 0; i < 2; ++i) {
    if (pin[i]) cudaFreeHost(pin[i]);
    if (ev[i]) cudaEventDestroy(ev[i]);
  }
  close(fd);
  if (rc) return rc;
  if (io_err.load()) {
    set_error("%s %s failed: %s", load ? "reading" : "writing", path, strerror(io_err.load()));
    return LIS_E_INVALID;
  }
  if (ce != cudaSuccess) {
    set_error("file i/o copy failed: %s", cudaGetErrorString(ce));
    return LIS_E_CUDA;
  }
  return LIS_OK;
}

int lis_index_load_rows(lis_index* ix, int plane, int64_t row0, int64_t n_rows, const char* path, int64_t file_offset,
                        int io_threads, void* stream) {
  return file_rows_io(ix, true, plane, row0, n_rows, path, file_offset, io_threads, (cudaStream_t)stream);
}

int lis_index_save_rows(const lis_index* ix, int plane, int64_t row0, int64_t n_rows, const char* path,
                        int64_t file_offset, int io_threads, void* stream) {
  return file_rows_io(const_cast<lis_index*>(ix), false, plane, row0, n_rows, path, file_offset, io_threads,
                      (cudaStream_t)stream);
}

int lis_index_tombstone(lis_index* ix, int64_t page, void* stream) {
  LIS_REQUIRE(ix, "null index");
  LIS_REQUIRE(page >= 0 && page < ix->n_pages, "page %lld out of range", (long long)page);
  std::lock_guard<std::mutex> lock(ix->mu);
  const int64_t gone = -1;
  cudaStream_t st = (cudaStream_t)stream;
  LIS_CUDA_CHECK(cudaSetDevice(ix->device));
  LIS_CUDA_CHECK(cudaMemcpyAsync(ix->ids + page, &gone, 8, cudaMemcpyHostToDevice, st));
  LIS_CUDA_CHECK(cudaStreamSynchronize(st));
  return LIS_OK;
}

int64_t lis_index_graph_stats(const lis_index* ix, int64_t* captures, int64_t* replays) {
  if (!ix) return 0;
  if (captures) *captures = ix->graph_captures;
  if (replays) *replays = ix->graph_replays;
  return (int64_t)ix->graphs.size();
}

int lis_index_page_lens(const lis_index* ix, int64_t first, int64_t n, int32_t* lens_host, void* stream) {
  LIS_REQUIRE(ix && lens_host, "null pointer");
  LIS_REQUIRE(first >= 0 && n >= 0 && first + n <= ix->n_pages, "page range out of bounds");
  if (n == 0) return LIS_OK;
  std::vector<int64_t> off((size_t)n + 1);
  cudaStream_t st = (cudaStream_t)stream;
  LIS_CUDA_CHECK(cudaMemcpyAsync(off.data(), ix->offsets + first, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, st));
  LIS_CUDA_CHECK(cudaStreamSynchronize(st));
  for (int64_t i = 0; i < n; ++i) lens_host[i] = (int32_t)(off[i + 1] - off[i]);
  return LIS_OK;
}

int lis_index_add_projected(lis_index* ix, const void* hidden, int64_t n_pages, int64_t seq, int64_t hidden_dim,
                            const void* weight, const void* bias, const void* mask, int mask_itemsize, int round_mode,
                            const int64_t* ids, void* stream) {
  LIS_REQUIRE(ix && hidden && weight && mask, "lis_index_add_projected: null pointer");
  LIS_REQUIRE(ix->dtype != LIS_F32X2, "lis_index_add_projected: the encoder head is 16-bit; not available for an f32x2 index");
  LIS_REQUIRE(mask_itemsize == 1 || mask_itemsize == 4 || mask_itemsize == 8, "mask_itemsize must be 1, 4 or 8");
  LIS_REQUIRE(n_pages > 0 && seq > 0 && n_pages * seq < (int64_t(1) << 31), "bad batch shape %lld x %lld", (long long)n_pages, (long long)seq);
  LIS_REQUIRE(ix->n_pages + n_pages <= ix->cap_pages, "page capacity exceeded: %lld + %lld > %lld", (long long)ix->n_pages,
              (long long)n_pages, (long long)ix->cap_pages);
  // worst case every token is kept; checked exactly after the lengths are known
  std::lock_guard<std::mutex> lock(ix->mu);
  cudaStream_t st = (cudaStream_t)stream;
  LIS_CUDA_CHECK(cudaSetDevice(ix->device));
  const int64_t n_tok = n_pages * seq;
  // scratch (reuses the search workspace: ingestion and search never overlap on one index -- the lock above)
  const int64_t need = al256(n_tok * 4) + al256(n_pages * 4) + al256(n_pages * 8);
  int rc = ensure(ix, &ix->topk_ws, &ix->topk_ws_bytes, need);
  if (rc) return rc;
  uint8_t* w = static_cast<uint8_t*>(ix->topk_ws);
  int32_t* dst_row = reinterpret_cast<int32_t*>(w);
  int32_t* lens = reinterpret_cast<int32_t*>(w + al256(n_tok * 4));
  int64_t* base = reinterpret_cast<int64_t*>(w + al256(n_tok * 4) + al256(n_pages * 4));
  page_prefix_kernel<<<(unsigned)n_pages, 256, 0, st>>>(mask, mask_itemsize, seq, dst_row, lens);
  page_tables_kernel<<<1, 256, 0, st>>>(lens, n_pages, seq, ix->n_rows, ix->offsets + ix->n_pages, ix->clamp + ix->n_pages, base);
  page_rebase_kernel<<<(unsigned)std::min<int64_t>((n_tok + 255) / 256, 4096), 256, 0, st>>>(dst_row, base, seq, n_tok);
  count_launch(3);
  LIS_CUDA_CHECK(cudaGetLastError());
  int64_t new_end = 0;
  LIS_CUDA_CHECK(cudaMemcpyAsync(&new_end, ix->offsets + ix->n_pages + n_pages, 8, cudaMemcpyDeviceToHost, st));
  LIS_CUDA_CHECK(cudaStreamSynchronize(st));
  if (new_end > ix->cap_rows) {
    // the tables past n_pages are not part of the index until the counters move: nothing to undo
    set_error("row capacity exceeded: %lld > %lld", (long long)new_end, (long long)ix->cap_rows);
    return LIS_E_INVALID;
  }
  std::vector<int64_t> idv((size_t)n_pages);
  for (int64_t i = 0; i < n_pages; ++i) {
    idv[i] = ids ? ids[i] : ix->n_pages + i;
    LIS_REQUIRE(idv[i] >= 0, "page ids must be non-negative");
  }
  LIS_CUDA_CHECK(cudaMemcpyAsync(ix->ids + ix->n_pages, idv.data(), (size_t)n_pages * 8, cudaMemcpyHostToDevice, st));
  rc = lis_project_normalize(hidden, n_tok, hidden_dim, weight, bias, mask, mask_itemsize, ix->dtype, round_mode, dst_row,
                             ix->tokens, stream);
  if (rc) return rc;
  LIS_CUDA_CHECK(cudaStreamSynchronize(st));
  ix->n_rows = new_end;
  ix->n_pages += n_pages;
  return LIS_OK;
}

}  // extern "C"
