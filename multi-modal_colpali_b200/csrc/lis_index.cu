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
#include <algorithm>
#include <vector>

#include "lis_common.h"
#include "lis_ptx.cuh"

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

static int ensure(void** p, int64_t* have, int64_t need) {
  if (*have >= need) return LIS_OK;
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
  *out = ix;
  return LIS_OK;
}

void lis_index_destroy(lis_index* ix) {
  if (!ix) return;
  cudaFree(ix->tokens);
  cudaFree(ix->offsets);
  cudaFree(ix->ids);
  cudaFree(ix->clamp);
  cudaFree(ix->seg_scores);
  cudaFree(ix->q_scores);
  cudaFree(ix->topk_ws);
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
  cudaStream_t st = (cudaStream_t)stream;
  LIS_CUDA_CHECK(cudaSetDevice(ix->device));
  int64_t new_rows = 0;
  int rc = append_tables(ix, lens, 0, ids, ix->n_pages, clamp, n, &new_rows, st);
  if (rc) return rc;
  if (new_rows > 0) {
    LIS_REQUIRE(tokens, "lis_index_add: null tokens");
    uint8_t* hi = static_cast<uint8_t*>(ix->tokens) + ix->n_rows * 256;
    if (ix->dtype == LIS_F32X2) {
      float* stage = nullptr;  // fp32 rows land in a staging buffer, then get split into the planes
      LIS_CUDA_CHECK(cudaMalloc((void**)&stage, (size_t)new_rows * 512));
      cudaError_t e = cudaMemcpyAsync(stage, tokens, (size_t)new_rows * 512, cudaMemcpyDefault, st);
      int rc2 = e == cudaSuccess ? lis_split_f32(stage, new_rows, hi, lo_plane(ix) + ix->n_rows * 256, stream) : LIS_E_CUDA;
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      cudaFree(stage);
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
    LIS_REQUIRE(ids[i] >= 0, "page ids must be non-negative");
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

int lis_index_search(lis_index* ix, const void* q, const void* q_lo, int64_t q_rows, const int32_t* seg_lo, const int32_t* seg_hi,
                     const int32_t* mt_seg, int64_t n_seg, int64_t n_mtiles, const int32_t* seg_first, int64_t nq,
                     int round_mode, int k, float* out_scores, int64_t* out_ids, void* stream) {
  LIS_REQUIRE(ix, "null index");
  LIS_REQUIRE(ix->n_pages > 0, "index is empty");
  LIS_REQUIRE(nq > 0 && n_seg >= nq, "bad query counts nq=%lld n_seg=%lld", (long long)nq, (long long)n_seg);
  LIS_REQUIRE(seg_first != nullptr || n_seg == nq, "split queries need seg_first");
  LIS_REQUIRE(k >= 1 && k <= LIS_MAX_K, "k out of range");
  const int64_t np = ix->n_pages;
  int rc = ensure((void**)&ix->seg_scores, &ix->seg_scores_bytes, n_seg * np * 4);
  if (rc) return rc;
  const int64_t ws_need = lis_topk_workspace_bytes(nq, np, k);
  rc = ensure(&ix->topk_ws, &ix->topk_ws_bytes, ws_need);
  if (rc) return rc;
  const int k1_round = (n_seg != nq) ? (round_mode | LIS_ROUND_DEFER_SUM) : round_mode;
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
  if (n_seg != nq) {
    rc = ensure((void**)&ix->q_scores, &ix->q_scores_bytes, nq * np * 4);
    if (rc) return rc;
    rc = lis_reduce_segments(ix->seg_scores, np, seg_first, nq, np, round_mode, elem_dtype(ix), ix->q_scores, np, stream);
    if (rc) return rc;
    scores = ix->q_scores;
  }
  return lis_topk(scores, np, nq, np, ix->ids, 0, k, out_scores, out_ids, ix->topk_ws, ix->topk_ws_bytes, stream);
}

}  // extern "C"
