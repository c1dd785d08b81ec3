// K2: per-query top-k page selection, replacing `query_scores.topk(top_k)` (05_experiment02.py:219)
// and the `limit=` of the Qdrant query (functions.py:894-904), with the tie order torch leaves
// unspecified pinned to (score descending, id ascending) so results are reproducible across
// shard counts.
//
// Tournament of block-wide bitonic sorts: every CTA sorts a chunk of candidates in shared memory and
// keeps its best k; passes repeat on the survivors until one chunk remains.  HBM traffic is one read
// of the score row plus a geometrically shrinking candidate list.  The chunk is 1024 candidates (256
// threads, 55 compare-exchange stages) for k <= 128 -- the search path: four times the CTAs and 30 % fewer,
// cheaper stages than a 4096-sort -- and 4096 (512 threads) for larger k, where a chunk must stay >> k.
#include <algorithm>
#include <cmath>

#include "lis_common.h"

namespace lis {

constexpr int64_t kPadId = INT64_MAX;
static inline int chunk_for(int k) { return k <= 128 ? 1024 : 4096; }

__device__ __forceinline__ bool better(float sa, int64_t ia, float sb, int64_t ib) {
  return sa > sb || (sa == sb && ia < ib);
}

// One tournament pass.  Input row q: n candidates (score s[q*ld_s + i]; id = ids ? ids[q*ld_ids + i]
// : id_base + i; ld_ids == 0 shares one id row between queries).  Output: chunk c of row q writes its
// best k to (out_s, out_id)[q*out_ld + c*k ...]; with final != 0 padding is converted to (-inf, -1).
// seg_len > 0: the input row is cut into blocks of seg_len candidates that lie seg_stride_s floats / seg_stride_i
// int64 apart (the all-gathered per-rank [nq, k] candidate blocks: block w of row q starts at w * stride + q * ld).
template <int kChunk, int kTopkThreads>
__global__ void __launch_bounds__(kTopkThreads)
topk_pass_kernel(const float* __restrict__ s, int64_t ld_s, const int64_t* __restrict__ ids, int64_t ld_ids,
                 int64_t id_base, int64_t n, int k, float* __restrict__ out_s, int64_t* __restrict__ out_id,
                 int64_t out_ld, int final_pass, int seg_len, int64_t seg_stride_s, int64_t seg_stride_i) {
  extern __shared__ __align__(16) uint8_t topk_smem[];
  int64_t* sid = reinterpret_cast<int64_t*>(topk_smem);                    // [kChunk]
  float* ssc = reinterpret_cast<float*>(topk_smem + sizeof(int64_t) * kChunk);  // [kChunk]

  const int64_t q = blockIdx.y;
  const int64_t c = blockIdx.x;
  const int64_t i0 = c * kChunk;
  for (int i = threadIdx.x; i < kChunk; i += kTopkThreads) {
    const int64_t col = i0 + i;
    float sc = -INFINITY;
    int64_t id = kPadId;
    if (col < n) {
      int64_t so = q * ld_s + col, io = q * ld_ids + col;
      if (seg_len > 0) {
        const int64_t w = col / seg_len, j = col - w * seg_len;
        so = w * seg_stride_s + q * ld_s + j;
        io = w * seg_stride_i + q * ld_ids + j;
      }
      sc = __ldg(s + so);
      id = ids ? __ldg(ids + io) : id_base + col;
      if (sc != sc) sc = -INFINITY;          // NaN never wins
      if (id < 0) { sc = -INFINITY; id = kPadId; }  // padding from a short shard
    }
    ssc[i] = sc;
    sid[i] = id;
  }
  __syncthreads();
  // bitonic sort, best first
  for (int size = 2; size <= kChunk; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < kChunk / 2; t += kTopkThreads) {
        const int lo = 2 * t - (t & (stride - 1));  // index with bit `stride` clear
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;         // this subsequence sorts best-first
        const float sa = ssc[lo], sb = ssc[hi];
        const int64_t ia = sid[lo], ib = sid[hi];
        const bool swap = desc ? better(sb, ib, sa, ia) : better(sa, ia, sb, ib);
        if (swap) {
          ssc[lo] = sb; ssc[hi] = sa;
          sid[lo] = ib; sid[hi] = ia;
        }
      }
      __syncthreads();
    }
  }
  for (int j = threadIdx.x; j < k; j += kTopkThreads) {
    float sc = j < kChunk ? ssc[j] : -INFINITY;
    int64_t id = j < kChunk ? sid[j] : kPadId;
    if (final_pass && id == kPadId) { sc = -INFINITY; id = -1; }
    out_s[q * out_ld + c * k + j] = sc;
    out_id[q * out_ld + c * k + j] = id;
  }
}

static inline int64_t chunks_of(int64_t n, int chunk) { return (n + chunk - 1) / chunk; }
static inline int64_t align256(int64_t x) { return (x + 255) & ~int64_t(255); }

int run_tournament(const float* s, int64_t ld_s, const int64_t* ids, int64_t ld_ids, int64_t id_base,
                   int64_t nq, int64_t n, int k, float* out_s, int64_t* out_id, void* ws,
                   int64_t ws_bytes, cudaStream_t st, int seg_len, int64_t seg_stride_s, int64_t seg_stride_i) {
  LIS_REQUIRE(k >= 1 && k <= LIS_MAX_K, "k=%d out of range 1..%d", k, LIS_MAX_K);
  LIS_REQUIRE(nq >= 1 && nq <= 65535, "nq=%lld out of range", (long long)nq);
  LIS_REQUIRE(n >= 1, "no candidates");
  LIS_REQUIRE(s && out_s && out_id, "null pointer");
  int chunk = chunk_for(k);
  // small inputs (the reference's own corpus sizes: a few hundred pages): one CTA per query sorts the next power of two,
  // not 1024 -- 36 / 45 compare-exchange stages instead of 55, and cheaper barriers with fewer threads
  if (n <= 256 && k <= 256) chunk = 256;
  else if (n <= 512 && k <= 512) chunk = 512;
  const int smem = chunk * (int)(sizeof(int64_t) + sizeof(float));
  static std::atomic<bool> configured[64];   // zero-initialised; setting the attribute twice is harmless
  int dev = 0;
  LIS_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
    LIS_CUDA_CHECK(cudaFuncSetAttribute(topk_pass_kernel<4096, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        4096 * (int)(sizeof(int64_t) + sizeof(float))));
    if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_release);
  }
  // ping-pong buffers carved from the workspace
  const int64_t c0 = chunks_of(n, chunk);
  const int64_t n1 = c0 * k;                  // survivors of pass 0 (per query)
  const int64_t n2 = chunks_of(n1, chunk) * k;       // survivors of pass 1
  const int64_t bytes_a = c0 > 1 ? align256(nq * n1 * 4) + align256(nq * n1 * 8) : 0;
  const int64_t bytes_b = chunks_of(n1, chunk) > 1 && c0 > 1 ? align256(nq * n2 * 4) + align256(nq * n2 * 8) : 0;
  LIS_REQUIRE(ws_bytes >= bytes_a + bytes_b && (bytes_a + bytes_b == 0 || ws), "top-k workspace too small: %lld < %lld",
              (long long)ws_bytes, (long long)(bytes_a + bytes_b));
  uint8_t* w = static_cast<uint8_t*>(ws);
  float* buf_s[2] = {reinterpret_cast<float*>(w), reinterpret_cast<float*>(w + bytes_a)};
  int64_t* buf_i[2] = {reinterpret_cast<int64_t*>(w + align256(nq * n1 * 4)),
                       reinterpret_cast<int64_t*>(w + bytes_a + align256(nq * n2 * 4))};

  const float* in_s = s;
  const int64_t* in_i = ids;
  int64_t in_ld = ld_s, in_ldi = ld_ids, in_n = n, base = id_base;
  int pass = 0;
  while (true) {
    const int64_t nc = chunks_of(in_n, chunk);
    const bool last = nc == 1;
    float* o_s = last ? out_s : buf_s[pass & 1];
    int64_t* o_i = last ? out_id : buf_i[pass & 1];
    const int64_t o_ld = last ? k : nc * k;
    dim3 grid((unsigned)nc, (unsigned)nq);
    if (chunk == 256)
      topk_pass_kernel<256, 64><<<grid, 64, smem, st>>>(in_s, in_ld, in_i, in_ldi, base, in_n, k, o_s, o_i, o_ld, last ? 1 : 0,
                                                        seg_len, seg_stride_s, seg_stride_i);
    else if (chunk == 512)
      topk_pass_kernel<512, 128><<<grid, 128, smem, st>>>(in_s, in_ld, in_i, in_ldi, base, in_n, k, o_s, o_i, o_ld, last ? 1 : 0,
                                                          seg_len, seg_stride_s, seg_stride_i);
    else if (chunk == 1024)
      topk_pass_kernel<1024, 256><<<grid, 256, smem, st>>>(in_s, in_ld, in_i, in_ldi, base, in_n, k, o_s, o_i, o_ld, last ? 1 : 0,
                                                           seg_len, seg_stride_s, seg_stride_i);
    else
      topk_pass_kernel<4096, 512><<<grid, 512, smem, st>>>(in_s, in_ld, in_i, in_ldi, base, in_n, k, o_s, o_i, o_ld, last ? 1 : 0,
                                                           seg_len, seg_stride_s, seg_stride_i);
    count_launch();
    LIS_CUDA_CHECK(cudaGetLastError());
    if (last) break;
    in_s = o_s; in_i = o_i; in_ld = o_ld; in_ldi = o_ld; in_n = o_ld; base = 0;
    seg_len = 0;   // survivors are plain rows
    ++pass;
  }
  return LIS_OK;
}

}  // namespace lis

using namespace lis;

extern "C" {

int64_t lis_topk_workspace_bytes(int64_t nq, int64_t np, int k) {
  if (nq < 1 || np < 1 || k < 1 || k > LIS_MAX_K) return 0;
  const int chunk = chunk_for(k);
  const int64_t c0 = chunks_of(np, chunk);
  if (c0 <= 1) return 256;
  const int64_t n1 = c0 * k;
  const int64_t n2 = chunks_of(n1, chunk) * k;
  int64_t bytes = align256(nq * n1 * 4) + align256(nq * n1 * 8);
  if (chunks_of(n1, chunk) > 1) bytes += align256(nq * n2 * 4) + align256(nq * n2 * 8);
  return bytes + 256;
}

int lis_topk(const float* scores, int64_t ld, int64_t nq, int64_t np, const int64_t* ids, int64_t id_base,
             int k, float* out_scores, int64_t* out_ids, void* workspace, int64_t workspace_bytes,
             void* stream) {
  LIS_REQUIRE(ld >= np, "lis_topk: ld < np");
  return run_tournament(scores, ld, ids, 0, id_base, nq, np, k, out_scores, out_ids, workspace,
                        workspace_bytes, (cudaStream_t)stream, 0, 0, 0);
}

int lis_merge_topk(const float* cand_scores, const int64_t* cand_ids, int64_t nq, int64_t n_cand, int k,
                   float* out_scores, int64_t* out_ids, void* workspace, int64_t workspace_bytes,
                   void* stream) {
  LIS_REQUIRE(cand_ids, "lis_merge_topk: cand_ids is null");
  return run_tournament(cand_scores, n_cand, cand_ids, n_cand, 0, nq, n_cand, k, out_scores, out_ids,
                        workspace, workspace_bytes, (cudaStream_t)stream, 0, 0, 0);
}

}  // extern "C"
