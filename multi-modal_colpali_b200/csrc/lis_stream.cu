// Surface 1 with the corpus in HOST memory, the way the reference calls it: `ps` is a CPU tensor (or a list of
// per-page CPU tensors) and colpali-engine moves it to the device 128 pages at a time inside the scoring loop
// (05_experiment02.py:213-214; HF processing_colpali.py:352-357).  Here the rows are cut into chunks of whole
// pages; chunk i+1 is gathered into a pinned buffer by a few host threads and DMA'd on a copy stream while K1
// scores chunk i on the compute stream.  Nothing is concatenated on the host and the corpus never has to fit HBM.
#include <algorithm>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "lis_common.h"

namespace lis {

struct StreamPool {          // per device; kept between calls (cudaHostAlloc of 2 x 128 MiB costs ~100 ms)
  uint8_t* h_tok[2] = {nullptr, nullptr};
  uint8_t* d_tok[2] = {nullptr, nullptr};
  int64_t tok_bytes = 0;
  int64_t* h_off[2] = {nullptr, nullptr};
  int64_t* d_off[2] = {nullptr, nullptr};
  int64_t off_entries = 0;
  uint8_t* d_clamp = nullptr;
  int64_t clamp_bytes = 0;
  cudaStream_t copy = nullptr;
  cudaEvent_t copied[2] = {nullptr, nullptr}, scored[2] = {nullptr, nullptr};
  std::mutex mu;
};
static StreamPool g_pools[64];

static void pool_free(StreamPool& p) {
  for (int i = 0; i < 2; ++i) {
    if (p.h_tok[i]) cudaFreeHost(p.h_tok[i]);
    if (p.d_tok[i]) cudaFree(p.d_tok[i]);
    if (p.h_off[i]) cudaFreeHost(p.h_off[i]);
    if (p.d_off[i]) cudaFree(p.d_off[i]);
    p.h_tok[i] = p.d_tok[i] = nullptr;
    p.h_off[i] = p.d_off[i] = nullptr;
  }
  if (p.d_clamp) cudaFree(p.d_clamp);
  p.d_clamp = nullptr;
  p.tok_bytes = p.off_entries = p.clamp_bytes = 0;
}

static int pool_ensure(StreamPool& p, int64_t tok_bytes, int64_t off_entries, int64_t clamp_bytes, bool need_host) {
  if (!p.copy) {
    LIS_CUDA_CHECK(cudaStreamCreateWithFlags(&p.copy, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      LIS_CUDA_CHECK(cudaEventCreateWithFlags(&p.copied[i], cudaEventDisableTiming));
      LIS_CUDA_CHECK(cudaEventCreateWithFlags(&p.scored[i], cudaEventDisableTiming));
    }
  }
  if (p.tok_bytes < tok_bytes || (need_host && !p.h_tok[0])) {
    tok_bytes = std::max(tok_bytes, p.tok_bytes);
    for (int i = 0; i < 2; ++i) {
      if (p.h_tok[i]) cudaFreeHost(p.h_tok[i]);
      if (p.d_tok[i]) cudaFree(p.d_tok[i]);
      p.h_tok[i] = p.d_tok[i] = nullptr;
    }
    p.tok_bytes = 0;
    for (int i = 0; i < 2; ++i) {
      if ((need_host && cudaHostAlloc((void**)&p.h_tok[i], (size_t)tok_bytes, cudaHostAllocDefault) != cudaSuccess) ||
          cudaMalloc((void**)&p.d_tok[i], (size_t)tok_bytes) != cudaSuccess) {
        cudaGetLastError();
        set_error("lis_stream_scores: could not allocate the %lld-byte staging buffers", (long long)tok_bytes);
        return LIS_E_NOMEM;
      }
    }
    p.tok_bytes = tok_bytes;
  }
  if (p.off_entries < off_entries) {
    for (int i = 0; i < 2; ++i) {
      if (p.h_off[i]) cudaFreeHost(p.h_off[i]);
      if (p.d_off[i]) cudaFree(p.d_off[i]);
      p.h_off[i] = p.d_off[i] = nullptr;
    }
    p.off_entries = 0;
    for (int i = 0; i < 2; ++i) {
      if (cudaHostAlloc((void**)&p.h_off[i], (size_t)off_entries * 8, cudaHostAllocDefault) != cudaSuccess ||
          cudaMalloc((void**)&p.d_off[i], (size_t)off_entries * 8) != cudaSuccess) {
        cudaGetLastError();
        set_error("lis_stream_scores: could not allocate the page tables");
        return LIS_E_NOMEM;
      }
    }
    p.off_entries = off_entries;
  }
  if (p.clamp_bytes < clamp_bytes) {
    if (p.d_clamp) cudaFree(p.d_clamp);
    p.d_clamp = nullptr;
    p.clamp_bytes = 0;
    if (cudaMalloc((void**)&p.d_clamp, (size_t)clamp_bytes) != cudaSuccess) {
      cudaGetLastError();
      set_error("lis_stream_scores: could not allocate the clamp flags");
      return LIS_E_NOMEM;
    }
    p.clamp_bytes = clamp_bytes;
  }
  return LIS_OK;
}

// Copy pages [p0, p1) into dst back to back with `threads` host threads (contiguous source or one pointer per page).
static void gather_pages(uint8_t* dst, const uint8_t* tokens_host, const void* const* page_ptrs, const int64_t* off,
                         int64_t p0, int64_t p1, int threads) {
  const int64_t row0 = off[p0], rows = off[p1] - row0;
  if (rows <= 0) return;
  auto work = [&](int64_t ra, int64_t rb) {          // chunk-local row range [ra, rb)
    if (!page_ptrs) {
      memcpy(dst + ra * 256, tokens_host + (row0 + ra) * 256, (size_t)(rb - ra) * 256);
      return;
    }
    // first page touching row ra
    int64_t p = std::upper_bound(off + p0, off + p1 + 1, row0 + ra) - off - 1;
    for (int64_t r = ra; r < rb;) {
      while (off[p + 1] - row0 <= r) ++p;            // skips empty pages
      const int64_t in_page = row0 + r - off[p];
      const int64_t n = std::min<int64_t>(rb - r, off[p + 1] - off[p] - in_page);
      memcpy(dst + r * 256, static_cast<const uint8_t*>(page_ptrs[p]) + in_page * 256, (size_t)n * 256);
      r += n;
    }
  };
  threads = (int)std::max<int64_t>(1, std::min<int64_t>(threads, rows / 4096));
  if (threads == 1) { work(0, rows); return; }
  std::vector<std::thread> pool;
  const int64_t per = (rows + threads - 1) / threads;
  for (int t = 0; t < threads; ++t) {
    const int64_t a = t * per, b = std::min(rows, a + per);
    if (a < b) pool.emplace_back(work, a, b);
  }
  for (auto& th : pool) th.join();
}

}  // namespace lis

using namespace lis;

extern "C" {

int lis_stream_scores(const void* q, int64_t q_rows, const int32_t* seg_lo, const int32_t* seg_hi, const int32_t* mt_seg,
                      int64_t n_seg, int64_t n_mtiles, const void* tokens_host, const void* const* page_ptrs_host,
                      int64_t n_rows, const int64_t* p_offsets_host, const uint8_t* p_clamp_host, int64_t np, int dtype,
                      int round_mode, float* out, int64_t ld_out, int64_t chunk_rows, int host_threads, void* stream) {
  LIS_REQUIRE(q && seg_lo && seg_hi && mt_seg && p_offsets_host && out, "lis_stream_scores: null pointer");
  LIS_REQUIRE((tokens_host != nullptr) != (page_ptrs_host != nullptr) || n_rows == 0,
              "lis_stream_scores: pass either one contiguous token matrix or one pointer per page");
  LIS_REQUIRE(np > 0 && n_rows >= 0 && p_offsets_host[0] == 0 && p_offsets_host[np] == n_rows,
              "lis_stream_scores: offsets must run from 0 to n_rows");
  LIS_REQUIRE(ld_out >= np, "ld_out < np");
  LIS_REQUIRE(chunk_rows >= 0, "chunk_rows < 0");
  // defaults from profiles/host_corpus_sweep_r2.jsonl: 64 MiB chunks gathered by up to 16 threads (0.85 of the PCIe floor from
  // pageable memory; 128 MiB / 8 threads: 0.5-0.7)
  if (chunk_rows == 0) chunk_rows = int64_t(1) << 18;
  if (host_threads <= 0) host_threads = (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
  int64_t max_page = 0;
  for (int64_t p = 0; p < np; ++p) {
    LIS_REQUIRE(p_offsets_host[p + 1] >= p_offsets_host[p], "offsets must be ascending (page %lld)", (long long)p);
    max_page = std::max(max_page, p_offsets_host[p + 1] - p_offsets_host[p]);
  }
  const int64_t buf_rows = std::max(chunk_rows, max_page);
  LIS_REQUIRE(buf_rows < (int64_t(1) << 31), "chunk too large");
  int dev = 0;
  LIS_CUDA_CHECK(cudaGetDevice(&dev));
  LIS_REQUIRE(dev >= 0 && dev < 64, "device index out of range");
  StreamPool& pool = g_pools[dev];
  std::lock_guard<std::mutex> lock(pool.mu);
  // a pinned (or registered) contiguous source is DMA'd from where it lies; anything else goes through the pinned pair
  bool src_pinned = false;
  if (tokens_host) {
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, tokens_host) == cudaSuccess) src_pinned = pa.type == cudaMemoryTypeHost;
    else cudaGetLastError();
  }
  int rc = pool_ensure(pool, std::max<int64_t>(buf_rows, 1) * 256, np + 1, np, !src_pinned);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const uint8_t* d_clamp = nullptr;
  if (p_clamp_host) {
    LIS_CUDA_CHECK(cudaMemcpyAsync(pool.d_clamp, p_clamp_host, (size_t)np, cudaMemcpyHostToDevice, st));
    d_clamp = pool.d_clamp;
  }
  // chunks of whole pages
  int64_t p0 = 0;
  int c = 0;
  bool used[2] = {false, false};
  while (p0 < np) {
    int64_t p1 = p0 + 1;
    while (p1 < np && p_offsets_host[p1 + 1] - p_offsets_host[p0] <= buf_rows) ++p1;
    const int b = c & 1;
    const int64_t rows = p_offsets_host[p1] - p_offsets_host[p0], pages = p1 - p0;
    // the pinned buffer is free once its previous upload has finished; the device buffer once K1 has read it
    if (used[b]) LIS_CUDA_CHECK(cudaEventSynchronize(pool.copied[b]));
    const uint8_t* src = pool.h_tok[b];
    if (src_pinned) src = static_cast<const uint8_t*>(tokens_host) + p_offsets_host[p0] * 256;
    else gather_pages(pool.h_tok[b], static_cast<const uint8_t*>(tokens_host), page_ptrs_host, p_offsets_host, p0, p1, host_threads);
    for (int64_t i = 0; i <= pages; ++i) pool.h_off[b][i] = p_offsets_host[p0 + i] - p_offsets_host[p0];
    if (used[b]) LIS_CUDA_CHECK(cudaStreamWaitEvent(pool.copy, pool.scored[b], 0));
    if (rows > 0)
      LIS_CUDA_CHECK(cudaMemcpyAsync(pool.d_tok[b], src, (size_t)rows * 256, cudaMemcpyHostToDevice, pool.copy));
    LIS_CUDA_CHECK(cudaMemcpyAsync(pool.d_off[b], pool.h_off[b], (size_t)(pages + 1) * 8, cudaMemcpyHostToDevice, pool.copy));
    LIS_CUDA_CHECK(cudaEventRecord(pool.copied[b], pool.copy));
    LIS_CUDA_CHECK(cudaStreamWaitEvent(st, pool.copied[b], 0));
    rc = lis_maxsim_scores(q, q_rows, seg_lo, seg_hi, mt_seg, n_seg, n_mtiles, pool.d_tok[b], rows, pool.d_off[b],
                           d_clamp ? d_clamp + p0 : nullptr, pages, dtype, round_mode, out + p0, ld_out, stream);
    if (rc) { cudaStreamSynchronize(st); cudaStreamSynchronize(pool.copy); return rc; }
    LIS_CUDA_CHECK(cudaEventRecord(pool.scored[b], st));
    used[b] = true;
    p0 = p1;
    ++c;
  }
  LIS_CUDA_CHECK(cudaStreamSynchronize(st));
  return LIS_OK;
}

int lis_memcpy2d_async(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t width_bytes, int64_t height,
                       void* stream) {
  LIS_REQUIRE(dst && src, "lis_memcpy2d_async: null pointer");
  LIS_REQUIRE(width_bytes >= 0 && height >= 0 && dst_pitch >= width_bytes && src_pitch >= width_bytes, "lis_memcpy2d_async: bad shape");
  if (width_bytes == 0 || height == 0) return LIS_OK;
  LIS_CUDA_CHECK(cudaMemcpy2DAsync(dst, (size_t)dst_pitch, src, (size_t)src_pitch, (size_t)width_bytes, (size_t)height,
                                   cudaMemcpyDefault, (cudaStream_t)stream));
  return LIS_OK;
}

void lis_stream_release(void) {
  int cur = 0;
  cudaGetDevice(&cur);
  for (int d = 0; d < 64; ++d) {
    StreamPool& p = g_pools[d];
    std::lock_guard<std::mutex> lock(p.mu);
    if (p.tok_bytes == 0 && p.off_entries == 0 && p.clamp_bytes == 0) continue;
    cudaSetDevice(d);
    pool_free(p);
  }
  cudaSetDevice(cur);
}

}  // extern "C"
