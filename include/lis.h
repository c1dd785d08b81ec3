/* lis.h -- C-ABI of the B200-native late-interaction (MaxSim) scoring engine.
 *
 * Plain C: pointers, sizes, an opaque index handle.  No torch / C++ types cross this line.
 * Every pointer documented "device" is a CUDA device pointer owned by the caller; `stream`
 * is a `cudaStream_t` passed as `void*` (NULL = default stream).  All calls are stream-ordered
 * and asynchronous with respect to the host unless stated otherwise.
 *
 * Return value of every function that returns `int`: 0 on success, a negative LIS_E_* code on
 * failure; `lis_last_error()` then returns a thread-local human-readable message.
 *
 * Reference interfaces each entry point replaces (paths relative to the reference tree):
 *   - lis_maxsim_scores      <- processor.score_multi_vector(qs, ps)         05_experiment02.py:214
 *                               (body: colpali-engine 0.3.13, == HF processing_colpali.py:350-364)
 *   - lis_topk               <- query_scores.topk(top_k)                     05_experiment02.py:219
 *   - lis_index_*            <- Qdrant multivector collection (MAX_SIM)      01_create_context_qdrant.py:208-222,
 *                               upsert functions.py:865, query_points functions.py:894-926
 *   - lis_project_normalize  <- retrieval head inside model(**batch)         functions.py:795,839,888
 *                               (body: HF modeling_colpali.py:148-155)
 *   - lis_merge_topk         <- (no reference equivalent; merges per-GPU candidates after the allgather)
 *   - lis_comm_*, lis_index_search_sharded
 *                            <- (no reference equivalent: the reference is single-GPU, functions.py:1472; this is
 *                               north_star's "each GPU scores its shard ... single NCCL allgather" search)
 *   - lis_index_add_projected <- the ingestion loop functions.py:838-865 (model head -> tolist() -> HTTP upsert)
 *   - lis_stream_scores      <- score_multi_vector with the corpus on the HOST, as the reference calls it
 *                               (05_experiment02.py:213-214: ps is a CPU tensor, moved per 128-page block)
 */
#ifndef LIS_H_
#define LIS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LIS_ABI_VERSION 2

/* embedding width: VECTOR_SIZE = 128 (01_create_context_qdrant.py:70) */
#define LIS_DIM 128
/* query-token rows per tensor-core M tile */
#define LIS_MTILE 128

enum lis_status {
  LIS_OK = 0,
  LIS_E_INVALID = -1,     /* bad argument */
  LIS_E_CUDA = -2,        /* CUDA runtime / driver error */
  LIS_E_UNSUPPORTED = -3, /* valid request this build cannot serve (e.g. not an sm_100 device) */
  LIS_E_NOMEM = -4,
  LIS_E_NCCL = -5         /* NCCL missing or a collective failed */
};

/* LIS_F32X2: fp32 embeddings held as two bf16 planes (hi = bf16(x), lo = bf16(x - hi)); index dtype only */
enum lis_dtype { LIS_BF16 = 0, LIS_F16 = 1, LIS_F32X2 = 2 };

/* Epilogue rounding.  LIS_ROUND_F32: fp32 max and fp32 sum (accuracy mode, checked at 1e-4
 * against the fp32-widened oracle).  LIS_ROUND_REFERENCE: reproduce what torch does to 16-bit
 * inputs -- per-token max rounded to the input dtype, summed in fp32, sum rounded to the input
 * dtype (checked against the reference's bf16 path at 1e-2). */
enum lis_round_mode { LIS_ROUND_F32 = 0, LIS_ROUND_REFERENCE = 1, LIS_ROUND_DEFER_SUM = 2 };

const char* lis_last_error(void);
int lis_abi_version(void);
/* 0 when device `device` can run the kernels (compute capability 10.x), else LIS_E_UNSUPPORTED. */
int lis_device_supported(int device);

/* ---------------------------------------------------------------------------------------------
 * Query packing (host-side, pure CPU).
 *
 * Queries are ragged lists of token rows, stored back to back ("packed rows": query q owns rows
 * [sum(len[<q]), sum(len[<=q]))).  The kernel reduces over rows inside one 128-row M tile, so a
 * query is cut into "segments" wherever it crosses a multiple of 64 rows (the CTA-pair kernel splits
 * the last M tile of an odd pass 64/64 rows over two SMs); the partial sums of a cut query are added
 * by lis_reduce_segments.  lis_plan_queries computes that segmentation.
 *   q_lens[nq]                tokens per query (>= 0; a 0-length query owns no segment, score 0)
 *   seg_query[cap]            out: owning query of each segment
 *   seg_lo[cap], seg_hi[cap]  out: packed row range [lo, hi) of the segment (inside one M tile)
 *   mt_seg[mt_cap]            out: mt_seg[t] .. mt_seg[t+1] = segments living in M tile t
 * Returns the number of segments (>= 0) or a negative status; *n_mtiles receives the tile count.
 * Call with cap = 0 to size the arrays (returns the segment count, fills *n_mtiles, writes nothing);
 * mt_cap must be >= *n_mtiles + 1.
 */
int64_t lis_plan_queries(const int32_t* q_lens, int64_t nq, int64_t cap, int32_t* seg_query,
                         int32_t* seg_lo, int32_t* seg_hi, int64_t mt_cap, int32_t* mt_seg,
                         int64_t* n_mtiles);

/* ---------------------------------------------------------------------------------------------
 * K1: fused MaxSim.  out[s, p] = sum over rows r of segment s of  max over tokens t of page p
 * of <q[r,:], tok[t,:]>.   Similarities never leave the SM (TMA -> smem -> tcgen05 -> TMEM ->
 * registers).
 *
 *   q            device [q_rows, 128] packed query rows, dtype; (n_mtiles-1)*128 < q_rows <= n_mtiles*128
 *                (rows past q_rows inside the last M tile are zero-filled by TMA)
 *   seg_lo/hi    device int32 [n_seg]        packed-row range of each segment
 *   mt_seg       device int32 [n_mtiles+1]   segment range of each M tile
 *   tokens       device [n_rows, 128]        page-token store, row-major, 16-byte aligned
 *   p_offsets    device int64 [np+1]         page p owns token rows [p_offsets[p], p_offsets[p+1])
 *                                            (ascending; rows are indices into `tokens`)
 *   p_clamp      device uint8 [np] or NULL   1 => clamp the per-token max of that page at >= 0.
 *                                            This is the reference's zero-padding semantics: a page
 *                                            shorter than the longest page of its 128-page block
 *                                            sees similarity 0 from the pad rows.
 *   out          device float [n_seg, ld_out], ld_out >= np
 */
int lis_maxsim_scores(const void* q, int64_t q_rows, const int32_t* seg_lo, const int32_t* seg_hi,
                      const int32_t* mt_seg, int64_t n_seg, int64_t n_mtiles, const void* tokens,
                      int64_t n_rows, const int64_t* p_offsets, const uint8_t* p_clamp, int64_t np,
                      int dtype, int round_mode, float* out, int64_t ld_out, void* stream);

/* How lis_maxsim_scores will cover n_mtiles query M tiles with the current tuning: one entry per pass over the
 * page store, +n = n tiles resident on one CTA per SM, -n = n tiles on CTA pairs.  Returns the number of passes
 * (writes at most `cap` entries; cap = 0 just counts).  Host-only; used by bench.py to state launches and bytes. */
int lis_maxsim_pass_plan(int64_t n_mtiles, int32_t* passes, int cap);

/* fp32 embeddings (ColFlor's default dtype, 05_experiment02.py:343-347) on the bf16 tensor pipe:
 * every operand is split into two bf16 planes, x = hi + lo, and a tile product is computed as
 * hi*hi + hi*lo + lo*hi in fp32 accumulators (error ~1e-6 on unit-norm rows).  lis_split_f32 produces
 * the planes of a [rows,128] fp32 matrix (device pointers, 16-byte aligned); the scoring call takes
 * the planes of the packed queries and of the token store and otherwise behaves like
 * lis_maxsim_scores with LIS_ROUND_F32 (torch computes fp32 inputs in fp32: nothing to emulate). */
int lis_split_f32(const float* src, int64_t rows, void* hi, void* lo, void* stream);
int lis_maxsim_scores_f32x2(const void* q_hi, const void* q_lo, int64_t q_rows, const int32_t* seg_lo,
                            const int32_t* seg_hi, const int32_t* mt_seg, int64_t n_seg, int64_t n_mtiles,
                            const void* tok_hi, const void* tok_lo, int64_t n_rows, const int64_t* p_offsets,
                            const uint8_t* p_clamp, int64_t np, float* out, int64_t ld_out, void* stream);

/* out[q, p] = sum of seg_scores[s, p] over the segments s of query q, in segment order.
 * Only needed when some query was cut (n_seg != nq); in that case pass round_mode |
 * LIS_ROUND_DEFER_SUM to lis_maxsim_scores so the final rounding of LIS_ROUND_REFERENCE happens
 * here, once, on the complete sum.  seg_first[nq+1] device int32. */
int lis_reduce_segments(const float* seg_scores, int64_t ld_seg, const int32_t* seg_first, int64_t nq,
                        int64_t np, int round_mode, int dtype, float* out, int64_t ld_out, void* stream);

/* Tuning / debug knobs (process-wide); 0 always means "auto", and the defaults are what ships.
 *   tile_n     {0, 128, 192, 256}  page-token rows per MMA tile
 *   group      {0, 1..10}          most query M tiles resident per pass over the page store (4..10: CTA pairs only)
 *   max_ctas   0 = one per SM
 *   epi_halves {0, 1, 2}           4 or 8 epilogue warps
 *   a_operand  {0, 1, 2, 3}        query operand of the MMA in shared memory (1) or tensor memory (2) on one
 *                                  CTA per SM, or 3 = CTA pairs (clusters of 2, tcgen05 cta_group::2: every page
 *                                  tile is loaded once per pair and shared by up to 10 query tiles).  Auto: the cheapest
 *                                  mix of passes by measured cost (one CTA per SM for 1-2 tiles, pairs from 3; see lis_set_pass_costs). */
int lis_set_tuning(int tile_n, int group, int max_ctas, int epi_halves, int a_operand);
/* The pass planner picks, for a batch of n query tiles, the cheapest mix of passes by the cost of one pass of every form
 * (single[1..3]: one CTA per SM with 1..3 resident tiles; pair[2..10]: CTA pairs with 2..10), in any common unit.  The
 * built-in table was measured on a power-capped B200 (profiles/pass_costs_r2.jsonl); a deployment can measure its own
 * (scoring.calibrate_pass_costs does) and install it here.  single has 4 entries, pair 11 (index = tile count; unused
 * leading entries are ignored); NULL restores the defaults for that form.  Process-wide. */
int lis_set_pass_costs(const float* single, const float* pair);
/* Timing experiments only (scores become invalid): 1 = K1's epilogue skips the TMEM read-out, 2 = it skips
 * the max arithmetic, 3 = the producer issues no TMA loads after the first ring fill (stale tiles are reused),
 * 4 = 3 and 1 together; 0 restores normal operation.  Used by scripts/gpu_ablate.py to attribute time. */
int lis_set_ablation(int mode);
/* Timing experiments: point K1 at 256 zeroed int64 on the device; CTA 0 then accumulates cycle counters there
 * ([0] MMA-warp loop, [1] its waits for page tiles, [2] its waits for a free accumulator, [3] uses,
 * [4+2w] epilogue warp w waiting for a full accumulator, [5+2w] holding it).  NULL switches them off. */
int lis_k1_stats(long long* device_buf);
/* Number of kernels this library launched since load (all entry points). */
int64_t lis_launch_count(void);

/* Debug: raw similarities of M tile 0 against the first `tile_n` token rows, out[128, tile_n]
 * (device float), straight from TMEM.  Exercises the exact TMA/UMMA path of lis_maxsim_scores
 * (a_in_tmem selects the TS form). */
int lis_debug_sim_tile(const void* q, int64_t q_rows, const void* tokens, int64_t n_rows, int dtype,
                       int tile_n, int a_in_tmem, float* out, void* stream);
/* Same for the CTA-pair form (cta_group::2): raw similarities of n_mt (3 or 4) query M tiles against the
 * first 256 token rows, out[n_mt * 128, 256].  n_mt = 3 ends with the 64/64-split M = 128 instruction, so
 * the dump pins both accumulator layouts. */
int lis_debug_sim_pair(const void* q, int64_t q_rows, const void* tokens, int64_t n_rows, int dtype, int n_mt,
                       float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K2: per-query top-k over a score matrix, deterministic order (score desc, id asc).
 *   scores       device float [nq, ld]
 *   ids          device int64 [np] or NULL (id = id_base + column index)
 *   out_scores   device float [nq, k]; out_ids device int64 [nq, k]
 *                (when np < k the tail is filled with -inf / -1)
 *   workspace    device, lis_topk_workspace_bytes(nq, np, k) bytes
 * k <= LIS_MAX_K.
 */
#define LIS_MAX_K 1024
int64_t lis_topk_workspace_bytes(int64_t nq, int64_t np, int k);
int lis_topk(const float* scores, int64_t ld, int64_t nq, int64_t np, const int64_t* ids,
             int64_t id_base, int k, float* out_scores, int64_t* out_ids, void* workspace,
             int64_t workspace_bytes, void* stream);

/* Merge candidate lists (e.g. the allgathered per-GPU top-k): cand_scores/cand_ids device
 * [nq, n_cand]; entries with id < 0 are padding.  Same ordering rule and outputs as lis_topk. */
int lis_merge_topk(const float* cand_scores, const int64_t* cand_ids, int64_t nq, int64_t n_cand,
                   int k, float* out_scores, int64_t* out_ids, void* workspace,
                   int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K3: retrieval head.  out[t,:] = mask[t] * normalize(W h[t,:] + b)   (no epsilon, like the
 * reference), written as 16-bit rows ready for the page store.
 *   hidden  device [n_tok, hidden_dim] dtype;  weight device [128, hidden_dim] dtype (row-major,
 *   i.e. torch Linear.weight);  bias device [128] dtype or NULL;  mask device [n_tok] integers of
 *   mask_itemsize bytes (1, 4 or 8: bool/uint8, int32, int64 attention masks are taken as they are)
 *   or NULL;  out device [n_tok,128] dtype.  hidden_dim % 64 == 0.
 *   round_mode  LIS_ROUND_F32: bias, norm and scale applied to the fp32 accumulators, one rounding at the store
 *               (the more accurate embedding).  LIS_ROUND_REFERENCE: what the reference's 16-bit model computes
 *               (HF modeling_colpali.py:148-155): Linear output rounded to dtype, norm of the rounded values
 *               rounded to dtype, quotient rounded to dtype.
 *   dst_row     device int32 [n_tok] or NULL.  When given, token t is written to row dst_row[t] of `out` instead
 *               of row t, and tokens with dst_row[t] < 0 are not written at all (ragged compaction of padded
 *               encoder batches straight into a page store; see lis_index_add_projected).
 */
int lis_project_normalize(const void* hidden, int64_t n_tok, int64_t hidden_dim, const void* weight,
                          const void* bias, const void* mask, int mask_itemsize, int dtype, int round_mode,
                          const int32_t* dst_row, void* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Page index: the GPU-resident replacement for the Qdrant multivector collection.
 * Owns a token store [cap_rows,128], page offsets, page ids and clamp flags on one device.
 */
typedef struct lis_index lis_index;

int lis_index_create(lis_index** out, int device, int dtype, int64_t cap_rows, int64_t cap_pages);
void lis_index_destroy(lis_index* idx);
/* Append n pages.  `tokens` [sum(lens),128] (16-bit elements, or float for an LIS_F32X2 index) may be
 * a host or device pointer (cudaMemcpyDefault);
 * lens host int32 [n]; ids host int64 [n] or NULL (then ids continue from the current count);
 * clamp host uint8 [n] or NULL (0).  Synchronous with respect to `stream` on return. */
int lis_index_add(lis_index* idx, const void* tokens, const int32_t* lens, const int64_t* ids,
                  const uint8_t* clamp, int64_t n, void* stream);
int64_t lis_index_num_pages(const lis_index* idx);
int64_t lis_index_num_rows(const lis_index* idx);
/* Raw views (device pointers) for zero-copy use by the stateless entry points. */
const void* lis_index_tokens(const lis_index* idx);    /* 16-bit rows; the hi plane of an LIS_F32X2 index */
const void* lis_index_tokens_lo(const lis_index* idx); /* lo plane of an LIS_F32X2 index, else NULL */
const int64_t* lis_index_offsets(const lis_index* idx);
const int64_t* lis_index_ids(const lis_index* idx);
const uint8_t* lis_index_clamp(const lis_index* idx);
/* Device-side fill for benchmarks: append n pages of unit-norm pseudo-random token rows generated
 * in place from a counter hash of (seed, global row index) -- a 130 GB shard never exists on the
 * host.  lens host int32 [n] or NULL (then every page has fixed_len rows).  Ids run from id_base. */
int lis_index_fill_synthetic(lis_index* idx, int64_t n, const int32_t* lens, int32_t fixed_len,
                             uint64_t seed, int64_t id_base, void* stream);
/* Copy token rows [row0, row0+n_rows) of the store to `dst` (host or device), synchronously
 * (16-bit rows; float rows hi+lo for an LIS_F32X2 index). */
int lis_index_read_rows(const lis_index* idx, int64_t row0, int64_t n_rows, void* dst, void* stream);
/* Raw persistence support (save / load of an index without re-ingesting):
 *   lis_index_read_plane   copy rows [row0, row0+n_rows) of plane 0 (16-bit rows / hi) or 1 (lo) to dst
 *   lis_index_write_rows   copy n_rows raw 16-bit rows from src (host or device) into a plane at row0
 *   lis_index_set_tables   install the page tables (host arrays) after the rows were written:
 *                          offsets[n_pages+1] ascending from 0, ids[n_pages] >= 0, clamp[n_pages] or NULL */
int lis_index_read_plane(const lis_index* idx, int plane, int64_t row0, int64_t n_rows, void* dst, void* stream);
int lis_index_write_rows(lis_index* idx, int plane, int64_t row0, int64_t n_rows, const void* src, void* stream);
int lis_index_set_tables(lis_index* idx, const int64_t* offsets, const int64_t* ids, const uint8_t* clamp,
                         int64_t n_pages, void* stream);
int lis_index_dtype(const lis_index* idx);
/* Stateless version of the generator: fill dst[n_rows,128] (device) with the rows the hash assigns
 * to global row indices row0 .. row0+n_rows-1. */
int lis_fill_synthetic_rows(void* dst, int64_t row0, int64_t n_rows, uint64_t seed, int dtype, void* stream);
/* Search: packed queries (as for lis_maxsim_scores; for an LIS_F32X2 index `q` is the hi plane and
 * `q_lo` the lo plane, else q_lo = NULL) -> top-k (score, id) per query.
 * seg_first == NULL states that segment s is query s (no query cut, none empty: QueryPlan.direct); otherwise
 * seg_first is device int32 [nq+1] and the segments of each query are added up.
 * Scratch is owned by the index and grown on demand (outside the timed path after warm-up).
 * ONE search in flight per index: the scratch is shared, so concurrent calls on the same index are serialised
 * by a lock inside the library and must use the same stream (different indexes are independent). */
int lis_index_search(lis_index* idx, const void* q, const void* q_lo, int64_t q_rows, const int32_t* seg_lo,
                     const int32_t* seg_hi, const int32_t* mt_seg, int64_t n_seg, int64_t n_mtiles,
                     const int32_t* seg_first, int64_t nq, int round_mode, int k, float* out_scores,
                     int64_t* out_ids, void* stream);


/* Ingestion fusion: K3 writes straight into the page store.  hidden device [n_pages, seq, hidden_dim]
 * (a padded encoder batch), mask device [n_pages, seq] integers of mask_itemsize bytes (required: it defines which
 * rows exist; left- and right-padded batches alike).  Row t of page b lands at
 * store_row(b) + (number of kept tokens before t in page b); pad rows are never written; page b gets
 * clamp = (kept < seq), i.e. the reference's zero-padding semantics for the dropped rows.  ids host int64
 * [n_pages] or NULL.  Page lengths are computed on the device; one 8-byte read-back tells the host the new
 * row count.  Not available for LIS_F32X2 indexes (the encoder head is 16-bit).  Synchronous on return. */
int lis_index_add_projected(lis_index* idx, const void* hidden, int64_t n_pages, int64_t seq, int64_t hidden_dim,
                            const void* weight, const void* bias, const void* mask, int mask_itemsize,
                            int round_mode, const int64_t* ids, void* stream);
/* Host copy of page lengths [first, first+n) (int32), e.g. after lis_index_add_projected. */
int lis_index_page_lens(const lis_index* idx, int64_t first, int64_t n, int32_t* lens_host, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU: one process per GPU, pages sharded by rank, ONE all-gather of the per-rank top-k per search.
 * NCCL is bound at run time (dlopen libnccl.so.2).  Rank 0 creates the id and ships it to the other ranks by
 * any means (torch.distributed broadcast, MPI, a file); every rank then calls lis_comm_init.
 */
#define LIS_COMM_ID_BYTES 128
typedef struct lis_comm lis_comm;
int lis_comm_unique_id(void* out, int bytes);            /* out: LIS_COMM_ID_BYTES bytes (an ncclUniqueId) */
int lis_comm_init(lis_comm** out, const void* unique_id, int rank, int world, int device);   /* collective */
void lis_comm_destroy(lis_comm* comm);
int lis_comm_rank(const lis_comm* comm);
int lis_comm_world(const lis_comm* comm);
int lis_nccl_version(void);                              /* e.g. 22809; 0 when NCCL is not loadable */

/* One-shot search, host in -> host out, synchronous: the call a server makes per request.
 *   q          packed query rows [q_rows,128] in HOST or device memory: 16-bit rows of the index dtype, or float
 *              rows for an LIS_F32X2 index (split into planes on the device)
 *   seg_lo/seg_hi/mt_seg/seg_first   HOST arrays from lis_plan_queries (seg_first NULL = direct, see above)
 *   comm       NULL or a 1-rank communicator: this index is the whole corpus.  Otherwise every rank calls with the
 *              same queries and k; each scores its shard, K2's last pass writes the local (score, id) candidates
 *              into the send buffer, ONE ncclAllGather, and the same tournament kernel merges world*k candidates
 *              on every rank.  A rank whose shard is empty contributes padding.  Page ids must be globally unique.
 *   out_scores HOST float [nq,k], out_ids HOST int64 [nq,k]: best first, ties by ascending id, (-inf,-1) padding.
 *   stream     the stream on which a device-resident q was produced (ordering only).
 * The whole device sequence (table/query upload, K1, segment sums, K2, all-gather, merge, result download) is
 * captured into a CUDA graph per (query shape, k, corpus size) the first time it is seen and replayed afterwards
 * on a stream owned by the index. */
int lis_index_search_sharded(lis_index* idx, lis_comm* comm, const void* q, int64_t q_rows, const int32_t* seg_lo,
                             const int32_t* seg_hi, const int32_t* mt_seg, int64_t n_seg, int64_t n_mtiles,
                             const int32_t* seg_first, int64_t nq, int round_mode, int k, float* out_scores,
                             int64_t* out_ids, void* stream);

/* Persistence at storage speed: copy n_rows raw 16-bit rows of a plane between the store (from row0) and a file
 * (from byte file_offset), through a pinned double buffer: every 64 MiB chunk is read by io_threads parallel
 * pread()s while the previous chunk is DMA'd (load), or written while the next one is downloaded (save).  The file
 * layout is exactly the HBM layout, so nothing is re-encoded.  Page tables travel separately (lis_index_set_tables).
 * io_threads = 0 picks a default.  Synchronous on return. */
int lis_index_load_rows(lis_index* idx, int plane, int64_t row0, int64_t n_rows, const char* path,
                        int64_t file_offset, int io_threads, void* stream);
int lis_index_save_rows(const lis_index* idx, int plane, int64_t row0, int64_t n_rows, const char* path,
                        int64_t file_offset, int io_threads, void* stream);

/* Remove page number `page` (position in the index, 0-based) from all future searches: its id becomes -1, which
 * K2 treats as padding.  The rows stay in the store (append-only layout); re-adding the content under the same
 * caller id is how an upsert replaces a point (functions.py:865 upserts by point id). */
int lis_index_tombstone(lis_index* idx, int64_t page, void* stream);

/* Introspection: number of cached search graphs; *captures / *replays (may be NULL) count since creation. */
int64_t lis_index_graph_stats(const lis_index* idx, int64_t* captures, int64_t* replays);

/* ---------------------------------------------------------------------------------------------
 * Surface 1 with a HOST-resident corpus, as the reference calls it (05_experiment02.py:213-214).  The token rows
 * stay where they are: they are cut into chunks of whole pages, each chunk goes through a pinned double buffer
 * to the device on a copy stream, and K1 scores chunk i while chunk i+1 is in flight.
 *   q, seg_*, mt_seg       device, as for lis_maxsim_scores
 *   tokens_host            HOST [n_rows,128] 16-bit rows, pageable or pinned (a pinned source is DMA'd in place), or NULL
 *   page_ptrs_host         HOST array of np HOST pointers, page p's rows at page_ptrs_host[p] (a Python list of per-page
 *                          tensors, as create_document_embeddings returns it), or NULL -- exactly one of the two
 *   p_offsets_host         HOST int64 [np+1] from 0 to n_rows, p_clamp_host HOST uint8 [np] or NULL
 *   out                    device float [n_seg, ld_out]
 *   chunk_rows             token rows per chunk (0 = default, 256 Ki rows = 64 MiB)
 *   host_threads           threads gathering pageable rows into the pinned buffer (0 = default)
 * Staging buffers (2 pinned + 2 device chunks) are kept per device between calls; lis_stream_release frees them.
 * Synchronous on return. */
int lis_stream_scores(const void* q, int64_t q_rows, const int32_t* seg_lo, const int32_t* seg_hi,
                      const int32_t* mt_seg, int64_t n_seg, int64_t n_mtiles, const void* tokens_host,
                      const void* const* page_ptrs_host, int64_t n_rows, const int64_t* p_offsets_host,
                      const uint8_t* p_clamp_host, int64_t np, int dtype, int round_mode, float* out,
                      int64_t ld_out, int64_t chunk_rows, int host_threads, void* stream);
void lis_stream_release(void);
/* Pitched copy between host and device in either direction (cudaMemcpy2DAsync, cudaMemcpyDefault): `height` rows of
 * `width_bytes` bytes.  Used to send column blocks of the [nq, np] score matrix to the caller's (pinned) result while K1 is
 * still scoring the next block of pages -- the D2H that colpali-engine does per block with .cpu(). */
int lis_memcpy2d_async(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t width_bytes,
                       int64_t height, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LIS_H_ */
