// K1: fused late-interaction (MaxSim) scoring kernel for sm_100a.
//
//   out[s, p] = sum_{r in segment s} max_{t in page p} <q[r,:], tok[t,:]>
//
// restating the arithmetic of score_multi_vector (reference call site 05_experiment02.py:214;
// body colpali-engine 0.3.13 == HF processing_colpali.py:360: einsum("bnd,csd->bcns").max(3).sum(2))
// without ever materialising the [B,C,N,S] similarity tensor.
//
// Layout / roles (one persistent CTA per SM, 192 or 320 threads):
//   warp 4*EH   TMA producer: query M tiles once (A operand, resident), then the CTA's slice of the
//               page-token store as flat NT-row tiles through an NS-stage mbarrier ring (B operand).
//   warp 4*EH+1 tcgen05.mma issuer (one elected lane).  For every B tile: G MMAs (one per resident M tile),
//               each 128 x NT x 128 (8 k-steps of 16), accumulators in a ring of 512/NT TMEM buffers.
//   warps 0..   epilogue (4 or 8 warps): tcgen05.ld the accumulator (thread = query-token row), running per-page
//               row max in registers (FMNMX3), page boundaries handled by column masks, then a
//               segmented sum over the rows of each query through shared memory -> one fp32 per
//               (query segment, page) to HBM.
//
// The page-token store is tiled FLAT (tiles ignore page boundaries); a CTA owns a contiguous range of
// whole pages, so no partial maxima ever cross CTAs and ragged pages cost no padding reads.
#pragma once
#include "lis_ptx.cuh"

// A/B experiments: -DLIS_MMA_ISSUE_LANE0 issues the MMAs from `lane == 0` instead of an elected lane.
// -DLIS_K1_STATS compiles the cycle counters of lis_k1_stats in (experiment builds only: they cost registers).
#ifdef LIS_K1_STATS
#define LIS_STATS_ON(args) ((args).stats != nullptr)
#else
#define LIS_STATS_ON(args) false
#endif
// Number of MMA-issuing warps (uses are dealt round-robin; with two, each owns one accumulator buffer).
#ifndef LIS_MMA_WARPS
#define LIS_MMA_WARPS 2
#endif
// Accumulator chunks (32 columns) an epilogue warp holds in registers at once.
#ifndef LIS_EPI_GRP
#define LIS_EPI_GRP 4
#endif
#ifdef LIS_MMA_ISSUE_LANE0
#define LIS_ISSUE_PRED (lane == 0)
#else
#define LIS_ISSUE_PRED elect_one_sync()
#endif

namespace lis {

constexpr int kDim = 128;        // embedding width (VECTOR_SIZE, 01_create_context_qdrant.py:70)
constexpr int kMTile = 128;      // query rows per UMMA
constexpr int kKHalf = 64;       // elements per 128-byte swizzle atom
constexpr int kATileBytes = kMTile * kDim * 2;  // 32 KB
constexpr int kNumThreads = 192;
constexpr int kEpiThreads = 128;
constexpr int kTmemCols = 512;
constexpr int kMmaWarps = LIS_MMA_WARPS;
constexpr int kCtrlThreads = 32 * (2 + kMmaWarps);   // reducer + producer + MMA warps

struct MaxSimArgs {
  const int64_t* p_offsets;  // [np+1]
  const uint8_t* p_clamp;    // [np] or null
  const int32_t* seg_lo;     // [n_seg]
  const int32_t* seg_hi;     // [n_seg]
  const int32_t* mt_seg;     // [n_mtiles_total+1]
  float* out;                // [n_seg, ld_out]
  float* dbg;                // debug: raw sims of (tile 0 of CTA 0), [G*128, NT]; normally null
  long long* stats;          // timing experiments: CTA 0 writes cycle counters here when non-null (see lis_k1_stats)
  const void* q;             // [q_rows, 128] packed query rows (read directly by the A-in-TMEM form)
  int64_t q_rows;
  int64_t ld_out;
  int64_t np;
  int32_t mt0;        // first M tile of this launch
  int32_t n_mt;       // M tiles in this launch (1..G)
  int32_t round_mode; // lis_round_mode bits
  int32_t is_bf16;    // 1 = bf16, 0 = fp16
  int32_t ablate;     // timing experiments only (results invalid): 1 = epilogue skips TMEM loads, 2 = skips the max,
                      // 3 = producer stops issuing TMA after the first ring fill, 4 = 3 + 1
};

__device__ __forceinline__ float round_to_input_dtype(float x, int is_bf16) {
  return is_bf16 ? __bfloat162float(__float2bfloat16_rn(x)) : __half2float(__float2half_rn(x));
}

// max of 32 accumulator columns and m: four independent FMNMX3 chains (the alu pipe has a 4-cycle
// dependent-issue latency and the epilogue runs one or two warps per scheduler, so ILP matters)
__device__ __forceinline__ float max32(const uint32_t (&v)[32], float m) {
  float m0 = m, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    m0 = fmax3(m0, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
    m1 = fmax3(m1, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
    m2 = fmax3(m2, __uint_as_float(v[i + 4]), __uint_as_float(v[i + 5]));
    m3 = fmax3(m3, __uint_as_float(v[i + 6]), __uint_as_float(v[i + 7]));
  }
  return fmax3(m0, m1, fmaxf(m2, m3));
}
// max over columns [lo, hi) of a 32-column chunk (hi <= lo: nothing)
__device__ __forceinline__ float max32_masked(const uint32_t (&v)[32], float m, int lo, int hi) {
  const uint32_t below_hi = hi >= 32 ? 0xffffffffu : (hi <= 0 ? 0u : ((1u << hi) - 1u));
  const uint32_t below_lo = lo >= 32 ? 0xffffffffu : (lo <= 0 ? 0u : ((1u << lo) - 1u));
  const uint32_t mask = below_hi & ~below_lo;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float x = ((mask >> i) & 1u) ? __uint_as_float(v[i]) : -INFINITY;
    m = fmaxf(m, x);
  }
  return m;
}

// one pass over a chunk in which a page ends at column b (0..32): returns max(m, v[0..b)) and puts
// max(v[b..32)) -- the start of the next page -- into m_next.  Branch-free (b is the same for the whole warp, but a
// switch on it compiles into a tree of indirect jumps that costs ~1000 cycles): per pair of columns two selects and
// one FMNMX3 per side, four independent chains.
__device__ __forceinline__ float max32_split(const uint32_t (&v)[32], float m, int b, float& m_next) {
  float a0 = m, a1 = -INFINITY, c0 = -INFINITY, c1 = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float x0 = __uint_as_float(v[i]), x1 = __uint_as_float(v[i + 1]);
    const float x2 = __uint_as_float(v[i + 2]), x3 = __uint_as_float(v[i + 3]);
    const bool p0 = i < b, p1 = i + 1 < b, p2 = i + 2 < b, p3 = i + 3 < b;
    a0 = fmax3(a0, p0 ? x0 : -INFINITY, p1 ? x1 : -INFINITY);
    c0 = fmax3(c0, p0 ? -INFINITY : x0, p1 ? -INFINITY : x1);
    a1 = fmax3(a1, p2 ? x2 : -INFINITY, p3 ? x3 : -INFINITY);
    c1 = fmax3(c1, p2 ? -INFINITY : x2, p3 ? -INFINITY : x3);
  }
  m_next = fmaxf(c0, c1);
  return fmaxf(a0, a1);
}

// first index i in [0, n] with off[i] >= target (off ascending, n+1 entries)
__device__ __forceinline__ int64_t lower_bound_off(const int64_t* off, int64_t n, int64_t target) {
  int64_t lo = 0, hi = n;  // answer in [lo, hi]; off[n] >= any target we pass
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(off + mid) >= target) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// Page reducer body shared by the single-CTA and the CTA-pair kernel: one finished (page, M tile) sits in
// shared memory as `nparts` column partitions of partial row maxima (row r of partition c at ex[c * pstride + r]);
// combine the partitions (max), clamp / round like the reference, and add up the rows of every query segment
// of the tile -> one fp32 per (segment, page).  The warp is cut into groups of G = 32 / 2^ceil(log2(seg_cnt))
// lanes, one group per segment (all segments of a typical tile in one round): a group strides over the
// segment's rows and finishes with a log2(G)-step shuffle tree.  The order of the additions depends only on
// the tile's segment table, so every kernel form produces the same bits.
// [rbase, rbase + rcnt): tile rows this CTA owns (the whole tile, or 64 rows of a split tile: segments of the
// other half are skipped; a segment straddling the boundary is a planning error and traps).
template <int NPARTS>
__device__ __forceinline__ void reduce_tile_segments(const float* ex, int pstride, const uint16_t* tab16,
                                                     const int32_t* g_lo, const int32_t* g_hi, int seg_first, int seg_cnt,
                                                     int tile_row0, int rbase, int rcnt, bool clamp, bool round_ref,
                                                     bool round_sum, int is_bf16, float* out_col, int64_t ld_out,
                                                     int lane) {
  int per_round = 1;
  while (per_round < seg_cnt && per_round < 32) per_round <<= 1;
  const int G = 32 / per_round;
  const int sub = lane & (G - 1), which = lane / G;
  for (int j0 = 0; j0 < seg_cnt; j0 += per_round) {
    const int j = j0 + which;
    int lo = 0, hi = 0;
    bool mine = false;
    if (j < seg_cnt) {
      if (j < 16) {
        const uint32_t lh = tab16[j];
        lo = (int)(lh & 0xffu); hi = (int)(lh >> 8);
      } else {
        lo = __ldg(g_lo + seg_first + j) - tile_row0;
        hi = __ldg(g_hi + seg_first + j) - tile_row0;
      }
      lo -= rbase; hi -= rbase;
      mine = hi > 0 && lo < rcnt;
      if (mine && (lo < 0 || hi > rcnt)) {
        printf("lis: a query segment straddles the 64-row midpoint of the M tile at row %d (plan with lis_plan_queries)\n", tile_row0);
        __trap();
      }
      if (!mine) { lo = 0; hi = 0; }
    }
    // Canonical order, independent of how many segments share the tile (so a query scores the same bits alone or
    // coalesced with others): eight interleaved partial sums A_v = x[lo+v] + x[lo+v+8] + ... (ascending rows),
    // combined as ((A0+A4)+(A2+A6)) + ((A1+A5)+(A3+A7)) -- the xor-4, xor-2, xor-1 butterfly.  A lane owns the partial
    // sums v = sub (mod G); it visits them in bit-reversed order and merges them pairwise (binary-counter stack), which
    // is the upper part of the butterfly; shuffles do the rest.  One instance of the row loop, whatever G is.
    const int nv = G >= 8 ? 1 : 8 / G;                       // partial sums per lane
    const int shift = nv == 8 ? 0 : (nv == 4 ? 1 : (nv == 2 ? 2 : 3));
    float acc = 0.f, s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 1
    for (int k = 0; k < nv; ++k) {
      const int v = sub + G * (int)(((0x73516240u >> (4 * k)) & 7u) >> shift);   // 3-bit reversal of k, cut to log2(nv) bits
      float a = 0.f;
      if (G < 8 || sub < 8) {
        // rows lo+v, lo+v+8, ...: all loads first (independent), then the additions in ascending row order; adding
        // the +0.0 of a row beyond the segment changes nothing, so a fixed trip count keeps the canonical value
        auto row = [&](int r) {
          float x = 0.f;
          if (r < hi) {
            x = ex[r];
#pragma unroll
            for (int c = 1; c < NPARTS; ++c) x = fmaxf(x, ex[c * pstride + r]);
            if (clamp) x = fmaxf(x, 0.f);
            if (round_ref) x = round_to_input_dtype(x, is_bf16);
          }
          return x;
        };
        const int r0 = lo + v;
        const float x0 = row(r0), x1 = row(r0 + 8), x2 = row(r0 + 16), x3 = row(r0 + 24);
        a = ((x0 + x1) + x2) + x3;       // == sequential accumulation from 0
        if (hi - lo > 32) {              // segments are at most 64 rows
          const float x4 = row(r0 + 32), x5 = row(r0 + 40), x6 = row(r0 + 48), x7 = row(r0 + 56);
          a = (((a + x4) + x5) + x6) + x7;
        }
      }
      if (k & 1) {
        a = s0 + a;
        if (k & 2) {
          a = s1 + a;
          if (k & 4) a = s2 + a; else s2 = a;
        } else s1 = a;
      } else s0 = a;
      acc = a;
    }
    for (int o = (G < 8 ? G : 8) >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (mine && sub == 0) {
      if (round_sum) acc = round_to_input_dtype(acc, is_bf16);
      out_col[(int64_t)(seg_first + j) * ld_out] = acc;
    }
  }
}

// ATM ("A in tensor memory"): the query M tiles are stored in TMEM (64 columns each, two 16-bit
// values per 32-bit cell) and the MMA takes its A operand from there (TS form).  Shared memory then
// serves only the page tiles: the SS form at M=128 x N=256 reads 96 B/clk of operands from smem,
// which is the measured operand-fetch limit (profiles/micro_mma_rate_r1.txt), so every TMA write
// competes with the tensor pipe; with A in TMEM the MMA reads 64 B/clk at any N.
//
// P ("planes"): fp32 embeddings are served as two bf16 planes per operand, x = hi + lo with
// hi = bf16(x), lo = bf16(x - hi) (|x - hi - lo| <= 2^-18 |x|).  The tile product is then
// hi*hi + hi*lo + lo*hi, three MMAs into the same fp32 accumulator (the dropped lo*lo term is of
// the order of the residual), which keeps ~fp32 accuracy on the bf16 tensor pipe.
template <int NT, int G, int EH, bool ATM, bool DBG, int P = 1>
__global__ void __launch_bounds__(kCtrlThreads + 128 * EH, 1)
maxsim_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_p,
              const __grid_constant__ CUtensorMap tmap_q2, const __grid_constant__ CUtensorMap tmap_p2,
              const MaxSimArgs args, const int NS) {
  static_assert(P == 1 || (P == 2 && !ATM), "planes");
  static_assert(NT == 128 || NT == 192 || NT == 256, "tile_n");
  static_assert(EH == 1 || EH == 2, "epilogue halves");
  constexpr int kACols = ATM ? 64 * G : 0;              // TMEM columns holding the query tiles
  constexpr int NACC = (kTmemCols - kACols) / NT;       // accumulator buffers
  static_assert(NACC >= 2 && NACC <= 4, "need at least two accumulator buffers");
  constexpr int kBPlaneBytes = NT * kDim * 2;
  constexpr int kBStageBytes = P * kBPlaneBytes;
  constexpr int kBHalfBytes = NT * 128;
  constexpr int kATile = P * kATileBytes;               // all planes of one M tile

  extern __shared__ __align__(1024) uint8_t smem[];  // 128-byte swizzle atoms need 1024-byte alignment
  uint8_t* smem_a = smem;                               // [G][P][2][128 rows x 128 B] (SS form only)
  uint8_t* smem_b = smem + (ATM ? 0 : G * kATile);      // [NS][P][2][NT rows x 128 B]
  uint8_t* tail = smem_b + (size_t)NS * kBStageBytes;
  uint64_t* q_full = reinterpret_cast<uint64_t*>(tail);      // 1
  uint64_t* b_full = q_full + 1;                             // [NS]
  uint64_t* b_empty = b_full + 8;                            // [NS]
  uint64_t* acc_full = b_empty + 8;                          // [NACC]
  uint64_t* acc_empty = acc_full + 4;                        // [NACC]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 4);
  int64_t* range = reinterpret_cast<int64_t*>(tmem_slot + 2);  // [0]=page begin [1]=page end [2]=row0
  float* srm = reinterpret_cast<float*>(range + 4);            // [2][EH*128] row-max exchange
  constexpr int kPW = 96;                                      // page-table window (pages)
  uint64_t* ex_full = reinterpret_cast<uint64_t*>(srm + 2 * EH * kMTile);  // [2] exchange slot published by all epilogue warps
  uint64_t* ex_empty = ex_full + 2;                                        // [2] ... consumed by the reducer warp
  int64_t* ex_meta = reinterpret_cast<int64_t*>(ex_empty + 2);             // [2][2] page index, (g | clamp << 8)
  int32_t* pw_end = reinterpret_cast<int32_t*>(ex_meta + 4);              // [kPW] end row (relative to row0) of page w0+i
  uint8_t* pw_clamp = reinterpret_cast<uint8_t*>(pw_end + kPW);          // [kPW] clamp flag of page w0+i
  uint16_t* segtab = reinterpret_cast<uint16_t*>(pw_clamp + kPW);        // [G][16] lo | hi<<8 of the tile's first segments
  int32_t* seginfo = reinterpret_cast<int32_t*>(segtab + G * 16);        // [G][2] first segment, segment count

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Roles by warp id: epilogue warps first, then the page reducer, the TMA producer, the MMA issuer last.  Within a
  // scheduler the highest warp id wins arbitration, so the two single-lane control warps are never
  // starved by the arithmetic of the epilogue warps they share an SM sub-partition with.
  constexpr int kReducerWarp = 4 * EH;
  constexpr int kProducerWarp = 4 * EH + 1;
  constexpr int kMmaWarp = 4 * EH + 2;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_p);
    if (P == 2) { tma_prefetch_desc(&tmap_q2); tma_prefetch_desc(&tmap_p2); }
    mbar_init(q_full, ATM ? 4 * EH : 1);
    for (int s = 0; s < NS; ++s) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, kMmaWarps); }
    for (int a = 0; a < NACC; ++a) { mbar_init(acc_full + a, 1); mbar_init(acc_empty + a, 4 * EH); }
    for (int i = 0; i < 2; ++i) { mbar_init(ex_full + i, 4 * EH); mbar_init(ex_empty + i, 1); }
    fence_barrier_init();
    // This CTA's contiguous range of whole pages, balanced by token rows.
    const int64_t np = args.np;
    const int64_t base = __ldg(args.p_offsets);
    const int64_t total = __ldg(args.p_offsets + np) - base;
    const int64_t per = (total + gridDim.x - 1) / gridDim.x;
    int64_t pa = np, pb = np;
    if (per > 0 && (int64_t)blockIdx.x * per < total) {
      pa = lower_bound_off(args.p_offsets, np, base + (int64_t)blockIdx.x * per);
      pb = (blockIdx.x + 1 == gridDim.x || (int64_t)(blockIdx.x + 1) * per >= total)
               ? np
               : lower_bound_off(args.p_offsets, np, base + (int64_t)(blockIdx.x + 1) * per);
    } else if (total == 0 && blockIdx.x == 0) {
      pa = 0; pb = np;  // only empty pages: CTA 0 emits them
    }
    range[0] = pa;
    range[1] = pb;
    range[2] = (pa < np) ? __ldg(args.p_offsets + pa) : 0;
    range[3] = (pa < pb) ? __ldg(args.p_offsets + pb) : range[2];
  }
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const uint32_t tmem_base = *tmem_slot;
  const int64_t pa = range[0], pb = range[1];
  const int64_t row0 = range[2];
  const int64_t rows = range[3] - row0;
  const int ntiles = (int)((rows + NT - 1) / NT);
  const int n_mt = args.n_mt;

  if (warp == kProducerWarp) {
    // ===================== TMA producer =====================
    if (pa < pb) {
      if (!ATM && ntiles > 0 && elect_one_sync()) {
        mbar_arrive_expect_tx(q_full, (uint32_t)n_mt * kATile);
        for (int g = 0; g < n_mt; ++g)
          for (int pl = 0; pl < P; ++pl)
            for (int h = 0; h < 2; ++h)
              tma_load_2d(smem_a + g * kATile + pl * kATileBytes + h * (kMTile * 128), pl ? &tmap_q2 : &tmap_q,
                          q_full, h * kKHalf, (args.mt0 + g) * kMTile, kPolicyEvictLast);
      }
      __syncwarp();
      for (int t = 0; t < ntiles; ++t) {
        const int s = t % NS;
        const uint32_t ph = (uint32_t)(t / NS) & 1u;
        mbar_wait(b_empty + s, ph ^ 1u);
        if (args.ablate >= 3 && t >= NS) {       // timing experiment: reuse stale tiles, no TMA traffic
          if (elect_one_sync()) mbar_arrive(b_full + s);
        } else if (elect_one_sync()) {
          mbar_arrive_expect_tx(b_full + s, kBStageBytes);
          uint8_t* dst = smem_b + (size_t)s * kBStageBytes;
          const int32_t r = (int32_t)(row0 + (int64_t)t * NT);
          for (int pl = 0; pl < P; ++pl) {
            const CUtensorMap* tm = pl ? &tmap_p2 : &tmap_p;
            tma_load_2d(dst + pl * kBPlaneBytes, tm, b_full + s, 0, r, kPolicyEvictFirst);
            tma_load_2d(dst + pl * kBPlaneBytes + kBHalfBytes, tm, b_full + s, kKHalf, r, kPolicyEvictFirst);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp >= kMmaWarp) {
    // ===================== MMA issuers =====================
    // kMmaWarps warps deal the (tile, M tile) uses round-robin.  The tensor pipe accepts only ~6-8
    // queued instructions (the 7th tcgen05.mma of a use stalls on the MIO queue), so a single issuing
    // thread spends ~900 cycles per use blocked in issue and only then starts the hand-shake for the
    // next accumulator; two issuers overlap one's issue with the other's hand-shake.
    // The whole warp runs the loop converged and one elected lane issues: in that form ptxas keeps
    // the (warp-uniform) descriptors in uniform registers; issuing from an `if (lane == 0)` branch
    // costs ~85 cycles per tcgen05.mma instead of ~50 (profiles/micro_mma_rate_r1.txt).
    if (pa < pb && ntiles > 0) {
      const uint32_t idesc = make_idesc_f16(args.is_bf16 ? 1u : 0u, kMTile, NT);
      const uint32_t a_base = smem_u32(smem_a);
      const uint32_t b_base = smem_u32(smem_b);
      mbar_wait(q_full, 0);
      uint32_t use = 0;
      const uint32_t my = (uint32_t)(warp - kMmaWarp);
      const bool st_on = LIS_STATS_ON(args) && blockIdx.x == 0 && my == 0;
      long long st_b = 0, st_acc = 0, st_issue = 0;
      const long long st_t0 = clock64();
      for (int t = 0; t < ntiles; ++t) {
        const int s = t % NS;
        long long c0 = st_on ? clock64() : 0;
        mbar_wait(b_full + s, (uint32_t)(t / NS) & 1u);
        if (st_on) st_b += clock64() - c0;
        tc_fence_after();
        bool issued = false;
        for (int g = 0; g < n_mt; ++g, ++use) {
          if (use % kMmaWarps != my) continue;
          issued = true;
          const uint32_t a = use % NACC;
          c0 = st_on ? clock64() : 0;
          const bool tl = LIS_STATS_ON(args) && blockIdx.x == 0 && use >= 1000 && use < 1008 && lane == 0;
          if (tl) args.stats[32 + (use - 1000) * 5 + 0] = clock64();      // MMA warp reaches the acc_empty wait
          mbar_wait(acc_empty + a, ((use / NACC) & 1u) ^ 1u);
          if (tl) args.stats[32 + (use - 1000) * 5 + 1] = clock64();      // ... passes it
          if (st_on) st_acc += clock64() - c0;
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + kACols + a * NT;
          const long long ic0 = st_on ? clock64() : 0;
          if (LIS_ISSUE_PRED) {
            // plane pairs (A plane, B plane): hi*hi only, or hi*hi + hi*lo + lo*hi for split fp32
#pragma unroll
            for (int pp = 0; pp < (P == 2 ? 3 : 1); ++pp) {
              const uint32_t pa_off = (pp == 2 ? 1u : 0u) * kATileBytes;
              const uint32_t pb_off = (pp == 1 ? 1u : 0u) * kBPlaneBytes;
#pragma unroll
              for (int k = 0; k < kDim / 16; ++k) {
                const uint32_t koff = (uint32_t)(k >> 2) * (kMTile * 128) + (uint32_t)(k & 3) * 32;
                const uint32_t koff_b = (uint32_t)(k >> 2) * kBHalfBytes + (uint32_t)(k & 3) * 32;
                const uint64_t bdesc = make_kmajor_sw128_desc(b_base + s * kBStageBytes + pb_off + koff_b);
                const uint32_t acc = (pp | k) ? 1u : 0u;
                if (ATM) {
                  // 16 K-elements of a 16-bit operand = 8 TMEM columns per k-step
                  umma_f16_ts(d_tmem, tmem_base + g * 64 + k * 8, bdesc, idesc, acc);
                } else {
                  const uint64_t adesc = make_kmajor_sw128_desc(a_base + g * kATile + pa_off + koff);
                  umma_f16(d_tmem, adesc, bdesc, idesc, acc);
                }
              }
            }
            umma_commit(acc_full + a);
          }
          __syncwarp();
          if (tl) args.stats[32 + (use - 1000) * 5 + 2] = clock64();      // ... has issued the MMAs + commit
          if (st_on) st_issue += clock64() - ic0;
        }
        // hand the page tile back: after this warp's MMAs on it have completed (or at once if it had none)
        if (LIS_ISSUE_PRED) {
          if (issued) umma_commit(b_empty + s);
          else mbar_arrive(b_empty + s);
        }
        __syncwarp();
      }
      if (st_on && lane == 0) {
        args.stats[0] = clock64() - st_t0;   // MMA warp: total cycles in the main loop
        args.stats[1] = st_b;                //           ... of which waiting for page tiles (TMA)
        args.stats[2] = st_acc;              //           ... of which waiting for a free accumulator
        args.stats[3] = use;
        args.stats[23] = st_issue;           //           ... of which issuing MMAs + commits
      }
    }
  } else if (warp == kReducerWarp) {
    // ===================== page reducer =====================
    // Every finished (page, M tile) arrives as a slot of partial row maxima published by the
    // epilogue warps (non-blocking for them).  This warp combines the column halves (max), clamps /
    // rounds like the reference and adds up the rows of every query segment with a fixed shuffle
    // tree (deterministic) -> one fp32 per (segment, page).  Keeping this off the epilogue warps
    // matters: a page end used to stall all of them at a barrier, and the MMA warp behind them.
    if (pa < pb) {
      const int is_bf16 = args.is_bf16;
      const bool round_ref = (args.round_mode & 1) != 0;
      const bool round_sum = round_ref && (args.round_mode & 2) == 0;
      const int64_t nfin = (pb - pa) * G;
      for (int64_t f = 0; f < nfin; ++f) {
        const int slot = (int)(f & 1);
        mbar_wait(ex_full + slot, (uint32_t)(f >> 1) & 1u);
        const int64_t p = ex_meta[2 * slot];
        const int g = (int)(ex_meta[2 * slot + 1] & 0xff);
        const bool clamp = (ex_meta[2 * slot + 1] >> 8) != 0;
        const float* ex = srm + slot * (EH * kMTile);
        const int mt = args.mt0 + g;
        const int seg_first = seginfo[2 * g], seg_cnt = seginfo[2 * g + 1];
        reduce_tile_segments<EH>(ex, kMTile, segtab + g * 16, args.seg_lo, args.seg_hi, seg_first, seg_cnt, mt * kMTile, 0,
                             kMTile, clamp, round_ref, round_sum, is_bf16, args.out + p, args.ld_out, lane);
        __syncwarp();
        if (lane == 0) mbar_arrive(ex_empty + slot);
      }
    }
  } else {
    // ===================== epilogue (warps 0 .. 4*EH-1) =====================
    // EH "column halves": with EH == 2 two warps share every TMEM lane quarter and each scans half
    // of the tile's columns; their partial row maxima meet in shared memory when a page ends.
    // The M-tile loop is deliberately NOT unrolled -- the hot loop must stay inside the instruction
    // cache -- so the G running maxima of a thread sit in a register ring that is rotated once per
    // M tile (the current tile's value is always rm[0]); the launcher guarantees n_mt == G.
    const int quarter = warp & 3;             // TMEM lane quarter this warp may read
    const int half = warp >> 2;               // 0 .. EH-1
    const int row = quarter * 32 + lane;      // query-token row inside the M tile
    const int etid = half * kMTile + row;     // 0 .. 128*EH-1
    constexpr int NCH = NT / 32;              // 32-column chunks per tile
    constexpr int NOWN = NCH / EH;            // chunks scanned by this warp
    const int c_lo = half * NOWN;
    if (ATM && pa < pb && ntiles > 0) {
      // Stage the query tiles in TMEM: thread = row, cell c of the row = its bytes [4c, 4c+4).
      // Rows past the end of the query matrix are zero (they belong to no segment).
      constexpr int kCols = 64 / EH;              // columns written by this warp per M tile
      for (int g = 0; g < G; ++g) {
        const int64_t qrow = (int64_t)(args.mt0 + g) * kMTile + row;
        const uint4* src = reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(args.q) + qrow * 256 +
                                                          half * (kCols * 4));
#pragma unroll
        for (int c = 0; c < kCols / 32; ++c) {
          uint32_t v[32];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            uint4 x = make_uint4(0, 0, 0, 0);
            if (qrow < args.q_rows) x = __ldg(src + c * 8 + i);
            v[4 * i] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
          }
          tmem_st32(tmem_base + ((uint32_t)(quarter * 32) << 16) + g * 64 + half * kCols + c * 32, v);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(q_full);
    }

    float rm[G];
#pragma unroll
    for (int g = 0; g < G; ++g) rm[g] = -INFINITY;
    auto rotate = [&](float cur) {
#pragma unroll
      for (int i = 0; i + 1 < G; ++i) rm[i] = rm[i + 1];
      rm[G - 1] = cur;
    };

    // All page bookkeeping below is 32-bit and relative to this CTA's range: page index pi = p - pa,
    // rows relative to row0 (a CTA's range is < 2^31 rows).  Barrier addresses are taken once.
    const uint32_t acc_full_u = smem_u32(acc_full), acc_empty_u = smem_u32(acc_empty);
    const uint32_t ex_full_u = smem_u32(ex_full), ex_empty_u = smem_u32(ex_empty);
    const int npages = (int)(pb - pa);

    // page-table window [w0, w0+kPW): relative end rows and clamp flags of the pages around the
    // cursor, in shared memory (L1 is ~4 KB here; a global read would be an L2 round trip).  Every
    // epilogue thread walks the pages in the same order, so refills are collective (two barriers).
    int w0 = 0;
    auto refill = [&](int base) {
      named_bar_sync(1, 128 * EH);            // nobody still reads the old window
      if (etid < kPW) {
        const int pg = base + etid;
        int e = 0;
        uint8_t c = 0;
        if (pg < npages) {
          e = (int)(__ldg(args.p_offsets + pa + pg + 1) - row0);
          if (args.p_clamp != nullptr) c = __ldg(args.p_clamp + pa + pg);
        }
        pw_end[etid] = e;
        pw_clamp[etid] = c;
      }
      named_bar_sync(1, 128 * EH);
      w0 = base;
    };
    if (etid < G * 16) {                       // segment tables of the resident M tiles
      const int g = etid >> 4, j = etid & 15;
      const int mt = args.mt0 + g;
      const int first = __ldg(args.mt_seg + mt), cnt = __ldg(args.mt_seg + mt + 1) - first;
      if (j == 0) { seginfo[2 * g] = first; seginfo[2 * g + 1] = cnt; }
      if (j < cnt)
        segtab[g * 16 + j] = (uint16_t)((__ldg(args.seg_lo + first + j) - mt * kMTile) |
                                        ((__ldg(args.seg_hi + first + j) - mt * kMTile) << 8));
    }

    // Publish one finished page of M tile g: the partial row maxima of this thread go into the next
    // exchange slot and the warp arrives on the slot's barrier -- nobody waits for anybody here;
    // the reducer warp takes over once all epilogue warps have arrived.
    uint32_t fin = 0;
    auto finish_page = [&](int g, int pi, float v) {
      const uint32_t slot = fin & 1u;
      mbar_wait_u32(ex_empty_u + slot * 8, ((fin >> 1) & 1u) ^ 1u);   // slot drained by the reducer (2 finishes ago)
      srm[slot * (EH * kMTile) + etid] = v;
      if (etid == 0) {
        ex_meta[2 * slot] = pa + pi;
        ex_meta[2 * slot + 1] = (int64_t)g | ((int64_t)pw_clamp[pi - w0] << 8);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_u32(ex_full_u + slot * 8);
      ++fin;
    };

    long long st_wait = 0, st_hold = 0;
    if (pa < pb) {
      int p = 0;                                        // current page (relative)
      refill(0);                                        // (also publishes the segment tables)
      int pend = pw_end[0];                             // its end row (relative to row0)
      uint32_t use = 0;
      // chunks held in registers at once: 4 (all of the warp's loads in flight before the release)
      // measured 2-3 % faster than 2 on the 3-tile pass despite sitting at the 168-register cap
      constexpr int GRP = (NOWN % LIS_EPI_GRP == 0) ? LIS_EPI_GRP : (NOWN % 3 == 0 ? 3 : 2);
      static_assert(NOWN % GRP == 0, "chunk grouping");
      const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16) + kACols + c_lo * 32;
      for (int t = 0; t < ntiles; ++t) {
        const int tcol = t * NT;                        // row (relative) of this tile's column 0
        // end column of a page relative to this tile (saturated; > NT: the page continues)
        auto rel_end = [&](int e) { const int d = e - tcol; return d > NT ? NT + 1 : d; };
        const int pe_tile = rel_end(pend);
        int p_next = p, pend_next = pend;
#pragma unroll 1
        for (int g = 0; g < G; ++g) {
          const uint32_t a = use % NACC;
          const bool st_on = LIS_STATS_ON(args) && blockIdx.x == 0;
          const long long ec0 = st_on ? clock64() : 0;
          mbar_wait_u32(acc_full_u + a * 8, (use / NACC) & 1u);
          const long long ec1 = st_on ? clock64() : 0;
          tc_fence_after();
          ++use;
          const uint32_t taddr = tlane + a * NT;
          int pp = p, ppend = pend;         // rewind the page cursor for every M tile
          bool live = pp < npages;
          int pe = pe_tile;
          float m = rm[0];

          auto next_page = [&]() {          // after a finish: advance the cursor
            m = -INFINITY;
            ++pp;
            if (pp >= npages) { live = false; pe = NT + 1; return; }
            if (pp >= w0 + kPW) refill(pp - p < kPW ? p : pp);
            ppend = pw_end[pp - w0];
            pe = rel_end(ppend);
          };
          // pages that end at or before column col_end without this warp scanning them
          auto skip_to = [&](int col_end) {
            while (live && pe <= col_end) { finish_page(g, pp, m); next_page(); }
          };
          // one 32-column chunk starting at tile column cb
          auto scan = [&](const uint32_t (&v)[32], int cb) {
            if (DBG) {
              if (blockIdx.x == 0 && t == 0 && args.dbg != nullptr) {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  args.dbg[(int64_t)(g * kMTile + row) * NT + cb + i] = __uint_as_float(v[i]);
              }
            }
            if (!live) return;
            if (pe - cb > 32) { m = max32(v, m); return; }
            // a page ends inside this chunk, at column b
            int lo = pe - cb;
            float m_next;
            m = max32_split(v, m, lo, m_next);
            finish_page(g, pp, m);
            next_page();
            if (!live) return;
            if (pe - cb > 32) { m = m_next; return; }     // the next page runs past the chunk: done
            // rare: further pages end inside the same chunk (pages shorter than 32 tokens)
            while (true) {
              const int rel = pe - cb;
              const int hi = rel < 32 ? rel : 32;
              m = max32_masked(v, m, lo, hi);
              if (rel > 32) break;
              finish_page(g, pp, m);
              next_page();
              if (!live) break;
              lo = hi;
              if (lo >= 32) break;
            }
          };

          // The accumulator buffer is the scarce resource (512/NT of them feed the tensor pipe): pull
          // this warp's columns into registers, hand the buffer back to the MMA warps as soon as the
          // last load has landed, and only then finish the arithmetic and the page logic.
#pragma unroll
          for (int grp = 0; grp < NOWN / GRP; ++grp) {
            uint32_t v0[32], v1[32], v2[32], v3[32];
            const uint32_t ta = taddr + grp * GRP * 32;
            if (args.ablate != 1 && args.ablate != 4) {
              tmem_ld32(ta, v0);
              if (GRP > 1) tmem_ld32(ta + 32, v1);
              if (GRP > 2) tmem_ld32(ta + 64, v2);
              if (GRP > 3) tmem_ld32(ta + 96, v3);
              tmem_ld_wait();
            }
            if (grp == NOWN / GRP - 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive_u32(acc_empty_u + a * 8);
              if (st_on) {
                st_wait += ec1 - ec0;          // epilogue warp: waiting for a full accumulator
                st_hold += clock64() - ec1;    //                holding it (wake -> release)
              }
            }
            const int cb = (c_lo + grp * GRP) * 32;
            if (grp == 0 && live && (p < w0 || p >= w0 + kPW)) refill(p);   // (rare) cursor rewound out of the window
            if (grp == 0 && EH == 2 && half == 1) skip_to(c_lo * 32);
            if (!DBG && (!live || pe > cb + GRP * 32)) {
              // fast path: no page ends inside these columns
              if (live && (args.ablate == 0 || args.ablate == 3)) {
                m = max32(v0, m);
                if (GRP > 1) m = max32(v1, m);
                if (GRP > 2) m = max32(v2, m);
                if (GRP > 3) m = max32(v3, m);
              }
            } else {
              scan(v0, cb);
              if (GRP > 1) scan(v1, cb + 32);
              if (GRP > 2) scan(v2, cb + 64);
              if (GRP > 3) scan(v3, cb + 96);
            }
          }
          // every warp closes ALL pages that end inside this tile before it moves on: a page ending exactly at the
          // tile's last column followed by empty pages would otherwise be closed in this tile by the warps of
          // column half 0 and in the next tile by those of half 1 -- different publication orders, mixed-up slots
          skip_to(NT);
          rotate(m);
          p_next = pp;
          pend_next = ppend;
        }
        p = p_next; pend = pend_next;
      }
      if (LIS_STATS_ON(args) && blockIdx.x == 0 && lane == 0) {
        args.stats[4 + 2 * warp] = st_wait;
        args.stats[5 + 2 * warp] = st_hold;
      }
      // pages not closed by any tile: trailing empty pages (or ntiles == 0)
      while (p < npages) {
        if (p < w0 || p >= w0 + kPW) refill(p);
#pragma unroll 1
        for (int g = 0; g < G; ++g) {
          finish_page(g, p, rm[0]);
          rotate(-INFINITY);
        }
        ++p;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace lis
