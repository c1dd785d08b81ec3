// K1P: the fused MaxSim kernel on CTA PAIRS (thread-block clusters of 2, tcgen05 cta_group::2).
//
// Same arithmetic and the same flat page tiling as maxsim_kernel.cuh (reference: score_multi_vector,
// 05_experiment02.py:214; einsum("bnd,csd->bcns").max(3).sum(2), HF processing_colpali.py:360), but
// two SMs work on every 256-row page tile together:
//
//   * each CTA of the pair TMA-loads HALF of the page tile (32 KB) -- the pair reads the page store ONCE
//     for up to 7 query M tiles (a single CTA can keep 3 resident), and the 32 KB stages leave room
//     for a 3..5-deep TMA ring next to the resident queries;
//   * one elected thread of the leader CTA issues tcgen05.mma.cta_group::2.  The accumulators form a
//     ring of FOUR 128-column TMEM slots, each filled by one "use" of 8 k-steps x 64 cycles:
//       - a query-tile pair (one tile per CTA) against half h of the page tile: D[256 x 128] = A[256 x 128] B_h^T
//         (each CTA supplies 64 of B_h's 128 rows: CTA r holds tile rows h*128 + r*64 .. +64);
//       - for an ODD tile count, the last query tile split 64/64 rows over the pair against the WHOLE page tile:
//         D[128 x 256] in 8 x 64 cycles -- the only shape of a split tile that runs at the full tensor rate (measured,
//         profiles/micro_mma_pair_shapes_r2.txt: M128 N256 64.0 cycles per instruction = 100 %, M128 N128 57.3 = 56 %: an
//         instruction never takes less than ~57 cycles).  cute's "2x2" layout (tmem_frg_2sm<M_MMA=64>) puts it into 128
//         columns: lanes 0-63 = rows x instruction columns [0,128), lanes 64-127 = rows x [128,256); instruction columns
//         [0,128) are the leader's B rows (tile rows 0-63 and 128-191), [128,256) the peer's (64-127 and 192-255).
//     Uses are dealt round-robin: use parity = issuing warp = draining warp set, and a slot (use & 3) always meets the
//     same set: consecutive phases of its barriers are waited for by the same warps, which is what makes parity-only
//     mbarrier waits unambiguous.  With an odd number of uses per page tile (2 NF + 1) the sets swap page-tile halves
//     from one tile to the next, and take turns at the split use.
//     Four slots matter: the round trip accumulator-free -> MMA issued -> MMAs done -> epilogue awake is
//     ~1000+ cycles on top of the MMA time, so a ring of two 256-column slots caps the tensor pipe at
//     ~83 % (measured); four 128-column slots cover it.
//   * the eight epilogue warps form two sets of four (one warp per TMEM lane quarter); set s drains the
//     uses of parity s, all 128 columns of the slot per warp.  The two warps of a scheduler are therefore
//     in different phases most of the time (one waits for tcgen05.ld while the other reduces).
//     A page tile in which no page ends (3 of 4 for ColPali's 1030-token pages) takes the FAST path: per use
//     wait -> 4 x tcgen05.ld -> slot handed back -> 64 FMNMX3 into a statically indexed running maximum, unrolled
//     over the query-tile groups, no page bookkeeping at all.  Tiles with a page end run the general path.
//
// Each CTA owns the scores of ITS query rows for all pages of the pair's range, so every output
// element still has exactly one writer.  Segments must not straddle the 64-row midpoint of a tile
// (lis_plan_queries cuts there).
#pragma once
#include "maxsim_kernel.cuh"

namespace lis {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster_u32(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion is signalled on an mbarrier that may live in the peer CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t bar_cluster,
                                                 int32_t c0, int32_t c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at the same shared-memory offset in both CTAs once the MMAs issued so far are done
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"((uint16_t)3)
      : "memory");
}

constexpr int kPairWindow = 64;   // page-table window (pages)
// exchange slots between the epilogue warps and the page reducer, and the bytes of barriers / tables / exchange
// slots behind the operand stages: 8 and 10 resident query tiles leave exactly 3 KB next to 3 / 2 page stages
__host__ __device__ constexpr int pair_ex_slots(int nf) { return nf >= 4 ? 2 : 8; }
__host__ __device__ constexpr int pair_tail_bytes(int nf) { return nf >= 4 ? 3072 : 9216; }

// NF: query-tile pairs (M = 256 uses, one tile per CTA); ODD: a final tile split 64/64 over the pair (M = 128 use).
template <int NF, bool ODD, bool DBG>
__global__ void __launch_bounds__(kCtrlThreads + 256, 1)
maxsim_pair_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_p,
                   const MaxSimArgs args, const int NS) {
  constexpr int NT = 256;                 // page rows per tile
  constexpr int U = NF + (ODD ? 1 : 0);   // query-tile groups per pass ("g")
  constexpr int NACC = 4;                 // accumulator slots
  constexpr int kSlotCols = 128;
  constexpr int kAFull = kMTile * kDim * 2;        // 32 KB: this CTA's tile of a pair
  constexpr int kAHalfTile = 64 * kDim * 2;        // 16 KB: this CTA's 64 rows of the split tile
  constexpr int kABytes = NF * kAFull + (ODD ? kAHalfTile : 0);
  constexpr int kBStage = (NT / 2) * kDim * 2;     // 32 KB: [K half][box h][64 rows x 128 B]
  constexpr int kBKHalf = (NT / 2) * 128;          // 16 KB: one K half (64 elements) of the stage
  constexpr int kBBox = 64 * 128;                  //  8 KB: the CTA's 64 rows of tile half h
  constexpr int kPW = kPairWindow;
  constexpr int kEx = pair_ex_slots(NF);
  constexpr int kPairTail = pair_tail_bytes(NF);

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kABytes;
  uint8_t* tail = smem_b + (size_t)NS * kBStage;
  uint64_t* q_full = reinterpret_cast<uint64_t*>(tail);      // 1   (leader's is used)
  uint64_t* b_full = q_full + 1;                             // [8] (leader's are used)
  uint64_t* b_empty = b_full + 8;                            // [8] per CTA
  uint64_t* acc_full = b_empty + 8;                          // [4] per CTA
  uint64_t* acc_empty = acc_full + NACC;                     // [4] (leader's are used; both CTAs' warps arrive)
  uint64_t* ex_full = acc_empty + NACC;                      // [kEx]
  uint64_t* ex_empty = ex_full + kEx;                        // [kEx]
  int64_t* ex_meta = reinterpret_cast<int64_t*>(ex_empty + kEx);   // [kEx][2]
  int64_t* range = ex_meta + 2 * kEx;                              // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(range + 4);    // [2]
  float* srm = reinterpret_cast<float*>(tmem_slot + 2);            // [kEx][256]
  int32_t* pw_end = reinterpret_cast<int32_t*>(srm + kEx * 2 * kMTile);   // [kPW]
  int32_t* seginfo = pw_end + kPW;                                       // [U][2]
  uint16_t* segtab = reinterpret_cast<uint16_t*>(seginfo + 2 * U);       // [U][16]
  uint8_t* pw_clamp = reinterpret_cast<uint8_t*>(segtab + 16 * U);       // [kPW]
  static_assert((1 + 8 + 8 + 2 * NACC + 2 * kEx + 2 * kEx + 4) * 8 + 8 + kEx * 2 * kMTile * 4 + kPW * 4 + U * 8 + U * 32 + kPW <=
                    kPairTail, "tail does not fit");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int kReducerWarp = 8;
  constexpr int kProducerWarp = 9;
  constexpr int kMmaWarp = 10;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int mt_split = args.mt0 + 2 * NF;       // the tile shared 64/64 by the pair (ODD only)

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_p);
    mbar_init(q_full, 1);
    for (int s = 0; s < NS; ++s) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, kMmaWarps); }
    for (int a = 0; a < NACC; ++a) { mbar_init(acc_full + a, 1); mbar_init(acc_empty + a, 2 * 4); }
    for (int i = 0; i < kEx; ++i) { mbar_init(ex_full + i, 8); mbar_init(ex_empty + i, 1); }
    fence_barrier_init();
    // The PAIR's contiguous range of whole pages, balanced by token rows.
    const int64_t np = args.np;
    const int64_t cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
    const int64_t base = __ldg(args.p_offsets);
    const int64_t total = __ldg(args.p_offsets + np) - base;
    const int64_t per = (total + ncl - 1) / ncl;
    int64_t pa = np, pb = np;
    if (per > 0 && cid * per < total) {
      pa = lower_bound_off(args.p_offsets, np, base + cid * per);
      pb = (cid + 1 == ncl || (cid + 1) * per >= total) ? np : lower_bound_off(args.p_offsets, np, base + (cid + 1) * per);
    } else if (total == 0 && cid == 0) {
      pa = 0; pb = np;
    }
    range[0] = pa;
    range[1] = pb;
    range[2] = (pa < np) ? __ldg(args.p_offsets + pa) : 0;
    range[3] = (pa < pb) ? __ldg(args.p_offsets + pb) : range[2];
  }
  if (warp == kMmaWarp) {
    tmem_alloc_pair(tmem_slot, kTmemCols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();        // both CTAs' barriers are initialised before anyone signals across the pair
  tc_fence_after();

  const uint32_t tmem_base = *tmem_slot;
  const int64_t pa = range[0], pb = range[1];
  const int64_t row0 = range[2];
  const int64_t rows = range[3] - row0;
  const int ntiles = (int)((rows + NT - 1) / NT);

  if (warp == kProducerWarp) {
    // ===================== TMA producer (both CTAs) =====================
    if (pa < pb && ntiles > 0) {
      const uint32_t q_full_l = mapa_u32(smem_u32(q_full), 0);
      const uint32_t b_full_l = mapa_u32(smem_u32(b_full), 0);
      if (elect_one_sync()) {
        if (leader) mbar_arrive_expect_tx(q_full, 2u * kABytes);
        const uint32_t a0 = smem_u32(smem_a);
        for (int u = 0; u < NF; ++u)
          for (int h = 0; h < 2; ++h)
            for (int j = 0; j < 2; ++j)
              tma_load_2d_pair(a0 + u * kAFull + h * (kMTile * 128) + j * 8192, &tmap_q, q_full_l, h * kKHalf,
                               (args.mt0 + 2 * u + (int)rank) * kMTile + j * 64, kPolicyEvictLast);
        if (ODD)
          for (int h = 0; h < 2; ++h)
            tma_load_2d_pair(a0 + NF * kAFull + h * 8192, &tmap_q, q_full_l, h * kKHalf,
                             mt_split * kMTile + (int)rank * 64, kPolicyEvictLast);
      }
      __syncwarp();
      const uint32_t b0 = smem_u32(smem_b);
      const bool st_on = LIS_STATS_ON(args) && blockIdx.x < 2;
      long long st_wait = 0;
      const long long st_t0 = clock64();
      for (int t = 0; t < ntiles; ++t) {
        const int s = t % NS;
        const uint32_t ph = (uint32_t)(t / NS) & 1u;
        const long long c0 = st_on ? clock64() : 0;
        mbar_wait(b_empty + s, ph ^ 1u);
        if (st_on) st_wait += clock64() - c0;
        if (elect_one_sync()) {
          if (leader) mbar_arrive_expect_tx(b_full + s, 2u * kBStage);
          const uint32_t dst = b0 + (uint32_t)s * kBStage;
          const int32_t r = (int32_t)(row0 + (int64_t)t * NT) + (int32_t)rank * 64;
#pragma unroll
          for (int kh = 0; kh < 2; ++kh)
#pragma unroll
            for (int h = 0; h < 2; ++h)    // this CTA's 64 rows of tile half h
              tma_load_2d_pair(dst + kh * kBKHalf + h * kBBox, &tmap_p, b_full_l + s * 8, kh * kKHalf, r + h * 128,
                               kPolicyEvictFirst);
        }
        __syncwarp();
      }
      if (st_on && lane == 0) {
        args.stats[rank * 64 + 24] = st_wait;              // producer: waiting for a free stage
        args.stats[rank * 64 + 25] = clock64() - st_t0;    //           whole loop
      }
      // tail: every release of a stage (a multicast commit issued by the leader) has landed here before
      // this CTA may leave -- wait for the hand-back of the last tile that used each stage
      for (int t = ntiles; t < ntiles + NS; ++t)
        if (t >= NS) mbar_wait(b_empty + t % NS, ((uint32_t)(t / NS) & 1u) ^ 1u);
    }
  } else if (warp >= kMmaWarp) {
    // ===================== MMA issuers (leader CTA only) =====================
    if (leader && pa < pb && ntiles > 0) {
      const uint32_t fmt = args.is_bf16 ? 1u : 0u;
      const uint32_t idesc_full = make_idesc_f16(fmt, 256, 128);     // tile pair x half page tile
      const uint32_t idesc_split = make_idesc_f16(fmt, 128, 256);    // split tile (64 rows per CTA) x whole page tile
      const uint32_t a_base = smem_u32(smem_a);
      const uint32_t b_base = smem_u32(smem_b);
      const uint32_t acc_full_u = smem_u32(acc_full), b_empty_u = smem_u32(b_empty);
      const uint32_t b_empty_peer = mapa_u32(b_empty_u, 1);
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
        // one use: a_off / b_off = operand offsets of k-step 0; a_kh = bytes between the K halves of A
        auto issue_use = [&](uint32_t a_off, uint32_t a_kh, uint32_t b_off, uint32_t idesc) {
          if (use % kMmaWarps == my) {
            issued = true;
            const uint32_t slot = use & (NACC - 1);
            const long long w0c = st_on ? clock64() : 0;
            mbar_wait(acc_empty + slot, ((use / NACC) & 1u) ^ 1u);
            const long long w1c = st_on ? clock64() : 0;
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + slot * kSlotCols;
            if (LIS_ISSUE_PRED) {
#pragma unroll
              for (int k = 0; k < kDim / 16; ++k) {
                const uint64_t adesc = make_kmajor_sw128_desc(a_base + a_off + (uint32_t)(k >> 2) * a_kh + (uint32_t)(k & 3) * 32);
                const uint64_t bdesc = make_kmajor_sw128_desc(b_base + s * kBStage + b_off + (uint32_t)(k >> 2) * kBKHalf +
                                                              (uint32_t)(k & 3) * 32);
                umma_f16_pair(d_tmem, adesc, bdesc, idesc, k ? 1u : 0u);
              }
              umma_commit_pair(acc_full_u + slot * 8);
            }
            __syncwarp();
            if (st_on) { st_acc += w1c - w0c; st_issue += clock64() - w1c; }
          }
          ++use;
        };
#pragma unroll
        for (int g = 0; g < NF; ++g) {
          issue_use(g * kAFull, kMTile * 128, 0, idesc_full);         // page-tile half 0 (even use: issuer 0, set 0)
          issue_use(g * kAFull, kMTile * 128, kBBox, idesc_full);     // half 1
        }
        if (ODD) issue_use(NF * kAFull, 8192, 0, idesc_split);       // split tile x whole page tile: one full-rate use
        // hand the page tile back to both producers once this warp's MMAs on it are done
        if (LIS_ISSUE_PRED) {
          if (issued) umma_commit_pair(b_empty_u + s * 8);
          else { mbar_arrive_u32(b_empty_u + s * 8); mbar_arrive_cluster_u32(b_empty_peer + s * 8); }
        }
        __syncwarp();
      }
      if (st_on && lane == 0) {
        args.stats[0] = clock64() - st_t0;   // MMA warp 0: whole loop
        args.stats[1] = st_b;                //   waiting for page tiles
        args.stats[2] = st_acc;              //   waiting for a free accumulator slot
        args.stats[3] = use;
        args.stats[23] = st_issue;           //   issuing MMAs + commit
      }
    }
  } else if (warp == kReducerWarp) {
    // ===================== page reducer (both CTAs, each for its own query rows) =====================
    if (pa < pb) {
      const int is_bf16 = args.is_bf16;
      const bool round_ref = (args.round_mode & 1) != 0;
      const bool round_sum = round_ref && (args.round_mode & 2) == 0;
      const int64_t nfin = (pb - pa) * U;
      for (int64_t f = 0; f < nfin; ++f) {
        const int slot = (int)(f % kEx);
        mbar_wait(ex_full + slot, (uint32_t)(f / kEx) & 1u);
        const int64_t p = ex_meta[2 * slot];
        const int g = (int)(ex_meta[2 * slot + 1] & 0xff);
        const bool clamp = (ex_meta[2 * slot + 1] >> 8) != 0;
        const float* ex = srm + slot * (2 * kMTile);
        const bool split = ODD && g == NF;
        const int mt = split ? mt_split : args.mt0 + 2 * g + (int)rank;
        const int rbase = split ? (int)rank * 64 : 0;       // first tile row this CTA owns in this group
        const int rcnt = split ? 64 : kMTile;
        const int seg_first = seginfo[2 * g], seg_cnt = seginfo[2 * g + 1];
        // partial maxima: [warp set][row] for a tile pair, [warp set][lane half][row] for the split tile
        if (split)
          reduce_tile_segments<4>(ex, 64, segtab + g * 16, args.seg_lo, args.seg_hi, seg_first, seg_cnt, mt * kMTile, rbase, rcnt,
                                  clamp, round_ref, round_sum, is_bf16, args.out + p, args.ld_out, lane);
        else
          reduce_tile_segments<2>(ex, kMTile, segtab + g * 16, args.seg_lo, args.seg_hi, seg_first, seg_cnt, mt * kMTile, rbase,
                                  rcnt, clamp, round_ref, round_sum, is_bf16, args.out + p, args.ld_out, lane);
        __syncwarp();
        if (lane == 0) mbar_arrive(ex_empty + slot);
      }
    }
  } else {
    // ===================== epilogue (warps 0..7, both CTAs) =====================
    // Warp = (TMEM lane quarter, set); thread = accumulator lane.  A use covers half h of the page tile and
    // belongs to set h: tile columns h*128 .. +128 with thread = query row of this CTA's tile (tile pair), or
    // tile columns h*128 + L*64 .. +64 for lane half L with thread = one of this CTA's 64 rows (split tile).
    const int quarter = warp & 3;
    const int set = warp >> 2;
    const int lhalf = quarter >> 1;
    const int etid = set * kMTile + quarter * 32 + lane;
    float rm[U];
#pragma unroll
    for (int g = 0; g < U; ++g) rm[g] = -INFINITY;
    auto rotate = [&](float cur) {
#pragma unroll
      for (int i = 0; i + 1 < U; ++i) rm[i] = rm[i + 1];
      rm[U - 1] = cur;
    };

    const uint32_t acc_full_u = smem_u32(acc_full);
    const uint32_t acc_empty_l = mapa_u32(smem_u32(acc_empty), 0);   // the leader's barriers, as cluster addresses
    const uint32_t ex_full_u = smem_u32(ex_full), ex_empty_u = smem_u32(ex_empty);
    const int npages = (int)(pb - pa);

    int w0 = 0;
    auto refill = [&](int base) {
      named_bar_sync(1, 256);
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
      named_bar_sync(1, 256);
      w0 = base;
    };
    if (etid < U * 16) {                       // segment tables of this CTA's tiles
      const int g = etid >> 4, j = etid & 15;
      const int mt = (ODD && g == NF) ? mt_split : args.mt0 + 2 * g + (int)rank;
      const int first = __ldg(args.mt_seg + mt), cnt = __ldg(args.mt_seg + mt + 1) - first;
      if (j == 0) { seginfo[2 * g] = first; seginfo[2 * g + 1] = cnt; }
      if (j < cnt)
        segtab[g * 16 + j] = (uint16_t)((__ldg(args.seg_lo + first + j) - mt * kMTile) |
                                        ((__ldg(args.seg_hi + first + j) - mt * kMTile) << 8));
    }

    // Publish one finished page of group g: this thread's partial row maximum goes into the next exchange
    // slot; the reducer takes over once all eight warps have arrived.  Nobody waits for anybody here
    // unless the ring of kEx slots is full.
    long long st_wait = 0, st_hold = 0, st_hold_split = 0, st_wait_split = 0, st_n_split = 0, st_slow = 0, st_nslow = 0, st_fin = 0, st_spe_take = 0, st_spe_comp = 0, st_spe_fin = 0;
    uint32_t fin = 0;
    auto finish_page = [&](int g, int pi, float v) {
      const uint32_t slot = fin % kEx;
      const long long fw0 = LIS_STATS_ON(args) ? clock64() : 0;
      mbar_wait_u32(ex_empty_u + slot * 8, ((fin / kEx) & 1u) ^ 1u);
      if (LIS_STATS_ON(args)) st_fin += clock64() - fw0;
      srm[slot * (2 * kMTile) + etid] = v;
      if (etid == 0) {
        ex_meta[2 * slot] = pa + pi;
        ex_meta[2 * slot + 1] = (int64_t)g | ((int64_t)pw_clamp[pi - w0] << 8);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_u32(ex_full_u + slot * 8);
      ++fin;
    };

    if (pa < pb) {
      int p = 0;
      refill(0);
      int pend = pw_end[0];
      uint32_t use_base = 0;                            // first use of the current (tile, group)
      const long long st_t0 = clock64();
      const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);
      const bool st_on = LIS_STATS_ON(args) && blockIdx.x < 2;
      for (int t = 0; t < ntiles; ++t) {
        const int tcol = t * NT;
        auto rel_end = [&](int e) { const int d = e - tcol; return d > NT ? NT + 1 : d; };
        const int pe_tile = rel_end(pend);

        // Take this set's use `my_use`: all 128 columns of the slot into registers, slot handed back at once.
        // A use is read out in two steps so that at most 96 accumulator values are in registers at once (with all 128
        // the page cursor no longer fits the 168 registers of a 384-thread CTA, and a spill is an L2 round trip here):
        // take_a waits for the slot and loads chunks rot+1, rot+2, rot+3 (mod 4) into v0..v2; after the caller has folded
        // them, take_b loads chunk rot into v3 and hands the slot back.  rot = 3 is the plain order.
        auto take_a = [&](const uint32_t my_use, const uint32_t rot, uint32_t (&v0)[32], uint32_t (&v1)[32], uint32_t (&v2)[32]) {
          const uint32_t slot = my_use & (NACC - 1);
          const long long ec0 = st_on ? clock64() : 0;
          mbar_wait_u32(acc_full_u + slot * 8, (my_use / NACC) & 1u);
          if (st_on) st_wait += clock64() - ec0;
          tc_fence_after();
          const uint32_t taddr = tlane + slot * kSlotCols;
          tmem_ld32(taddr + 32u * ((rot + 1u) & 3u), v0);
          tmem_ld32(taddr + 32u * ((rot + 2u) & 3u), v1);
          tmem_ld32(taddr + 32u * ((rot + 3u) & 3u), v2);
          tmem_ld_wait();
        };
        auto take_b = [&](const uint32_t my_use, const uint32_t rot, uint32_t (&v3)[32]) {
          const uint32_t slot = my_use & (NACC - 1);
          const long long hc0 = st_on ? clock64() : 0;
          tmem_ld32(tlane + slot * kSlotCols + 32u * rot, v3);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_u32(acc_empty_l + slot * 8);    // slot back to the MMA warps
          if (st_on) st_hold += clock64() - hc0;
        };

        const bool live_tile = p < npages;
        // (votes make the path selection a warp-uniform branch: no convergence-barrier bookkeeping around the loop body)
        if (!DBG && __all_sync(0xffffffffu, !live_tile || pe_tile > NT)) {
          // ---------- FAST path: no page ends inside this tile.  A runtime loop over the groups (ONE copy of the
          // body; the running maxima rotate through rm[]): unrolling it per group made the hot loop ~1000 instructions
          // and the page-end paths, entered every fourth tile, then started from a cold instruction cache.
#pragma unroll 1
          for (int g = 0; g < U; ++g) {
            const bool split = ODD && g == NF;
            // a tile pair: uses 2g (tile half 0) and 2g+1 (half 1), the one whose parity equals `set` is ours;
            // the split tile: ONE use, which the sets take in turns
            const uint32_t my_use = split ? use_base : use_base + ((use_base ^ (uint32_t)set) & 1u);
            const bool have = !split || (use_base & 1u) == (uint32_t)set;
            use_base += split ? 1u : 2u;
            float m = rm[0];
            if (have) {
              {
                uint32_t v0[32], v1[32], v2[32];
                take_a(my_use, 3u, v0, v1, v2);
                if (live_tile) {
                  m = max32(v0, m);
                  m = max32(v1, m);
                  m = max32(v2, m);
                }
              }
              uint32_t v3[32];
              take_b(my_use, 3u, v3);
              if (live_tile) m = max32(v3, m);
            }
            rotate(m);
          }
          continue;
        }

        // ---------- ONE page ends inside this tile (at tile column e) and the next page reaches beyond it: the only
        // kind of page-end tile a corpus of pages longer than 256 tokens produces (ColPali: every fourth tile).  The page
        // structure is the same for all groups, so it is resolved once here; per group a warp folds its chunks into the
        // ending page (columns < e) or into the next one (columns >= e), publishes the former and keeps the latter.
        if (!DBG && __all_sync(0xffffffffu, pe_tile > 0 && p + 1 < npages && p + 1 < w0 + kPW && p >= w0 &&
                                                pw_end[p + 1 < w0 + kPW && p + 1 >= w0 ? p + 1 - w0 : 0] - tcol > NT)) {
          const long long sl0 = st_on ? clock64() : 0;
          const int e = pe_tile;
#pragma unroll 1
          for (int g = 0; g < U; ++g) {
            const bool split = ODD && g == NF;
            uint32_t my_use = use_base;
            int cb0, cb1;
            bool have = true;
            if (!split) {
              const int h = (int)((use_base ^ (uint32_t)set) & 1u);
              my_use = use_base + (uint32_t)h;
              cb0 = h * 128; cb1 = cb0 + 64;
            } else {
              have = (use_base & 1u) == (uint32_t)set;
              cb0 = lhalf * 64; cb1 = 128 + lhalf * 64;
            }
            use_base += split ? 1u : 2u;
            float m_old = rm[0], m_new = -INFINITY;
            if (have) {
              // The warp's chunks start at tile columns cb0, cb0+32, cb1, cb1+32 (ascending).  Chunk c is the first one
              // that does not lie entirely before the page end; it is cut at column b (0: it belongs to the next page
              // altogether, 32: all four chunks belong to the page that ends).  The chunks are loaded ROTATED so that
              // chunk c always sits in v3: one copy of the cut code serves every position of the page end.
              int n_old = (e >= cb0 + 32) + (e >= cb0 + 64) + (e >= cb1 + 32) + (e >= cb1 + 64);
              int b = 32;
              if (n_old < 4) {
                const int start = n_old < 2 ? cb0 + 32 * n_old : cb1 + 32 * (n_old - 2);
                b = e - start;
                if (b < 0) b = 0;
              } else n_old = 3;
              const uint32_t c = (uint32_t)n_old;
              const long long q0 = st_on ? clock64() : 0;
              float f0, f1, f2;
              {
                uint32_t v0[32], v1[32], v2[32];
                take_a(my_use, c, v0, v1, v2);
                // v0, v1, v2 = chunks c+1, c+2, c+3 (mod 4): behind c -> next page, before c (wrapped around) -> ending page
                f0 = max32(v0, -INFINITY); f1 = max32(v1, -INFINITY); f2 = max32(v2, -INFINITY);
              }
              uint32_t v3[32];
              take_b(my_use, c, v3);
              const long long q1 = st_on ? clock64() : 0;
              const bool n0 = c + 1u < 4u, n1 = c + 2u < 4u, n2 = c + 3u < 4u;
              m_new = fmax3(m_new, n0 ? f0 : -INFINITY, n1 ? f1 : -INFINITY);
              m_new = fmaxf(m_new, n2 ? f2 : -INFINITY);
              m_old = fmax3(m_old, n0 ? -INFINITY : f0, n1 ? -INFINITY : f1);
              m_old = fmaxf(m_old, n2 ? -INFINITY : f2);
              {   // branch-free cut (a switch on b compiles into a tree of indirect jumps, ~1000 cycles)
                float cut_new;
                m_old = max32_split(v3, m_old, b, cut_new);
                m_new = fmaxf(m_new, cut_new);
              }
              if (st_on) { st_spe_take += q1 - q0; st_spe_comp += clock64() - q1; }
            }
            const long long q2 = st_on ? clock64() : 0;
            finish_page(g, p, m_old);
            rotate(m_new);
            if (st_on) st_spe_fin += clock64() - q2;
          }
          ++p;
          pend = pw_end[p - w0];
          if (st_on) { st_slow += clock64() - sl0; ++st_nslow; }
          continue;
        }

        // ---------- general path: pages end inside this tile.  One query-tile group at a time (ONE instance of the
        // code: a runtime loop over the groups); every warp walks the pages that end inside the tile (it must publish its
        // partial maxima for each of them); at most one use of the group is its own.  Each group consumes rm[0] and
        // pushes the new running maximum to the back, so after the U groups of a tile the ring is back in the order
        // the fast path indexes.
        const long long sl0 = st_on ? clock64() : 0;
        int pp = p, ppend = pend;
        if (live_tile && (p < w0 || p >= w0 + kPW)) refill(p);   // (rare) cursor rewound out of the window
#pragma unroll 1
        for (int g = 0; g < U; ++g) {
          const bool split = ODD && g == NF;
          pp = p; ppend = pend;
          bool live = live_tile;
          int pe = pe_tile;
          float m = rm[0];

          auto next_page = [&]() {
            m = -INFINITY;
            ++pp;
            if (pp >= npages) { live = false; pe = NT + 1; return; }
            if (pp >= w0 + kPW) refill(pp - p < kPW ? p : pp);
            ppend = pw_end[pp - w0];
            pe = rel_end(ppend);
          };
          auto skip_to = [&](int col_end) {          // pages that end at or before tile column col_end
            while (live && pe <= col_end) { finish_page(g, pp, m); next_page(); }
          };
          auto scan = [&](const uint32_t (&v)[32], int cb) {      // one 32-column chunk starting at tile column cb
            if (DBG) {
              if (blockIdx.x < 2 && t == 0 && args.dbg != nullptr) {
                const int qrow = split ? mt_split * kMTile + (int)rank * 64 + (quarter & 1) * 32 + lane
                                       : (args.mt0 + 2 * g + (int)rank) * kMTile + quarter * 32 + lane;
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  args.dbg[(int64_t)(qrow - args.mt0 * kMTile) * NT + cb + i] = __uint_as_float(v[i]);
              }
            }
            if (!live) return;
            if (pe - cb > 32) { m = max32(v, m); return; }   // no page ends inside this chunk
            int lo = 0;                                      // (pe > cb here: earlier pages were closed already)
            while (true) {
              const int rel = pe - cb;                       // > 32: the page continues behind this chunk
              const int hi = rel < 32 ? rel : 32;
              m = max32_masked(v, m, lo, hi);
              if (rel > 32) break;
              finish_page(g, pp, m);                         // the page ended inside (or exactly at the end of) the chunk
              next_page();
              if (!live) break;
              lo = hi;
              if (lo >= 32) break;
            }
          };

          // Which use of this group, if any, belongs to this warp's set, and which tile columns it covers: a tile pair
          // has one use per half of the page tile (128 consecutive tile columns); the split tile has ONE use over the
          // whole page tile, in which this warp's lanes hold tile columns [lhalf*64, +64) and [128 + lhalf*64, +64).
          uint32_t my_use = use_base;
          int cb0, cb1;
          bool have = true;
          if (!split) {
            const int h = (int)((use_base ^ (uint32_t)set) & 1u);
            my_use = use_base + (uint32_t)h;
            cb0 = h * 128; cb1 = cb0 + 64;
          } else {
            have = (use_base & 1u) == (uint32_t)set;
            cb0 = lhalf * 64; cb1 = 128 + lhalf * 64;
          }
          use_base += split ? 1u : 2u;
          if (have) {
            {
              uint32_t v0[32], v1[32], v2[32];
              take_a(my_use, 3u, v0, v1, v2);
              skip_to(cb0);
              scan(v0, cb0);
              scan(v1, cb0 + 32);
              skip_to(cb1);
              scan(v2, cb1);
            }
            uint32_t v3[32];
            take_b(my_use, 3u, v3);
            scan(v3, cb1 + 32);
          }
          skip_to(NT);
          rotate(m);
        }
        p = pp; pend = ppend;
        if (st_on) { st_slow += clock64() - sl0; ++st_nslow; }
      }
      if (LIS_STATS_ON(args) && blockIdx.x < 2 && lane == 0) {
        args.stats[rank * 64 + 4 + 2 * warp] = st_wait;     // epilogue warp: waiting for a full accumulator slot
        args.stats[rank * 64 + 5 + 2 * warp] = st_hold;     //                wake -> release
        if (warp == 0) args.stats[rank * 64 + 26] = clock64() - st_t0;
        if (warp == 0 || warp == 4) { args.stats[rank * 64 + 28 + warp] = st_hold_split; args.stats[rank * 64 + 29 + warp] = st_n_split; args.stats[rank * 64 + 30 + warp] = st_wait_split; }
        if (warp == 0 || warp == 4) { args.stats[rank * 64 + 40 + warp] = st_slow; args.stats[rank * 64 + 41 + warp] = st_nslow; args.stats[rank * 64 + 42 + warp] = st_fin; args.stats[rank * 64 + 43 + warp] = ntiles; }
        if (warp == 0 || warp == 4) { args.stats[rank * 64 + 48 + warp] = st_spe_take; args.stats[rank * 64 + 49 + warp] = st_spe_comp; args.stats[rank * 64 + 50 + warp] = st_spe_fin; }
      }
      while (p < npages) {          // pages not closed by any tile: trailing empty pages (or ntiles == 0)
        if (p < w0 || p >= w0 + kPW) refill(p);
#pragma unroll 1
        for (int g = 0; g < U; ++g) {
          finish_page(g, p, rm[0]);
          rotate(-INFINITY);
        }
        ++p;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();        // nobody leaves while the peer may still read its shared memory or signal its barriers
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

}  // namespace lis
