// C-ABI entry points for the scoring path: query packing, K1 launcher, segment reduction.
// Interfaces replaced: see include/lis.h.
#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

#include "lis_common.h"
#include "maxsim_kernel.cuh"

namespace lis {

// Queries are cut into segments at multiples of 64 packed rows: an M tile holds 128 rows, and the CTA-pair
// kernel splits the last tile of an odd pass 64/64 over two SMs (each sums the segments of its half).
constexpr int kSegCut = 64;
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
std::atomic<int64_t> g_launches{0};
Tuning g_tuning;   // experiment knobs: written under g_tuning_mu, read through tuning_snapshot()
static std::mutex g_tuning_mu;
Tuning tuning_snapshot() {
  std::lock_guard<std::mutex> lock(g_tuning_mu);
  return g_tuning;
}
static long long* g_stats = nullptr;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

int encode_rows_tmap(CUtensorMap* map, const void* base, int64_t rows, int box_rows, int dtype) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return LIS_E_CUDA;
  }
  LIS_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor base %p is not 16-byte aligned", base);
  LIS_REQUIRE(rows > 0 && rows < (int64_t(1) << 31), "row count %lld out of range for one tensor map",
              (long long)rows);
  cuuint64_t dims[2] = {(cuuint64_t)kDim, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)kDim * 2};
  cuuint32_t box[2] = {(cuuint32_t)kKHalf, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, dtype == LIS_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                   2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld box_rows=%d)", (int)r,
              (long long)rows, box_rows);
    return LIS_E_CUDA;
  }
  return LIS_OK;
}

int sm_count(int device) {
  static int cached[64] = {0};
  if (device >= 0 && device < 64 && cached[device]) return cached[device];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 0;
  if (device >= 0 && device < 64) cached[device] = n;
  return n;
}

// ---------------------------------------------------------------------------------------------
constexpr int kSmemBudget = 232448;  // 227 KB opt-in maximum per CTA on sm_100
constexpr int kSmemTail = 3072;      // barriers (240 B) + row-max exchange (<= 2048 B)

struct Maps {
  CUtensorMap q, p, q2, p2;  // q2/p2: low planes of split-fp32 operands (copies of q/p otherwise)
};

template <int NT, int G, int EH, bool ATM, bool DBG, int P = 1>
static int launch_maxsim(const Maps& m, const MaxSimArgs& a, int grid, cudaStream_t st) {
  const int stage = P * NT * kDim * 2;
  const int a_bytes = ATM ? 0 : P * G * kATileBytes;
  int ns = (kSmemBudget - kSmemTail - a_bytes) / stage;
  ns = std::min(ns, 8);
  if (ns < 1) {
    set_error("tile_n=%d group=%d does not fit shared memory", NT, G);
    return LIS_E_INVALID;
  }
  const int smem = a_bytes + ns * stage + kSmemTail;
  if (a.n_mt != G) {
    set_error("internal: n_mt=%d must equal the instantiated group %d", a.n_mt, G);
    return LIS_E_INVALID;
  }
  auto kern = maxsim_kernel<NT, G, EH, ATM, DBG, P>;
  static std::atomic<bool> configured[64];  // per template instantiation and device
  int dev = 0;
  LIS_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
    LIS_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_release);
  }
  kern<<<grid, kCtrlThreads + 128 * EH, smem, st>>>(m.q, m.p, m.q2, m.p2, a, ns);
  count_launch();
  LIS_CUDA_CHECK(cudaGetLastError());
  return LIS_OK;
}

// Instantiations.  SS form (A in shared memory): NT 256 x G 1..3, NT 128 x G 1..5.
// TS form (A in tensor memory): 64*G + NACC*NT <= 512 columns -> NT 128 x G 1..4, NT 192 x G 1..2.
static int dispatch_maxsim(const Tuning& tn, int nt, int g, bool atm, const Maps& m, const MaxSimArgs& a, int grid,
                           cudaStream_t st, bool dbg = false, int planes = 1) {
  if (planes == 2) {  // split fp32: one resident M tile (2 planes) + 2 stages of 2-plane 128-row tiles
    if (g == 1) return launch_maxsim<128, 1, 2, false, false, 2>(m, a, grid, st);
    set_error("split-fp32 scoring runs one M tile per pass");
    return LIS_E_INVALID;
  }
  if (dbg) {
    if (nt == 256 && g == 1 && !atm) return launch_maxsim<256, 1, 2, false, true>(m, a, grid, st);
    if (nt == 128 && g == 1 && !atm) return launch_maxsim<128, 1, 1, false, true>(m, a, grid, st);
    if (nt == 128 && g == 1 && atm) return launch_maxsim<128, 1, 2, true, true>(m, a, grid, st);
    if (nt == 192 && g == 1 && atm) return launch_maxsim<192, 1, 2, true, true>(m, a, grid, st);
  }
  const int eh = tn.epi_halves ? tn.epi_halves : 2;
#define LIS_CASE(NT_, G_, ATM_)                                                                \
  if (nt == NT_ && g == G_ && atm == ATM_)                                                     \
    return eh == 2 ? launch_maxsim<NT_, G_, 2, ATM_, false>(m, a, grid, st)                    \
                   : launch_maxsim<NT_, G_, 1, ATM_, false>(m, a, grid, st);
  LIS_CASE(256, 1, false) LIS_CASE(256, 2, false) LIS_CASE(256, 3, false)
  LIS_CASE(128, 1, false) LIS_CASE(128, 2, false) LIS_CASE(128, 3, false) LIS_CASE(128, 4, false)
  LIS_CASE(128, 5, false)
  LIS_CASE(128, 1, true) LIS_CASE(128, 2, true) LIS_CASE(128, 3, true) LIS_CASE(128, 4, true)
  LIS_CASE(192, 1, true) LIS_CASE(192, 2, true)
#undef LIS_CASE
  set_error("unsupported tiling tile_n=%d group=%d a_in_tmem=%d", nt, g, (int)atm);
  return LIS_E_INVALID;
}

// CTA-pair form (lis_maxsim_pair.cu)
int dispatch_maxsim_pair(const CUtensorMap& q, const CUtensorMap& p, const MaxSimArgs& a, int grid, cudaStream_t st,
                         bool dbg);

static int max_group(int nt, bool atm) {
  if (atm) return nt == 128 ? 4 : (nt == 192 ? 2 : 0);
  return nt == 256 ? 3 : (nt == 128 ? 5 : 0);
}

}  // namespace lis

using namespace lis;

extern "C" {

const char* lis_last_error(void) { return g_err; }
int lis_abi_version(void) { return LIS_ABI_VERSION; }
int64_t lis_launch_count(void) { return g_launches.load(); }
int lis_k1_stats(long long* device_buf) {
  g_stats = device_buf;  // 8 x int64 on the device, zeroed by the caller; null switches the counters off
  return LIS_OK;
}
int lis_set_pass_costs(const float* single, const float* pair) {
  // single[1..3], pair[2..10]: cost of one pass with that many resident query tiles (same unit for both); NULL = defaults
  std::lock_guard<std::mutex> lock(g_tuning_mu);
  const Tuning def;
  for (int i = 0; i < 4; ++i) g_tuning.cost_single[i] = def.cost_single[i];
  for (int i = 0; i < 11; ++i) g_tuning.cost_pair[i] = def.cost_pair[i];
  if (single)
    for (int i = 1; i < 4; ++i) {
      LIS_REQUIRE(single[i] > 0.f, "lis_set_pass_costs: single[%d] must be positive", i);
      g_tuning.cost_single[i] = single[i];
    }
  if (pair)
    for (int i = 2; i < 11; ++i) {
      LIS_REQUIRE(pair[i] > 0.f, "lis_set_pass_costs: pair[%d] must be positive", i);
      g_tuning.cost_pair[i] = pair[i];
    }
  return LIS_OK;
}

int lis_set_ablation(int mode) {
  LIS_REQUIRE(mode >= 0 && mode <= 4, "ablation mode must be in 0..4");
  std::lock_guard<std::mutex> lock(g_tuning_mu);
  g_tuning.ablate = mode;
  return LIS_OK;
}

int lis_device_supported(int device) {
  int major = 0;
  LIS_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  if (major != 10) {
    set_error("device %d has compute capability %d.x; this library is built for sm_100a only", device, major);
    return LIS_E_UNSUPPORTED;
  }
  return LIS_OK;
}

int lis_set_tuning(int tile_n, int group, int max_ctas, int epi_halves, int a_operand) {
  LIS_REQUIRE(tile_n == 0 || tile_n == 128 || tile_n == 192 || tile_n == 256, "tile_n must be 0, 128, 192 or 256");
  LIS_REQUIRE(a_operand >= 0 && a_operand <= 3,
              "a_operand must be 0 (auto), 1 (shared memory), 2 (tensor memory) or 3 (CTA pairs)");
  LIS_REQUIRE(epi_halves >= 0 && epi_halves <= 2, "epi_halves must be 0 (auto), 1 or 2");
  LIS_REQUIRE(max_ctas >= 0, "max_ctas must be >= 0");
  LIS_REQUIRE(group >= 0 && group <= 10, "group must be in 0..10");
  LIS_REQUIRE(a_operand == 3 || (a_operand == 0 && tile_n == 0) || group <= 5, "group > 5 needs the CTA-pair form");
  LIS_REQUIRE(a_operand != 3 || group == 0 || group >= 2, "the CTA-pair form keeps 2..10 query tiles per pass");
  LIS_REQUIRE(a_operand != 3 || tile_n == 0 || tile_n == 256, "the CTA-pair form uses 256-row page tiles");
  if (tile_n && a_operand && a_operand != 3) {
    const int gm = max_group(tile_n, a_operand == 2);
    LIS_REQUIRE(gm > 0, "tile_n=%d is not available with a_operand=%d", tile_n, a_operand);
    LIS_REQUIRE(group <= gm, "tile_n=%d a_operand=%d supports group <= %d", tile_n, a_operand, gm);
  }
  LIS_REQUIRE(!(tile_n == 256 && group > 3 && a_operand != 3), "tile_n=256 supports group <= 3 on a single CTA");
  std::lock_guard<std::mutex> lock(g_tuning_mu);
  g_tuning.tile_n = tile_n;
  g_tuning.group = group;
  g_tuning.max_ctas = max_ctas;
  g_tuning.epi_halves = epi_halves;
  g_tuning.a_operand = a_operand;
  return LIS_OK;
}

int64_t lis_plan_queries(const int32_t* q_lens, int64_t nq, int64_t cap, int32_t* seg_query, int32_t* seg_lo,
                         int32_t* seg_hi, int64_t mt_cap, int32_t* mt_seg, int64_t* n_mtiles) {
  if (nq < 0 || (nq > 0 && q_lens == nullptr)) {
    set_error("lis_plan_queries: bad arguments");
    return LIS_E_INVALID;
  }
  const bool write = cap > 0;
  int64_t n_seg = 0;
  int64_t row = 0;  // next packed row
  for (int64_t q = 0; q < nq; ++q) {
    int64_t len = q_lens[q];
    if (len < 0) {
      set_error("lis_plan_queries: negative length for query %lld", (long long)q);
      return LIS_E_INVALID;
    }
    while (len > 0) {
      const int64_t room = kSegCut - (row % kSegCut);  // rows left before the next cut
      const int64_t take = std::min<int64_t>(len, room);
      if (write) {
        if (n_seg >= cap) {
          set_error("lis_plan_queries: segment capacity %lld too small", (long long)cap);
          return LIS_E_INVALID;
        }
        seg_query[n_seg] = (int32_t)q;
        seg_lo[n_seg] = (int32_t)row;
        seg_hi[n_seg] = (int32_t)(row + take);
      }
      ++n_seg;
      row += take;
      len -= take;
    }
    if (row >= (int64_t(1) << 31) - kMTile) {
      set_error("lis_plan_queries: too many query rows");
      return LIS_E_INVALID;
    }
  }
  const int64_t tiles = (row + kMTile - 1) / kMTile;
  if (n_mtiles) *n_mtiles = tiles;
  if (write) {
    if (mt_cap < tiles + 1) {
      set_error("lis_plan_queries: mt_seg capacity %lld too small (need %lld)", (long long)mt_cap,
                (long long)(tiles + 1));
      return LIS_E_INVALID;
    }
    int64_t s = 0;
    for (int64_t t = 0; t <= tiles; ++t) {
      while (s < n_seg && seg_lo[s] < t * kMTile) ++s;
      mt_seg[t] = (int32_t)s;
    }
  }
  return n_seg;
}

}  // extern "C"

namespace lis {
__global__ void reduce_segments_kernel(const float* __restrict__ seg, int64_t ld_seg,
                                       const int32_t* __restrict__ seg_first, int64_t np, int round_mode,
                                       int is_bf16, float* __restrict__ out, int64_t ld_out) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t q = blockIdx.y;
  if (p >= np) return;
  const int a = __ldg(seg_first + q), b = __ldg(seg_first + q + 1);
  float acc = 0.f;
  for (int s = a; s < b; ++s) acc += __ldg(seg + (int64_t)s * ld_seg + p);
  if (round_mode & LIS_ROUND_REFERENCE) acc = round_to_input_dtype(acc, is_bf16);
  out[q * ld_out + p] = acc;
}
}  // namespace lis

extern "C" {

int lis_reduce_segments(const float* seg_scores, int64_t ld_seg, const int32_t* seg_first, int64_t nq,
                        int64_t np, int round_mode, int dtype, float* out, int64_t ld_out, void* stream) {
  LIS_REQUIRE(seg_scores && seg_first && out, "lis_reduce_segments: null pointer");
  LIS_REQUIRE(nq > 0 && np > 0 && nq < 65536, "lis_reduce_segments: bad shape nq=%lld np=%lld",
              (long long)nq, (long long)np);
  dim3 grid((unsigned)((np + 255) / 256), (unsigned)nq);
  reduce_segments_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(seg_scores, ld_seg, seg_first, np, round_mode,
                                                                dtype == LIS_BF16, out, ld_out);
  count_launch();
  LIS_CUDA_CHECK(cudaGetLastError());
  return LIS_OK;
}

static int current_device_sm_count() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  return sm_count(dev);
}

// Tiling policy (measured on B200, profiles/).  The op's arithmetic intensity is (query rows) FLOP per
// page byte: up to 2 query M tiles a pass is HBM-bound, beyond that tensor-bound.  In both regimes the
// SS form with NT = 256 measured best: 64 KB TMA tiles keep HBM saturated, and N = 256 is the only
// shape whose operand fetch (96 B/clk) does not exceed what shared memory delivers to the tensor
// pipe.  Up to 3 M tiles stay resident (A 96 KB + 2 B stages of 64 KB); more tiles -> several
// balanced passes (5 -> 3+2).  The TS form (queries in tensor memory, a_operand = 2) is functionally
// identical and kept selectable; on a power-capped B200 it measured 3-8 % slower (sweep_r1_v4).
static void choose_tiling(const Tuning& tn, int64_t n_mtiles, int* nt, int* g, bool* atm) {
  if (tn.a_operand) *atm = tn.a_operand == 2;
  else *atm = false;
  if (tn.tile_n) *nt = tn.tile_n;
  else *nt = *atm ? 128 : 256;
  int gmax = max_group(*nt, *atm);
  if (gmax == 0) {  // inconsistent override: fall back to the default tile for this form
    *nt = *atm ? 128 : 256;
    gmax = max_group(*nt, *atm);
  }
  if (tn.group) {
    *g = std::min(tn.group, gmax);
    return;
  }
  const int64_t passes = (n_mtiles + gmax - 1) / gmax;
  *g = (int)((n_mtiles + passes - 1) / passes);
}

// Pass plan: which kernel form takes how many query tiles in each pass over the store.
struct Pass { bool pair; int n; };
static void build_pass_plan(const Tuning& tn, int64_t n_mtiles, bool must_single, int g_single, std::vector<Pass>& plan) {
  plan.clear();
  const bool forced_single = must_single || tn.a_operand == 1 || tn.a_operand == 2 ||
                             tn.tile_n != 0 || tn.group == 1;
  const int g = g_single;
  if (forced_single) {
    for (int64_t t = 0; t < n_mtiles; t += g) plan.push_back({false, (int)std::min<int64_t>(g, n_mtiles - t)});
  } else if (tn.a_operand == 3) {
    // experiments: CTA pairs only, balanced passes of at most `group` tiles (a leftover single tile runs on one CTA)
    const int gmax = tn.group ? std::min(std::max(tn.group, 2), 10) : 6;
    const int64_t passes = (n_mtiles + gmax - 1) / gmax;
    const int gp = (int)((n_mtiles + passes - 1) / passes);
    for (int64_t t = 0; t < n_mtiles; t += gp) {
      const int n = (int)std::min<int64_t>(gp, n_mtiles - t);
      plan.push_back({n >= 2, n});
    }
  } else {
    // auto: the cheapest decomposition of n_mtiles by the measured steady-state cost of one pass of each
    // form (ms per 60 000 ColPali pages on a power-capped B200, scripts/gpu_pass_costs.py -> profiles/).
    // One CTA per SM for 1 and 2 tiles (a single tile is HBM-bound; at 2 tiles the two forms tie); CTA pairs from 3
    // tiles on, every count 3..10 in ONE pass (an odd count ends with the full-rate N = 256 use of the split tile);
    // 8 and 10 tiles amortise the page stream a little more.
    const float* cost_single = tn.cost_single;
    const float* cost_pair = tn.cost_pair;     // 0 = shape not available
    const int gcap = tn.group ? tn.group : 10;    // group = most tiles a pass may hold
    std::vector<float> best((size_t)n_mtiles + 1, 1e30f);
    std::vector<int8_t> choice((size_t)n_mtiles + 1, 0);       // +n = single pass of n tiles, -n = pair pass
    best[0] = 0.f;
    for (int64_t t = 1; t <= n_mtiles; ++t) {
      for (int n = 1; n <= 3 && n <= t && n <= gcap; ++n)
        if (best[t - n] + cost_single[n] < best[t]) { best[t] = best[t - n] + cost_single[n]; choice[t] = (int8_t)n; }
      for (int n = 2; n <= 10 && n <= t && n <= gcap; ++n)
        if (cost_pair[n] > 0.f && best[t - n] + cost_pair[n] < best[t]) { best[t] = best[t - n] + cost_pair[n]; choice[t] = (int8_t)-n; }
    }
    for (int64_t t = n_mtiles; t > 0;) {
      const int c = choice[t];
      plan.push_back({c < 0, c < 0 ? -c : c});
      t -= c < 0 ? -c : c;
    }
    std::sort(plan.begin(), plan.end(), [](const Pass& a, const Pass& b) { return a.pair != b.pair ? a.pair : a.n > b.n; });
  }
}

static int maxsim_impl(const void* q, const void* q_lo, int64_t q_rows, const int32_t* seg_lo, const int32_t* seg_hi,
                       const int32_t* mt_seg, int64_t n_seg, int64_t n_mtiles, const void* tokens,
                       const void* tokens_lo, int64_t n_rows, const int64_t* p_offsets, const uint8_t* p_clamp,
                       int64_t np, int dtype, int round_mode, float* out, int64_t ld_out, void* stream) {
  const int planes = q_lo ? 2 : 1;
  const Tuning tn = tuning_snapshot();   // one consistent view of the experiment knobs per call
  LIS_REQUIRE(q && seg_lo && seg_hi && mt_seg && p_offsets && out, "maxsim: null pointer");
  LIS_REQUIRE(dtype == LIS_BF16 || dtype == LIS_F16, "maxsim: dtype must be bf16 or f16");
  LIS_REQUIRE(round_mode >= 0 && round_mode <= 3, "bad round_mode");
  LIS_REQUIRE(n_seg > 0 && n_mtiles > 0 && np > 0, "maxsim: empty problem (n_seg=%lld np=%lld)",
              (long long)n_seg, (long long)np);
  LIS_REQUIRE(q_rows > 0 && q_rows > (n_mtiles - 1) * kMTile && q_rows <= n_mtiles * kMTile,
              "q_rows=%lld inconsistent with n_mtiles=%lld", (long long)q_rows, (long long)n_mtiles);
  LIS_REQUIRE(ld_out >= np, "ld_out < np");
  LIS_REQUIRE(n_rows >= 0, "n_rows < 0");
  LIS_REQUIRE(n_rows == 0 || tokens, "maxsim: null token store");
  LIS_REQUIRE(planes == 1 || n_rows == 0 || tokens_lo, "maxsim: null low plane");
  cudaStream_t st = (cudaStream_t)stream;
  int sms = current_device_sm_count();
  LIS_REQUIRE(sms > 0, "no CUDA device");

  int nt, g;
  bool atm;
  if (planes == 2) { nt = 128; g = 1; atm = false; }
  else choose_tiling(tn, n_mtiles, &nt, &g, &atm);
  std::vector<Pass> plan;
  build_pass_plan(tn, n_mtiles, planes == 2 || n_rows == 0 || sms < 2, g, plan);
  if (planes != 2 && !(tn.a_operand == 1 || tn.a_operand == 2 || tn.tile_n != 0)) { nt = 256; atm = false; }
  bool pair = false;
  for (const Pass& ps : plan) pair = pair || ps.pair;
  Maps m;   // single-CTA form: 128-row query box, nt-row page box
  Maps mp;  // pair form: 64-row boxes for queries and pages
  bool have_single = false;
  auto single_maps = [&]() -> int {
    if (have_single) return LIS_OK;
    int rc = encode_rows_tmap(&m.q, q, q_rows, kMTile, dtype);
    if (rc) return rc;
    // an empty token store still needs a valid map: point it at the query rows (never loaded)
    rc = n_rows > 0 ? encode_rows_tmap(&m.p, tokens, n_rows, nt, dtype) : encode_rows_tmap(&m.p, q, q_rows, nt, dtype);
    if (rc) return rc;
    m.q2 = m.q;
    m.p2 = m.p;
    if (planes == 2) {
      rc = encode_rows_tmap(&m.q2, q_lo, q_rows, kMTile, dtype);
      if (rc) return rc;
      if (n_rows > 0) {
        rc = encode_rows_tmap(&m.p2, tokens_lo, n_rows, nt, dtype);
        if (rc) return rc;
      }
    }
    have_single = true;
    return LIS_OK;
  };
  int rc = LIS_OK;
  if (pair) {
    rc = encode_rows_tmap(&mp.q, q, q_rows, 64, dtype);
    if (rc) return rc;
    rc = encode_rows_tmap(&mp.p, tokens, n_rows, 64, dtype);
    if (rc) return rc;
  }

  int grid = sms;
  if (tn.max_ctas > 0) grid = std::min(grid, tn.max_ctas);
  grid = (int)std::min<int64_t>(grid, std::max<int64_t>(np, 1));
  int grid_pair = sms & ~1;
  if (tn.max_ctas > 0) grid_pair = std::max(2, std::min(grid_pair, tn.max_ctas & ~1));
  grid_pair = (int)std::min<int64_t>(grid_pair, 2 * std::max<int64_t>(np, 1));

  int64_t mt0 = 0;
  for (const Pass& ps : plan) {
    MaxSimArgs a;
    a.p_offsets = p_offsets;
    a.p_clamp = p_clamp;
    a.seg_lo = seg_lo;
    a.seg_hi = seg_hi;
    a.mt_seg = mt_seg;
    a.out = out;
    a.dbg = nullptr;
    a.q = q;
    a.q_rows = q_rows;
    a.ld_out = ld_out;
    a.np = np;
    a.mt0 = (int32_t)mt0;
    a.n_mt = ps.n;
    a.round_mode = round_mode;
    a.is_bf16 = dtype == LIS_BF16;
    a.ablate = tn.ablate;
    a.stats = g_stats;
    if (ps.pair) {
      rc = dispatch_maxsim_pair(mp.q, mp.p, a, grid_pair, st, false);
    } else {
      // the instantiation whose group equals this pass's tile count
      rc = single_maps();
      if (rc) return rc;
      rc = dispatch_maxsim(tn, nt, a.n_mt, atm, m, a, grid, st, false, planes);
    }
    if (rc) return rc;
    mt0 += ps.n;
  }
  return LIS_OK;
}

int lis_maxsim_pass_plan(int64_t n_mtiles, int32_t* passes, int cap) {
  LIS_REQUIRE(n_mtiles > 0 && (cap == 0 || passes), "lis_maxsim_pass_plan: bad arguments");
  const Tuning tn = tuning_snapshot();
  int nt, g;
  bool atm;
  choose_tiling(tn, n_mtiles, &nt, &g, &atm);
  std::vector<Pass> plan;
  build_pass_plan(tn, n_mtiles, false, g, plan);
  for (size_t i = 0; i < plan.size() && (int)i < cap; ++i) passes[i] = plan[i].pair ? -plan[i].n : plan[i].n;
  return (int)plan.size();
}

int lis_maxsim_scores(const void* q, int64_t q_rows, const int32_t* seg_lo, const int32_t* seg_hi,
                      const int32_t* mt_seg, int64_t n_seg, int64_t n_mtiles, const void* tokens,
                      int64_t n_rows, const int64_t* p_offsets, const uint8_t* p_clamp, int64_t np,
                      int dtype, int round_mode, float* out, int64_t ld_out, void* stream) {
  return maxsim_impl(q, nullptr, q_rows, seg_lo, seg_hi, mt_seg, n_seg, n_mtiles, tokens, nullptr, n_rows, p_offsets,
                     p_clamp, np, dtype, round_mode, out, ld_out, stream);
}

int lis_maxsim_scores_f32x2(const void* q_hi, const void* q_lo, int64_t q_rows, const int32_t* seg_lo,
                            const int32_t* seg_hi, const int32_t* mt_seg, int64_t n_seg, int64_t n_mtiles,
                            const void* tok_hi, const void* tok_lo, int64_t n_rows, const int64_t* p_offsets,
                            const uint8_t* p_clamp, int64_t np, float* out, int64_t ld_out, void* stream) {
  LIS_REQUIRE(q_lo, "lis_maxsim_scores_f32x2: null low plane");
  return maxsim_impl(q_hi, q_lo, q_rows, seg_lo, seg_hi, mt_seg, n_seg, n_mtiles, tok_hi, tok_lo, n_rows, p_offsets,
                     p_clamp, np, LIS_BF16, LIS_ROUND_F32, out, ld_out, stream);
}

namespace lis {
// x = hi + lo + O(2^-18 |x|):  hi = bf16(x),  lo = bf16(x - hi)
__global__ void split_f32_kernel(const float4* __restrict__ src, int64_t n4, uint2* __restrict__ hi,
                                 uint2* __restrict__ lo) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 x = __ldg(src + i);
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x.x), h1 = __float2bfloat16_rn(x.y);
    const __nv_bfloat16 h2 = __float2bfloat16_rn(x.z), h3 = __float2bfloat16_rn(x.w);
    const __nv_bfloat162 ha = __halves2bfloat162(h0, h1), hb = __halves2bfloat162(h2, h3);
    const __nv_bfloat162 la = __floats2bfloat162_rn(x.x - __bfloat162float(h0), x.y - __bfloat162float(h1));
    const __nv_bfloat162 lb = __floats2bfloat162_rn(x.z - __bfloat162float(h2), x.w - __bfloat162float(h3));
    hi[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&ha), *reinterpret_cast<const uint32_t*>(&hb));
    lo[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&la), *reinterpret_cast<const uint32_t*>(&lb));
  }
}
}  // namespace lis

int lis_split_f32(const float* src, int64_t rows, void* hi, void* lo, void* stream) {
  LIS_REQUIRE(rows >= 0, "negative row count");
  if (rows == 0) return LIS_OK;
  LIS_REQUIRE(src && hi && lo, "lis_split_f32: null pointer");
  LIS_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 15) == 0,
              "lis_split_f32: pointers must be 16-byte aligned");
  const int64_t n4 = rows * (kDim / 4);
  const int sms = current_device_sm_count();
  const unsigned grid = (unsigned)std::min<int64_t>((n4 + 255) / 256, (int64_t)std::max(sms, 1) * 32);
  split_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(src), n4,
                                                          static_cast<uint2*>(hi), static_cast<uint2*>(lo));
  count_launch();
  LIS_CUDA_CHECK(cudaGetLastError());
  return LIS_OK;
}

namespace {
__global__ void fill_iota_offsets(int64_t* off, int32_t* seg_lo, int32_t* seg_hi, int32_t* mt_seg, int64_t rows) {
  // one page covering all rows; one segment covering M tile 0
  off[0] = 0; off[1] = rows;
  seg_lo[0] = 0; seg_hi[0] = 128;
  mt_seg[0] = 0; mt_seg[1] = 1;
}
}  // namespace

int lis_debug_sim_tile(const void* q, int64_t q_rows, const void* tokens, int64_t n_rows, int dtype,
                       int tile_n, int a_in_tmem, float* out, void* stream) {
  LIS_REQUIRE(q && tokens && out, "lis_debug_sim_tile: null pointer");
  LIS_REQUIRE(a_in_tmem ? (tile_n == 128 || tile_n == 192) : (tile_n == 128 || tile_n == 256),
              "tile_n must be 128/256 (shared-memory A) or 128/192 (tensor-memory A)");
  LIS_REQUIRE(q_rows > 0 && n_rows > 0, "empty input");
  cudaStream_t st = (cudaStream_t)stream;
  Maps m;
  int rc = encode_rows_tmap(&m.q, q, q_rows, kMTile, dtype);
  if (rc) return rc;
  rc = encode_rows_tmap(&m.p, tokens, n_rows, tile_n, dtype);
  if (rc) return rc;
  m.q2 = m.q;
  m.p2 = m.p;
  // scratch tables + a dummy score
  char* scratch = nullptr;
  LIS_CUDA_CHECK(cudaMalloc(&scratch, 256));
  int64_t* off = (int64_t*)scratch;
  int32_t* seg_lo = (int32_t*)(scratch + 32);
  int32_t* seg_hi = (int32_t*)(scratch + 48);
  int32_t* mt_seg = (int32_t*)(scratch + 64);
  float* dummy = (float*)(scratch + 128);
  fill_iota_offsets<<<1, 1, 0, st>>>(off, seg_lo, seg_hi, mt_seg, std::min<int64_t>(n_rows, tile_n));
  count_launch();
  MaxSimArgs a;
  a.p_offsets = off; a.p_clamp = nullptr; a.seg_lo = seg_lo; a.seg_hi = seg_hi; a.mt_seg = mt_seg;
  a.out = dummy; a.dbg = out; a.ld_out = 1; a.np = 1; a.mt0 = 0; a.n_mt = 1; a.round_mode = 0;
  a.is_bf16 = dtype == LIS_BF16;
  a.ablate = 0;
  a.stats = nullptr;
  a.q = q;
  a.q_rows = q_rows;
  rc = dispatch_maxsim(tuning_snapshot(), tile_n, 1, a_in_tmem != 0, m, a, 1, st, true);
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(scratch);
  if (rc) return rc;
  LIS_CUDA_CHECK(e);
  return LIS_OK;
}


namespace {
__global__ void fill_pair_debug_tables(int64_t* off, int32_t* mt_seg, int64_t rows, int n_mt) {
  off[0] = 0; off[1] = rows;                       // one page covering all rows
  for (int t = 0; t <= n_mt; ++t) mt_seg[t] = 0;   // no segments: nothing is written but the dump
}
}  // namespace

int lis_debug_sim_pair(const void* q, int64_t q_rows, const void* tokens, int64_t n_rows, int dtype, int n_mt,
                       float* out, void* stream) {
  LIS_REQUIRE(q && tokens && out, "lis_debug_sim_pair: null pointer");
  LIS_REQUIRE(n_mt == 3 || n_mt == 4, "lis_debug_sim_pair: n_mt must be 3 or 4");
  LIS_REQUIRE(q_rows > (int64_t)(n_mt - 1) * kMTile && q_rows <= (int64_t)n_mt * kMTile && n_rows > 0,
              "lis_debug_sim_pair: q_rows must fill %d M tiles and the store must not be empty", n_mt);
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap mq, mp;
  int rc = encode_rows_tmap(&mq, q, q_rows, 64, dtype);
  if (rc) return rc;
  rc = encode_rows_tmap(&mp, tokens, n_rows, 64, dtype);
  if (rc) return rc;
  char* scratch = nullptr;
  LIS_CUDA_CHECK(cudaMalloc(&scratch, 256));
  int64_t* off = (int64_t*)scratch;
  int32_t* mt_seg = (int32_t*)(scratch + 64);
  float* dummy = (float*)(scratch + 128);
  fill_pair_debug_tables<<<1, 1, 0, st>>>(off, mt_seg, std::min<int64_t>(n_rows, 256), n_mt);
  count_launch();
  MaxSimArgs a;
  a.p_offsets = off; a.p_clamp = nullptr; a.seg_lo = mt_seg; a.seg_hi = mt_seg; a.mt_seg = mt_seg;
  a.out = dummy; a.dbg = out; a.ld_out = 1; a.np = 1; a.mt0 = 0; a.n_mt = n_mt; a.round_mode = 0;
  a.is_bf16 = dtype == LIS_BF16;
  a.ablate = 0;
  a.stats = nullptr;
  a.q = q;
  a.q_rows = q_rows;
  rc = dispatch_maxsim_pair(mq, mp, a, 2, st, true);
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(scratch);
  if (rc) return rc;
  LIS_CUDA_CHECK(e);
  return LIS_OK;
}

}  // extern "C"
