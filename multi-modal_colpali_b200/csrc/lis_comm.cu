// Multi-GPU plumbing of the sharded search: one process per GPU, one NCCL communicator per process.
// The reference has no multi-GPU path (single device: functions.py:1472, 05_experiment02.py:338); pages are
// independent, so the only exchange step is ONE all-gather of the per-rank top-k candidates.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): a single-GPU user never needs it, and inside a torch
// process the already loaded library is reused, so both sides talk through the same NCCL build.
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "lis_common.h"

namespace lis {

// The slice of nccl.h this file needs (NCCL 2.x ABI: ncclUniqueId is 128 opaque bytes passed by value).
struct NcclUniqueId { char internal[128]; };
typedef struct ncclComm* ncclComm_t;
enum { kNcclInt8 = 0 };   // ncclInt8 / ncclChar
typedef int (*PFN_ncclGetUniqueId)(NcclUniqueId*);
typedef int (*PFN_ncclCommInitRank)(ncclComm_t*, int, NcclUniqueId, int);
typedef int (*PFN_ncclCommDestroy)(ncclComm_t);
typedef int (*PFN_ncclAllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
typedef const char* (*PFN_ncclGetErrorString)(int);
typedef int (*PFN_ncclGetVersion)(int*);

struct NcclApi {
  void* handle = nullptr;
  PFN_ncclGetUniqueId GetUniqueId = nullptr;
  PFN_ncclCommInitRank CommInitRank = nullptr;
  PFN_ncclCommDestroy CommDestroy = nullptr;
  PFN_ncclAllGather AllGather = nullptr;
  PFN_ncclGetErrorString GetErrorString = nullptr;
  PFN_ncclGetVersion GetVersion = nullptr;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    api.GetUniqueId = (PFN_ncclGetUniqueId)dlsym(h, "ncclGetUniqueId");
    api.CommInitRank = (PFN_ncclCommInitRank)dlsym(h, "ncclCommInitRank");
    api.CommDestroy = (PFN_ncclCommDestroy)dlsym(h, "ncclCommDestroy");
    api.AllGather = (PFN_ncclAllGather)dlsym(h, "ncclAllGather");
    api.GetErrorString = (PFN_ncclGetErrorString)dlsym(h, "ncclGetErrorString");
    api.GetVersion = (PFN_ncclGetVersion)dlsym(h, "ncclGetVersion");
    if (api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather) api.handle = h;
  });
  return api.handle ? &api : nullptr;
}

static int nccl_fail(const char* what, int rc) {
  NcclApi* n = nccl_api();
  set_error("%s failed: %s", what, (n && n->GetErrorString) ? n->GetErrorString(rc) : "NCCL error");
  return LIS_E_NCCL;
}

int comm_all_gather(lis_comm* c, const void* send, void* recv, size_t bytes, cudaStream_t st);

}  // namespace lis

struct lis_comm {
  lis::ncclComm_t nccl = nullptr;
  int rank = 0, world = 1, device = 0;
};

namespace lis {
int comm_all_gather(lis_comm* c, const void* send, void* recv, size_t bytes, cudaStream_t st) {
  NcclApi* n = nccl_api();
  if (!n || !c || !c->nccl) {
    set_error("all-gather without an initialised communicator");
    return LIS_E_NCCL;
  }
  const int rc = n->AllGather(send, recv, bytes, kNcclInt8, c->nccl, st);
  if (rc != 0) return nccl_fail("ncclAllGather", rc);
  return LIS_OK;
}
}  // namespace lis

using namespace lis;

extern "C" {

int lis_comm_unique_id(void* out, int bytes) {
  LIS_REQUIRE(out && bytes >= LIS_COMM_ID_BYTES, "lis_comm_unique_id: need a %d-byte buffer", LIS_COMM_ID_BYTES);
  NcclApi* n = nccl_api();
  if (!n) {
    set_error("libnccl.so.2 could not be loaded: %s", dlerror() ? dlerror() : "symbols missing");
    return LIS_E_NCCL;
  }
  NcclUniqueId id;
  const int rc = n->GetUniqueId(&id);
  if (rc != 0) return nccl_fail("ncclGetUniqueId", rc);
  static_assert(sizeof(id) == LIS_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
  memcpy(out, &id, sizeof(id));
  return LIS_OK;
}

int lis_comm_init(lis_comm** out, const void* unique_id, int rank, int world, int device) {
  LIS_REQUIRE(out, "lis_comm_init: null out");
  *out = nullptr;
  LIS_REQUIRE(world >= 1 && rank >= 0 && rank < world, "lis_comm_init: rank %d of %d", rank, world);
  int rc = lis_device_supported(device);
  if (rc) return rc;
  LIS_REQUIRE(world == 1 || unique_id, "lis_comm_init: a communicator of %d ranks needs the shared unique id", world);
  lis_comm* c = new lis_comm();
  c->rank = rank;
  c->world = world;
  c->device = device;
  if (world > 1) {
    NcclApi* n = nccl_api();
    if (!n) {
      delete c;
      set_error("libnccl.so.2 could not be loaded; multi-GPU search needs NCCL");
      return LIS_E_NCCL;
    }
    if (cudaSetDevice(device) != cudaSuccess) {
      delete c;
      set_error("cudaSetDevice(%d) failed", device);
      return LIS_E_CUDA;
    }
    NcclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    const int nrc = n->CommInitRank(&c->nccl, world, id, rank);
    if (nrc != 0) {
      delete c;
      return nccl_fail("ncclCommInitRank", nrc);
    }
    // one throw-away collective: NCCL sets its channels up on first use, which must not happen inside a
    // stream capture later (lis_index_search_sharded captures its all-gather into a CUDA graph)
    uint8_t* tmp = nullptr;
    cudaError_t e = cudaMalloc((void**)&tmp, (size_t)256 * (world + 1));
    if (e == cudaSuccess) e = cudaMemset(tmp, 0, (size_t)256 * (world + 1));
    int grc = e == cudaSuccess ? n->AllGather(tmp, tmp + 256, 256, kNcclInt8, c->nccl, nullptr) : 0;
    if (e == cudaSuccess) e = cudaStreamSynchronize(nullptr);
    cudaFree(tmp);
    if (e != cudaSuccess || grc != 0) {
      n->CommDestroy(c->nccl);
      delete c;
      if (grc != 0) return nccl_fail("ncclAllGather (warm-up)", grc);
      set_error("lis_comm_init warm-up failed: %s", cudaGetErrorString(e));
      return LIS_E_CUDA;
    }
  }
  *out = c;
  return LIS_OK;
}

void lis_comm_destroy(lis_comm* c) {
  if (!c) return;
  index_release_comm(c);   // graphs that captured this communicator's all-gather must go first
  NcclApi* n = nccl_api();
  if (c->nccl && n) n->CommDestroy(c->nccl);
  delete c;
}

int lis_comm_rank(const lis_comm* c) { return c ? c->rank : 0; }
int lis_comm_world(const lis_comm* c) { return c ? c->world : 1; }

int lis_nccl_version(void) {
  NcclApi* n = nccl_api();
  int v = 0;
  if (!n || !n->GetVersion || n->GetVersion(&v) != 0) return 0;
  return v;
}

}  // extern "C"
