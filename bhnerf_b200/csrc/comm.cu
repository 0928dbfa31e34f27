// Gradient exchange below the C ABI: jax.lax.pmean(grads, 'batch') (bhnerf/network.py:620, :680) as ONE NCCL all-reduce
// on the kernel stream (220 KB: latency-bound, no fused compute+collective kernel is warranted), and the all-reduce SUM
// that ray sharding needs for partial lightcurves / visibilities.  NCCL is resolved at run time with
// dlopen -- the instance torch already loaded when there is one -- so the library itself links against nothing but the
// CUDA runtime and still loads (and exports every symbol) on a machine without NCCL.
#include <dlfcn.h>
#include <nccl.h>
#include <mutex>
#include "common.cuh"

namespace {
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

int load_nccl() {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.handle) return 0;
  const char* names[] = {getenv("BHNERF_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    if (!n) continue;
    h = dlopen(n, RTLD_NOW | RTLD_NOLOAD);        // the copy the host process (torch) already mapped, if any
    if (!h) h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  BH_REQUIRE(h, "NCCL not found (dlopen libnccl.so.2 failed: %s); set BHNERF_NCCL_LIB", dlerror());
#define BH_NCCL_SYM(field, name) \
  *(void**)(&g_nccl.field) = dlsym(h, name); BH_REQUIRE(g_nccl.field, "NCCL symbol %s missing", name);
  BH_NCCL_SYM(GetUniqueId, "ncclGetUniqueId") BH_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  BH_NCCL_SYM(CommDestroy, "ncclCommDestroy") BH_NCCL_SYM(AllReduce, "ncclAllReduce")
  BH_NCCL_SYM(GetErrorString, "ncclGetErrorString") BH_NCCL_SYM(GetVersion, "ncclGetVersion")
  g_nccl.handle = h;
  return 0;
}
#define BH_CHECK_NCCL(expr)                                                                                  \
  do {                                                                                                       \
    ncclResult_t _r = (expr);                                                                                \
    if (_r != ncclSuccess) { bh_set_error("%s:%d NCCL error %s: %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r)); return 3; } \
  } while (0)
}  // namespace

extern "C" int bhnerf_comm_unique_id(void* id_host) {
  BH_REQUIRE(id_host, "comm_unique_id: NULL argument");
  if (int r = load_nccl()) return r;
  static_assert(sizeof(ncclUniqueId) == BHNERF_COMM_ID_BYTES, "ncclUniqueId size");
  BH_CHECK_NCCL(g_nccl.GetUniqueId((ncclUniqueId*)id_host));
  return 0;
}

extern "C" int bhnerf_comm_init(int32_t rank, int32_t world, const void* id_host, void** comm_out) {
  BH_REQUIRE(id_host && comm_out && world >= 1 && rank >= 0 && rank < world, "comm_init: bad argument");
  if (int r = load_nccl()) return r;
  ncclUniqueId id;
  memcpy(&id, id_host, sizeof(id));
  ncclComm_t c = nullptr;
  BH_CHECK_NCCL(g_nccl.CommInitRank(&c, world, id, rank));
  *comm_out = (void*)c;
  return 0;
}

extern "C" int bhnerf_comm_destroy(void* comm) {
  if (!comm) return 0;
  if (int r = load_nccl()) return r;
  BH_CHECK_NCCL(g_nccl.CommDestroy((ncclComm_t)comm));
  return 0;
}

extern "C" int bhnerf_allreduce_mean(float* buf, int64_t n, void* comm, void* stream) {
  BH_REQUIRE(buf && comm && n > 0, "allreduce_mean: bad argument");
  if (int r = load_nccl()) return r;
  BhProfScope ps(BH_CAT_COMM, 1, (cudaStream_t)stream);
  BH_CHECK_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, ncclFloat32, ncclAvg, (ncclComm_t)comm, (cudaStream_t)stream));
  return 0;
}

extern "C" int bhnerf_allreduce_sum(float* buf, int64_t n, void* comm, void* stream) {
  BH_REQUIRE(buf && comm && n > 0, "allreduce_sum: bad argument");
  if (int r = load_nccl()) return r;
  BhProfScope ps(BH_CAT_COMM, 1, (cudaStream_t)stream);
  BH_CHECK_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, ncclFloat32, ncclSum, (ncclComm_t)comm, (cudaStream_t)stream));
  return 0;
}
