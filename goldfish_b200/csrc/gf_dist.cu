// Multi-GPU exchange steps of the patch-sharded path: NCCL over NVLink / NVSwitch, called from inside the
// library on the compute stream (no host callback between the kernels of a Krylov iteration).
//
// One process per GPU.  The communicator is created here from a unique id that the host
// broadcasts over its own process group (torch.distributed); NCCL is resolved at run time from the process
// image (torch has loaded libnccl.so.2 already) so the library has no link-time dependency on it and still
// loads on a CPU-only box for the ABI checks.
#include "gf_common.cuh"
#include <dlfcn.h>
#include <string.h>

namespace gf {
// the few NCCL declarations needed (nccl.h of NCCL 2.x; these have been ABI-stable since 2.0)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0, ncclInt32 = 2, ncclInt64 = 4, ncclFloat32 = 7, ncclFloat64 = 8 };
enum { ncclSum = 0, ncclMax = 2 };
struct Nccl {
  int (*GetUniqueId)(ncclUniqueId*);
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  int (*CommDestroy)(ncclComm_t);
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
  int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*GroupStart)();
  int (*GroupEnd)();
  const char* (*GetErrorString)(int);
  bool ok = false;
};
static Nccl g_nccl;

static int nccl_load() {
  if (g_nccl.ok) return GF_OK;
  void* h = RTLD_DEFAULT;
  if (!dlsym(h, "ncclAllReduce")) {
    h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return set_error(GF_ERR_BADARG, "gf_dist: NCCL (libnccl.so.2) is not loaded in this process and cannot be opened");
  }
#define GF_SYM(name) \
  *(void**)(&g_nccl.name) = dlsym(h, "nccl" #name); \
  if (!g_nccl.name) return set_error(GF_ERR_BADARG, "gf_dist: symbol nccl" #name " not found")
  GF_SYM(GetUniqueId); GF_SYM(CommInitRank); GF_SYM(CommDestroy); GF_SYM(AllReduce); GF_SYM(AllGather);
  GF_SYM(Broadcast); GF_SYM(GroupStart); GF_SYM(GroupEnd); GF_SYM(GetErrorString);
#undef GF_SYM
  g_nccl.ok = true;
  return GF_OK;
}
static int nccl_check(int rc, const char* where) {
  if (rc == ncclSuccess) return GF_OK;
  char buf[256];
  snprintf(buf, sizeof(buf), "%s: NCCL error %d (%s)", where, rc, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
  return set_error(GF_ERR_CUDA, buf);
}

int dist_allreduce(const GfDist* d, double* buf, int64_t n, cudaStream_t st) {
  if (!d || d->world <= 1 || n <= 0) return GF_OK;
  if (!d->comm) return set_error(GF_ERR_BADARG, "gf_dist: communicator not initialised (gf_dist_init)");
  return nccl_check(g_nccl.AllReduce(buf, buf, (size_t)n, ncclFloat64, ncclSum, (ncclComm_t)d->comm, st), "ncclAllReduce");
}
}  // namespace gf

using namespace gf;

extern "C" int gf_dist_unique_id(void* id128) {
  if (!id128) return set_error(GF_ERR_BADARG, "gf_dist_unique_id: null argument");
  int rc = nccl_load();
  if (rc) return rc;
  ncclUniqueId id;
  rc = nccl_check(g_nccl.GetUniqueId(&id), "ncclGetUniqueId");
  if (rc) return rc;
  memcpy(id128, &id, sizeof(id));
  return GF_OK;
}

extern "C" int gf_dist_init(GfDist* d, const void* id128, int rank, int world) {
  if (!d || !id128 || world < 1 || rank < 0 || rank >= world) return set_error(GF_ERR_BADARG, "gf_dist_init: bad argument");
  int rc = nccl_load();
  if (rc) return rc;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm = nullptr;
  rc = nccl_check(g_nccl.CommInitRank(&comm, world, id, rank), "ncclCommInitRank");
  if (rc) return rc;
  d->comm = comm; d->rank = rank; d->world = world;
  return GF_OK;
}

extern "C" int gf_dist_destroy(GfDist* d) {
  if (d && d->comm) { g_nccl.CommDestroy((ncclComm_t)d->comm); d->comm = nullptr; }
  return GF_OK;
}

extern "C" int gf_dist_allreduce(const GfDist* d, double* buf, int64_t n, void* stream) {
  return dist_allreduce(d, buf, n, (cudaStream_t)stream);
}
