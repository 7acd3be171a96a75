// comm.cu -- the one exchange step of the path: a sum all-reduce of the 2k+10 cross-gene sums
// (plus a handful of L-BFGS scalars) between the GPUs that hold different gene shards.
// NCCL is resolved at run time with dlopen so the C-ABI library has no link-time dependency on it
// (an R session loads it the same way a torch process does).
#include <dlfcn.h>
#include <string.h>

#include "bvg_internal.cuh"

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat64 = 8 };  // ncclDataType_t: int8 0, uint8 1, int32 2, uint32 3, int64 4, uint64 5, f16 6, f32 7, f64 8
enum { ncclSum = 0 };

struct NcclApi {
  void *h;
  int (*GetUniqueId)(ncclUniqueId *);
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*CommDestroy)(ncclComm_t);
  const char *(*GetErrorString)(int);
};

static NcclApi *load_nccl() {
  static NcclApi api;
  static bool tried = false, ok = false;
  if (tried) return ok ? &api : nullptr;
  tried = true;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.h) break;
  }
  if (!api.h) { bvg_set_error("cannot dlopen libnccl.so.2: %s", dlerror()); return nullptr; }
  api.GetUniqueId = (int (*)(ncclUniqueId *))dlsym(api.h, "ncclGetUniqueId");
  api.CommInitRank = (int (*)(ncclComm_t *, int, ncclUniqueId, int))dlsym(api.h, "ncclCommInitRank");
  api.AllReduce = (int (*)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(api.h, "ncclAllReduce");
  api.CommDestroy = (int (*)(ncclComm_t))dlsym(api.h, "ncclCommDestroy");
  api.GetErrorString = (const char *(*)(int))dlsym(api.h, "ncclGetErrorString");
  if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) {
    bvg_set_error("libnccl is missing a required symbol");
    return nullptr;
  }
  ok = true;
  return &api;
}

extern "C" int bvg_comm_unique_id(void *out, size_t cap, size_t *len) {
  if (!out || cap < sizeof(ncclUniqueId)) { bvg_set_error("unique id buffer must hold %zu bytes", sizeof(ncclUniqueId)); return BVG_ERR_ARG; }
  NcclApi *api = load_nccl();
  if (!api) return BVG_ERR_COMM;
  ncclUniqueId id;
  const int rc = api->GetUniqueId(&id);
  if (rc != ncclSuccess) { bvg_set_error("ncclGetUniqueId: %s", api->GetErrorString ? api->GetErrorString(rc) : "error"); return BVG_ERR_COMM; }
  memcpy(out, &id, sizeof(id));
  if (len) *len = sizeof(id);
  return BVG_OK;
}

extern "C" int bvg_fit_comm_init(bvg_fit *f, const void *unique_id, size_t len) {
  if (!f || !unique_id || len != sizeof(ncclUniqueId)) { bvg_set_error("bad unique id"); return BVG_ERR_ARG; }
  if (f->cfg.world_size <= 1) return BVG_OK;
  NcclApi *api = load_nccl();
  if (!api) return BVG_ERR_COMM;
  CUDA_TRY(cudaSetDevice(f->cfg.device));
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  ncclComm_t comm;
  const int rc = api->CommInitRank(&comm, f->cfg.world_size, id, f->cfg.rank);
  if (rc != ncclSuccess) { bvg_set_error("ncclCommInitRank: %s", api->GetErrorString ? api->GetErrorString(rc) : "error"); return BVG_ERR_COMM; }
  f->nccl = api;
  f->comm = comm;
  return BVG_OK;
}

int comm_allreduce(bvg_fit *f, double *buf, int n) {
  if (!f->comm) return BVG_OK;
  const int rc = f->nccl->AllReduce(buf, buf, (size_t)n, ncclFloat64, ncclSum, (ncclComm_t)f->comm, f->stream);
  if (rc != ncclSuccess) { bvg_set_error("ncclAllReduce: %s", f->nccl->GetErrorString ? f->nccl->GetErrorString(rc) : "error"); return BVG_ERR_COMM; }
  return BVG_OK;
}

void comm_destroy(bvg_fit *f) {
  if (f->comm) f->nccl->CommDestroy((ncclComm_t)f->comm);
  f->comm = nullptr;
}
