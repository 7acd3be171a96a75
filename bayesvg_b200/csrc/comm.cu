// comm.cu -- the one exchange step of the path: a sum all-reduce of the 2k+10 cross-gene sums
// (plus a handful of L-BFGS scalars) between the GPUs that hold different gene shards.
// Two transports:
//   peer (default) -- every rank maps every peer's exchange buffer (CUDA IPC over NVLink / NVSwitch) and
//       the all-reduce runs INSIDE the kernel that produced the sums (peer_allreduce_cta in
//       bvg_internal.cuh, called from k_gene's last CTA): no extra launch, one NVLink store per value.
//   nccl -- ncclAllReduce between two kernels (the baseline the fused variant is measured against).
// NCCL is resolved at run time with dlopen so the C-ABI library has no link-time dependency on it
// (an R session loads it the same way a torch process does).
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#define BVG_NO_POOL   // the exchange buffer is shared through CUDA IPC: a plain cudaMalloc block
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
  if (f->comm_mode) { bvg_set_error("a transport is already initialised for this fit"); return BVG_ERR_STATE; }
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
  f->comm_mode = 1;
  return BVG_OK;
}

// ---- peer-memory transport ------------------------------------------------------------------------
static size_t peer_buf_bytes(int world) {
  const size_t b = (size_t)2 * world * 2 * BVG_PEER_CAP * sizeof(uint2) + 256;  // cells + the sequence counter
  return (b + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);                // own 2 MB granule: one IPC handle = this buffer
}

extern "C" int bvg_fit_peer_handle(bvg_fit *f, void *out, size_t cap, size_t *len) {
  if (!f || !out || cap < sizeof(cudaIpcMemHandle_t)) { bvg_set_error("peer handle buffer must hold %zu bytes", sizeof(cudaIpcMemHandle_t)); return BVG_ERR_ARG; }
  if (f->cfg.world_size > BVG_MAX_PEERS) { bvg_set_error("peer exchange supports up to %d GPUs of one node", BVG_MAX_PEERS); return BVG_ERR_ARG; }
  CUDA_TRY(cudaSetDevice(f->cfg.device));
  if (!f->peer_buf) {
    CUDA_TRY(cudaMalloc(&f->peer_buf, peer_buf_bytes(f->cfg.world_size)));
    CUDA_TRY(cudaMemset(f->peer_buf, 0, peer_buf_bytes(f->cfg.world_size)));
    CUDA_TRY(cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, f->peer_buf));
  memcpy(out, &h, sizeof(h));
  if (len) *len = sizeof(h);
  return BVG_OK;
}

// handles: world_size handles of bvg_fit_peer_handle, in rank order (the host all-gathers them)
extern "C" int bvg_fit_peer_init(bvg_fit *f, const void *handles, size_t len) {
  if (!f || !handles) { bvg_set_error("null argument"); return BVG_ERR_ARG; }
  const int W = f->cfg.world_size;
  if (W <= 1) return BVG_OK;
  if (W > BVG_MAX_PEERS || len != (size_t)W * sizeof(cudaIpcMemHandle_t)) { bvg_set_error("expected %d handles of %zu bytes", W, sizeof(cudaIpcMemHandle_t)); return BVG_ERR_ARG; }
  if (!f->peer_buf) { bvg_set_error("call bvg_fit_peer_handle first"); return BVG_ERR_STATE; }
  if (f->comm_mode) { bvg_set_error("a transport is already initialised for this fit"); return BVG_ERR_STATE; }
  CUDA_TRY(cudaSetDevice(f->cfg.device));
  PeerCtx &p = f->m.peer;
  memset(&p, 0, sizeof(p));
  for (int r = 0; r < W; ++r) {
    if (r == f->cfg.rank) { p.cells[r] = (uint2 *)f->peer_buf; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char *)handles + (size_t)r * sizeof(h), sizeof(h));
    const cudaError_t e = cudaIpcOpenMemHandle(&f->peer_map[r], h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      bvg_set_error("cudaIpcOpenMemHandle(rank %d): %s (the GPUs of one fit must be NVLink/P2P peers in one node)", r, cudaGetErrorString(e));
      cudaGetLastError();
      return BVG_ERR_COMM;
    }
    p.cells[r] = (uint2 *)f->peer_map[r];
  }
  p.seq = (uint32_t *)((char *)f->peer_buf + (size_t)2 * W * 2 * BVG_PEER_CAP * sizeof(uint2));
  p.world = W;
  p.rank = f->cfg.rank;
  {  // how long a poll waits for a peer: legitimate host-side skew (module load, cudaMalloc, a paused rank) can be seconds
    double sec = 120.0;
    if (const char *e = getenv("BVG_PEER_TIMEOUT_S")) { const double v = atof(e); if (v > 0) sec = v; }
    int khz = 2000000;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, f->cfg.device);
    p.timeout = (long long)(sec * 1e3 * (double)khz);
  }
  p.fail = f->m.ctl ? &f->m.ctl->comm_fail : nullptr;
  f->comm_mode = 2;
  return BVG_OK;
}

// stand-alone form of the exchange for the host-driven reductions (L-BFGS dot products, entropy term)
__global__ void __launch_bounds__(1024) k_peer_allreduce(PeerCtx p, double *buf, int n) {
  extern __shared__ double scratch[];  // (world - 1) * n doubles
  peer_allreduce_cta(p, buf, n, scratch);
}

int comm_allreduce(bvg_fit *f, double *buf, int n) {
  if (f->comm_mode == 2) {
    const int step = 512;  // (world - 1) * 512 doubles of polling scratch stay under the 48 KB default
    for (int o = 0; o < n; o += step) {
      const int nn = n - o < step ? n - o : step;
      k_peer_allreduce<<<1, 1024, sizeof(double) * (f->m.peer.world - 1) * nn, f->stream>>>(f->m.peer, buf + o, nn);
      f->st.kernel_launches++;
    }
    CUDA_TRY(cudaGetLastError());
    return BVG_OK;
  }
  if (!f->comm) return BVG_OK;
  const int rc = f->nccl->AllReduce(buf, buf, (size_t)n, ncclFloat64, ncclSum, (ncclComm_t)f->comm, f->stream);
  if (rc != ncclSuccess) { bvg_set_error("ncclAllReduce: %s", f->nccl->GetErrorString ? f->nccl->GetErrorString(rc) : "error"); return BVG_ERR_COMM; }
  return BVG_OK;
}

void comm_destroy(bvg_fit *f) {
  if (f->comm) f->nccl->CommDestroy((ncclComm_t)f->comm);
  f->comm = nullptr;
  for (int r = 0; r < BVG_MAX_PEERS; ++r)
    if (f->peer_map[r]) { cudaIpcCloseMemHandle(f->peer_map[r]); f->peer_map[r] = nullptr; }
  if (f->peer_buf) { cudaFree(f->peer_buf); f->peer_buf = nullptr; }
  f->m.peer.world = 0;
  f->comm_mode = 0;
}
