// Data parallelism inside the C ABI: an NCCL communicator owned by the library and the
// sharded CD step   STATS on this rank's rows -> all-reduce of the packed statistics -> identical APPLY on every rank.
//
// The packed buffer is [v0^T ph - nv^T nh (V*H) | sum(ph-nh) (H) | sum(v0-nv) (V) | cost numerator | rows]
// (src/rbm.py:411-417 are sums over minibatch rows, so the shards add).  It is reduced in two chunks on a side stream:
// the V*H block as soon as the statistics GEMM has produced it (the bias / cost / pseudo-likelihood kernels of the tail
// keep running on the caller's stream meanwhile), then the small tail; no second collective for the cost.
//
// NCCL is the one torch ships (nvidia/nccl): resolved with dlopen at run time, no link-time dependency, so a build
// without NCCL still loads and only mdbn_comm_* fail.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>
#include "ctx.h"

struct mdbn_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
  cudaStream_t stream = nullptr;          // side stream of the collectives
  cudaEvent_t ev_w = nullptr, ev_tail = nullptr, ev_done = nullptr;
};

namespace mdbn {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  const char* (*GetErrorString)(ncclResult_t);
  bool ok = false;
};

static NcclApi* nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.ok ? &api : nullptr;
  tried = true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);      // already in the process when torch is
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    set_error("NCCL not found: %s", dlerror());
    return nullptr;
  }
  api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
  api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
  api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GetErrorString;
  if (!api.ok) set_error("NCCL symbols missing in libnccl");
  return api.ok ? &api : nullptr;
}

#define MDBN_NCCL(x)                                                                           \
  do {                                                                                         \
    ncclResult_t r_ = (x);                                                                     \
    if (r_ != ncclSuccess) {                                                                   \
      mdbn::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #x, n->GetErrorString(r_));        \
      return 4;                                                                                \
    }                                                                                          \
  } while (0)

// The sharded step.  `a.B` / `a.indices` are THIS rank's rows, `a.B_total` the rows of the whole minibatch.
int comm_cd_step(mdbn_ctx* c, const mdbn_cd_args& a, cudaStream_t st) {
  NcclApi* n = nccl();
  if (!n) return 4;
  mdbn_comm* cm = a.comm;
  MDBN_CHECK(a.B_total >= a.B, "cd_step with comm: B_total (rows of the whole minibatch) must be given");
  const long long VH = (long long)a.V * a.H, tail = (long long)a.H + a.V + 2;
  float* G = (float*)ws_get(c, WS_G, (size_t)(VH + tail) * sizeof(float));
  if (!G) return 3;
  mdbn_cd_args s = a;
  s.comm = nullptr;
  s.phase = MDBN_PHASE_STATS;
  s.stats_buf = G;
  if (s.path == MDBN_PATH_SKINNY || s.path == MDBN_PATH_TINY) s.path = MDBN_PATH_AUTO;   // whole-step kernels: no cut
  c->ev_stats_w = cm->ev_w;          // recorded by the statistics path once the V*H block is complete
  c->ev_stats_w_done = false;
  int rc = mdbn_cd_step(c, &s, (void*)st);
  c->ev_stats_w = nullptr;
  if (rc) return rc;
  if (!c->ev_stats_w_done) MDBN_CUDA(cudaEventRecord(cm->ev_w, st));
  // chunk 1: the V*H block, overlapping the tail kernels still queued on `st`
  MDBN_CUDA(cudaStreamWaitEvent(cm->stream, cm->ev_w, 0));
  MDBN_NCCL(n->AllReduce(G, G, (size_t)VH, ncclFloat, ncclSum, cm->comm, cm->stream));
  // chunk 2: bias sums, cost numerator and row count
  MDBN_CUDA(cudaEventRecord(cm->ev_tail, st));
  MDBN_CUDA(cudaStreamWaitEvent(cm->stream, cm->ev_tail, 0));
  MDBN_NCCL(n->AllReduce(G + VH, G + VH, (size_t)tail, ncclFloat, ncclSum, cm->comm, cm->stream));
  MDBN_CUDA(cudaEventRecord(cm->ev_done, cm->stream));
  MDBN_CUDA(cudaStreamWaitEvent(st, cm->ev_done, 0));
  return apply_update(c, a, G, a.B_total, st);
}

}  // namespace mdbn

using namespace mdbn;

extern "C" {

int mdbn_comm_unique_id(unsigned char* id_out) {
  NcclApi* n = nccl();
  if (!n) return 4;
  MDBN_CHECK(id_out != nullptr, "comm_unique_id: id_out is NULL");
  static_assert(sizeof(ncclUniqueId) == MDBN_COMM_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  MDBN_NCCL(n->GetUniqueId(&id));
  memcpy(id_out, &id, sizeof(id));
  return 0;
}

int mdbn_comm_init(mdbn_comm** out, const unsigned char* id, int rank, int world, int device) {
  NcclApi* n = nccl();
  if (!n) return 4;
  MDBN_CHECK(out && id, "comm_init: NULL argument");
  MDBN_CHECK(world >= 1 && rank >= 0 && rank < world, "comm_init: bad rank %d / world %d", rank, world);
  MDBN_CUDA(cudaSetDevice(device));
  mdbn_comm* cm = new mdbn_comm();
  cm->rank = rank; cm->world = world; cm->device = device;
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  ncclResult_t r = n->CommInitRank(&cm->comm, world, uid, rank);
  if (r != ncclSuccess) {
    set_error("ncclCommInitRank(rank %d of %d) -> %s", rank, world, n->GetErrorString(r));
    delete cm;
    return 4;
  }
  MDBN_CUDA(cudaStreamCreateWithFlags(&cm->stream, cudaStreamNonBlocking));
  MDBN_CUDA(cudaEventCreateWithFlags(&cm->ev_w, cudaEventDisableTiming));
  MDBN_CUDA(cudaEventCreateWithFlags(&cm->ev_tail, cudaEventDisableTiming));
  MDBN_CUDA(cudaEventCreateWithFlags(&cm->ev_done, cudaEventDisableTiming));
  *out = cm;
  return 0;
}

int mdbn_comm_destroy(mdbn_comm* cm) {
  if (!cm) return 0;
  NcclApi* n = nccl();
  cudaSetDevice(cm->device);
  if (cm->stream) cudaStreamSynchronize(cm->stream);
  if (n && cm->comm) n->CommDestroy(cm->comm);
  if (cm->ev_w) cudaEventDestroy(cm->ev_w);
  if (cm->ev_tail) cudaEventDestroy(cm->ev_tail);
  if (cm->ev_done) cudaEventDestroy(cm->ev_done);
  if (cm->stream) cudaStreamDestroy(cm->stream);
  delete cm;
  return 0;
}

int mdbn_comm_all_reduce(mdbn_comm* cm, float* buf, unsigned long long count, void* stream) {
  NcclApi* n = nccl();
  if (!n) return 4;
  MDBN_CHECK(cm && buf, "comm_all_reduce: NULL argument");
  MDBN_NCCL(n->AllReduce(buf, buf, (size_t)count, ncclFloat, ncclSum, cm->comm, (cudaStream_t)stream));
  return 0;
}

}  // extern "C"
