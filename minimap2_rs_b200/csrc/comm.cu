// comm.cu — the NCCL communicator owned by libmm2b200.so ("mm2_ctx_t owns devices, streams, NCCL comms").
// NCCL is bound at run time with dlopen("libnccl.so.2"): inside a PyTorch process that resolves to the NCCL torch has
// already loaded (one copy per process), in the stand-alone mm2rs binary to the system library.  The single-GPU paths of
// the library never touch it.  One communicator rank per mm2_ctx (one process per GPU under torchrun, or one host thread
// per GPU inside mm2_index_build_multi).
#include "comm.cuh"

#include <dlfcn.h>

#include <mutex>

namespace {
NcclApi g_api;
std::once_flag g_once;
std::string g_load_err;

template <class F>
bool bind(void* h, const char* name, F& fn) {
  fn = reinterpret_cast<F>(dlsym(h, name));
  if (!fn) { g_load_err = std::string("libnccl: missing symbol ") + name; return false; }
  return true;
}
void load_nccl() {
  void* h = nullptr;
  for (const char* nm : {"libnccl.so.2", "libnccl.so"}) {
    h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) { g_load_err = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "?"); return; }
  bool ok = bind(h, "ncclGetUniqueId", g_api.GetUniqueId) && bind(h, "ncclCommInitRank", g_api.CommInitRank) &&
            bind(h, "ncclCommDestroy", g_api.CommDestroy) && bind(h, "ncclGetErrorString", g_api.GetErrorString) &&
            bind(h, "ncclAllReduce", g_api.AllReduce) && bind(h, "ncclAllGather", g_api.AllGather) &&
            bind(h, "ncclBroadcast", g_api.Broadcast) && bind(h, "ncclSend", g_api.Send) && bind(h, "ncclRecv", g_api.Recv) &&
            bind(h, "ncclGroupStart", g_api.GroupStart) && bind(h, "ncclGroupEnd", g_api.GroupEnd);
  g_api.ok = ok;
}
}  // namespace

const NcclApi* nccl_api() {
  std::call_once(g_once, load_nccl);
  if (!g_api.ok) { mm2_set_error("%s", g_load_err.c_str()); return nullptr; }
  return &g_api;
}

extern "C" int mm2_comm_get_unique_id(void* id128) {
  if (!id128) { mm2_set_error("mm2_comm_get_unique_id: NULL argument"); return MM2_E_ARG; }
  const NcclApi* N = nccl_api();
  if (!N) return MM2_E_UNSUPPORTED;
  ncclUniqueId id;
  NCCL_TRY(N, N->GetUniqueId(&id));
  memcpy(id128, &id, sizeof id);
  return MM2_OK;
}

extern "C" int mm2_comm_create(mm2_ctx_t* ctx, const void* id128, int nranks, int rank, mm2_comm_t** out) {
  if (!ctx || !id128 || !out || nranks < 1 || rank < 0 || rank >= nranks) { mm2_set_error("mm2_comm_create: bad argument"); return MM2_E_ARG; }
  const NcclApi* N = nccl_api();
  if (!N) return MM2_E_UNSUPPORTED;
  CUDA_TRY(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  mm2_comm* c = new mm2_comm();
  c->ctx = ctx; c->nranks = nranks; c->rank = rank;
  const ncclResult_t r = N->CommInitRank(&c->comm, nranks, id, rank);
  if (r != ncclSuccess) { mm2_set_error("ncclCommInitRank: %s", N->GetErrorString(r)); delete c; return MM2_E_CUDA; }
  if (cudaMalloc(&c->d_small, 64 * 8) != cudaSuccess) { N->CommDestroy(c->comm); delete c; mm2_set_error("cudaMalloc failed"); return MM2_E_OOM; }
  *out = c;
  return MM2_OK;
}

extern "C" void mm2_comm_destroy(mm2_comm_t* c) {
  if (!c) return;
  cudaSetDevice(c->ctx->device);
  cudaStreamSynchronize(c->ctx->stream);
  if (c->d_small) cudaFree(c->d_small);
  const NcclApi* N = nccl_api();
  if (N && c->comm) N->CommDestroy(c->comm);
  delete c;
}

extern "C" int mm2_comm_rank(const mm2_comm_t* c, int* rank, int* nranks) {
  if (!c) { mm2_set_error("NULL comm"); return MM2_E_ARG; }
  if (rank) *rank = c->rank;
  if (nranks) *nranks = c->nranks;
  return MM2_OK;
}

// all ranks have finished everything they enqueued on their context's stream before anyone returns
extern "C" int mm2_comm_barrier(mm2_comm_t* c) {
  if (!c) { mm2_set_error("NULL comm"); return MM2_E_ARG; }
  const NcclApi* N = nccl_api();
  if (!N) return MM2_E_UNSUPPORTED;
  CUDA_TRY(cudaSetDevice(c->ctx->device));
  CUDA_TRY(cudaMemsetAsync(c->d_small, 0, 8, c->ctx->stream));
  NCCL_TRY(N, N->AllReduce(c->d_small, c->d_small, 1, ncclUint64, ncclSum, c->comm, c->ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(c->ctx->stream));
  return MM2_OK;
}
