// comm.cuh — run-time binding of NCCL and the communicator object (see comm.cu).
#pragma once
#include <nccl.h>   // types and enums only: the library is bound with dlopen, not linked

#include "mm2_internal.cuh"

struct NcclApi {
  bool ok = false;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
};
const NcclApi* nccl_api();   // NULL (and mm2_last_error set) when libnccl cannot be loaded

struct mm2_comm {
  mm2_ctx* ctx = nullptr;
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  u64* d_small = nullptr;   // 64 u64 of device scratch for tiny collectives
};

#define NCCL_TRY(N, expr)                                                                      \
  do {                                                                                          \
    ncclResult_t r__ = (expr);                                                                  \
    if (r__ != ncclSuccess) {                                                                   \
      mm2_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, (N)->GetErrorString(r__));    \
      return MM2_E_CUDA;                                                                        \
    }                                                                                           \
  } while (0)
