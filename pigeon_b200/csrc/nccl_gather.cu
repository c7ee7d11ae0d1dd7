// Multi-GPU exchange step of the path: the all-gather of per-rank embeddings / head outputs that the
// reference does with accelerate's `gather` (preprocessing/embed.py:36-37) — here on a caller-supplied NCCL communicator.
//
// NCCL is resolved at run time (dlopen by soname): the library keeps no link-time dependency on it, a single-GPU process never
// loads it, and inside a PyTorch process the copy torch already mapped is the one that answers.  The Python mirror of the
// path uses torch.distributed for the same exchange (pigeon_b200/dist.py); these entry points are the C-ABI seam for hosts
// that own their communicator.
#include "pigeon_b200.h"
#include "tma_host.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstring>
#include <mutex>

namespace {

struct NcclUniqueId { char internal[128]; };   // ncclUniqueId of nccl.h (NCCL_UNIQUE_ID_BYTES = 128)
using ncclComm_t = void*;

struct NcclApi {
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

NcclApi* nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(h, "ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.GetErrorString;
  });
  if (!api.ok) {
    pg::set_last_error("NCCL is not available: libnccl.so.2 could not be loaded or lacks a symbol");
    return nullptr;
  }
  return &api;
}

int fail(NcclApi* a, const char* what, int rc) {
  pg::set_last_error("%s: %s", what, a->GetErrorString(rc));
  return 1;
}

}  // namespace

extern "C" {

int pg_nccl_unique_id(void* id128) {
  NcclApi* a = nccl();
  if (!a) return 1;
  if (!id128) { pg::set_last_error("pg_nccl_unique_id: null argument"); return 1; }
  NcclUniqueId id;
  const int rc = a->GetUniqueId(&id);
  if (rc) return fail(a, "ncclGetUniqueId", rc);
  memcpy(id128, &id, sizeof(id));
  return 0;
}

int pg_nccl_comm_create(const void* id128, int32_t n_ranks, int32_t rank, void** comm) {
  NcclApi* a = nccl();
  if (!a) return 1;
  if (!id128 || !comm || n_ranks <= 0 || rank < 0 || rank >= n_ranks) {
    pg::set_last_error("pg_nccl_comm_create: bad argument");
    return 1;
  }
  NcclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t c = nullptr;
  const int rc = a->CommInitRank(&c, n_ranks, id, rank);
  if (rc) return fail(a, "ncclCommInitRank", rc);
  *comm = c;
  return 0;
}

void pg_nccl_comm_destroy(void* comm) {
  NcclApi* a = nccl();
  if (a && comm) a->CommDestroy(comm);
}

int pg_allgather_embeddings(void* comm, const void* send, void* recv, size_t bytes_per_rank, void* stream) {
  NcclApi* a = nccl();
  if (!a) return 1;
  if (!comm || !send || !recv) { pg::set_last_error("pg_allgather_embeddings: null argument"); return 1; }
  if (bytes_per_rank == 0) return 0;
  const int rc = a->AllGather(send, recv, bytes_per_rank, /* ncclInt8 */ 0, comm, reinterpret_cast<cudaStream_t>(stream));
  if (rc) return fail(a, "ncclAllGather", rc);
  return 0;
}

}  // extern "C"
