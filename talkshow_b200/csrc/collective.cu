// talkshow_b200 — the ONE collective of the path inside the library: an NCCL all-gather of the pose tensor over
// NVLink / NVSwitch (SURVEY.md §8b `ts_allgather`, §8e).  The reference has no multi-GPU inference path; its batch /
// diversity-sample loop (scripts/demo.py:195-204) is what gets sharded.
//
// NCCL is resolved at run time with dlopen (the library torch ships: nvidia/nccl/lib/libnccl.so.2; the host shim passes
// the path), so libtalkshow_b200.so keeps libcudart as its only link-time dependency.  Communicator bootstrap: rank 0
// calls ts_nccl_unique_id, the 128-byte id travels to the other ranks through whatever the host already has
// (torch.distributed's store / a gloo broadcast), every rank calls ts_nccl_init.
#include "pixelcnn.h"

#include <dlfcn.h>

#include <memory>

namespace ts {

struct NcclId { char internal[128]; };   // ncclUniqueId

struct NcclApi {
  void* lib = nullptr;
  void* comm = nullptr;
  int rank = 0, world = 1;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;   // ncclUniqueId is passed by value
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

static NcclApi* nccl_open(ts_engine* e, const char* path) {
  if (e->nccl) return (NcclApi*)e->nccl;
  std::unique_ptr<NcclApi> a(new NcclApi());
  a->lib = dlopen(path && path[0] ? path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!a->lib) fail(TS_ERR_UNSUPPORTED, "NCCL: dlopen(%s) failed: %s", path ? path : "libnccl.so.2", dlerror());
  auto sym = [&](const char* n) {
    void* p = dlsym(a->lib, n);
    if (!p) fail(TS_ERR_UNSUPPORTED, "NCCL: symbol %s missing", n);
    return p;
  };
  a->GetUniqueId = (int (*)(void*))sym("ncclGetUniqueId");
  a->CommInitRank = (int (*)(void**, int, NcclId, int))sym("ncclCommInitRank");
  a->AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))sym("ncclAllGather");
  a->CommDestroy = (int (*)(void*))sym("ncclCommDestroy");
  a->GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
  e->nccl = a.release();
  return (NcclApi*)e->nccl;
}

void nccl_destroy(ts_engine* e) {
  NcclApi* a = (NcclApi*)e->nccl;
  if (!a) return;
  if (a->comm && a->CommDestroy) a->CommDestroy(a->comm);
  delete a;
  e->nccl = nullptr;
}

}  // namespace ts

using namespace ts;

// rank 0: fill id[128] (host memory)
extern "C" int ts_nccl_unique_id(ts_engine* e, const char* libnccl_path, void* id128) {
  TS_API_BEGIN(e)
  if (!id128) fail(TS_ERR_INVALID, "ts_nccl_unique_id: null id buffer");
  NcclApi* a = nccl_open(e, libnccl_path);
  const int rc = a->GetUniqueId(id128);
  if (rc) fail(TS_ERR_CUDA, "ncclGetUniqueId: %s", a->GetErrorString(rc));
  TS_API_END(e)
}

// every rank, with the same id: creates this engine's communicator on the engine's device
extern "C" int ts_nccl_init(ts_engine* e, const char* libnccl_path, const void* id128, int rank, int world) {
  TS_API_BEGIN(e)
  if (e->host_only) fail(TS_ERR_UNSUPPORTED, "host-only engine has no device");
  if (!id128 || world < 1 || rank < 0 || rank >= world) fail(TS_ERR_INVALID, "ts_nccl_init: rank %d of %d", rank, world);
  NcclApi* a = nccl_open(e, libnccl_path);
  if (a->comm) { a->CommDestroy(a->comm); a->comm = nullptr; }
  NcclId id;
  memcpy(&id, id128, sizeof id);
  const int rc = a->CommInitRank(&a->comm, world, id, rank);
  if (rc) fail(TS_ERR_CUDA, "ncclCommInitRank: %s", a->GetErrorString(rc));
  a->rank = rank; a->world = world;
  TS_API_END(e)
}

// out[world * count] = concatenation over ranks of in[count] (fp32, device pointers), enqueued on `stream`.
// This is the single collective of the generation path: the [b,F,265] pose shards -> [B,F,265] on every rank.
extern "C" int ts_allgather(ts_engine* e, const float* in, float* out, int64_t count, void* stream) {
  TS_API_BEGIN(e)
  NcclApi* a = (NcclApi*)e->nccl;
  if (!a || !a->comm) fail(TS_ERR_NOT_LOADED, "ts_allgather: communicator not initialised (ts_nccl_init)");
  if (count < 0 || !in || !out) fail(TS_ERR_INVALID, "ts_allgather: bad arguments");
  const int rc = a->AllGather(in, out, (size_t)count, /* ncclFloat32 */ 7, a->comm, (cudaStream_t)stream);
  if (rc) fail(TS_ERR_CUDA, "ncclAllGather: %s", a->GetErrorString(rc));
  e->launches++;
  TS_API_END(e)
}
