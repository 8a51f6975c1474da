// vvae_comm_*: the gradient exchange of the data-parallel step behind the C ABI.
//
// The reference gets its gradient all-reduce from XLA SPMD inside the jitted step
// (claude_distributed/distributed_train.py:107-109,378-380,412) and its start-up replication from
// multihost_utils.broadcast_one_to_all (:339).  A host that is not PyTorch (the jax.ffi binding of INTEGRATION.md) needs
// the same two collectives on the flat gradient / parameter buffers without torch.distributed: these entry points are a
// thin, allocation-free layer over NCCL, which is resolved at run time with dlopen so that libvvae.so has no link-time
// dependency on it (a torch process already has libnccl.so.2 mapped; dlopen returns that copy).
//
// One communicator per process (= per GPU).  Rank 0 makes the 128-byte rendezvous token with vvae_comm_unique_id and the
// host carries it to the other ranks (environment, file, MPI, a torch store -- not this library's business).
#include <dlfcn.h>
#include <string.h>

#include <mutex>

#include "common.cuh"

namespace vvae {
namespace {

// the slice of NCCL's C API this file uses (nccl.h: ncclUniqueId is 128 opaque bytes, passed by value)
struct NcclId { char internal[128]; };
using nccl_comm = void*;
enum { kNcclSum = 0, kNcclAvg = 4, kNcclFloat32 = 7, kNcclBfloat16 = 9 };

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(nccl_comm*, int, NcclId, int) = nullptr;
  int (*CommDestroy)(nccl_comm) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  const char* why = nullptr;   // non-null: loading failed
};

const NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) {
      api.why = "libnccl.so.2 not found (dlopen)";
      return;
    }
    auto sym = [](const char* n) { return dlsym(api.handle, n); };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce || !api.Broadcast || !api.GetErrorString)
      api.why = "libnccl.so.2 lacks a required symbol";
  });
  return api;
}

int nccl_status(const NcclApi& n, int rc, const char* what) {
  if (rc == 0) return VVAE_OK;
  set_error("%s: NCCL error %d (%s)", what, rc, n.GetErrorString ? n.GetErrorString(rc) : "?");
  return VVAE_ERR_CUDA;
}

int nccl_dtype(int dtype) { return dtype == VVAE_F32 ? kNcclFloat32 : dtype == VVAE_BF16 ? kNcclBfloat16 : -1; }

constexpr unsigned kCommMagic = 0x76766165u;   // "vvae": a stale or foreign handle is rejected instead of dereferenced blindly

}  // namespace
}  // namespace vvae

struct vvae_comm {
  unsigned magic;
  vvae::nccl_comm comm;
  int rank, world, device;
};

using namespace vvae;

#define VVAE_COMM_PROLOGUE(what)                                         \
  const NcclApi& n = nccl();                                             \
  if (n.why) {                                                           \
    set_error(what ": %s", n.why);                                       \
    return VVAE_ERR_UNSUPPORTED;                                         \
  }

extern "C" {

// One front door for the three caller-provided scratch buffers of the library (declared next to vvae_comm_* in vvae.h).
long long vvae_workspace_bytes(int op, const long long* dims, int ndims) {
  auto positive = [&](int n, long long limit) {
    if (!dims || ndims != n) return false;
    for (int i = 0; i < n; ++i)
      if (dims[i] <= 0 || dims[i] > limit) return false;
    return true;
  };
  constexpr long long kInt = 0x7fffffffLL;
  switch (op) {
    case VVAE_WS_CONVT122:
      if (!positive(5, kInt)) return -1;
      return vvae_convT122_workspace_bytes((int)dims[0], (int)dims[1], (int)dims[2], (int)dims[3], (int)dims[4]);
    case VVAE_WS_SUMSQ_DET:
      if (!positive(1, 1LL << 40)) return -1;
      return 4LL * vvae_sumsq_partials(dims[0]);
    case VVAE_WS_ATTN_BWD:
      if (!positive(3, kInt) || dims[0] * dims[1] > (1LL << 40) / dims[2]) return -1;
      return 4LL * dims[0] * dims[1] * dims[2];
    default:
      return -1;
  }
}

int vvae_comm_unique_id(void* id128) {
  VVAE_REQUIRE(id128, "vvae_comm_unique_id: null pointer");
  VVAE_COMM_PROLOGUE("vvae_comm_unique_id");
  NcclId id;
  int rc = nccl_status(n, n.GetUniqueId(&id), "vvae_comm_unique_id");
  if (rc) return rc;
  memcpy(id128, id.internal, sizeof(id.internal));
  return VVAE_OK;
}

int vvae_comm_init(vvae_comm_t* comm, const void* id128, int rank, int world) {
  VVAE_REQUIRE(comm && id128, "vvae_comm_init: null pointer");
  *comm = nullptr;
  VVAE_REQUIRE(world >= 1 && rank >= 0 && rank < world, "vvae_comm_init: rank %d of %d", rank, world);
  if (!vvae_device_ok()) {
    set_error("vvae_comm_init: no sm_100 CUDA device");
    return VVAE_ERR_CUDA;
  }
  VVAE_COMM_PROLOGUE("vvae_comm_init");
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_error("vvae_comm_init: cudaGetDevice: %s", cudaGetErrorString(cudaGetLastError()));
    return VVAE_ERR_CUDA;
  }
  NcclId id;
  memcpy(id.internal, id128, sizeof(id.internal));
  nccl_comm c = nullptr;
  int rc = nccl_status(n, n.CommInitRank(&c, world, id, rank), "vvae_comm_init");   // collective: every rank calls it
  if (rc) return rc;
  *comm = new vvae_comm{kCommMagic, c, rank, world, dev};   // host-side handle only; no device memory is allocated here
  return VVAE_OK;
}

int vvae_comm_rank(vvae_comm_t comm, int* rank, int* world) {
  VVAE_REQUIRE(comm && comm->magic == kCommMagic, "vvae_comm_rank: not a communicator");
  if (rank) *rank = comm->rank;
  if (world) *world = comm->world;
  return VVAE_OK;
}

int vvae_comm_allreduce(vvae_comm_t comm, void* buf, long long count, int dtype, int average, vvae_stream_t stream) {
  VVAE_REQUIRE(comm && comm->magic == kCommMagic, "vvae_comm_allreduce: not a communicator");
  VVAE_REQUIRE(count >= 0 && (buf || count == 0), "vvae_comm_allreduce: null buffer");
  const int dt = nccl_dtype(dtype);
  VVAE_REQUIRE(dt >= 0, "vvae_comm_allreduce: unsupported dtype %d", dtype);
  if (count == 0) return VVAE_OK;
  VVAE_COMM_PROLOGUE("vvae_comm_allreduce");
  return nccl_status(n, n.AllReduce(buf, buf, (size_t)count, dt, average ? kNcclAvg : kNcclSum, comm->comm, as_stream(stream)),
                     "vvae_comm_allreduce");
}

int vvae_comm_broadcast(vvae_comm_t comm, void* buf, long long count, int dtype, int root, vvae_stream_t stream) {
  VVAE_REQUIRE(comm && comm->magic == kCommMagic, "vvae_comm_broadcast: not a communicator");
  VVAE_REQUIRE(count >= 0 && (buf || count == 0), "vvae_comm_broadcast: null buffer");
  VVAE_REQUIRE(root >= 0 && root < comm->world, "vvae_comm_broadcast: root %d of %d", root, comm->world);
  const int dt = nccl_dtype(dtype);
  VVAE_REQUIRE(dt >= 0, "vvae_comm_broadcast: unsupported dtype %d", dtype);
  if (count == 0) return VVAE_OK;
  VVAE_COMM_PROLOGUE("vvae_comm_broadcast");
  return nccl_status(n, n.Broadcast(buf, buf, (size_t)count, dt, root, comm->comm, as_stream(stream)), "vvae_comm_broadcast");
}

int vvae_comm_destroy(vvae_comm_t comm) {
  if (!comm) return VVAE_OK;
  VVAE_REQUIRE(comm->magic == kCommMagic, "vvae_comm_destroy: not a communicator");
  VVAE_COMM_PROLOGUE("vvae_comm_destroy");
  int rc = nccl_status(n, n.CommDestroy(comm->comm), "vvae_comm_destroy");
  comm->magic = 0;
  delete comm;
  return rc;
}

}  // extern "C"
