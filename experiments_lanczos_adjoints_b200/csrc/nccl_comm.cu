// NCCL directly from the library (no PyTorch in the multi-GPU path): `bl_dist_nccl_*`.
//
// north_star: "independent Hutchinson/SLQ probe vectors are sharded across the GPUs of one 8xB200 box, finishing
// with a single NCCL allreduce over NVLink.  For one large operator, rows are sharded, with an NCCL allreduce of the
// per-step dot products and an allgather of the Lanczos vector."  libnccl.so.2 is loaded at run time with dlopen
// (the library has no link-time dependency on it; BL_NCCL_LIB overrides the name): one process per GPU, the
// 128-byte unique id travels over the host communicator of the Python layer (comm.py: a socket rendezvous built
// from MASTER_ADDR / MASTER_PORT / RANK), every collective is enqueued on the caller's stream.
#include <dlfcn.h>

#include <mutex>

#include "common.cuh"

namespace {

// the part of nccl.h this file needs (ABI-stable since NCCL 2.0)
typedef struct ncclComm* ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat32 = 7, ncclFloat64 = 8 };
enum { ncclSum = 0, ncclMax = 2 };

struct Api {
  void* handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
  bool ok = false;
};

Api& api() {
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* name = std::getenv("BL_NCCL_LIB");
    a.handle = dlopen(name && name[0] ? name : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!a.handle) return;
    auto sym = [&](const char* s) { return dlsym(a.handle, s); };
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
    a.Send = reinterpret_cast<decltype(a.Send)>(sym("ncclSend"));
    a.Recv = reinterpret_cast<decltype(a.Recv)>(sym("ncclRecv"));
    a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
    a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
    a.GetVersion = reinterpret_cast<decltype(a.GetVersion)>(sym("ncclGetVersion"));
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.AllGather && a.Send && a.Recv &&
           a.GroupStart && a.GroupEnd;
  });
  return a;
}

int nccl_type(int dtype) { return dtype == BL_F32 ? ncclFloat32 : ncclFloat64; }

#define BL_NCCL(expr)                                                                                        \
  do {                                                                                                       \
    int _r = (expr);                                                                                         \
    if (_r != ncclSuccess) {                                                                                 \
      ::bl::set_error(std::string(#expr) + ": " + (api().GetErrorString ? api().GetErrorString(_r) : "NCCL error")); \
      return BL_ECUDA;                                                                                       \
    }                                                                                                        \
  } while (0)

}  // namespace

struct bl_nccl {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
};

namespace {
// row sharding: the cross-rank sum of a reduction of the Krylov loops (bl_dist_set_reduce_hook)
int nccl_reduce_hook(void* user, double* values, int count, void* stream) {
  auto* c = static_cast<bl_nccl*>(user);
  if (c->world == 1) return 0;
  return api().AllReduce(values, values, (size_t)count, ncclFloat64, ncclSum, c->comm, bl::as_stream(stream)) ==
                 ncclSuccess
             ? 0
             : 1;
}
}  // namespace

extern "C" {

int bl_dist_nccl_available(int* yes, int* version) {
  BL_REQUIRE(yes != nullptr, "NULL argument");
  *yes = api().ok ? 1 : 0;
  if (version) {
    *version = 0;
    if (api().ok && api().GetVersion) api().GetVersion(version);
  }
  return BL_OK;
}

int bl_dist_nccl_unique_id(void* id_128) {
  BL_REQUIRE(id_128 != nullptr, "NULL argument");
  BL_REQUIRE(api().ok, "libnccl.so.2 could not be loaded");
  ncclUniqueId id;
  BL_NCCL(api().GetUniqueId(&id));
  std::memcpy(id_128, &id, sizeof(id));
  return BL_OK;
}

int bl_dist_nccl_init(const void* id_128, int rank, int world, bl_nccl_t** out) {
  BL_REQUIRE(id_128 && out && world >= 1 && rank >= 0 && rank < world, "bad communicator arguments");
  BL_REQUIRE(api().ok, "libnccl.so.2 could not be loaded");
  ncclUniqueId id;
  std::memcpy(&id, id_128, sizeof(id));
  auto* c = new bl_nccl();
  c->rank = rank;
  c->world = world;
  const int rc = api().CommInitRank(&c->comm, world, id, rank);  // on the calling thread's current device
  if (rc != ncclSuccess) {
    bl::set_error(std::string("ncclCommInitRank: ") + (api().GetErrorString ? api().GetErrorString(rc) : "NCCL error"));
    delete c;
    return BL_ECUDA;
  }
  *out = c;
  return BL_OK;
}

int bl_dist_nccl_allreduce(bl_nccl_t* comm, void* buf, int64_t count, int dtype, int op, void* stream) {
  BL_REQUIRE(comm && buf && count >= 0 && (dtype == BL_F32 || dtype == BL_F64) && (op == 0 || op == 1),
             "bad all-reduce arguments");
  if (count == 0) return BL_OK;
  BL_NCCL(api().AllReduce(buf, buf, (size_t)count, nccl_type(dtype), op == 0 ? ncclSum : ncclMax, comm->comm,
                          bl::as_stream(stream)));
  return BL_OK;
}

int bl_dist_nccl_allgather(bl_nccl_t* comm, const void* send, void* recv, int64_t count, int dtype, void* stream) {
  BL_REQUIRE(comm && send && recv && count >= 0 && (dtype == BL_F32 || dtype == BL_F64), "bad all-gather arguments");
  if (count == 0) return BL_OK;
  BL_NCCL(api().AllGather(send, recv, (size_t)count, nccl_type(dtype), comm->comm, bl::as_stream(stream)));
  return BL_OK;
}

int bl_dist_nccl_sendrecv(bl_nccl_t* comm, const void* send, int send_peer, void* recv, int recv_peer, int64_t count,
                          int dtype, void* stream) {
  BL_REQUIRE(comm && count >= 0 && (dtype == BL_F32 || dtype == BL_F64), "bad send/recv arguments");
  BL_REQUIRE(send_peer < comm->world && recv_peer < comm->world, "peer out of range");
  if (count == 0 || (send_peer < 0 && recv_peer < 0)) return BL_OK;
  BL_NCCL(api().GroupStart());
  int rc = ncclSuccess;
  if (send_peer >= 0 && send)
    rc = api().Send(send, (size_t)count, nccl_type(dtype), send_peer, comm->comm, bl::as_stream(stream));
  if (rc == ncclSuccess && recv_peer >= 0 && recv)
    rc = api().Recv(recv, (size_t)count, nccl_type(dtype), recv_peer, comm->comm, bl::as_stream(stream));
  const int rc_end = api().GroupEnd();
  BL_NCCL(rc);
  BL_NCCL(rc_end);
  return BL_OK;
}

int bl_dist_nccl_reduce_hook(bl_nccl_t* comm) {
  if (comm == nullptr) return bl_dist_set_reduce_hook(nullptr, nullptr);
  return bl_dist_set_reduce_hook(nccl_reduce_hook, comm);
}

int bl_dist_nccl_destroy(bl_nccl_t* comm) {
  if (comm == nullptr) return BL_OK;
  if (comm->comm && api().ok) api().CommDestroy(comm->comm);
  delete comm;
  return BL_OK;
}

}  // extern "C"
