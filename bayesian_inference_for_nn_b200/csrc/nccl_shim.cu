// nccl_shim.cu — NCCL is resolved at run time with dlopen/dlsym, so libpyesian_b200.so carries no link
// dependency on it: only sharded SVGD (gradient / particle all-to-all, Gram and histogram all-reduce) and the sharded
// predictive need it.  Every stream synchronisation behind a collective is followed by nccl_check_async(): an
// asynchronous communicator error (a peer that died, a link fault) aborts the communicator and surfaces as PYB_ERR_CUDA
// instead of a hang in the next collective.
// The Python side preloads the libnccl.so.2 that ships with the CUDA stack (torch's bundled copy) so the
// soname lookup below finds the already-mapped library.
#include "common.cuh"
#include <dlfcn.h>

namespace pyb {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { NCCL_UINT64 = 5, NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8 };
enum { NCCL_SUM = 0, NCCL_MIN = 3 };

struct NcclApi {
  int (*GetUniqueId)(ncclUniqueId*);
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  int (*CommDestroy)(ncclComm_t);
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*GroupStart)();
  int (*GroupEnd)();
  int (*CommGetAsyncError)(ncclComm_t, int*);
  int (*CommAbort)(ncclComm_t);
  int (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, void*);
  const char* (*GetErrorString)(int);
  bool ok = false;
};

static NcclApi& api() {
  static NcclApi a;
  if (a.ok) return a;
  void* lib = nullptr;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  const char* env = getenv("PYB_NCCL_LIB");
  if (!lib && env) lib = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
  PYB_REQUIRE(lib != nullptr, PYB_ERR_UNSUPPORTED, "libnccl.so.2 not found (preload it or set PYB_NCCL_LIB)");
#define PYB_SYM(field, name)                                                               \
  a.field = (decltype(a.field))dlsym(lib, name);                                           \
  PYB_REQUIRE(a.field != nullptr, PYB_ERR_UNSUPPORTED, "NCCL symbol missing: " name);
  PYB_SYM(GetUniqueId, "ncclGetUniqueId")
  PYB_SYM(CommInitRank, "ncclCommInitRank")
  PYB_SYM(CommDestroy, "ncclCommDestroy")
  PYB_SYM(AllGather, "ncclAllGather")
  PYB_SYM(AllReduce, "ncclAllReduce")
  PYB_SYM(Broadcast, "ncclBroadcast")
  PYB_SYM(Send, "ncclSend")
  PYB_SYM(Recv, "ncclRecv")
  PYB_SYM(GroupStart, "ncclGroupStart")
  PYB_SYM(GroupEnd, "ncclGroupEnd")
  PYB_SYM(CommGetAsyncError, "ncclCommGetAsyncError")
  PYB_SYM(CommAbort, "ncclCommAbort")
  PYB_SYM(CommSplit, "ncclCommSplit")
  PYB_SYM(GetErrorString, "ncclGetErrorString")
#undef PYB_SYM
  a.ok = true;
  return a;
}

static void nccl_check(int rc, const char* what) {
  if (rc != 0) throw Error(PYB_ERR_CUDA, std::string(what) + " failed: " + api().GetErrorString(rc));
}

void nccl_unique_id(void* out_128) {
  ncclUniqueId id;
  nccl_check(api().GetUniqueId(&id), "ncclGetUniqueId");
  memcpy(out_128, id.internal, 128);
}
void* nccl_comm_init(int rank, int world, const void* id_128) {
  ncclUniqueId id;
  memcpy(id.internal, id_128, 128);
  ncclComm_t c = nullptr;
  nccl_check(api().CommInitRank(&c, world, id, rank), "ncclCommInitRank");
  return (void*)c;
}
void nccl_comm_destroy(void* comm) {
  if (comm) api().CommDestroy((ncclComm_t)comm);
}
void nccl_all_gather_f32(void* comm, const float* send, float* recv, size_t count_per_rank, cudaStream_t s) {
  nccl_check(api().AllGather(send, recv, count_per_rank, NCCL_FLOAT32, (ncclComm_t)comm, s), "ncclAllGather");
}
void nccl_all_reduce_u64(void* comm, unsigned long long* buf, size_t count, cudaStream_t s) {
  nccl_check(api().AllReduce(buf, buf, count, NCCL_UINT64, NCCL_SUM, (ncclComm_t)comm, s), "ncclAllReduce");
}
void nccl_all_reduce_min_u64(void* comm, unsigned long long* buf, size_t count, cudaStream_t s) {
  nccl_check(api().AllReduce(buf, buf, count, NCCL_UINT64, NCCL_MIN, (ncclComm_t)comm, s), "ncclAllReduce(min)");
}
// a second communicator over the same ranks (same rank order): collectives issued on ANOTHER stream run concurrently with
// those of the first one without sharing its internal ordering
void* nccl_comm_dup(void* comm, int rank) {
  ncclComm_t c = nullptr;
  nccl_check(api().CommSplit((ncclComm_t)comm, 0, rank, &c, nullptr), "ncclCommSplit");
  return (void*)c;
}
void nccl_all_reduce_f64(void* comm, double* buf, size_t count, cudaStream_t s) {
  nccl_check(api().AllReduce(buf, buf, count, NCCL_FLOAT64, NCCL_SUM, (ncclComm_t)comm, s), "ncclAllReduce");
}
void nccl_all_reduce_f32(void* comm, float* buf, size_t count, cudaStream_t s) {
  nccl_check(api().AllReduce(buf, buf, count, NCCL_FLOAT32, NCCL_SUM, (ncclComm_t)comm, s), "ncclAllReduce");
}
// personalised exchange: block q of `send` (count_per_peer floats) goes to rank q, block q of `recv` comes from rank q
void nccl_all_to_all_f32(void* comm, const float* send, float* recv, size_t count_per_peer, int world, cudaStream_t s) {
  nccl_check(api().GroupStart(), "ncclGroupStart");
  for (int q = 0; q < world; ++q) {
    nccl_check(api().Send(send + (size_t)q * count_per_peer, count_per_peer, NCCL_FLOAT32, q, (ncclComm_t)comm, s), "ncclSend");
    nccl_check(api().Recv(recv + (size_t)q * count_per_peer, count_per_peer, NCCL_FLOAT32, q, (ncclComm_t)comm, s), "ncclRecv");
  }
  nccl_check(api().GroupEnd(), "ncclGroupEnd");
}
// block q of `send` (pitch send_stride) goes to rank q, rank q's block lands in block q of `recv` (pitch recv_stride)
void nccl_all_to_all_f32_strided(void* comm, const float* send, size_t send_stride, float* recv, size_t recv_stride,
                                 size_t count, int world, cudaStream_t s) {
  nccl_check(api().GroupStart(), "ncclGroupStart");
  for (int q = 0; q < world; ++q) {
    nccl_check(api().Send(send + (size_t)q * send_stride, count, NCCL_FLOAT32, q, (ncclComm_t)comm, s), "ncclSend");
    nccl_check(api().Recv(recv + (size_t)q * recv_stride, count, NCCL_FLOAT32, q, (ncclComm_t)comm, s), "ncclRecv");
  }
  nccl_check(api().GroupEnd(), "ncclGroupEnd");
}
// part of an all-to-all: block q of `send` to every rank q in send_to, block q of `recv` from every rank q in recv_from
// (the lists of all ranks must pair up: q in send_to(r) <=> r in recv_from(q))
void nccl_exchange_f32(void* comm, const float* send, float* recv, size_t stride, size_t count, const int* send_to, int n_send,
                       const int* recv_from, int n_recv, cudaStream_t s) {
  if (n_send == 0 && n_recv == 0) return;
  nccl_check(api().GroupStart(), "ncclGroupStart");
  for (int i = 0; i < n_send; ++i)
    nccl_check(api().Send(send + (size_t)send_to[i] * stride, count, NCCL_FLOAT32, send_to[i], (ncclComm_t)comm, s), "ncclSend");
  for (int i = 0; i < n_recv; ++i)
    nccl_check(api().Recv(recv + (size_t)recv_from[i] * stride, count, NCCL_FLOAT32, recv_from[i], (ncclComm_t)comm, s), "ncclRecv");
  nccl_check(api().GroupEnd(), "ncclGroupEnd");
}
// after a stream synchronisation: has the communicator recorded an asynchronous error?  Abort it and fail cleanly.
void nccl_check_async(void** comm) {
  if (!comm || !*comm) return;
  int err = 0;
  nccl_check(api().CommGetAsyncError((ncclComm_t)*comm, &err), "ncclCommGetAsyncError");
  if (err != 0) {
    std::string msg = std::string("NCCL asynchronous error: ") + api().GetErrorString(err) + " (communicator aborted)";
    api().CommAbort((ncclComm_t)*comm);
    *comm = nullptr;
    throw Error(PYB_ERR_CUDA, msg);
  }
}
void nccl_broadcast_f32(void* comm, float* buf, size_t count, int root, cudaStream_t s) {
  nccl_check(api().Broadcast(buf, buf, count, NCCL_FLOAT32, root, (ncclComm_t)comm, s), "ncclBroadcast");
}

}  // namespace pyb
