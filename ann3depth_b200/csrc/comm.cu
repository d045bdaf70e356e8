// comm.cu -- data-parallel gradient exchange: NCCL (dlopen'ed, no link-time dependency) sum-allreduce
// of flat gradient buckets over NVLink 5 / NVSwitch.  Replaces the parameter-server replication of
// src/ann3depth.py:78-92 (every session.run pulled weights / pushed gradients over gRPC).
#include "common.cuh"
#include <dlfcn.h>
#include <string.h>

namespace {
typedef struct { char internal[128]; } nccl_uid;
typedef int (*fn_get_uid)(nccl_uid*);
typedef int (*fn_init_rank)(void** comm, int nranks, nccl_uid id, int rank);
typedef int (*fn_destroy)(void* comm);
typedef int (*fn_allreduce)(const void* send, void* recv, size_t count, int dtype, int op, void* comm, cudaStream_t st);
typedef const char* (*fn_errstr)(int);
typedef int (*fn_reduce_scatter)(const void* send, void* recv, size_t recvcount, int dtype, int op, void* comm, cudaStream_t st);
typedef int (*fn_allgather)(const void* send, void* recv, size_t sendcount, int dtype, void* comm, cudaStream_t st);
typedef int (*fn_comm_rank)(void* comm, int* rank);

void* open_nccl(const char* path) {
  void* h = dlopen(path && path[0] ? path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) a3d_set_error("dlopen(%s) failed: %s", path && path[0] ? path : "libnccl.so.2", dlerror());
  return h;
}
}  // namespace

extern "C" int a3d_comm_unique_id(const char* nccl_path, void* id128) {
  A3D_REQUIRE(id128, "comm_unique_id: null argument");
  void* h = open_nccl(nccl_path);
  if (!h) return A3D_ENCCL;
  fn_get_uid f = (fn_get_uid)dlsym(h, "ncclGetUniqueId");
  if (!f) { a3d_set_error("ncclGetUniqueId not found"); return A3D_ENCCL; }
  nccl_uid id;
  int r = f(&id);
  if (r) { a3d_set_error("ncclGetUniqueId -> %d", r); return A3D_ENCCL; }
  memcpy(id128, &id, 128);
  return 0;
}

extern "C" int a3d_comm_init(a3d_ctx* ctx, const char* nccl_path, const void* id128, int rank, int nranks) {
  A3D_REQUIRE(ctx && id128 && nranks > 0 && rank >= 0 && rank < nranks, "comm_init: bad argument");
  if (!ctx->nccl_lib) ctx->nccl_lib = open_nccl(nccl_path);
  if (!ctx->nccl_lib) return A3D_ENCCL;
  fn_init_rank f = (fn_init_rank)dlsym(ctx->nccl_lib, "ncclCommInitRank");
  if (!f) { a3d_set_error("ncclCommInitRank not found"); return A3D_ENCCL; }
  nccl_uid id;
  memcpy(&id, id128, 128);
  A3D_CHECK_CUDA(cudaSetDevice(ctx->device));
  int r = f(&ctx->nccl_comm, nranks, id, rank);
  if (r) {
    fn_errstr es = (fn_errstr)dlsym(ctx->nccl_lib, "ncclGetErrorString");
    a3d_set_error("ncclCommInitRank -> %d (%s)", r, es ? es(r) : "?");
    ctx->nccl_comm = nullptr;
    return A3D_ENCCL;
  }
  ctx->nranks = nranks;
  return 0;
}

extern "C" int a3d_comm_destroy(a3d_ctx* ctx) {
  if (ctx && ctx->nccl_comm && ctx->nccl_lib) {
    fn_destroy f = (fn_destroy)dlsym(ctx->nccl_lib, "ncclCommDestroy");
    if (f) f(ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
  }
  return 0;
}

extern "C" int a3d_allreduce_sum(a3d_ctx* ctx, void* buf, size_t count, int dtype, void* stream) {
  A3D_REQUIRE(ctx && buf, "allreduce: null argument");
  if (!ctx->nccl_comm) { a3d_set_error("allreduce: communicator not initialised"); return A3D_ENCCL; }
  fn_allreduce f = (fn_allreduce)dlsym(ctx->nccl_lib, "ncclAllReduce");
  if (!f) { a3d_set_error("ncclAllReduce not found"); return A3D_ENCCL; }
  // ncclDataType_t: ncclFloat32 = 7, ncclBfloat16 = 9 ; ncclRedOp_t: ncclSum = 0
  int nd = dtype == A3D_BF16 ? 9 : 7;
  int r = f(buf, buf, count, nd, 0, ctx->nccl_comm, as_stream(stream));
  if (r) { a3d_set_error("ncclAllReduce -> %d", r); return A3D_ENCCL; }
  return 0;
}

static int comm_rank(a3d_ctx* ctx, int* rank) {
  fn_comm_rank f = (fn_comm_rank)dlsym(ctx->nccl_lib, "ncclCommUserRank");
  if (!f || f(ctx->nccl_comm, rank)) { a3d_set_error("ncclCommUserRank failed"); return A3D_ENCCL; }
  return 0;
}

extern "C" int a3d_reduce_scatter_sum(a3d_ctx* ctx, void* buf, size_t chunk, int dtype, void* stream) {
  A3D_REQUIRE(ctx && buf, "reduce_scatter: null argument");
  if (!ctx->nccl_comm) { a3d_set_error("reduce_scatter: communicator not initialised"); return A3D_ENCCL; }
  fn_reduce_scatter f = (fn_reduce_scatter)dlsym(ctx->nccl_lib, "ncclReduceScatter");
  if (!f) { a3d_set_error("ncclReduceScatter not found"); return A3D_ENCCL; }
  int rank = 0, rc = comm_rank(ctx, &rank);
  if (rc) return rc;
  const size_t es = dtype == A3D_BF16 ? 2 : 4;
  int r = f(buf, reinterpret_cast<char*>(buf) + (size_t)rank * chunk * es, chunk, dtype == A3D_BF16 ? 9 : 7, 0,
            ctx->nccl_comm, as_stream(stream));
  if (r) { a3d_set_error("ncclReduceScatter -> %d", r); return A3D_ENCCL; }
  return 0;
}

extern "C" int a3d_allgather(a3d_ctx* ctx, void* buf, size_t chunk, int dtype, void* stream) {
  A3D_REQUIRE(ctx && buf, "allgather: null argument");
  if (!ctx->nccl_comm) { a3d_set_error("allgather: communicator not initialised"); return A3D_ENCCL; }
  fn_allgather f = (fn_allgather)dlsym(ctx->nccl_lib, "ncclAllGather");
  if (!f) { a3d_set_error("ncclAllGather not found"); return A3D_ENCCL; }
  int rank = 0, rc = comm_rank(ctx, &rank);
  if (rc) return rc;
  const size_t es = dtype == A3D_BF16 ? 2 : 4;
  int r = f(reinterpret_cast<char*>(buf) + (size_t)rank * chunk * es, buf, chunk, dtype == A3D_BF16 ? 9 : 7,
            ctx->nccl_comm, as_stream(stream));
  if (r) { a3d_set_error("ncclAllGather -> %d", r); return A3D_ENCCL; }
  return 0;
}

// Several all-gathers as ONE NCCL group (one fused launch, one latency): bufs[i] holds nranks * chunks[i] elements and
// is gathered in place like a3d_allgather.  Used for the activation matrices of all dense layers (dp.py).
extern "C" int a3d_allgather_multi(a3d_ctx* ctx, void* const* bufs, const size_t* chunks, int n, int dtype, void* stream) {
  A3D_REQUIRE(ctx && bufs && chunks && n > 0, "allgather_multi: bad argument");
  if (!ctx->nccl_comm) { a3d_set_error("allgather_multi: communicator not initialised"); return A3D_ENCCL; }
  typedef int (*fn_group)(void);
  fn_allgather f = (fn_allgather)dlsym(ctx->nccl_lib, "ncclAllGather");
  fn_group gs = (fn_group)dlsym(ctx->nccl_lib, "ncclGroupStart");
  fn_group ge = (fn_group)dlsym(ctx->nccl_lib, "ncclGroupEnd");
  if (!f || !gs || !ge) { a3d_set_error("ncclAllGather / ncclGroupStart / ncclGroupEnd not found"); return A3D_ENCCL; }
  int rank = 0, rc = comm_rank(ctx, &rank);
  if (rc) return rc;
  const size_t es = dtype == A3D_BF16 ? 2 : 4;
  int r = gs();
  for (int i = 0; i < n && !r; ++i)
    r = f(reinterpret_cast<char*>(bufs[i]) + (size_t)rank * chunks[i] * es, bufs[i], chunks[i], dtype == A3D_BF16 ? 9 : 7,
          ctx->nccl_comm, as_stream(stream));
  int r2 = ge();
  if (r || r2) { a3d_set_error("grouped ncclAllGather -> %d / %d", r, r2); return A3D_ENCCL; }
  return 0;
}
