// NCCL plumbing for the feature-parallel minibatch path.  libnccl is resolved at run time
// (dlopen) so the single-GPU library has no link-time dependency on it; in a torchrun process the copy torch
// already loaded is reused.
#include "common.cuh"

#include <dlfcn.h>
#include <mutex>

namespace fmwr {

typedef struct { char internal[128]; } NcclUniqueId;
typedef int (*fn_get_unique_id)(NcclUniqueId*);
typedef int (*fn_comm_init_rank)(void**, int, NcclUniqueId, int);
typedef int (*fn_comm_destroy)(void*);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*fn_error_string)(int);

static struct {
  void* handle = nullptr;
  fn_get_unique_id get_unique_id = nullptr;
  fn_comm_init_rank comm_init_rank = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_all_reduce all_reduce = nullptr;
  fn_error_string error_string = nullptr;
} g_nccl;

static void load_nccl()
{
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (g_nccl.handle) return;
  const char* names[] = {getenv("FMWR_NCCL_LIB"), "libnccl.so.2", "libnccl.so", nullptr};
  for (int i = 0; i < 4 && !g_nccl.handle; ++i) {
    if (!names[i]) continue;
    g_nccl.handle = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  }
  FMWR_REQUIRE(g_nccl.handle, FMWR_ERR_COMM, "cannot load libnccl.so.2 (set FMWR_NCCL_LIB)");
  g_nccl.get_unique_id = (fn_get_unique_id)dlsym(g_nccl.handle, "ncclGetUniqueId");
  g_nccl.comm_init_rank = (fn_comm_init_rank)dlsym(g_nccl.handle, "ncclCommInitRank");
  g_nccl.comm_destroy = (fn_comm_destroy)dlsym(g_nccl.handle, "ncclCommDestroy");
  g_nccl.all_reduce = (fn_all_reduce)dlsym(g_nccl.handle, "ncclAllReduce");
  g_nccl.error_string = (fn_error_string)dlsym(g_nccl.handle, "ncclGetErrorString");
  FMWR_REQUIRE(g_nccl.get_unique_id && g_nccl.comm_init_rank && g_nccl.comm_destroy && g_nccl.all_reduce, FMWR_ERR_COMM,
               "libnccl lacks an expected symbol");
}

static void nccl_check(int rc, const char* what)
{
  if (rc != 0) {
    std::string msg = std::string("NCCL error in ") + what + ": " + (g_nccl.error_string ? g_nccl.error_string(rc) : "?");
    throw Error(FMWR_ERR_COMM, msg);
  }
}

void comm_allreduce_sum(fmwr_ctx* ctx, void* buf, size_t count, bool f64)
{
  FMWR_REQUIRE(ctx->nccl_comm, FMWR_ERR_COMM, "no communicator");
  nccl_check(g_nccl.all_reduce(buf, buf, count, f64 ? 8 /*ncclFloat64*/ : 7 /*ncclFloat32*/, 0 /*ncclSum*/, ctx->nccl_comm, ctx->stream),
             "ncclAllReduce");
  ctx->launches++;
}

void comm_allreduce_sum_u32(fmwr_ctx* ctx, uint32_t* buf, size_t count)
{
  FMWR_REQUIRE(ctx->nccl_comm, FMWR_ERR_COMM, "no communicator");
  nccl_check(g_nccl.all_reduce(buf, buf, count, 3 /*ncclUint32*/, 0 /*ncclSum*/, ctx->nccl_comm, ctx->stream), "ncclAllReduce");
  ctx->launches++;
}

void comm_allreduce_max_u32(fmwr_ctx* ctx, uint32_t* buf, size_t count)
{
  FMWR_REQUIRE(ctx->nccl_comm, FMWR_ERR_COMM, "no communicator");
  nccl_check(g_nccl.all_reduce(buf, buf, count, 3 /*ncclUint32*/, 2 /*ncclMax*/, ctx->nccl_comm, ctx->stream), "ncclAllReduce");
  ctx->launches++;
}

}  // namespace fmwr

using namespace fmwr;

// A peer that never reached an in-kernel barrier leaves PEER_ERR set in our window (common.cuh: peer_wait).  Read it after the
// final stream sync of a trainer, clear it (one timeout must not poison later calls) and report FMWR_ERR_COMM.
namespace fmwr {
void peer_check_error(fmwr_ctx* ctx)
{
  if (!ctx->peer.ready || !ctx->peer.base[ctx->rank]) return;
  uint32_t* word = reinterpret_cast<uint32_t*>(ctx->peer.base[ctx->rank]) + PEER_ERR;
  uint32_t err = 0;
  FMWR_CUDA(cudaMemcpy(&err, word, 4, cudaMemcpyDeviceToHost));
  if (!err) return;
  FMWR_CUDA(cudaMemset(word, 0, 4));
  throw Error(FMWR_ERR_COMM, "a peer rank did not reach the in-kernel barrier (timeout); the model is invalid");
}
}  // namespace fmwr

extern "C" {

int fmwr_comm_unique_id(uint8_t* id128)
{
  return guarded([&] {
    FMWR_REQUIRE(id128, FMWR_ERR_ARG, "null argument");
    load_nccl();
    NcclUniqueId id;
    nccl_check(g_nccl.get_unique_id(&id), "ncclGetUniqueId");
    memcpy(id128, id.internal, 128);
  });
}

int fmwr_comm_init(fmwr_ctx* ctx, const uint8_t* id128, int32_t rank, int32_t world)
{
  return guarded([&] {
    FMWR_REQUIRE(ctx && id128 && world >= 1 && rank >= 0 && rank < world, FMWR_ERR_ARG, "bad argument");
    FMWR_REQUIRE(!ctx->nccl_comm, FMWR_ERR_ARG, "communicator already initialised");
    load_nccl();
    FMWR_CUDA(cudaSetDevice(ctx->device));
    NcclUniqueId id;
    memcpy(id.internal, id128, 128);
    void* comm = nullptr;
    nccl_check(g_nccl.comm_init_rank(&comm, world, id, rank), "ncclCommInitRank");
    ctx->nccl_comm = comm; ctx->rank = rank; ctx->world = world;
  });
}

int64_t fmwr_comm_peer_bytes(int64_t batch_size, int32_t k, int32_t world)
{
  if (batch_size <= 0 || k < 0 || world < 1) return 0;
  // stride <= 2k + 8 covers the padding of either precision; slabs: world x ceil(B / world) rows, S cache: B rows, mult: B
  const int64_t stride = 2 * (int64_t)k + 8;
  const int64_t rpo = (batch_size + world - 1) / world;
  return PEER_CTL_BYTES + 8 * (world * rpo * stride + batch_size * stride + batch_size) + 4 * 256 + 8192;   // + the multiplier-sum slots
}

int fmwr_comm_peer_alloc(fmwr_ctx* ctx, int64_t bytes, uint8_t* handle64)
{
  return guarded([&] {
    FMWR_REQUIRE(ctx && handle64 && bytes >= PEER_CTL_BYTES, FMWR_ERR_ARG, "bad argument");
    FMWR_REQUIRE(ctx->world > 1 && ctx->world <= 8, FMWR_ERR_ARG, "peer windows need an initialised communicator of 2..8 ranks");
    FMWR_REQUIRE(!ctx->peer.base[ctx->rank], FMWR_ERR_ARG, "peer window already allocated");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    FMWR_CUDA(cudaSetDevice(ctx->device));
    void* p = nullptr;
    FMWR_CUDA(cudaMalloc(&p, (size_t)bytes));                  // a plain allocation: IPC handles name whole allocations
    FMWR_CUDA(cudaMemset(p, 0, (size_t)bytes));
    FMWR_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    FMWR_CUDA(cudaIpcGetMemHandle(&h, p));
    memcpy(handle64, &h, 64);
    ctx->peer.base[ctx->rank] = p;
    ctx->peer.bytes = (size_t)bytes;
  });
}

int fmwr_comm_peer_open(fmwr_ctx* ctx, const uint8_t* handles)
{
  return guarded([&] {
    FMWR_REQUIRE(ctx && handles, FMWR_ERR_ARG, "null argument");
    FMWR_REQUIRE(ctx->peer.base[ctx->rank] && !ctx->peer.ready, FMWR_ERR_ARG, "allocate the local window first (once)");
    FMWR_CUDA(cudaSetDevice(ctx->device));
    for (int r = 0; r < ctx->world; ++r) {
      if (r == ctx->rank) continue;
      cudaIpcMemHandle_t h;
      memcpy(&h, handles + 64 * r, 64);
      void* p = nullptr;
      FMWR_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
      ctx->peer.base[r] = p;
    }
    ctx->peer.ready = true;
  });
}

static void peer_close(fmwr_ctx* ctx)
{
  for (int r = 0; r < 8; ++r) {
    if (!ctx->peer.base[r]) continue;
    if (r == ctx->rank) cudaFree(ctx->peer.base[r]);
    else cudaIpcCloseMemHandle(ctx->peer.base[r]);
    ctx->peer.base[r] = nullptr;
  }
  ctx->peer.ready = false; ctx->peer.bytes = 0;
}

int fmwr_comm_destroy(fmwr_ctx* ctx)
{
  return guarded([&] {
    FMWR_REQUIRE(ctx, FMWR_ERR_ARG, "null ctx");
    cudaStreamSynchronize(ctx->stream);
    peer_close(ctx);
    if (ctx->nccl_comm) {
      cudaStreamSynchronize(ctx->stream);
      g_nccl.comm_destroy(ctx->nccl_comm);
      ctx->nccl_comm = nullptr; ctx->rank = 0; ctx->world = 1;
    }
  });
}

}  // extern "C"
