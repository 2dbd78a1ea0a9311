// vi_comm.cu -- the collectives of the multi-rank build, owned by the library: NCCL over NVLink 5 / NVSwitch, enqueued
// on the context's own stream (no host synchronisation around a collective).  libnccl.so.2 is bound at run time with
// dlopen/dlsym -- the copy already loaded in the process (torch ships one) or the system one -- so the library has no
// link-time dependency and a single-GPU host needs no NCCL at all.
//
// The reference has no counterpart (single process); this is north_star's "per-level range statistics
// are combined with a single NCCL all-reduce over NVLink".  The callback form (vi_set_collective) remains for hosts that
// bring their own transport (the CPU tests run it over gloo).
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include "vi_common.cuh"

namespace
{
struct NcclApi
{
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi g_nccl;

bool load_nccl(std::string& why)
{
  if (g_nccl.ok) return true;
  const char* names[] = {"libnccl.so.2", "/usr/lib/x86_64-linux-gnu/libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names)
  {
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h)
  {
    why = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "");
    return false;
  }
  g_nccl.handle = h;
#define VI_SYM(field, name)                                                 \
  *(void**)(&g_nccl.field) = dlsym(h, name);                                \
  if (!g_nccl.field) { why = std::string("NCCL symbol missing: ") + name; return false; }
  VI_SYM(GetUniqueId, "ncclGetUniqueId")
  VI_SYM(CommInitRank, "ncclCommInitRank")
  VI_SYM(CommDestroy, "ncclCommDestroy")
  VI_SYM(AllReduce, "ncclAllReduce")
  VI_SYM(AllGather, "ncclAllGather")
  VI_SYM(Broadcast, "ncclBroadcast")
  VI_SYM(Send, "ncclSend")
  VI_SYM(Recv, "ncclRecv")
  VI_SYM(GroupStart, "ncclGroupStart")
  VI_SYM(GroupEnd, "ncclGroupEnd")
  VI_SYM(GetErrorString, "ncclGetErrorString")
#undef VI_SYM
  g_nccl.ok = true;
  return true;
}

int nccl_fail(vi_ctx* ctx, ncclResult_t r, const char* what)
{
  return ctx->fail(VI_ERR_CUDA, std::string("NCCL error: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?") + " at " + what);
}
}  // namespace

#define VI_NCCL_TRY(expr)                                       \
  do                                                            \
  {                                                             \
    ncclResult_t _r = (expr);                                   \
    if (_r != ncclSuccess) return nccl_fail(ctx, _r, #expr);    \
  } while (0)

// ---- internal collective layer used by vi_build.cu (NCCL when the library owns a communicator, else the callbacks) ------
// All of them are enqueued on ctx->stream when NCCL is used; with callbacks they synchronise the stream first (the
// callback contract is host-synchronous).
bool vi_coll_in_stream(const vi_ctx* ctx) { return ctx->nccl != nullptr; }

int vi_coll_allreduce_u64(vi_ctx* ctx, void* d_buf, int64_t count)
{
  if (count <= 0) return VI_OK;
  if (ctx->nccl)
  {
    VI_NCCL_TRY(g_nccl.AllReduce(d_buf, d_buf, (size_t)count, ncclUint64, ncclSum, (ncclComm_t)ctx->nccl, ctx->stream));
    ++ctx->coll_calls[0];
    ctx->coll_bytes[0] += count * 8;
    return VI_OK;
  }
  VI_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  if (!ctx->allreduce || ctx->allreduce(ctx->coll_user, d_buf, count) != 0) return ctx->fail(VI_ERR_CUDA, "all-reduce callback failed");
  ++ctx->coll_calls[0];
  ctx->coll_bytes[0] += count * 8;
  return VI_OK;
}

// rank r's send_bytes[d] bytes (consecutive in d_send) go to rank d; recv_bytes[r] arrive from rank r, ordered by source
int vi_coll_alltoallv(vi_ctx* ctx, const void* d_send, const int64_t* send_bytes, void* d_recv, const int64_t* recv_bytes)
{
  const int G = ctx->world;
  if (ctx->nccl)
  {
    VI_NCCL_TRY(g_nccl.GroupStart());
    int64_t so = 0, ro = 0;
    for (int g = 0; g < G; ++g)
    {
      if (send_bytes[g] > 0)
        VI_NCCL_TRY(g_nccl.Send((const char*)d_send + so, (size_t)send_bytes[g], ncclUint8, g, (ncclComm_t)ctx->nccl, ctx->stream));
      if (recv_bytes[g] > 0)
        VI_NCCL_TRY(g_nccl.Recv((char*)d_recv + ro, (size_t)recv_bytes[g], ncclUint8, g, (ncclComm_t)ctx->nccl, ctx->stream));
      so += send_bytes[g];
      ro += recv_bytes[g];
    }
    VI_NCCL_TRY(g_nccl.GroupEnd());
    ++ctx->coll_calls[1];
    ctx->coll_bytes[1] += so;
    return VI_OK;
  }
  VI_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  if (!ctx->alltoallv || ctx->alltoallv(ctx->coll_user, d_send, send_bytes, d_recv, recv_bytes) != 0)
    return ctx->fail(VI_ERR_CUDA, "all-to-all callback failed");
  ++ctx->coll_calls[1];
  for (int g = 0; g < G; ++g) ctx->coll_bytes[1] += send_bytes[g];
  return VI_OK;
}

// every rank contributes `bytes` bytes at d_send; d_recv receives world * bytes, ordered by rank
int vi_coll_allgather(vi_ctx* ctx, const void* d_send, void* d_recv, int64_t bytes)
{
  if (bytes <= 0) return VI_OK;
  const int G = ctx->world;
  if (ctx->nccl)
  {
    VI_NCCL_TRY(g_nccl.AllGather(d_send, d_recv, (size_t)bytes, ncclUint8, (ncclComm_t)ctx->nccl, ctx->stream));
    ++ctx->coll_calls[2];
    ctx->coll_bytes[2] += bytes;
    return VI_OK;
  }
  // callback hosts: an all-to-all with equal pieces
  std::vector<int64_t> sb((size_t)G, bytes), rb((size_t)G, bytes);
  // the send buffer of the all-to-all holds one copy per destination
  char* tmp = nullptr;
  VI_CUDA_TRY(cudaMalloc((void**)&tmp, (size_t)bytes * G));
  for (int g = 0; g < G; ++g)
    cudaMemcpyAsync(tmp + (size_t)g * bytes, d_send, (size_t)bytes, cudaMemcpyDeviceToDevice, ctx->stream);
  int rc = vi_coll_alltoallv(ctx, tmp, sb.data(), d_recv, rb.data());
  cudaStreamSynchronize(ctx->stream);
  cudaFree(tmp);
  return rc;
}

// variable-size all-gather: rank g contributes bytes[g] at d_buf + offset[g] (in place: every rank's buffer has the
// whole layout)
int vi_coll_allgatherv_inplace(vi_ctx* ctx, void* d_buf, const int64_t* offset, const int64_t* bytes)
{
  const int G = ctx->world;
  if (ctx->nccl)
  {
    // point-to-point pairs inside one group (the connections exist since the build's all-to-all; a broadcast per rank
    // would set up new ring channels on first use: measured 600 ms)
    const int me = ctx->rank;
    VI_NCCL_TRY(g_nccl.GroupStart());
    for (int g = 0; g < G; ++g)
    {
      if (g == me) continue;
      if (bytes[me] > 0)
        VI_NCCL_TRY(g_nccl.Send((const char*)d_buf + offset[me], (size_t)bytes[me], ncclUint8, g, (ncclComm_t)ctx->nccl, ctx->stream));
      if (bytes[g] > 0)
        VI_NCCL_TRY(g_nccl.Recv((char*)d_buf + offset[g], (size_t)bytes[g], ncclUint8, g, (ncclComm_t)ctx->nccl, ctx->stream));
    }
    VI_NCCL_TRY(g_nccl.GroupEnd());
    ++ctx->coll_calls[2];
    ctx->coll_bytes[2] += bytes[me] * (G - 1);
    return VI_OK;
  }
  // callback hosts: all-to-all where everybody sends its piece to everybody
  std::vector<int64_t> sb((size_t)G, bytes[ctx->rank]), rb((size_t)G);
  int64_t total = 0;
  for (int g = 0; g < G; ++g) { rb[g] = bytes[g]; total += bytes[g]; }
  char *tmp = nullptr, *out = nullptr;
  VI_CUDA_TRY(cudaMalloc((void**)&tmp, (size_t)std::max<int64_t>(bytes[ctx->rank] * G, 16)));
  VI_CUDA_TRY(cudaMalloc((void**)&out, (size_t)std::max<int64_t>(total, 16)));
  for (int g = 0; g < G; ++g)
    cudaMemcpyAsync(tmp + (size_t)g * bytes[ctx->rank], (char*)d_buf + offset[ctx->rank], (size_t)bytes[ctx->rank],
                    cudaMemcpyDeviceToDevice, ctx->stream);
  int rc = vi_coll_alltoallv(ctx, tmp, sb.data(), out, rb.data());
  if (rc == VI_OK)
  {
    int64_t ro = 0;
    for (int g = 0; g < G; ++g)
    {
      cudaMemcpyAsync((char*)d_buf + offset[g], out + ro, (size_t)bytes[g], cudaMemcpyDeviceToDevice, ctx->stream);
      ro += bytes[g];
    }
  }
  cudaStreamSynchronize(ctx->stream);
  cudaFree(tmp);
  cudaFree(out);
  return rc;
}

void vi_comm_release(vi_ctx* ctx)
{
  if (ctx->nccl && g_nccl.ok) g_nccl.CommDestroy((ncclComm_t)ctx->nccl);
  ctx->nccl = nullptr;
}

extern "C" {

int vi_comm_unique_id(void* out, int32_t bytes)
{
  if (!out || bytes < (int32_t)NCCL_UNIQUE_ID_BYTES) return VI_ERR_INVALID_ARG;
  std::string why;
  if (!load_nccl(why)) return VI_ERR_STATE;
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return VI_ERR_CUDA;
  memcpy(out, id.internal, NCCL_UNIQUE_ID_BYTES);
  return VI_OK;
}

int vi_comm_init(vi_ctx* ctx, const void* unique_id, int32_t bytes, int32_t rank, int32_t world)
{
  if (!ctx) return VI_ERR_INVALID_ARG;
  if (world < 1 || world > 64 || rank < 0 || rank >= world || (world > 1 && (!unique_id || bytes < (int32_t)NCCL_UNIQUE_ID_BYTES)))
    return ctx->fail(VI_ERR_INVALID_ARG, "bad communicator arguments");
  vi_comm_release(ctx);
  ctx->rank = rank;
  ctx->world = world;
  ctx->allreduce = nullptr;
  ctx->alltoallv = nullptr;
  ctx->coll_user = nullptr;
  if (world == 1) return VI_OK;
  std::string why;
  if (!load_nccl(why)) return ctx->fail(VI_ERR_STATE, why);
  VI_CUDA_TRY(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(id.internal, unique_id, NCCL_UNIQUE_ID_BYTES);
  ncclComm_t comm = nullptr;
  VI_NCCL_TRY(g_nccl.CommInitRank(&comm, world, id, rank));
  ctx->nccl = comm;
  return VI_OK;
}

int vi_comm_stats(const vi_ctx* ctx, int64_t* calls3, int64_t* bytes3)
{
  if (!ctx) return VI_ERR_INVALID_ARG;
  for (int i = 0; i < 3; ++i)
  {
    if (calls3) calls3[i] = ctx->coll_calls[i];
    if (bytes3) bytes3[i] = ctx->coll_bytes[i];
  }
  return VI_OK;
}

}  // extern "C"
