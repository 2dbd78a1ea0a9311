// vi_table.cu -- the data formats on either side of the hot path:
//
//   * range-table import: rows (RangeID, Dimension, Mid, Id) as the reference's consumers keep them (a
//     Dictionary<long, RangeValue>, Program.cs:18-26, or the CSV "RangeID,Dimension,Mid,ID", Program.cs:80,145-149)
//     become a searchable device table: sort by RangeID, link each row to 2r+1 / 2r+2 (an absent child row stays
//     absent, DDL.sql:251-262), pack the traversal rows.  The inverse direction is vi_ranges_copy / vi_textindex_copy.
//   * record ingest: the FileRangeStore record [int64 id][dims x float32] (FileRangeStore.cs:127-165), from a host
//     buffer or streamed from a file through two pinned buffers (read of batch k+1 overlaps the H2D copy and the
//     de-interleave kernel of batch k).
#include <errno.h>
#include <stdio.h>
#include <string.h>

#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cub/device/device_radix_sort.cuh>
#include <thread>
#include <vector>

#include "vi_common.cuh"

// ---- range-table import --------------------------------------------------------------------------------------------
__global__ void k_iota(u32* p, u32 n)
{
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}

__global__ void k_gather_rows(const u32* __restrict__ order, const i64* __restrict__ rid_sorted, const int* __restrict__ dim,
                              const float* __restrict__ mid, const i64* __restrict__ id, u32 n, int dims, i64* t_rid,
                              int* t_dim, float* t_mid, i64* t_id, int* t_src, u32* __restrict__ bad)
{
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u32 o = order[i];
  const i64 r = rid_sorted[i];
  const int d = dim[o];
  t_rid[i] = r;
  t_dim[i] = d;
  t_mid[i] = mid[o];
  t_id[i] = id[o];
  t_src[i] = -1;  // no vectors behind an imported table: no candidate verification
  if (r < 0 || (d < -1 && d != VI_DIM_NULL) || d >= dims) atomicOr(bad, 1u);  // not a RangeID / Dimension of this index
  if (i > 0 && rid_sorted[i - 1] == r) atomicOr(bad, 2u);          // duplicate RangeID
  if (i == 0 && r != 0) atomicOr(bad, 4u);                         // no root row
}

__device__ __forceinline__ int find_row(const i64* __restrict__ rid, u32 n, i64 key)
{
  u32 lo = 0, hi = n;
  while (lo < hi)
  {
    const u32 m = (lo + hi) >> 1;
    if (rid[m] < key) lo = m + 1;
    else hi = m;
  }
  return (lo < n && rid[lo] == key) ? (int)lo : -1;
}

__global__ void k_link_children(const i64* __restrict__ t_rid, const int* __restrict__ t_dim, u32 n, int* t_low, int* t_high)
{
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const i64 r = t_rid[i];
  int lo = -1, hi = -1;
  if ((t_dim[i] >= 0 || t_dim[i] == VI_DIM_NULL) && r <= (0x7fffffffffffffffLL - 2) / 2)  // children 2r+1, 2r+2 (IndexBuilder.cs:99,104)
  {
    lo = find_row(t_rid, n, 2 * r + 1);
    hi = find_row(t_rid, n, 2 * r + 2);
  }
  t_low[i] = lo;
  t_high[i] = hi;
}

__global__ void k_pack_nodes_t(const int* __restrict__ t_dim, const float* __restrict__ t_mid, const i64* __restrict__ t_id,
                               const int* __restrict__ t_low, const int* __restrict__ t_high, int4* __restrict__ node, u32 n)
{
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int d = t_dim[i];
  int4 v;
  v.x = d == VI_DIM_NULL ? VI_NODE_BOTH : d;
  v.y = __float_as_int(t_mid[i]);
  if (d < 0 && d != VI_DIM_NULL)
  {
    const u64 id = (u64)t_id[i];  // leaf: the TextID rides in the child slots (same packing as vi_build.cu)
    v.z = (int)(u32)id;
    v.w = (int)(u32)(id >> 32);
  }
  else
  {
    v.z = t_low[i];
    v.w = t_high[i];
  }
  node[i] = v;
}

int vi_alloc_table_rows(vi_ctx* ctx, int64_t rows);  // vi_build.cu

int vi_ranges_load_impl(vi_ctx* ctx, const int64_t* rid, const int32_t* dim, const float* mid, const int64_t* id, int64_t n)
{
  cudaStream_t st = ctx->stream;
  ctx->built = false;
  ctx->replicated = false;
  ctx->levels.clear();
  ctx->info = vi_build_info();
  if (n == 0)
  {
    ctx->t_rows = 0;
    ctx->built = true;
    return VI_OK;
  }
  int rc = vi_alloc_table_rows(ctx, n);
  if (rc != VI_OK) return rc;
  const u32 N = (u32)n;
  // staging: unsorted columns + sort buffers
  i64 *d_rid = nullptr, *d_rid_s = nullptr, *d_id = nullptr;
  int* d_dim = nullptr;
  float* d_mid = nullptr;
  u32 *d_ord = nullptr, *d_ord_s = nullptr, *d_bad = nullptr;
  void* d_tmp = nullptr;
  size_t tmp_bytes = 0;
  auto cleanup = [&]()
  {
    cudaFree(d_rid); cudaFree(d_rid_s); cudaFree(d_id); cudaFree(d_dim); cudaFree(d_mid);
    cudaFree(d_ord); cudaFree(d_ord_s); cudaFree(d_bad); cudaFree(d_tmp);
  };
#define TRY_OR_CLEAN(call)                                               \
  do                                                                     \
  {                                                                      \
    cudaError_t e_ = (call);                                             \
    if (e_ != cudaSuccess)                                               \
    {                                                                    \
      cleanup();                                                         \
      return ctx->fail_cuda(e_, #call, __FILE__, __LINE__);              \
    }                                                                    \
  } while (0)
  TRY_OR_CLEAN(cudaMalloc((void**)&d_rid, (size_t)n * 8));
  TRY_OR_CLEAN(cudaMalloc((void**)&d_rid_s, (size_t)n * 8));
  TRY_OR_CLEAN(cudaMalloc((void**)&d_id, (size_t)n * 8));
  TRY_OR_CLEAN(cudaMalloc((void**)&d_dim, (size_t)n * 4));
  TRY_OR_CLEAN(cudaMalloc((void**)&d_mid, (size_t)n * 4));
  TRY_OR_CLEAN(cudaMalloc((void**)&d_ord, (size_t)n * 4));
  TRY_OR_CLEAN(cudaMalloc((void**)&d_ord_s, (size_t)n * 4));
  TRY_OR_CLEAN(cudaMalloc((void**)&d_bad, 4));
  TRY_OR_CLEAN(cudaMemcpyAsync(d_rid, rid, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  TRY_OR_CLEAN(cudaMemcpyAsync(d_id, id, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  TRY_OR_CLEAN(cudaMemcpyAsync(d_dim, dim, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  TRY_OR_CLEAN(cudaMemcpyAsync(d_mid, mid, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  TRY_OR_CLEAN(cudaMemsetAsync(d_bad, 0, 4, st));
  k_iota<<<(N + 255) / 256, 256, 0, st>>>(d_ord, N);
  // RangeIDs are non-negative (checked below): an unsigned 63-bit radix sort orders them
  TRY_OR_CLEAN(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const u64*)d_rid, (u64*)d_rid_s, d_ord, d_ord_s, (int)N, 0,
                                               64, st));
  TRY_OR_CLEAN(cudaMalloc(&d_tmp, tmp_bytes + 16));
  TRY_OR_CLEAN(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, (const u64*)d_rid, (u64*)d_rid_s, d_ord, d_ord_s, (int)N, 0, 64,
                                               st));
  k_gather_rows<<<(N + 255) / 256, 256, 0, st>>>(d_ord_s, d_rid_s, d_dim, d_mid, d_id, N, ctx->dims, ctx->t_rid, ctx->t_dim,
                                                 ctx->t_mid, ctx->t_id, ctx->t_src, d_bad);
  k_link_children<<<(N + 255) / 256, 256, 0, st>>>(ctx->t_rid, ctx->t_dim, N, ctx->t_low, ctx->t_high);
  k_pack_nodes_t<<<(N + 255) / 256, 256, 0, st>>>(ctx->t_dim, ctx->t_mid, ctx->t_id, ctx->t_low, ctx->t_high, ctx->t_node, N);
  u32 bad = 0;
  TRY_OR_CLEAN(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, st));
  TRY_OR_CLEAN(cudaStreamSynchronize(st));
  TRY_OR_CLEAN(cudaGetLastError());
#undef TRY_OR_CLEAN
  cleanup();
  if (bad & 1u) return ctx->fail(VI_ERR_INVALID_ARG, "range table: negative RangeID or Dimension outside [-1, dims) (VI_DIM_NULL = -3 is a null Dimension)");
  if (bad & 2u) return ctx->fail(VI_ERR_INVALID_ARG, "range table: duplicate RangeID");
  if (bad & 4u) return ctx->fail(VI_ERR_INVALID_ARG, "range table: no row for RangeID 0");
  ctx->t_rows = n;
  ctx->info.ranges = n;
  ctx->replicated = true;  // (same consequences as a replicated table: rows only, no vectors to verify against)
  ctx->built = true;
  return VI_OK;
}

// ---- record ingest -------------------------------------------------------------------------------------------------
// one thread per (record, 4-byte word): word 0..1 = id, 2.. = vector
// idw = 2: FileRangeStore records; idw = 0: bare rows (an HDF5 data set), ids are first_id, first_id + 1, ...
__global__ void k_split_records(const u32* __restrict__ rec, u32 words, u32 n, int dims, int ld, i64* __restrict__ ids,
                                float* __restrict__ rows, u32 idw = 2, i64 first_id = 0)
{
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  const u64 r = t / words;
  const u32 w = (u32)(t % words);
  if (r >= n) return;
  const u32 x = rec[t];
  if (w < idw) reinterpret_cast<u32*>(ids + r)[w] = x;
  else rows[r * ld + (w - idw)] = __uint_as_float(x);
  if (w == idw)
  {
    if (idw == 0) ids[r] = first_id + (i64)r;
    for (int c = dims; c < ld; ++c) rows[r * ld + c] = 0.f;  // padding columns
  }
}

int vi_reserve_for_add(vi_ctx* ctx, int64_t n);  // vi_abi.cu: grows the point store for n more points

static int split_batch(vi_ctx* ctx, const u32* d_rec, int64_t n, cudaStream_t st)
{
  const u32 words = 2u + (u32)ctx->dims;
  const u64 total = (u64)n * words;
  k_split_records<<<(u32)((total + 255) / 256), 256, 0, st>>>(d_rec, words, (u32)n, ctx->dims, ctx->ld, ctx->ids + ctx->n,
                                                             ctx->rows + (size_t)ctx->n * ctx->ld);
  return VI_OK;
}

int vi_points_add_records_impl(vi_ctx* ctx, const void* records, int64_t n)
{
  const size_t rec_bytes = 8 + 4 * (size_t)ctx->dims;
  int rc = vi_reserve_for_add(ctx, n);
  if (rc != VI_OK) return rc;
  const int64_t batch = std::max<int64_t>(1, (int64_t)((64u << 20) / rec_bytes));
  u32* d_rec = nullptr;
  VI_CUDA_TRY(cudaMalloc((void**)&d_rec, (size_t)std::min(batch, n) * rec_bytes + 16));
  for (int64_t done = 0; done < n; done += batch)
  {
    const int64_t k = std::min(batch, n - done);
    cudaError_t e = cudaMemcpyAsync(d_rec, (const char*)records + (size_t)done * rec_bytes, (size_t)k * rec_bytes,
                                    cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
    {
      split_batch(ctx, d_rec, k, ctx->stream);
      e = cudaStreamSynchronize(ctx->stream);
    }
    if (e != cudaSuccess)
    {
      cudaFree(d_rec);
      return ctx->fail_cuda(e, "record ingest", __FILE__, __LINE__);
    }
    ctx->n += k;
  }
  cudaFree(d_rec);
  return VI_OK;
}

// reads [offset, offset + bytes) of fd into dst with a few threads (one thread copies out of the page cache at
// ~8 GB/s, well below PCIe); false on a short read
static bool parallel_pread(int fd, char* dst, size_t bytes, off_t offset, unsigned nthreads)
{
  std::atomic<bool> ok(true);
  auto part = [&](size_t a, size_t b)
  {
    while (a < b)
    {
      const ssize_t g = pread(fd, dst + a, b - a, offset + (off_t)a);
      if (g <= 0) { ok = false; return; }
      a += (size_t)g;
    }
  };
  nthreads = std::max(1u, std::min<unsigned>(nthreads, (unsigned)(bytes >> 20) + 1u));
  std::vector<std::thread> th;
  const size_t chunk = ((bytes + nthreads - 1) / nthreads + 4095) & ~(size_t)4095;
  for (unsigned t = 1; t < nthreads; ++t)
    if ((size_t)t * chunk < bytes) th.emplace_back(part, (size_t)t * chunk, std::min(bytes, (size_t)(t + 1) * chunk));
  part(0, std::min(bytes, chunk));
  for (auto& t : th) t.join();
  return ok;
}

static int add_file_impl(vi_ctx* ctx, const char* path, int64_t offset_bytes, int64_t n, double* read_ms, double* total_ms,
                         u32 idw, i64 first_id);

int vi_points_add_file_impl(vi_ctx* ctx, const char* path, int64_t offset_bytes, int64_t n, double* read_ms, double* total_ms)
{
  return add_file_impl(ctx, path, offset_bytes, n, read_ms, total_ms, 2u, 0);
}

// bare float32 rows (the contiguous storage of an HDF5 data set, vi_hdf5.cu): ids first_id, first_id + 1, ...
int vi_points_add_rows_file_impl(vi_ctx* ctx, const char* path, int64_t offset_bytes, int64_t n, int64_t first_id,
                                 double* read_ms, double* total_ms)
{
  return add_file_impl(ctx, path, offset_bytes, n, read_ms, total_ms, 0u, first_id);
}

static int add_file_impl(vi_ctx* ctx, const char* path, int64_t offset_bytes, int64_t n, double* read_ms, double* total_ms,
                         u32 idw, i64 first_id)
{
  const size_t rec_bytes = 4 * (size_t)idw + 4 * (size_t)ctx->dims;
  FILE* f = fopen(path, "rb");
  if (!f) return ctx->fail(VI_ERR_INVALID_ARG, std::string("cannot open ") + path + ": " + strerror(errno));
  if (n < 0)  // to the end of the file
  {
    fseeko(f, 0, SEEK_END);
    const int64_t size = (int64_t)ftello(f);
    n = size > offset_bytes ? (size - offset_bytes) / (int64_t)rec_bytes : 0;
  }
  int rc = vi_reserve_for_add(ctx, n);
  if (rc != VI_OK) { fclose(f); return rc; }
  const int64_t batch = std::max<int64_t>(1, (int64_t)((64u << 20) / rec_bytes));
  const unsigned nread = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
  void* h[2] = {nullptr, nullptr};
  u32* d[2] = {nullptr, nullptr};
  cudaStream_t cs[2] = {nullptr, nullptr};
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  auto cleanup = [&]()
  {
    for (int i = 0; i < 2; ++i)
    {
      if (cs[i]) { cudaStreamSynchronize(cs[i]); cudaStreamDestroy(cs[i]); }
      cudaFreeHost(h[i]);
      cudaFree(d[i]);
    }
    if (t0) cudaEventDestroy(t0);
    if (t1) cudaEventDestroy(t1);
    fclose(f);
  };
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < 2 && e == cudaSuccess; ++i)
  {
    e = cudaMallocHost(&h[i], (size_t)batch * rec_bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d[i], (size_t)batch * rec_bytes + 16);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&cs[i], cudaStreamNonBlocking);
  }
  if (e == cudaSuccess) e = cudaEventCreate(&t0);
  if (e == cudaSuccess) e = cudaEventCreate(&t1);
  if (e != cudaSuccess)
  {
    cleanup();
    return ctx->fail_cuda(e, "record file ingest: buffers", __FILE__, __LINE__);
  }
  cudaStreamSynchronize(ctx->stream);
  cudaEventRecord(t0, cs[0]);
  double rd_ms = 0;
  int64_t done = 0;
  int b = 0;
  const int64_t n0 = ctx->n;
  while (done < n)
  {
    const int64_t k = std::min(batch, n - done);
    e = cudaStreamSynchronize(cs[b]);  // the buffer pair b is free again
    if (e != cudaSuccess) break;
    timespec a, z;
    clock_gettime(CLOCK_MONOTONIC, &a);
    const bool got = parallel_pread(fileno(f), (char*)h[b], (size_t)k * rec_bytes,
                                    (off_t)offset_bytes + (off_t)((size_t)done * rec_bytes), nread);
    clock_gettime(CLOCK_MONOTONIC, &z);
    rd_ms += (z.tv_sec - a.tv_sec) * 1e3 + (z.tv_nsec - a.tv_nsec) * 1e-6;
    if (!got)
    {
      cleanup();
      ctx->n = n0;
      return ctx->fail(VI_ERR_INVALID_ARG, "record file shorter than the requested number of records");
    }
    e = cudaMemcpyAsync(d[b], h[b], (size_t)k * rec_bytes, cudaMemcpyHostToDevice, cs[b]);
    if (e != cudaSuccess) break;
    // (ctx->n is only advanced below: the kernel of this batch writes behind the points of the batches before it)
    const u32 words = idw + (u32)ctx->dims;
    const u64 total = (u64)k * words;
    k_split_records<<<(u32)((total + 255) / 256), 256, 0, cs[b]>>>(d[b], words, (u32)k, ctx->dims, ctx->ld,
                                                                  ctx->ids + n0 + done,
                                                                  ctx->rows + (size_t)(n0 + done) * ctx->ld, idw,
                                                                  first_id + done);
    done += k;
    b ^= 1;
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(cs[0]);
  if (e == cudaSuccess) e = cudaStreamSynchronize(cs[1]);
  if (e == cudaSuccess) e = cudaGetLastError();
  float ms = 0;
  if (e == cudaSuccess)
  {
    cudaEventRecord(t1, cs[0]);
    cudaEventSynchronize(t1);
    cudaEventElapsedTime(&ms, t0, t1);
  }
  cleanup();
  if (e != cudaSuccess)
  {
    ctx->n = n0;
    return ctx->fail_cuda(e, "record file ingest", __FILE__, __LINE__);
  }
  ctx->n = n0 + n;
  if (read_ms) *read_ms = rd_ms;
  if (total_ms) *total_ms = ms;
  return VI_OK;
}
