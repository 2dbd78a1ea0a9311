// vi_abi.cu -- the extern "C" boundary declared in include/vi_b200.h.  Host-side bookkeeping only; all compute is in
// vi_build.cu / vi_search.cu.  There is no CPU fallback anywhere: without a usable CUDA device vi_create fails.
#include <math.h>
#include <string.h>

#include <new>

#include "vi_common.cuh"

static const char* kNullCtx = "null context";

template <typename T>
static int grow(vi_ctx* ctx, T** buf, int64_t* cap, int64_t need)
{
  if (*cap >= need && *buf) return VI_OK;
  cudaFree(*buf);
  *buf = nullptr;
  *cap = 0;
  int64_t ncap = need + need / 4 + 256;
  VI_CUDA_TRY(cudaMalloc((void**)buf, (size_t)ncap * sizeof(T)));
  *cap = ncap;
  return VI_OK;
}

int vi_ranges_load_impl(vi_ctx* ctx, const int64_t* rid, const int32_t* dim, const float* mid, const int64_t* id, int64_t n);
int vi_points_add_records_impl(vi_ctx* ctx, const void* records, int64_t n);
int vi_points_add_file_impl(vi_ctx* ctx, const char* path, int64_t offset_bytes, int64_t n, double* read_ms, double* total_ms);
int vi_search_topk_impl(vi_ctx* ctx, const float* d_queries, int64_t nq, float proximity, int32_t k, int32_t metric,
                        int64_t* ids, float* dist, int32_t* count, int64_t* candidates);

extern "C" {

int vi_abi_version(void) { return VI_ABI_VERSION; }

int vi_create(int32_t device, vi_ctx** out)
{
  if (!out) return VI_ERR_INVALID_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count)
  {
    cudaGetLastError();
    return VI_ERR_CUDA;  // no CPU fallback
  }
  vi_ctx* ctx = new (std::nothrow) vi_ctx();
  if (!ctx) return VI_ERR_OOM;
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMalloc((void**)&ctx->counters, 256) != cudaSuccess)
  {
    cudaGetLastError();
    delete ctx;
    return VI_ERR_CUDA;
  }
  cudaMemset(ctx->counters, 0, 256);
  *out = ctx;
  return VI_OK;
}

static void free_points(vi_ctx* ctx)
{
  cudaFree(ctx->rows); ctx->rows = nullptr;
  cudaFree(ctx->ids); ctx->ids = nullptr;
  ctx->capacity = 0;
  ctx->n = 0;
}

void vi_destroy(vi_ctx* ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  vi_free_workspace(ctx);
  vi_free_table(ctx);
  free_points(ctx);
  cudaFree(ctx->q_buf); cudaFree(ctx->off_buf); cudaFree(ctx->ids_buf); cudaFree(ctx->off2_buf);
  cudaFree(ctx->ids2_buf); cudaFree(ctx->search_src); cudaFree(ctx->verify_keep);
  cudaFree(ctx->sw_pool); cudaFree(ctx->sw_head); cudaFree(ctx->sw_spill);
  cudaFree(ctx->own_rows); cudaFree(ctx->own_ids); cudaFree(ctx->send_rows); cudaFree(ctx->send_ids);
  vi_comm_release(ctx);
  cudaFree(ctx->sh_dev);
  if (ctx->sh_host) cudaFreeHost(ctx->sh_host);
  cudaFree(ctx->counters);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* vi_last_error(const vi_ctx* ctx) { return ctx ? ctx->err.c_str() : kNullCtx; }

static int reserve_points(vi_ctx* ctx, int64_t capacity)
{
  // grows the point store, keeping the points already added
  if (capacity <= ctx->capacity && ctx->rows) return VI_OK;
  float* nrows = nullptr;
  i64* nids = nullptr;
  const size_t rbytes = (size_t)capacity * ctx->ld * sizeof(float) + 256;
  VI_CUDA_TRY(cudaMalloc((void**)&nrows, rbytes));
  cudaError_t e = cudaMalloc((void**)&nids, (size_t)capacity * sizeof(i64) + 256);
  if (e != cudaSuccess) { cudaFree(nrows); return ctx->fail_cuda(e, "cudaMalloc(ids)", __FILE__, __LINE__); }
  if (ctx->n > 0)
  {
    e = cudaMemcpyAsync(nrows, ctx->rows, (size_t)ctx->n * ctx->ld * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(nids, ctx->ids, (size_t)ctx->n * sizeof(i64), cudaMemcpyDeviceToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess)
    {
      cudaFree(nrows);  // the old store stays valid
      cudaFree(nids);
      return ctx->fail_cuda(e, "growing the point store", __FILE__, __LINE__);
    }
  }
  cudaFree(ctx->rows);
  cudaFree(ctx->ids);
  ctx->rows = nrows;
  ctx->ids = nids;
  ctx->capacity = capacity;
  return VI_OK;
}

int vi_points_reserve(vi_ctx* ctx, int64_t capacity, int32_t dims)
{
  if (!ctx) return VI_ERR_INVALID_ARG;
  if (capacity < 0 || dims <= 0 || dims > 32767)  // `short dimensions`, FileRangeStore.cs:18
    return ctx->fail(VI_ERR_INVALID_ARG, "Invalid capacity or dimensions.");
  VI_CUDA_TRY(cudaSetDevice(ctx->device));
  ctx->built = false;
  if (dims != ctx->dims)
  {
    // buffers are sized per dimension count: start over
    free_points(ctx);
    vi_free_table(ctx);
    vi_free_workspace(ctx);
    ctx->dims = dims;
    ctx->ld = (dims + 3) & ~3;
  }
  ctx->n = 0;  // drop the points; device buffers (points, work space, table) are kept and reused when large enough
  if (capacity == 0) return VI_OK;
  return reserve_points(ctx, capacity);
}

}  // extern "C"

// room for n more points (vi_table.cu's record ingest)
int vi_reserve_for_add(vi_ctx* ctx, int64_t n)
{
  if (ctx->n + n > ctx->capacity)
  {
    int64_t want = ctx->capacity * 2;
    if (want < ctx->n + n) want = ctx->n + n;
    return reserve_points(ctx, want);
  }
  return VI_OK;
}

extern "C" {

static int add_points(vi_ctx* ctx, const int64_t* ids, const float* rows, int64_t n, int32_t dims, cudaMemcpyKind kind)
{
  if (!ctx) return VI_ERR_INVALID_ARG;
  if (ctx->dims == 0) return ctx->fail(VI_ERR_STATE, "vi_points_reserve must be called first");
  if (dims != ctx->dims) return ctx->fail(VI_ERR_INVALID_ARG, "Invalid length of vector.");  // FileRangeStore.cs:59-64
  if (n < 0 || (n > 0 && (!ids || !rows))) return ctx->fail(VI_ERR_INVALID_ARG, "null points");
  if (n == 0) return VI_OK;
  VI_CUDA_TRY(cudaSetDevice(ctx->device));
  if (ctx->n + n > ctx->capacity)
  {
    int64_t want = ctx->capacity * 2;
    if (want < ctx->n + n) want = ctx->n + n;
    int rc = reserve_points(ctx, want);
    if (rc != VI_OK) return rc;
  }
  ctx->built = false;
  float* dst = ctx->rows + (size_t)ctx->n * ctx->ld;
  if (ctx->ld == dims)
    VI_CUDA_TRY(cudaMemcpyAsync(dst, rows, (size_t)n * dims * sizeof(float), kind, ctx->stream));
  else
  {
    VI_CUDA_TRY(cudaMemsetAsync(dst, 0, (size_t)n * ctx->ld * sizeof(float), ctx->stream));
    VI_CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)ctx->ld * sizeof(float), rows, (size_t)dims * sizeof(float),
                                  (size_t)dims * sizeof(float), (size_t)n, kind, ctx->stream));
  }
  VI_CUDA_TRY(cudaMemcpyAsync(ctx->ids + ctx->n, ids, (size_t)n * sizeof(i64), kind, ctx->stream));
  VI_CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // caller's buffers are reusable on return
  ctx->n += n;
  return VI_OK;
}

int vi_points_add(vi_ctx* ctx, const int64_t* ids, const float* rows, int64_t n, int32_t dims)
{
  return add_points(ctx, ids, rows, n, dims, cudaMemcpyHostToDevice);
}

int vi_points_add_device(vi_ctx* ctx, const int64_t* d_ids, const float* d_rows, int64_t n, int32_t dims)
{
  return add_points(ctx, d_ids, d_rows, n, dims, cudaMemcpyDeviceToDevice);
}

int vi_points_add_records(vi_ctx* ctx, const void* records, int64_t n, int32_t dims)
{
  if (!ctx) return VI_ERR_INVALID_ARG;
  if (ctx->dims == 0) return ctx->fail(VI_ERR_STATE, "vi_points_reserve must be called first");
  if (dims != ctx->dims) return ctx->fail(VI_ERR_INVALID_ARG, "Invalid length of vector.");  // FileRangeStore.cs:59-64
  if (n < 0 || (n > 0 && !records)) return ctx->fail(VI_ERR_INVALID_ARG, "null records");
  if (n == 0) return VI_OK;
  VI_CUDA_TRY(cudaSetDevice(ctx->device));
  ctx->built = false;
  return vi_points_add_records_impl(ctx, records, n);
}

int vi_points_add_file(vi_ctx* ctx, const char* path, int64_t offset_bytes, int64_t n, int32_t dims, double* read_ms,
                       double* total_ms)
{
  if (!ctx) return VI_ERR_INVALID_ARG;
  if (ctx->dims == 0) return ctx->fail(VI_ERR_STATE, "vi_points_reserve must be called first");
  if (dims != ctx->dims) return ctx->fail(VI_ERR_INVALID_ARG, "Invalid length of vector.");
  if (!path || offset_bytes < 0) return ctx->fail(VI_ERR_INVALID_ARG, "bad record file arguments");
  VI_CUDA_TRY(cudaSetDevice(ctx->device));
  ctx->built = false;
  return vi_points_add_file_impl(ctx, path, offset_bytes, n, read_ms, total_ms);
}

int64_t vi_points_count(const vi_ctx* ctx) { return ctx ? ctx->n : 0; }

int vi_build(vi_ctx* ctx, int32_t mode, vi_build_info* info)
{
  if (!ctx) return VI_ERR_INVALID_ARG;
  if (mode != VI_MODE_EXACT && mode != VI_MODE_FAST && mode != VI_MODE_SQL)
    return ctx->fail(VI_ERR_INVALID_ARG, "unknown build mode");
  if (ctx->dims == 0) return ctx->fail(VI_ERR_STATE, "no points: vi_points_reserve / vi_points_add first");
  VI_CUDA_TRY(cudaSetDevice(ctx->device));
  ctx->pending_nq = -1;
  int rc = vi_build_impl(ctx, mode);
  if (info) *info = ctx->info;
  return rc;
}

// vi_build + vi_ranges_copy in one call: the dense row blocks the build's last kernel finishes first travel to the host
// while it is still running (vi_build.cu run_subtrees / copy_out_rows); the rest follows when the build is done.
int vi_build_copy(vi_ctx* ctx, int32_t mode, vi_build_info* info, int64_t* range_id, int32_t* dimension, float* mid,
                  int64_t* id, int64_t cap, int64_t* rows)
{
  if (!ctx || !rows) return VI_ERR_INVALID_ARG;
  *rows = 0;
  if (mode != VI_MODE_EXACT && mode != VI_MODE_FAST && mode != VI_MODE_SQL)
    return ctx->fail(VI_ERR_INVALID_ARG, "unknown build mode");
  if (ctx->dims == 0) return ctx->fail(VI_ERR_STATE, "no points: vi_points_reserve / vi_points_add first");
  VI_CUDA_TRY(cudaSetDevice(ctx->device));
  ctx->pending_nq = -1;
  ctx->out = vi_ctx::CopyOut();
  ctx->out.rid = range_id;
  ctx->out.dim = dimension;
  ctx->out.mid = mid;
  ctx->out.id = id;
  ctx->out.cap = cap;
  ctx->out.active = ctx->world == 1 && cap > 0;
  int rc = vi_build_impl(ctx, mode);
  const vi_ctx::CopyOut o = ctx->out;
  ctx->out = vi_ctx::CopyOut();
  if (info) *info = ctx->info;
  cudaError_t e = cudaSuccess;
  if (rc == VI_OK)
  {
    const int64_t k = ctx->t_rows;
    *rows = k;
    if (cap < k) rc = ctx->fail(VI_ERR_CAPACITY, "range buffers too small");
    else if (k > 0)
    {
      // the build has synchronised its stream: what was not sent on the way goes now, [0, lo) and [hi, k)
      cudaStream_t cs = ctx->copy_stream ? ctx->copy_stream : ctx->stream;
      const int64_t lo = o.copied_hi > o.copied_lo ? o.copied_lo : k, hi = o.copied_hi > o.copied_lo ? o.copied_hi : k;
      const int64_t part[2][2] = {{0, lo}, {hi, k}};
      for (int i = 0; i < 2 && e == cudaSuccess; ++i)
      {
        const int64_t a = part[i][0];
        const size_t n = part[i][1] > a ? (size_t)(part[i][1] - a) : 0;
        if (!n) continue;
        if (range_id) e = cudaMemcpyAsync(range_id + a, ctx->t_rid + a, n * 8, cudaMemcpyDeviceToHost, cs);
        if (dimension && e == cudaSuccess) e = cudaMemcpyAsync(dimension + a, ctx->t_dim + a, n * 4, cudaMemcpyDeviceToHost, cs);
        if (mid && e == cudaSuccess) e = cudaMemcpyAsync(mid + a, ctx->t_mid + a, n * 4, cudaMemcpyDeviceToHost, cs);
        if (id && e == cudaSuccess) e = cudaMemcpyAsync(id + a, ctx->t_id + a, n * 8, cudaMemcpyDeviceToHost, cs);
      }
    }
  }
  // copies may be in flight into the caller's buffers whatever the outcome
  if (ctx->copy_stream)
  {
    const cudaError_t e2 = cudaStreamSynchronize(ctx->copy_stream);
    if (e == cudaSuccess) e = e2;
  }
  {
    const cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = e2;
  }
  if (e != cudaSuccess && rc == VI_OK) return ctx->fail_cuda(e, "vi_build_copy", __FILE__, __LINE__);
  return rc;
}

int vi_build_levels(const vi_ctx* ctx, vi_level_info* out, int32_t cap, int32_t* n)
{
  if (!ctx || !n) return VI_ERR_INVALID_ARG;
  *n = (int32_t)ctx->levels.size();
  if (out)
    for (int32_t i = 0; i < *n && i < cap; ++i) out[i] = ctx->levels[i];
  return VI_OK;
}

int64_t vi_range_count(const vi_ctx* ctx) { return (ctx && ctx->built) ? ctx->t_rows : 0; }

int vi_ranges_copy(const vi_ctx* cctx, int64_t* range_id, int32_t* dimension, float* mid, int64_t* id, int64_t cap)
{
  vi_ctx* ctx = const_cast<vi_ctx*>(cctx);
  if (!ctx) return VI_ERR_INVALID_ARG;
  if (!ctx->built) return ctx->fail(VI_ERR_STATE, "no built index");
  const int64_t k = ctx->t_rows;
  if (cap < k) return ctx->fail(VI_ERR_CAPACITY, "range buffers too small");
  if (k == 0) return VI_OK;
  VI_CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  if (range_id) VI_CUDA_TRY(cudaMemcpyAsync(range_id, ctx->t_rid, (size_t)k * 8, cudaMemcpyDeviceToHost, st));
  if (dimension) VI_CUDA_TRY(cudaMemcpyAsync(dimension, ctx->t_dim, (size_t)k * 4, cudaMemcpyDeviceToHost, st));
  if (mid) VI_CUDA_TRY(cudaMemcpyAsync(mid, ctx->t_mid, (size_t)k * 4, cudaMemcpyDeviceToHost, st));
  if (id) VI_CUDA_TRY(cudaMemcpyAsync(id, ctx->t_id, (size_t)k * 8, cudaMemcpyDeviceToHost, st));
  VI_CUDA_TRY(cudaStreamSynchronize(st));
  return VI_OK;
}

int vi_ranges_load(vi_ctx* ctx, const int64_t* range_id, const int32_t* dimension, const float* mid, const int64_t* id,
                   int64_t n, int32_t dims)
{
  if (!ctx) return VI_ERR_INVALID_ARG;
  if (dims <= 0 || dims > 32767) return ctx->fail(VI_ERR_INVALID_ARG, "Invalid capacity or dimensions.");
  if (ctx->dims != 0 && ctx->dims != dims) return ctx->fail(VI_ERR_INVALID_ARG, "Invalid vector size.");
  if (n < 0 || n >= (int64_t)0x7fffffff || (n > 0 && (!range_id || !dimension || !mid || !id)))
    return ctx->fail(VI_ERR_INVALID_ARG, "bad range table");
  if (ctx->world > 1) return ctx->fail(VI_ERR_STATE, "vi_ranges_load on a multi-rank context: load on every rank of a world of 1");
  VI_CUDA_TRY(cudaSetDevice(ctx->device));
  if (ctx->dims == 0)
  {
    ctx->dims = dims;
    ctx->ld = (dims + 3) & ~3;
  }
  ctx->pending_nq = -1;
  return vi_ranges_load_impl(ctx, range_id, dimension, mid, id, n);
}

int vi_textindex_copy(const vi_ctx* cctx, int64_t* range_id, int16_t* dimension, float* mid, int64_t* low_range_id,
                      int64_t* high_range_id, int64_t* text_id, int64_t cap)
{
  vi_ctx* ctx = const_cast<vi_ctx*>(cctx);
  if (!ctx) return VI_ERR_INVALID_ARG;
  if (!ctx->built) return ctx->fail(VI_ERR_STATE, "no built index");
  const int64_t k = ctx->t_rows;
  if (cap < k) return ctx->fail(VI_ERR_CAPACITY, "range buffers too small");
  if (k == 0) return VI_OK;
  std::vector<i64> rid((size_t)k), id((size_t)k);
  std::vector<int> dim((size_t)k), lo((size_t)k), hi((size_t)k);
  std::vector<float> m((size_t)k);
  int rc = vi_ranges_copy(ctx, rid.data(), dim.data(), m.data(), id.data(), k);
  if (rc != VI_OK) return rc;
  VI_CUDA_TRY(cudaMemcpy(lo.data(), ctx->t_low, (size_t)k * 4, cudaMemcpyDeviceToHost));
  VI_CUDA_TRY(cudaMemcpy(hi.data(), ctx->t_high, (size_t)k * 4, cudaMemcpyDeviceToHost));
  // a table built by VI_MODE_SQL names both children of every range of more than one point, present or not, as
  // dbo.BuildIndex does (DDL.sql:195-196 `iif(Count = 1, null, RangeID * 2 + 1)`)
  const bool sql = ctx->info.mode == VI_MODE_SQL;
  for (int64_t i = 0; i < k; ++i)
  {
    const bool leaf = dim[i] == -1;
    const bool null_dim = dim[i] < 0;  // leaves and Stdev = 0 rows (VI_DIM_NULL): DDL.sql:193-194
    if (range_id) range_id[i] = rid[i];
    if (dimension) dimension[i] = null_dim ? (int16_t)-1 : (int16_t)dim[i];
    if (mid) mid[i] = null_dim ? NAN : m[i];
    if (low_range_id) low_range_id[i] = (sql && !leaf) ? rid[i] * 2 + 1 : (lo[i] >= 0 ? rid[lo[i]] : -1);
    if (high_range_id) high_range_id[i] = (sql && !leaf) ? rid[i] * 2 + 2 : (hi[i] >= 0 ? rid[hi[i]] : -1);
    if (text_id) text_id[i] = leaf ? id[i] : -1;  // DDL.sql:195-197: internal rows carry no TextID
  }
  return VI_OK;
}

static int searchable(vi_ctx* ctx)
{
  if (!ctx->built) return ctx->fail(VI_ERR_STATE, "no built index");
  // after a multi-rank build a context holds the shared top rows plus its own sub-trees; the level-L rows of ranges
  // owned elsewhere are Dimension == -2 placeholders without Id / Mid / source row: not a searchable table
  if (ctx->world > 1 && !ctx->replicated)
    return ctx->fail(VI_ERR_STATE, "multi-rank build: call vi_table_replicate before searching (this rank holds only "
                                   "the sub-trees it owns)");
  return VI_OK;
}

int vi_search_device(vi_ctx* ctx, const float* d_queries, int64_t nq, int32_t dims, float proximity, int64_t* d_offsets,
                     int64_t* d_ids, int64_t cap, int64_t* total, int64_t* visits)
{
  if (!ctx || !total) return VI_ERR_INVALID_ARG;
  { const int rs = searchable(ctx); if (rs != VI_OK) return rs; }
  if (dims != ctx->dims) return ctx->fail(VI_ERR_INVALID_ARG, "Invalid vector size.");  // MemoryVectorIndex.cs:254
  if (nq < 0 || nq >= (int64_t)0x7fffffff || (nq > 0 && !d_queries) || !d_offsets)
    return ctx->fail(VI_ERR_INVALID_ARG, "bad query batch");
  VI_CUDA_TRY(cudaSetDevice(ctx->device));
  ctx->pending_nq = -1;  // a pending vi_search_begin shares the candidate pool with this call
  int* keep_src = ctx->search_src;
  ctx->search_src = nullptr;  // plain search does not record source rows
  int rc = vi_search_impl(ctx, d_queries, nq, proximity, d_offsets, d_ids, cap, total, visits, false);
  ctx->search_src = keep_src;
  return rc;
}

static int stage_queries(vi_ctx* ctx, const float* queries, int64_t nq)
{
  ctx->pending_nq = -1;  // the staged queries / offsets of a vi_search_begin are overwritten
  int rc = grow(ctx, &ctx->q_buf, &ctx->q_cap, nq * ctx->dims + 4);
  if (rc != VI_OK) return rc;
  rc = grow(ctx, &ctx->off_buf, &ctx->off_cap, nq + 2);
  if (rc != VI_OK) return rc;
  if (nq > 0)
    VI_CUDA_TRY(cudaMemcpyAsync(ctx->q_buf, queries, (size_t)nq * ctx->dims * sizeof(float), cudaMemcpyHostToDevice,
                                ctx->stream));
  return VI_OK;
}

int vi_search(vi_ctx* ctx, const float* queries, int64_t nq, int32_t dims, float proximity, int64_t* offsets,
              int64_t* ids, int64_t cap, int64_t* total)
{
  if (!ctx || !total || !offsets) return VI_ERR_INVALID_ARG;
  { const int rs = searchable(ctx); if (rs != VI_OK) return rs; }
  if (dims != ctx->dims) return ctx->fail(VI_ERR_INVALID_ARG, "Invalid vector size.");
  if (nq < 0 || nq >= (int64_t)0x7fffffff || (nq > 0 && !queries)) return ctx->fail(VI_ERR_INVALID_ARG, "bad query batch");
  VI_CUDA_TRY(cudaSetDevice(ctx->device));
  int rc = stage_queries(ctx, queries, nq);
  if (rc != VI_OK) return rc;
  int* keep_src = ctx->search_src;
  ctx->search_src = nullptr;
  // count pass
  rc = vi_search_impl(ctx, ctx->q_buf, nq, proximity, ctx->off_buf, nullptr, 0, total, nullptr, false);
  if (rc == VI_OK)
  {
    VI_CUDA_TRY(cudaMemcpyAsync(offsets, ctx->off_buf, (size_t)(nq + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    VI_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (ids)
    {
      if (cap < *total) rc = ctx->fail(VI_ERR_CAPACITY, "ids capacity smaller than the number of candidates");
      else if (*total > 0)
      {
        rc = grow(ctx, &ctx->ids_buf, &ctx->ids_cap, *total);
        if (rc == VI_OK) rc = vi_search_impl(ctx, ctx->q_buf, nq, proximity, ctx->off_buf, ctx->ids_buf, *total, total, nullptr, true);
        if (rc == VI_OK)
        {
          VI_CUDA_TRY(cudaMemcpyAsync(ids, ctx->ids_buf, (size_t)*total * 8, cudaMemcpyDeviceToHost, ctx->stream));
          VI_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        }
      }
    }
  }
  ctx->search_src = keep_src;
  return rc;
}

// Two-step form of vi_search that walks the table only twice in all: begin = count pass (the staged queries and the
// scanned offsets stay in the context), fetch = fill pass + copies.
int vi_search_begin(vi_ctx* ctx, const float* queries, int64_t nq, int32_t dims, float proximity, int64_t* total)
{
  if (!ctx || !total) return VI_ERR_INVALID_ARG;
  ctx->pending_nq = -1;
  { const int rs = searchable(ctx); if (rs != VI_OK) return rs; }
  if (dims != ctx->dims) return ctx->fail(VI_ERR_INVALID_ARG, "Invalid vector size.");
  if (nq < 0 || nq >= (int64_t)0x7fffffff || (nq > 0 && !queries)) return ctx->fail(VI_ERR_INVALID_ARG, "bad query batch");
  VI_CUDA_TRY(cudaSetDevice(ctx->device));
  int rc = stage_queries(ctx, queries, nq);
  if (rc != VI_OK) return rc;
  int* keep_src = ctx->search_src;
  ctx->search_src = nullptr;
  rc = vi_search_impl(ctx, ctx->q_buf, nq, proximity, ctx->off_buf, nullptr, 0, total, nullptr, false);
  ctx->search_src = keep_src;
  if (rc != VI_OK) return rc;
  ctx->pending_nq = nq;
  ctx->pending_total = *total;
  ctx->pending_prox = proximity;
  return VI_OK;
}

int vi_search_fetch(vi_ctx* ctx, int64_t* offsets, int64_t* ids, int64_t cap)
{
  if (!ctx || !offsets) return VI_ERR_INVALID_ARG;
  if (ctx->pending_nq < 0 || !ctx->built) return ctx->fail(VI_ERR_STATE, "vi_search_fetch without vi_search_begin");
  const int64_t nq = ctx->pending_nq, total = ctx->pending_total;
  if (total > 0 && (!ids || cap < total)) return ctx->fail(VI_ERR_CAPACITY, "ids capacity smaller than the number of candidates");
  VI_CUDA_TRY(cudaSetDevice(ctx->device));
  VI_CUDA_TRY(cudaMemcpyAsync(offsets, ctx->off_buf, (size_t)(nq + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  int rc = VI_OK;
  if (total > 0)
  {
    rc = grow(ctx, &ctx->ids_buf, &ctx->ids_cap, total);
    int* keep_src = ctx->search_src;
    ctx->search_src = nullptr;
    int64_t t2 = 0;
    if (rc == VI_OK) rc = vi_search_impl(ctx, ctx->q_buf, nq, ctx->pending_prox, ctx->off_buf, ctx->ids_buf, total, &t2, nullptr, true);
    ctx->search_src = keep_src;
    if (rc == VI_OK) VI_CUDA_TRY(cudaMemcpyAsync(ids, ctx->ids_buf, (size_t)total * 8, cudaMemcpyDeviceToHost, ctx->stream));
  }
  VI_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  ctx->pending_nq = -1;
  return rc;
}

int vi_search_topk(vi_ctx* ctx, const float* queries, int64_t nq, int32_t dims, float proximity, int32_t k, int32_t metric,
                   int64_t* ids, float* dist, int32_t* count, int64_t* candidates)
{
  if (!ctx) return VI_ERR_INVALID_ARG;
  { const int rs = searchable(ctx); if (rs != VI_OK) return rs; }
  if (dims != ctx->dims) return ctx->fail(VI_ERR_INVALID_ARG, "Invalid vector size.");
  if (nq < 0 || nq >= (int64_t)0x7fffffff || (nq > 0 && !queries)) return ctx->fail(VI_ERR_INVALID_ARG, "bad query batch");
  if (k <= 0 || k > 1024 || (metric != 0 && metric != 1)) return ctx->fail(VI_ERR_INVALID_ARG, "k must be 1..1024, metric 0 or 1");
  if (ctx->replicated || !ctx->src_rows)
    return ctx->fail(VI_ERR_STATE, "top-k needs the vectors: not on a replicated or imported table");
  VI_CUDA_TRY(cudaSetDevice(ctx->device));
  int rc = stage_queries(ctx, queries, nq);
  if (rc != VI_OK) return rc;
  return vi_search_topk_impl(ctx, ctx->q_buf, nq, proximity, k, metric, ids, dist, count, candidates);
}

int vi_search_verify(vi_ctx* ctx, const float* queries, int64_t nq, int32_t dims, float proximity, float distance,
                     int64_t* offsets, int64_t* ids, int64_t cap, int64_t* total)
{
  if (!ctx || !total || !offsets) return VI_ERR_INVALID_ARG;
  { const int rs = searchable(ctx); if (rs != VI_OK) return rs; }
  if (dims != ctx->dims) return ctx->fail(VI_ERR_INVALID_ARG, "Invalid vector size.");
  if (nq < 0 || nq >= (int64_t)0x7fffffff || (nq > 0 && !queries)) return ctx->fail(VI_ERR_INVALID_ARG, "bad query batch");
  VI_CUDA_TRY(cudaSetDevice(ctx->device));
  if (ctx->replicated || !ctx->src_rows)
    return ctx->fail(VI_ERR_STATE, "candidate verification needs the vectors: not on a replicated or imported table");
  int rc = stage_queries(ctx, queries, nq);
  if (rc != VI_OK) return rc;
  int64_t cand = 0;
  int* keep_src = ctx->search_src;
  ctx->search_src = nullptr;
  ctx->search_want_src = true;  // the fill pass records source rows: count only, no candidate pool
  rc = vi_search_impl(ctx, ctx->q_buf, nq, proximity, ctx->off_buf, nullptr, 0, &cand, nullptr, false);
  ctx->search_want_src = false;
  ctx->search_src = keep_src;
  if (rc != VI_OK) return rc;
  rc = grow(ctx, &ctx->ids_buf, &ctx->ids_cap, cand + 1);
  if (rc == VI_OK) rc = grow(ctx, &ctx->search_src, &ctx->src_cap, cand + 1);
  if (rc == VI_OK) rc = grow(ctx, &ctx->verify_keep, &ctx->keep_cap, cand + 2);
  if (rc == VI_OK) rc = grow(ctx, &ctx->off2_buf, &ctx->off2_cap, nq + 2);
  if (rc == VI_OK) rc = grow(ctx, &ctx->ids2_buf, &ctx->ids2_cap, cand + 1);
  if (rc != VI_OK) return rc;
  rc = vi_search_impl(ctx, ctx->q_buf, nq, proximity, ctx->off_buf, ctx->ids_buf, cand, &cand, nullptr, true);
  if (rc != VI_OK) return rc;
  rc = vi_verify_impl(ctx, ctx->q_buf, nq, distance, ctx->off_buf, ctx->ids_buf, cand, ctx->off2_buf,
                      ids ? ctx->ids2_buf : nullptr, total);
  if (rc != VI_OK) return rc;
  VI_CUDA_TRY(cudaMemcpyAsync(offsets, ctx->off2_buf, (size_t)(nq + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  VI_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  if (ids)
  {
    if (cap < *total) return ctx->fail(VI_ERR_CAPACITY, "ids capacity smaller than the number of matches");
    if (*total > 0)
    {
      VI_CUDA_TRY(cudaMemcpyAsync(ids, ctx->ids2_buf, (size_t)*total * 8, cudaMemcpyDeviceToHost, ctx->stream));
      VI_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
  }
  return VI_OK;
}

int vi_set_collective(vi_ctx* ctx, int32_t rank, int32_t world, vi_allreduce_u64_fn allreduce,
                      vi_alltoallv_fn alltoallv, void* user)
{
  if (!ctx) return VI_ERR_INVALID_ARG;
  if (world < 1 || world > 64 || rank < 0 || rank >= world || (world > 1 && (!allreduce || !alltoallv)))  // (2^L <= VI_SH_MAXR)
    return ctx->fail(VI_ERR_INVALID_ARG, "bad collective");
  vi_comm_release(ctx);  // the host's transport replaces a library-owned communicator
  ctx->rank = rank;
  ctx->world = world;
  ctx->allreduce = allreduce;
  ctx->alltoallv = alltoallv;
  ctx->coll_user = user;
  return VI_OK;
}

int vi_table_replicate(vi_ctx* ctx)
{
  if (!ctx) return VI_ERR_INVALID_ARG;
  if (!ctx->built) return ctx->fail(VI_ERR_STATE, "no built index");
  VI_CUDA_TRY(cudaSetDevice(ctx->device));
  return vi_table_replicate_impl(ctx);
}

int vi_shared_rows(const vi_ctx* ctx, int64_t* shared_rows)
{
  if (!ctx || !shared_rows) return VI_ERR_INVALID_ARG;
  *shared_rows = ctx->built ? ctx->shared_rows : 0;
  return VI_OK;
}

int vi_table_device(const vi_ctx* ctx, const int64_t** range_id, const int32_t** dimension, const float** mid,
                    const int64_t** id, const int32_t** low_row, const int32_t** high_row)
{
  if (!ctx) return VI_ERR_INVALID_ARG;
  if (!ctx->built) return VI_ERR_STATE;
  if (range_id) *range_id = ctx->t_rid;
  if (dimension) *dimension = ctx->t_dim;
  if (mid) *mid = ctx->t_mid;
  if (id) *id = ctx->t_id;
  if (low_row) *low_row = ctx->t_low;
  if (high_row) *high_row = ctx->t_high;
  return VI_OK;
}

void* vi_stream(const vi_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int vi_debug_divcheck(vi_ctx* ctx, uint64_t seed, int64_t samples, int64_t* mismatches)
{
  if (!ctx || !mismatches || samples < 0) return VI_ERR_INVALID_ARG;
  VI_CUDA_TRY(cudaSetDevice(ctx->device));
  return vi_debug_divcheck_impl(ctx, seed, samples, mismatches);
}

}  // extern "C"
