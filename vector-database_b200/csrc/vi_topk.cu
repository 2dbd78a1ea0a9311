// vi_topk.cu -- top-k over the candidates of dbo.Search: the quality layer the reference's README
// aims at (README.md:102-103) on top of its own candidate generator.  The traversal is vi_search.cu's; here every
// candidate gets its distance to the query (float32 accumulation in index order, as the reference's test helper does,
// MemoryVectorIndexTests.cs:209-217) and each query keeps its k nearest, ties in traversal order.
//   metric 0: Euclidean  sqrt(sum (a-b)^2)            metric 1: angular  1 - a.b / (|a| |b|)
#include <cub/device/device_segmented_sort.cuh>

#include "vi_common.cuh"

int vi_search_impl(vi_ctx* ctx, const float* d_queries, int64_t nq, float proximity, i64* d_offsets, i64* d_ids,
                   int64_t cap, int64_t* total, int64_t* visits, bool have_offsets);

// one thread per candidate; the summation order is the oracle's
__global__ void __launch_bounds__(128)
k_candidate_distance(const float* __restrict__ rows, int ld, int dims, const float* __restrict__ queries, int ldq,
                     const i64* __restrict__ offsets, u32 nq, const int* __restrict__ src, i64 total, int metric,
                     float* __restrict__ dist, u32* __restrict__ index)
{
  const i64 c = (i64)blockIdx.x * 128 + threadIdx.x;
  if (c >= total) return;
  u32 lo = 0, hi = nq;  // owning query: largest q with offsets[q] <= c
  while (hi - lo > 1)
  {
    const u32 m = (lo + hi) >> 1;
    if (offsets[m] <= c) lo = m; else hi = m;
  }
  const float* a = rows + (size_t)src[c] * ld;
  const float* b = queries + (size_t)lo * ldq;
  float r;
  if (metric == 0)
  {
    float s = 0.f;
    for (int i = 0; i < dims; ++i)
    {
      const float t = __fsub_rn(a[i], b[i]);
      s = __fadd_rn(s, __fmul_rn(t, t));
    }
    r = __fsqrt_rn(s);
  }
  else
  {
    float dot = 0.f, na = 0.f, nb = 0.f;
    for (int i = 0; i < dims; ++i)
    {
      dot = __fadd_rn(dot, __fmul_rn(a[i], b[i]));
      na = __fadd_rn(na, __fmul_rn(a[i], a[i]));
      nb = __fadd_rn(nb, __fmul_rn(b[i], b[i]));
    }
    r = __fsub_rn(1.f, __fdiv_rn(dot, __fmul_rn(__fsqrt_rn(na), __fsqrt_rn(nb))));
  }
  dist[c] = r;
  index[c] = (u32)c;
}

__global__ void __launch_bounds__(256)
k_take_topk(const i64* __restrict__ offsets, u32 nq, int k, const float* __restrict__ dist_sorted,
            const u32* __restrict__ index_sorted, const i64* __restrict__ cand_ids, i64* __restrict__ out_ids,
            float* __restrict__ out_dist, int* __restrict__ out_count)
{
  const u64 t = (u64)blockIdx.x * 256 + threadIdx.x;
  const u32 q = (u32)(t / (u32)k);
  const int j = (int)(t % (u32)k);
  if (q >= nq) return;
  const i64 o = offsets[q], cnt = offsets[q + 1] - o;
  if (j == 0) out_count[q] = (int)(cnt < k ? cnt : k);
  if (j < cnt)
  {
    out_ids[t] = cand_ids[index_sorted[o + j]];
    out_dist[t] = dist_sorted[o + j];
  }
  else
  {
    out_ids[t] = -1;
    out_dist[t] = __int_as_float(0x7f800000);  // +inf
  }
}

template <typename T>
static cudaError_t ensure(T** p, int64_t* cap, int64_t need)
{
  if (*cap >= need && *p) return cudaSuccess;
  cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  const int64_t ncap = need + need / 4 + 256;
  cudaError_t e = cudaMalloc((void**)p, (size_t)ncap * sizeof(T));
  if (e == cudaSuccess) *cap = ncap;
  return e;
}

// d_queries: nq x dims on the device.  Host outputs ids[nq*k], dist[nq*k], count[nq].
int vi_search_topk_impl(vi_ctx* ctx, const float* d_queries, int64_t nq, float proximity, int32_t k, int32_t metric,
                        int64_t* ids, float* dist, int32_t* count, int64_t* candidates)
{
  cudaStream_t st = ctx->stream;
  if (candidates) *candidates = 0;
  if (nq == 0) return VI_OK;
  VI_CUDA_TRY(ensure(&ctx->off_buf, &ctx->off_cap, nq + 2));
  int64_t cand = 0;
  int* keep_src = ctx->search_src;
  ctx->search_src = nullptr;
  ctx->search_want_src = true;  // the fill pass records source rows: count only, no candidate pool
  int rc = vi_search_impl(ctx, d_queries, nq, proximity, ctx->off_buf, nullptr, 0, &cand, nullptr, false);
  ctx->search_want_src = false;
  ctx->search_src = keep_src;
  if (rc != VI_OK) return rc;
  if (cand >= (int64_t)0x7fffffff) return ctx->fail(VI_ERR_CAPACITY, "too many candidates for one top-k call: split the batch");
  if (candidates) *candidates = cand;
  // scratch: candidate ids + source rows (traversal), distances + indexes and their sorted copies, outputs
  float *d_dist = nullptr, *d_dist_s = nullptr, *d_out_dist = nullptr;
  u32 *d_idx = nullptr, *d_idx_s = nullptr;
  i64* d_out_ids = nullptr;
  int* d_out_cnt = nullptr;
  void* d_tmp = nullptr;
  auto cleanup = [&]()
  {
    cudaFree(d_dist); cudaFree(d_dist_s); cudaFree(d_out_dist); cudaFree(d_idx); cudaFree(d_idx_s);
    cudaFree(d_out_ids); cudaFree(d_out_cnt); cudaFree(d_tmp);
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
  const size_t nc = (size_t)cand + 1;
  TRY_OR_CLEAN(ensure(&ctx->ids_buf, &ctx->ids_cap, cand + 1));
  TRY_OR_CLEAN(ensure(&ctx->search_src, &ctx->src_cap, cand + 1));
  TRY_OR_CLEAN(cudaMalloc((void**)&d_dist, nc * 4));
  TRY_OR_CLEAN(cudaMalloc((void**)&d_dist_s, nc * 4));
  TRY_OR_CLEAN(cudaMalloc((void**)&d_idx, nc * 4));
  TRY_OR_CLEAN(cudaMalloc((void**)&d_idx_s, nc * 4));
  TRY_OR_CLEAN(cudaMalloc((void**)&d_out_ids, (size_t)nq * k * 8));
  TRY_OR_CLEAN(cudaMalloc((void**)&d_out_dist, (size_t)nq * k * 4));
  TRY_OR_CLEAN(cudaMalloc((void**)&d_out_cnt, (size_t)nq * 4));
  if (cand > 0)
  {
    rc = vi_search_impl(ctx, d_queries, nq, proximity, ctx->off_buf, ctx->ids_buf, cand, &cand, nullptr, true);
    if (rc != VI_OK) { cleanup(); return rc; }
    k_candidate_distance<<<(u32)((cand + 127) / 128), 128, 0, st>>>(ctx->src_rows, ctx->ld, ctx->dims,
                                                                    d_queries, ctx->dims, ctx->off_buf, (u32)nq,
                                                                    ctx->search_src, cand, metric, d_dist, d_idx);
    size_t tmp_bytes = 0;
    TRY_OR_CLEAN(cub::DeviceSegmentedSort::StableSortPairs(nullptr, tmp_bytes, d_dist, d_dist_s, d_idx, d_idx_s, (int)cand,
                                                           (int)nq, ctx->off_buf, ctx->off_buf + 1, st));
    TRY_OR_CLEAN(cudaMalloc(&d_tmp, tmp_bytes + 16));
    TRY_OR_CLEAN(cub::DeviceSegmentedSort::StableSortPairs(d_tmp, tmp_bytes, d_dist, d_dist_s, d_idx, d_idx_s, (int)cand,
                                                           (int)nq, ctx->off_buf, ctx->off_buf + 1, st));
  }
  const u64 threads = (u64)nq * (u64)k;
  k_take_topk<<<(u32)((threads + 255) / 256), 256, 0, st>>>(ctx->off_buf, (u32)nq, k, d_dist_s, d_idx_s, ctx->ids_buf, d_out_ids,
                                                           d_out_dist, d_out_cnt);
  if (ids) TRY_OR_CLEAN(cudaMemcpyAsync(ids, d_out_ids, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, st));
  if (dist) TRY_OR_CLEAN(cudaMemcpyAsync(dist, d_out_dist, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, st));
  if (count) TRY_OR_CLEAN(cudaMemcpyAsync(count, d_out_cnt, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
  TRY_OR_CLEAN(cudaStreamSynchronize(st));
  TRY_OR_CLEAN(cudaGetLastError());
#undef TRY_OR_CLEAN
  cleanup();
  return VI_OK;
}
