// vi_search.cu -- batched proximity search over the device-resident range table.
//
// Replaces the recursive CTE dbo.Search (DDL.sql:234-295): from RangeID 0 follow LowRangeID iff
// Mid >= v[Dimension] - domain, HighRangeID iff Mid <= v[Dimension] + domain, emit TextID of every leaf reached.
// One thread walks one query with an explicit stack (depth <= 63 => at most 64 pending entries); rows are the
// packed 16-byte traversal rows built by k_pack_nodes.  Two passes (count, scan, fill) give a CSR result in the
// oracle's DFS order (low branch first).
//
// Candidate verification (the predicate half of Find, MemoryVectorIndex.cs:237-241,336-342): Euclidean distance
// with float32 accumulation in index order, as MemoryVectorIndexTests.cs:209-217, one thread per candidate so the
// summation order is the oracle's.
#include "vi_common.cuh"

template <typename T>
__device__ __forceinline__ T warp_incl(T v)
{
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1)
  {
    T t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

// single-CTA exclusive scan of an int64 array in place, a[n] <- total (nq is at most a few million)
__global__ void __launch_bounds__(1024) k_scan_i64(i64* a, u32 n)
{
  __shared__ i64 wsum[32];
  __shared__ i64 carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int ITEMS = 8;
  for (u32 base = 0; base < n; base += 1024 * ITEMS)
  {
    const u32 i0 = base + threadIdx.x * ITEMS;
    i64 v[ITEMS];
    i64 run = 0;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k)
    {
      i64 t = (i0 + k < n) ? a[i0 + k] : 0;
      v[k] = run;
      run += t;
    }
    i64 incl = warp_incl(run);
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    i64 woff = 0, total = 0;
    for (int w = 0; w < 32; ++w)
    {
      i64 y = wsum[w];
      if (w < warp) woff += y;
      total += y;
    }
    const i64 off = carry_s + woff + incl - run;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k)
      if (i0 + k < n) a[i0 + k] = v[k] + off;
    __syncthreads();
    if (threadIdx.x == 0) carry_s += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) a[n] = carry_s;
}

template <bool FILL>
__global__ void __launch_bounds__(128)
k_search(const int4* __restrict__ node, const int* __restrict__ t_src, const float* __restrict__ queries, int ldq,
         u32 nq, float prox, i64* __restrict__ offsets, i64* __restrict__ ids_out, int* __restrict__ src_out,
         i64 cap, unsigned long long* __restrict__ visits)
{
  const u32 i = blockIdx.x * 128u + threadIdx.x;
  unsigned long long v = 0;
  if (i < nq)
  {
    const float* q = queries + (size_t)i * ldq;
    u32 stack[64];
    int sp = 0;
    stack[sp++] = 0u;
    i64 cnt = 0;
    const i64 base = FILL ? offsets[i] : 0;
    while (sp > 0)
    {
      const u32 r = stack[--sp];
      const int4 nd = __ldg(node + r);
      ++v;
      if (nd.x < 0)
      {
        // leaf: TextID = RangeValue.Id (DDL.sql:295 "where TextID is not null")
        if (FILL && base + cnt < cap)
        {
          ids_out[base + cnt] = (i64)(((u64)(u32)nd.w << 32) | (u64)(u32)nd.z);
          if (src_out) src_out[base + cnt] = t_src[r];
        }
        ++cnt;
        continue;
      }
      const float x = __ldg(q + nd.x);
      const float lo = __fsub_rn(x, prox);  // MinValue = value - @domain, DDL.sql:249
      const float hi = __fadd_rn(x, prox);  // MaxValue = value + @domain, DDL.sql:250
      const float mid = __int_as_float(nd.y);
      if (mid <= hi && nd.w >= 0) stack[sp++] = (u32)nd.w;  // DDL.sql:280-293 (pushed first, visited second)
      if (mid >= lo && nd.z >= 0) stack[sp++] = (u32)nd.z;  // DDL.sql:265-278
    }
    if (!FILL) offsets[i] = cnt;
  }
  if (!FILL && visits)
  {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(visits, v);
  }
}

// have_offsets: d_offsets already holds the scanned counts of this very batch (skip the count pass)
int vi_search_impl(vi_ctx* ctx, const float* d_queries, int64_t nq, float proximity, i64* d_offsets, i64* d_ids,
                   int64_t cap, int64_t* total, int64_t* visits, bool have_offsets)
{
  cudaStream_t st = ctx->stream;
  *total = 0;
  if (visits) *visits = 0;
  if (nq == 0)
  {
    VI_CUDA_TRY(cudaMemsetAsync(d_offsets, 0, sizeof(i64), st));
    VI_CUDA_TRY(cudaStreamSynchronize(st));
    return VI_OK;
  }
  if (ctx->t_rows == 0)
  {
    VI_CUDA_TRY(cudaMemsetAsync(d_offsets, 0, sizeof(i64) * (size_t)(nq + 1), st));
    VI_CUDA_TRY(cudaStreamSynchronize(st));
    return VI_OK;
  }
  unsigned long long* d_vis = (unsigned long long*)ctx->counters + 1;  // counters[2..3]
  const u32 grid = (u32)((nq + 127) / 128);
  i64 tot = 0;
  unsigned long long vis = 0;
  if (!have_offsets)
  {
    VI_CUDA_TRY(cudaMemsetAsync(d_vis, 0, 8, st));
    k_search<false><<<grid, 128, 0, st>>>(ctx->t_node, ctx->t_src, d_queries, ctx->dims, (u32)nq, proximity, d_offsets,
                                          nullptr, nullptr, 0, d_vis);
    k_scan_i64<<<1, 1024, 0, st>>>(d_offsets, (u32)nq);
    VI_CUDA_TRY(cudaMemcpyAsync(&vis, d_vis, 8, cudaMemcpyDeviceToHost, st));
  }
  VI_CUDA_TRY(cudaMemcpyAsync(&tot, d_offsets + nq, sizeof(i64), cudaMemcpyDeviceToHost, st));
  VI_CUDA_TRY(cudaStreamSynchronize(st));
  *total = tot;
  if (visits && !have_offsets) *visits = (int64_t)vis;
  if (d_ids == nullptr) return VI_OK;
  if (cap < tot) return ctx->fail(VI_ERR_CAPACITY, "ids capacity smaller than the number of candidates");
  if (tot > 0)
  {
    k_search<true><<<grid, 128, 0, st>>>(ctx->t_node, ctx->t_src, d_queries, ctx->dims, (u32)nq, proximity, d_offsets,
                                         d_ids, ctx->search_src, cap, nullptr);
    VI_CUDA_TRY(cudaStreamSynchronize(st));
  }
  VI_CUDA_TRY(cudaGetLastError());
  return VI_OK;
}

// ---- verification ------------------------------------------------------------------------------------------------
// one thread per candidate: float32 sum of squares in index order, sqrt, compare (MemoryVectorIndexTests.cs:209-217)
__global__ void __launch_bounds__(128)
k_verify_flags(const float* __restrict__ rows, int ld, int dims, const float* __restrict__ queries, int ldq,
               const i64* __restrict__ offsets, u32 nq, const int* __restrict__ src, i64 total, float distance,
               u32* __restrict__ keep)
{
  const i64 c = (i64)blockIdx.x * 128 + threadIdx.x;
  if (c >= total) return;
  // owning query: largest q with offsets[q] <= c
  u32 lo = 0, hi = nq;
  while (hi - lo > 1)
  {
    const u32 m = (lo + hi) >> 1;
    if (offsets[m] <= c) lo = m; else hi = m;
  }
  const float* a = rows + (size_t)src[c] * ld;
  const float* b = queries + (size_t)lo * ldq;
  float s = 0.f;
  for (int i = 0; i < dims; ++i)
  {
    const float t = __fsub_rn(a[i], b[i]);
    s = __fadd_rn(s, __fmul_rn(t, t));
  }
  keep[c] = (__fsqrt_rn(s) <= distance) ? 1u : 0u;
}

__global__ void __launch_bounds__(1024) k_scan_u32_single(u32* a, u32 n)
{
  __shared__ u32 wsum[32];
  __shared__ u32 carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int ITEMS = 8;
  for (u32 base = 0; base < n; base += 1024 * ITEMS)
  {
    const u32 i0 = base + threadIdx.x * ITEMS;
    u32 v[ITEMS];
    u32 run = 0;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k)
    {
      u32 t = (i0 + k < n) ? a[i0 + k] : 0;
      v[k] = run;
      run += t;
    }
    u32 incl = warp_incl(run);
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    u32 woff = 0, total = 0;
    for (int w = 0; w < 32; ++w)
    {
      u32 y = wsum[w];
      if (w < warp) woff += y;
      total += y;
    }
    const u32 off = carry_s + woff + incl - run;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k)
      if (i0 + k < n) a[i0 + k] = v[k] + off;
    __syncthreads();
    if (threadIdx.x == 0) carry_s += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) a[n] = carry_s;
}

__global__ void __launch_bounds__(256)
k_verify_compact(const i64* __restrict__ ids_in, const u32* __restrict__ keep_scan, i64 total, i64* __restrict__ ids_out,
                 i64 cap)
{
  const i64 c = (i64)blockIdx.x * 256 + threadIdx.x;
  if (c >= total) return;
  const u32 a = keep_scan[c], b = keep_scan[c + 1];
  if (b != a && (i64)a < cap) ids_out[a] = ids_in[c];
}

__global__ void __launch_bounds__(256)
k_verify_offsets(const i64* __restrict__ offsets_in, const u32* __restrict__ keep_scan, u32 nq, i64* __restrict__ offsets_out)
{
  const u32 q = blockIdx.x * 256u + threadIdx.x;
  if (q > nq) return;
  offsets_out[q] = (i64)keep_scan[offsets_in[q]];
}

// d_offsets_in/d_ids_in/src: result of a traversal with src rows; writes the filtered CSR.
int vi_verify_impl(vi_ctx* ctx, const float* d_queries, int64_t nq, float distance, const i64* d_offsets_in,
                   const i64* d_ids_in, int64_t total_in, i64* d_offsets_out, i64* d_ids_out, int64_t* total_out)
{
  cudaStream_t st = ctx->stream;
  *total_out = 0;
  if (total_in == 0)
  {
    VI_CUDA_TRY(cudaMemsetAsync(d_offsets_out, 0, sizeof(i64) * (size_t)(nq + 1), st));
    VI_CUDA_TRY(cudaStreamSynchronize(st));
    return VI_OK;
  }
  if (total_in >= (int64_t)0xfffffff0) return ctx->fail(VI_ERR_CAPACITY, "too many candidates to verify in one call");
  u32* keep = ctx->verify_keep;
  k_verify_flags<<<(u32)((total_in + 127) / 128), 128, 0, st>>>(ctx->src_rows, ctx->ld, ctx->dims, d_queries, ctx->dims,
                                                                d_offsets_in, (u32)nq, ctx->search_src, total_in, distance,
                                                                keep);
  k_scan_u32_single<<<1, 1024, 0, st>>>(keep, (u32)total_in);
  u32 kept = 0;
  VI_CUDA_TRY(cudaMemcpyAsync(&kept, keep + total_in, 4, cudaMemcpyDeviceToHost, st));
  k_verify_offsets<<<(u32)((nq + 1 + 255) / 256), 256, 0, st>>>(d_offsets_in, keep, (u32)nq, d_offsets_out);
  VI_CUDA_TRY(cudaStreamSynchronize(st));
  *total_out = kept;
  if (d_ids_out)
  {
    k_verify_compact<<<(u32)((total_in + 255) / 256), 256, 0, st>>>(d_ids_in, keep, total_in, d_ids_out, (i64)kept);
    VI_CUDA_TRY(cudaStreamSynchronize(st));
  }
  VI_CUDA_TRY(cudaGetLastError());
  return VI_OK;
}
