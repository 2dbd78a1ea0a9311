// vi_scan.cuh -- exclusive scans used by the partition pass (in place allowed; data[n] receives the total).
#pragma once
#include "vi_common.cuh"

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename T>
__device__ __forceinline__ T warp_inclusive_scan(T v)
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

template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile(const T* in, T* out, T* bsum, u32 n, T* total_if_single)
{
  __shared__ T wsum[SCAN_THREADS / 32];
  const u32 base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  T v[SCAN_ITEMS];
  T run = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i)
  {
    T t = (base + i < n) ? in[base + i] : T(0);
    v[i] = run;
    run += t;
  }
  T incl = warp_inclusive_scan(run);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  T woff = 0, total = 0;
#pragma unroll
  for (int w = 0; w < SCAN_THREADS / 32; ++w)
  {
    T x = wsum[w];
    if (w < warp) woff += x;
    total += x;
  }
  const T off = woff + incl - run;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i)
    if (base + i < n) out[base + i] = v[i] + off;
  if (threadIdx.x == 0)
  {
    bsum[blockIdx.x] = total;
    if (total_if_single) *total_if_single = total;  // single-tile scan: no second kernel needed
  }
}

template <typename T>
__global__ void __launch_bounds__(1024) k_scan_bsums(T* bsum, u32 nb, T* total_out)
{
  __shared__ T wsum[32];
  __shared__ T carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (u32 base = 0; base < nb; base += 1024)
  {
    const u32 i = base + threadIdx.x;
    T x = i < nb ? bsum[i] : T(0);
    T incl = warp_inclusive_scan(x);
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    T woff = 0, total = 0;
    for (int w = 0; w < 32; ++w)
    {
      T y = wsum[w];
      if (w < warp) woff += y;
      total += y;
    }
    const T carry = carry_s;
    if (i < nb) bsum[i] = carry + woff + incl - x;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = carry_s;
}

template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(T* out, const T* bsum, u32 n)
{
  const T off = bsum[blockIdx.x];
  const u32 base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i)
    if (base + i < n) out[base + i] += off;
}

template <typename T>
static void scan_exclusive(vi_ctx* ctx, T* data, u32 n, int64_t& launches)
{
  const u32 nb = (n + SCAN_TILE - 1) / SCAN_TILE;
  T* bsum = (T*)ctx->scan_tmp;
  if (nb == 0)
  {
    cudaMemsetAsync(data, 0, sizeof(T), ctx->stream);
    return;
  }
  if (nb == 1)
  {
    k_scan_tile<T><<<1, SCAN_THREADS, 0, ctx->stream>>>(data, data, bsum, n, data + n);
    ++launches;
    return;
  }
  k_scan_tile<T><<<nb, SCAN_THREADS, 0, ctx->stream>>>(data, data, bsum, n, nullptr);
  k_scan_bsums<T><<<1, 1024, 0, ctx->stream>>>(bsum, nb, data + n);
  k_scan_add<T><<<nb, SCAN_THREADS, 0, ctx->stream>>>(data, bsum, n);
  launches += 3;
}

// ---- in-kernel scans of the partition pass (vi_partition.cuh) ------------------------------------------------------
// Exclusive scan of a[0..m) in place by ONE CTA of NT threads, a[m] <- total.  Used by the last CTA of a kernel over
// the per-tile aggregates the other CTAs published (threadfence + ticket): loads bypass L1.
template <typename T, int NT>
__device__ __forceinline__ void cta_scan_inplace(T* a, u32 m)
{
  constexpr int ITEMS = 4;
  __shared__ T s_wsum[NT / 32];
  __shared__ T s_carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (u32 base = 0; base < m; base += NT * ITEMS)
  {
    const u32 i0 = base + threadIdx.x * ITEMS;
    T v[ITEMS];
    T run = 0;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k)
    {
      const T t = (i0 + k < m) ? __ldcg(a + i0 + k) : T(0);
      v[k] = run;
      run += t;
    }
    const T incl = warp_inclusive_scan(run);
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    T woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w)
    {
      const T y = s_wsum[w];
      if (w < warp) woff += y;
      total += y;
    }
    const T off = s_carry + woff + incl - run;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k)
      if (i0 + k < m) a[i0 + k] = v[k] + off;
    __syncthreads();
    if (threadIdx.x == 0) s_carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) a[m] = s_carry;
}
