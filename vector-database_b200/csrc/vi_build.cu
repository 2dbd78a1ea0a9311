// vi_build.cu -- level-synchronous split-tree builder (replaces the sequential walker IndexBuilder.Build,
// VectorIndex/IndexBuilder.cs:23-157).  One pass per tree level over every open range at once:
//
//   statistics   IndexBuilder.cs:55-68,159-197  ->  k_stats_small_* / k_stats_big_* (+ k_finalize_big_*)
//   split choice IndexBuilder.cs:75-88          ->  fused into the statistics kernels (warp/team arg-max)
//   partition    IndexBuilder.cs:99-129         ->  k_flags + scan + k_seg_children + scans + k_emit_children
//                                                    + k_scatter (stable: order inside low and high is kept)
//
// Points never move: a range is a contiguous slice [start, start+count) of a position space; perm[] maps a
// position to its row, pid[] carries the point id along.  Leaves (count == 1) get their row written when
// their parent is partitioned and drop out of the position space (it is compacted every level).
#include <math.h>

#include "vi_common.cuh"

// =============================================================================================================
// scans (exclusive, in place allowed; out[n] receives the total)
// =============================================================================================================
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
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile(const T* in, T* out, T* bsum, u32 n)
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
  if (threadIdx.x == 0) bsum[blockIdx.x] = total;
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
  // in place; data[n] <- total
  const u32 nb = (n + SCAN_TILE - 1) / SCAN_TILE;
  T* bsum = (T*)ctx->scan_tmp;
  if (nb == 0)
  {
    cudaMemsetAsync(data, 0, sizeof(T), ctx->stream);
    return;
  }
  k_scan_tile<T><<<nb, SCAN_THREADS, 0, ctx->stream>>>(data, data, bsum, n);
  k_scan_bsums<T><<<1, 1024, 0, ctx->stream>>>(bsum, nb, data + n);
  launches += 2;
  if (nb > 1)
  {
    k_scan_add<T><<<nb, SCAN_THREADS, 0, ctx->stream>>>(data, bsum, n);
    ++launches;
  }
}

// =============================================================================================================
// level 0 set-up
// =============================================================================================================
__global__ void k_init_level0(u32* perm, i64* pid, const i64* ids, u32* seg_of, u32 n, SegLevel sg, u32* big_list,
                              i64* t_rid, int* t_low, int* t_high)
{
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
  {
    perm[i] = i;
    pid[i] = ids[i];
    seg_of[i] = 0;
  }
  if (i == 0)
  {
    sg.start[0] = 0;
    sg.count[0] = n;
    sg.rid[0] = 0;
    sg.row[0] = 0;
    big_list[0] = 0;
    t_rid[0] = 0;
    t_low[0] = -1;
    t_high[0] = -1;
  }
}

__global__ void k_single_point(const i64* ids, i64* t_rid, int* t_dim, float* t_mid, i64* t_id, int* t_low, int* t_high,
                               int* t_src)
{
  t_src[0] = 0;
  // IndexBuilder.cs:81-82: count == 1 -> Dimension -1, Mid default 0, Id = the point id
  t_rid[0] = 0;
  t_dim[0] = -1;
  t_mid[0] = 0.0f;
  t_id[0] = ids[0];
  t_low[0] = -1;
  t_high[0] = -1;
}

__global__ void k_absmax(const float4* __restrict__ rows, size_t n4, u32* out)
{
  float m = 0.0f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
  {
    float4 v = ldg_f4_stream(rows + i);
    float a = fabsf(v.x), b = fabsf(v.y), c = fabsf(v.z), d = fabsf(v.w);
    if (a > m) m = a;  // NaN never wins
    if (b > m) m = b;
    if (c > m) m = c;
    if (d > m) m = d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));  // non-negative floats order like their bits
}

// =============================================================================================================
// shared pieces of the statistics kernels
// =============================================================================================================
struct StatsOut
{
  int* t_dim;
  float* t_mid;
  i64* t_id;
};

__device__ __forceinline__ void write_split(const SegLevel& sg, const StatsOut& o, u32 s, int dim, float mid, i64 pivot)
{
  const u32 row = sg.row[s];
  o.t_dim[row] = dim;
  o.t_mid[row] = mid;
  o.t_id[row] = pivot;
  sg.dim[s] = dim;
  sg.mid[s] = mid;
  sg.pivot[s] = pivot;
}

// (long)(IdN / Count), Int128 division truncating toward zero (IndexBuilder.cs:87)
__device__ __forceinline__ i64 mean_id(u64 slo, i64 shi, u32 n)
{
  // IdN = shi * 2^32 + slo  (slo = sum of low 32-bit halves, shi = sum of arithmetic-shifted high halves)
  const i128 idn = ((i128)shi << 32) + (i128)slo;
  if (idn >= (i128)INT64_MIN && idn <= (i128)INT64_MAX) return (i64)idn / (i64)n;
  return (i64)(idn / (i128)n);
}

// ---- the literal float32 recurrence (IndexBuilder.cs:159-197), one chain per (range, dim) ----------------------
__device__ __forceinline__ void welford_step(float& mean, float& q, float value, float c)
{
  // var a = pa + (value - pa) / count;  var q = pq + (value - pa) * (value - a);   IndexBuilder.cs:186-187
  const float d1 = __fsub_rn(value, mean);
  const float a = __fadd_rn(mean, __fdiv_rn(d1, c));
  q = __fadd_rn(q, __fmul_rn(d1, __fsub_rn(value, a)));
  mean = a;
}

struct ExBest
{
  float key;
  float mean;
  int idx;
};

__device__ __forceinline__ bool ex_better(float k, int i, float bk, int bi)
{
  if (i == 0x7fffffff) return false;
  if (bi == 0x7fffffff) return true;
  const int c = cmp_float_dotnet(k, bk);
  if (c != 0) return c > 0;
  return i < bi;
}

template <int W>
__device__ __forceinline__ ExBest ex_reduce_w(ExBest b, u32 mask)
{
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1)
  {
    ExBest t;
    t.key = __shfl_xor_sync(mask, b.key, o);
    t.mean = __shfl_xor_sync(mask, b.mean, o);
    t.idx = __shfl_xor_sync(mask, b.idx, o);
    if (ex_better(t.key, t.idx, b.key, b.idx)) b = t;
  }
  return b;
}

__device__ __forceinline__ ExBest ex_reduce(ExBest b) { return ex_reduce_w<32>(b, 0xffffffffu); }

// Fast-mode fallback for a poorly resolved range (no dimension spreads over 2^10 quantisation steps): the
// reference's own float32 statistics, computed by the team that owns the range, rows in position order.
// Lane tl owns the float4 column chunks c0 + k*TS + tl.  pp = perm + start of the range.
constexpr int VI_Q30_MIN_RES_BITS = 10;

template <int TS, int CH>
__device__ __noinline__ ExBest welford_team(const float* __restrict__ rows, int ld, int dims, const u32* __restrict__ pp,
                                            u32 n, int tl, u32 tmask, bool mx)
{
  ExBest best;
  best.key = 0.f;
  best.mean = 0.f;
  best.idx = 0x7fffffff;
  const int C4 = ld >> 2;
  for (int c0 = 0; c0 < C4; c0 += TS * CH)
  {
    float mean[CH * 4], q[CH * 4];
    {
      const float4* rp = reinterpret_cast<const float4*>(rows + (size_t)pp[0] * ld);
#pragma unroll
      for (int k = 0; k < CH; ++k)
      {
        const int c = c0 + k * TS + tl;
        const float4 x = (c < C4) ? rp[c] : make_float4(0.f, 0.f, 0.f, 0.f);
        mean[k * 4 + 0] = x.x; mean[k * 4 + 1] = x.y; mean[k * 4 + 2] = x.z; mean[k * 4 + 3] = x.w;
        q[k * 4 + 0] = 0.f; q[k * 4 + 1] = 0.f; q[k * 4 + 2] = 0.f; q[k * 4 + 3] = 0.f;
      }
    }
    for (u32 j = 1; j < n; ++j)
    {
      const float4* rp = reinterpret_cast<const float4*>(rows + (size_t)pp[j] * ld);
      const float cnt = (float)(j + 1u);
#pragma unroll
      for (int k = 0; k < CH; ++k)
      {
        const int c = c0 + k * TS + tl;
        const float4 x = (c < C4) ? rp[c] : make_float4(0.f, 0.f, 0.f, 0.f);
        welford_step(mean[k * 4 + 0], q[k * 4 + 0], x.x, cnt);
        welford_step(mean[k * 4 + 1], q[k * 4 + 1], x.y, cnt);
        welford_step(mean[k * 4 + 2], q[k * 4 + 2], x.z, cnt);
        welford_step(mean[k * 4 + 3], q[k * 4 + 3], x.w, cnt);
      }
    }
#pragma unroll
    for (int k = 0; k < CH; ++k)
#pragma unroll
      for (int e = 0; e < 4; ++e)
      {
        const int d = (c0 + k * TS + tl) * 4 + e;
        if (d < dims)
        {
          const float key = mx ? q[k * 4 + e] : -q[k * 4 + e];
          if (ex_better(key, d, best.key, best.idx))
          {
            best.key = key;
            best.mean = mean[k * 4 + e];
            best.idx = d;
          }
        }
      }
  }
  return ex_reduce_w<TS>(best, tmask);
}

// ---- fast mode (q30): exact integer sums of xi = rint(x * 2^(30-E)) ----------------------------------------
__device__ __forceinline__ void q30_acc(u64& s1, u64& s2lo, u32& s2hi, float x, float k)
{
  const int xi = __float2int_rn(__fmul_rn(x, k));
  s1 += (u64)(i64)xi;
  const u64 sq = (u64)((i64)xi * (i64)xi);
  const u64 t = s2lo + sq;
  s2hi += (t < sq) ? 1u : 0u;
  s2lo = t;
}

struct Q30Best
{
  u128 key;
  i64 s1;
  int idx;  // INT_MAX = none
};

__device__ __forceinline__ bool q30_better(bool mx, u128 k, int i, u128 bk, int bi)
{
  if (i == 0x7fffffff) return false;
  if (bi == 0x7fffffff) return true;
  if (k != bk) return mx ? (k > bk) : (k < bk);
  return i < bi;  // lowest index wins ties (MaxBy keeps the first maximum)
}

__device__ __forceinline__ u128 q30_key(u32 n, i64 s1, u64 s2lo, u64 s2hi)
{
  const u128 s2 = ((u128)s2hi << 64) | (u128)s2lo;
  const i128 a = (i128)s1;
  return (u128)n * s2 - (u128)(a * a);  // n * S2 - S1^2 >= 0 (Cauchy-Schwarz), exact
}

template <int W>
__device__ __forceinline__ Q30Best q30_reduce(Q30Best b, bool mx, u32 mask)
{
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1)
  {
    Q30Best t;
    u32 k0 = (u32)b.key, k1 = (u32)(b.key >> 32), k2 = (u32)(b.key >> 64), k3 = (u32)(b.key >> 96);
    k0 = __shfl_xor_sync(mask, k0, o);
    k1 = __shfl_xor_sync(mask, k1, o);
    k2 = __shfl_xor_sync(mask, k2, o);
    k3 = __shfl_xor_sync(mask, k3, o);
    t.key = ((u128)k3 << 96) | ((u128)k2 << 64) | ((u128)k1 << 32) | (u128)k0;
    t.s1 = __shfl_xor_sync(mask, b.s1, o);
    t.idx = __shfl_xor_sync(mask, b.idx, o);
    if (q30_better(mx, t.key, t.idx, b.key, b.idx)) b = t;
  }
  return b;
}

__device__ __forceinline__ float q30_mid(i64 s1, u32 n, double qinv)
{
  return __double2float_rn(__dmul_rn(__ddiv_rn(__ll2double_rn(s1), (double)n), qinv));
}

// One team of TS lanes owns one small range; a lane owns CH float4 column chunks (TS*CH*4 dims per pass).
template <int TS, int CH>
__global__ void __launch_bounds__(256)
k_stats_small_q30(SegLevel sg, u32 R, const u32* __restrict__ perm, const i64* __restrict__ pid,
                  const float* __restrict__ rows, int ld, int dims, float qk, double qinv, int mx, StatsOut out)
{
  const int lane = threadIdx.x & 31;
  const int tl = lane % TS;
  const u32 team = (blockIdx.x * 256u + threadIdx.x) / TS;
  const u32 tmask = (TS == 32) ? 0xffffffffu : (((1u << TS) - 1u) << (lane - tl));
  u32 n = 0, S = 0;
  if (team < R)
  {
    n = sg.count[team];
    S = sg.start[team];
    if (n >= VI_BIG) n = 0;
  }
  if (__all_sync(tmask, n == 0)) return;  // n is uniform inside a team
  const int C4 = ld >> 2;
  const u128 thr = ((u128)n * (u128)n) << (2 * VI_Q30_MIN_RES_BITS);
  bool ok = false;
  Q30Best best;
  best.key = 0;
  best.s1 = 0;
  best.idx = 0x7fffffff;

  for (int c0 = 0; c0 < C4; c0 += TS * CH)
  {
    u64 s1[CH * 4], s2lo[CH * 4];
    u32 s2hi[CH * 4];
#pragma unroll
    for (int i = 0; i < CH * 4; ++i) { s1[i] = 0; s2lo[i] = 0; s2hi[i] = 0; }

    for (u32 jb = 0; jb < n; jb += TS)
    {
      const u32 mine = (jb + tl < n) ? perm[S + jb + tl] : 0u;
      const u32 m = min((u32)TS, n - jb);
#pragma unroll 2
      for (u32 jj = 0; jj < m; ++jj)
      {
        const u32 r = __shfl_sync(tmask, mine, jj, TS);
        const float4* rp = reinterpret_cast<const float4*>(rows + (size_t)r * ld);
        float4 x[CH];
#pragma unroll
        for (int k = 0; k < CH; ++k)
        {
          const int c = c0 + k * TS + tl;
          x[k] = (c < C4) ? ldg_f4_stream(rp + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < CH; ++k)
        {
          q30_acc(s1[k * 4 + 0], s2lo[k * 4 + 0], s2hi[k * 4 + 0], x[k].x, qk);
          q30_acc(s1[k * 4 + 1], s2lo[k * 4 + 1], s2hi[k * 4 + 1], x[k].y, qk);
          q30_acc(s1[k * 4 + 2], s2lo[k * 4 + 2], s2hi[k * 4 + 2], x[k].z, qk);
          q30_acc(s1[k * 4 + 3], s2lo[k * 4 + 3], s2hi[k * 4 + 3], x[k].w, qk);
        }
      }
    }
    if (n > 0)
    {
#pragma unroll
      for (int k = 0; k < CH; ++k)
#pragma unroll
        for (int e = 0; e < 4; ++e)
        {
          const int d = (c0 + k * TS + tl) * 4 + e;
          if (d < dims)
          {
            const u128 key = q30_key(n, (i64)s1[k * 4 + e], s2lo[k * 4 + e], s2hi[k * 4 + e]);
            ok |= key >= thr;
            if (q30_better(mx != 0, key, d, best.key, best.idx))
            {
              best.key = key;
              best.s1 = (i64)s1[k * 4 + e];
              best.idx = d;
            }
          }
        }
    }
  }
  best = q30_reduce<TS>(best, mx != 0, tmask);
  int dim = best.idx;
  float mid = q30_mid(best.s1, n, qinv);
  if (!__any_sync(tmask, ok))
  {
    const ExBest eb = welford_team<TS, CH>(rows, ld, dims, perm + S, n, tl, tmask, mx != 0);
    dim = eb.idx;
    mid = eb.mean;
  }

  // id sum (Stats.IdN, IndexBuilder.cs:170,194)
  u64 slo = 0;
  i64 shi = 0;
  for (u32 j = tl; j < n; j += TS)
  {
    const i64 id = pid[S + j];
    slo += (u32)id;
    shi += (id >> 32);
  }
#pragma unroll
  for (int o = TS / 2; o > 0; o >>= 1)
  {
    slo += __shfl_xor_sync(tmask, slo, o);
    shi += __shfl_xor_sync(tmask, shi, o);
  }
  if (tl == 0) write_split(sg, out, team, dim, mid, mean_id(slo, shi, n));
}

// One CTA owns one chunk (VI_CHUNK rows) of one big range; partial sums go to gacc with integer atomics
// (order-independent, so the result does not depend on scheduling).
template <int TS, int CH>
__global__ void __launch_bounds__(256)
k_stats_big_q30(SegLevel sg, const u32* __restrict__ big_list, const u32* __restrict__ chunk_first, u32 nbig,
                const u32* __restrict__ perm, const i64* __restrict__ pid, const float* __restrict__ rows, int ld,
                int dims, float qk, u64* __restrict__ gacc)
{
  constexpr int NT = 256 / TS;          // teams per CTA
  constexpr int PD = TS * CH * 4;       // dims per pass
  __shared__ u64 sacc[PD * 4];
  __shared__ u32 sperm[VI_CHUNK];
  const u32 bid = blockIdx.x;
  u32 lo = 0, hi = nbig;
  while (hi - lo > 1)
  {
    const u32 m = (lo + hi) >> 1;
    if (chunk_first[m] <= bid) lo = m; else hi = m;
  }
  const u32 slot = lo;
  const u32 s = big_list[slot];
  const u32 S = sg.start[s], n = sg.count[s];
  const u32 a = (bid - chunk_first[slot]) * VI_CHUNK;
  const u32 b = min(n, a + VI_CHUNK);
  const u32 m = b - a;
  for (u32 i = threadIdx.x; i < m; i += 256) sperm[i] = perm[S + a + i];
  const int tl = threadIdx.x % TS, team = threadIdx.x / TS;
  const int C4 = ld >> 2;
  const size_t gstride = (size_t)ld * 4 + 2;
  u64* g = gacc + (size_t)slot * gstride;

  for (int c0 = 0; c0 < C4; c0 += TS * CH)
  {
    for (int i = threadIdx.x; i < PD * 4; i += 256) sacc[i] = 0;
    __syncthreads();
    u64 s1[CH * 4], s2lo[CH * 4];
    u32 s2hi[CH * 4];
#pragma unroll
    for (int i = 0; i < CH * 4; ++i) { s1[i] = 0; s2lo[i] = 0; s2hi[i] = 0; }
#pragma unroll 2
    for (u32 j = team; j < m; j += NT)
    {
      const u32 r = sperm[j];
      const float4* rp = reinterpret_cast<const float4*>(rows + (size_t)r * ld);
      float4 x[CH];
#pragma unroll
      for (int k = 0; k < CH; ++k)
      {
        const int c = c0 + k * TS + tl;
        x[k] = (c < C4) ? ldg_f4_stream(rp + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int k = 0; k < CH; ++k)
      {
        q30_acc(s1[k * 4 + 0], s2lo[k * 4 + 0], s2hi[k * 4 + 0], x[k].x, qk);
        q30_acc(s1[k * 4 + 1], s2lo[k * 4 + 1], s2hi[k * 4 + 1], x[k].y, qk);
        q30_acc(s1[k * 4 + 2], s2lo[k * 4 + 2], s2hi[k * 4 + 2], x[k].z, qk);
        q30_acc(s1[k * 4 + 3], s2lo[k * 4 + 3], s2hi[k * 4 + 3], x[k].w, qk);
      }
    }
#pragma unroll
    for (int k = 0; k < CH; ++k)
#pragma unroll
      for (int e = 0; e < 4; ++e)
      {
        const int dl = (k * TS + tl) * 4 + e;  // dim inside this pass
        if ((c0 + k * TS + tl) < C4)
        {
          atomicAdd(&sacc[dl * 4 + 0], s1[k * 4 + e]);
          atomicAdd(&sacc[dl * 4 + 1], s2lo[k * 4 + e] & 0xffffffffull);
          atomicAdd(&sacc[dl * 4 + 2], s2lo[k * 4 + e] >> 32);
          atomicAdd(&sacc[dl * 4 + 3], (u64)s2hi[k * 4 + e]);
        }
      }
    __syncthreads();
    const int pass_dims = min(PD, (C4 - c0) * 4);
    for (int i = threadIdx.x; i < pass_dims * 4; i += 256)
    {
      const u64 v = sacc[i];
      if (v) atomicAdd(&g[(size_t)c0 * 16 + i], v);
    }
    __syncthreads();
  }
  // id sums
  u64 slo = 0;
  i64 shi = 0;
  for (u32 j = a + threadIdx.x; j < b; j += 256)
  {
    const i64 id = pid[S + j];
    slo += (u32)id;
    shi += (id >> 32);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
  {
    slo += __shfl_xor_sync(0xffffffffu, slo, o);
    shi += __shfl_xor_sync(0xffffffffu, shi, o);
  }
  if ((threadIdx.x & 31) == 0)
  {
    atomicAdd(&g[(size_t)ld * 4 + 0], slo);
    atomicAdd(&g[(size_t)ld * 4 + 1], (u64)shi);
  }
}

// warp per big range: arg-max over the combined sums
__global__ void __launch_bounds__(256)
k_finalize_big_q30(SegLevel sg, const u32* __restrict__ big_list, u32 nbig, const u64* __restrict__ gacc, int ld,
                   int dims, double qinv, int mx, StatsOut out, const float* __restrict__ rows,
                   const u32* __restrict__ perm)
{
  const u32 warp = (blockIdx.x * 256u + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= nbig) return;
  const u32 s = big_list[warp];
  const u32 n = sg.count[s];
  const u128 thr = ((u128)n * (u128)n) << (2 * VI_Q30_MIN_RES_BITS);
  bool ok = false;
  const u64* g = gacc + (size_t)warp * ((size_t)ld * 4 + 2);
  Q30Best best;
  best.key = 0;
  best.s1 = 0;
  best.idx = 0x7fffffff;
  for (int d = lane; d < dims; d += 32)
  {
    const i64 s1 = (i64)g[d * 4 + 0];
    const u128 s2 = (u128)g[d * 4 + 1] + ((u128)g[d * 4 + 2] << 32) + ((u128)g[d * 4 + 3] << 64);
    const i128 a = (i128)s1;
    const u128 key = (u128)n * s2 - (u128)(a * a);
    ok |= key >= thr;
    if (q30_better(mx != 0, key, d, best.key, best.idx))
    {
      best.key = key;
      best.s1 = s1;
      best.idx = d;
    }
  }
  best = q30_reduce<32>(best, mx != 0, 0xffffffffu);
  int dim = best.idx;
  float mid = q30_mid(best.s1, n, qinv);
  if (!__any_sync(0xffffffffu, ok))
  {
    // >= VI_BIG points that the quantisation cannot tell apart: reference arithmetic, one warp (rare, slow)
    const ExBest eb = welford_team<32, 1>(rows, ld, dims, perm + sg.start[s], n, lane, 0xffffffffu, mx != 0);
    dim = eb.idx;
    mid = eb.mean;
  }
  if (lane == 0) write_split(sg, out, s, dim, mid, mean_id(g[(size_t)ld * 4], (i64)g[(size_t)ld * 4 + 1], n));
}

__global__ void k_big_chunks(const u32* __restrict__ count, const u32* __restrict__ big_list, const u32* counters,
                             u32* chunks, u32 bound)
{
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= bound) return;
  const u32 nbig = counters[0];
  chunks[i] = i < nbig ? (count[big_list[i]] + VI_CHUNK - 1) / VI_CHUNK : 0u;
}

// ---- exact mode -------------------------------------------------------------------------------------------------
// warp per small range; a lane runs CHX chains (dims lane, lane+32, ...) per column pass
template <int CHX>
__global__ void __launch_bounds__(256)
k_stats_small_exact(SegLevel sg, u32 R, const u32* __restrict__ perm, const i64* __restrict__ pid,
                    const float* __restrict__ rows, int ld, int dims, int mx, StatsOut out)
{
  const u32 s = (blockIdx.x * 256u + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (s >= R) return;
  const u32 n = sg.count[s];
  if (n >= VI_BIG) return;
  const u32 S = sg.start[s];
  ExBest best;
  best.key = 0.f;
  best.mean = 0.f;
  best.idx = 0x7fffffff;
  for (int c0 = 0; c0 < dims; c0 += 32 * CHX)
  {
    float mean[CHX], q[CHX];
    {
      const float* rp = rows + (size_t)perm[S] * ld;
#pragma unroll
      for (int k = 0; k < CHX; ++k)
      {
        const int c = c0 + k * 32 + lane;
        mean[k] = (c < dims) ? ldg_f_stream(rp + c) : 0.f;  // InitStats, IndexBuilder.cs:159-173
        q[k] = 0.f;
      }
    }
    for (u32 jb = 0; jb < n; jb += 32)
    {
      const u32 mine = (jb + lane < n) ? perm[S + jb + lane] : 0u;
      const u32 m = min(32u, n - jb);
#pragma unroll 4
      for (u32 jj = (jb == 0 ? 1u : 0u); jj < m; ++jj)
      {
        const u32 r = __shfl_sync(0xffffffffu, mine, jj);
        const float* rp = rows + (size_t)r * ld;
        float v[CHX];
#pragma unroll
        for (int k = 0; k < CHX; ++k)
        {
          const int c = c0 + k * 32 + lane;
          v[k] = (c < dims) ? ldg_f_stream(rp + c) : 0.f;
        }
        const float cnt = (float)(jb + jj + 1u);  // (float)(Count + 1), IndexBuilder.cs:185-186
#pragma unroll
        for (int k = 0; k < CHX; ++k) welford_step(mean[k], q[k], v[k], cnt);
      }
    }
#pragma unroll
    for (int k = 0; k < CHX; ++k)
    {
      const int d = c0 + k * 32 + lane;
      if (d < dims)
      {
        const float key = mx ? q[k] : -q[k];  // IndexBuilder.cs:79
        if (ex_better(key, d, best.key, best.idx))
        {
          best.key = key;
          best.mean = mean[k];
          best.idx = d;
        }
      }
    }
  }
  best = ex_reduce(best);
  u64 slo = 0;
  i64 shi = 0;
  for (u32 j = lane; j < n; j += 32)
  {
    const i64 id = pid[S + j];
    slo += (u32)id;
    shi += (id >> 32);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
  {
    slo += __shfl_xor_sync(0xffffffffu, slo, o);
    shi += __shfl_xor_sync(0xffffffffu, shi, o);
  }
  if (lane == 0) write_split(sg, out, s, best.idx, best.mean, mean_id(slo, shi, n));
}

// one warp (= one CTA) per (big range, block of 32 dims): one chain per thread, deep register prefetch
constexpr int EXU = 32;

__global__ void __launch_bounds__(32)
k_stats_big_exact(SegLevel sg, const u32* __restrict__ big_list, u32 nblk, const u32* __restrict__ perm,
                  const float* __restrict__ rows, int ld, int dims, float2* __restrict__ gstats)
{
  const u32 slot = blockIdx.x / nblk;
  const int col = (int)(blockIdx.x % nblk) * 32 + threadIdx.x;
  const int lane = threadIdx.x;
  const u32 s = big_list[slot];
  const u32 S = sg.start[s], n = sg.count[s];
  const bool act = col < dims;
  const float* base = rows + (act ? col : 0);
  const u32* pp = perm + S;

  float mean = ldg_f_stream(base + (size_t)pp[0] * ld), q = 0.f;
  float bufA[EXU], bufB[EXU];
  u32 pA, pB;
  auto load_perm = [&](u32 j0) -> u32 { return (j0 + lane < n) ? pp[j0 + lane] : 0u; };
  auto load_rows = [&](float(&buf)[EXU], u32 pv, u32 j0)
  {
#pragma unroll
    for (int u = 0; u < EXU; ++u)
    {
      const u32 r = __shfl_sync(0xffffffffu, pv, u);
      buf[u] = (j0 + u < n) ? ldg_f_stream(base + (size_t)r * ld) : 0.f;
    }
  };
  auto compute = [&](const float(&buf)[EXU], u32 j0)
  {
#pragma unroll
    for (int u = 0; u < EXU; ++u)
      if (j0 + u < n) welford_step(mean, q, buf[u], (float)(j0 + u + 1u));
  };
  pA = load_perm(1);
  pB = load_perm(1 + EXU);
  load_rows(bufA, pA, 1);
  pA = load_perm(1 + 2 * EXU);
  for (u32 j0 = 1; j0 < n; j0 += 2 * EXU)
  {
    load_rows(bufB, pB, j0 + EXU);
    pB = load_perm(j0 + 3 * EXU);
    compute(bufA, j0);
    load_rows(bufA, pA, j0 + 2 * EXU);
    pA = load_perm(j0 + 4 * EXU);
    compute(bufB, j0 + EXU);
  }
  if (act) gstats[(size_t)slot * dims + col] = make_float2(mean, q);
}

__global__ void __launch_bounds__(256)
k_finalize_big_exact(SegLevel sg, const u32* __restrict__ big_list, u32 nbig, const float2* __restrict__ gstats,
                     const i64* __restrict__ pid, int dims, int mx, StatsOut out)
{
  const u32 warp = (blockIdx.x * 256u + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= nbig) return;
  const u32 s = big_list[warp];
  const u32 S = sg.start[s], n = sg.count[s];
  ExBest best;
  best.key = 0.f;
  best.mean = 0.f;
  best.idx = 0x7fffffff;
  for (int d = lane; d < dims; d += 32)
  {
    const float2 st = gstats[(size_t)warp * dims + d];
    const float key = mx ? st.y : -st.y;
    if (ex_better(key, d, best.key, best.idx))
    {
      best.key = key;
      best.mean = st.x;
      best.idx = d;
    }
  }
  best = ex_reduce(best);
  u64 slo = 0;
  i64 shi = 0;
  for (u32 j = lane; j < n; j += 32)
  {
    const i64 id = pid[S + j];
    slo += (u32)id;
    shi += (id >> 32);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
  {
    slo += __shfl_xor_sync(0xffffffffu, slo, o);
    shi += __shfl_xor_sync(0xffffffffu, shi, o);
  }
  if (lane == 0) write_split(sg, out, s, best.idx, best.mean, mean_id(slo, shi, n));
}

// =============================================================================================================
// partition (IndexBuilder.cs:111-124): high iff value > Mid || (value == Mid && id > Id); stable
// =============================================================================================================
__global__ void __launch_bounds__(256)
k_flags(SegLevel sg, const u32* __restrict__ seg_of, const u32* __restrict__ perm, const i64* __restrict__ pid,
        const float* __restrict__ rows, int ld, u32 A, u32* __restrict__ fbits, u32* __restrict__ wcnt)
{
  const u32 p = blockIdx.x * 256u + threadIdx.x;
  bool hi = false;
  if (p < A)
  {
    const u32 s = seg_of[p];
    const int dim = sg.dim[s];
    const float mid = sg.mid[s];
    const float v = rows[(size_t)perm[p] * ld + dim];
    hi = v > mid || (v == mid && pid[p] > sg.pivot[s]);
  }
  const u32 b = __ballot_sync(0xffffffffu, hi);
  if ((threadIdx.x & 31) == 0 && (p >> 5) <= ((A + 31) >> 5))
  {
    fbits[p >> 5] = b;
    wcnt[p >> 5] = __popc(b);
  }
}

__global__ void __launch_bounds__(256)
k_seg_children(SegLevel sg, u32 R, const u32* __restrict__ wpre, const u32* __restrict__ fbits, u32* seg_nlo,
               u32* seg_hbase, u32* c_rows, u64* c_actpos)
{
  const u32 s = blockIdx.x * 256u + threadIdx.x;
  if (s >= R) return;
  const u32 S = sg.start[s], n = sg.count[s];
  const u32 hb = hi_before(wpre, fbits, S);
  const u32 nhi = hi_before(wpre, fbits, S + n) - hb;
  const u32 nlo = n - nhi;
  seg_nlo[s] = nlo;
  seg_hbase[s] = hb;
  c_rows[s] = (nlo > 0) + (nhi > 0);
  const u32 act = (nlo >= 2) + (nhi >= 2);
  const u32 pos = (nlo >= 2 ? nlo : 0) + (nhi >= 2 ? nhi : 0);
  c_actpos[s] = ((u64)act << 32) | (u64)pos;
}

struct TableOut
{
  i64* t_rid;
  int* t_dim;
  float* t_mid;
  i64* t_id;
  int* t_low;
  int* t_high;
};

// counters: [0] next-level big count, [1] error flag
__global__ void __launch_bounds__(256)
k_emit_children(SegLevel sg, u32 R, const u32* __restrict__ seg_nlo, const u32* __restrict__ c_rows,
                const u64* __restrict__ c_actpos, SegLevel nx, u32 row_base_next, u32 t_cap, TableOut t,
                u32* big_list_next, u32* counters)
{
  const u32 s = blockIdx.x * 256u + threadIdx.x;
  if (s >= R) return;
  if ((u64)row_base_next + c_rows[R] > (u64)t_cap)
  {
    if (s == 0) counters[1] = 1;
    return;
  }
  const u32 n = sg.count[s], nlo = seg_nlo[s], nhi = n - nlo;
  const i64 rid = sg.rid[s];
  const u32 row = sg.row[s];
  const u32 r0 = row_base_next + c_rows[s];
  const u32 a0 = (u32)(c_actpos[s] >> 32), p0 = (u32)c_actpos[s];
  const int lo_row = nlo > 0 ? (int)r0 : -1;
  const int hi_row = nhi > 0 ? (int)(r0 + (nlo > 0 ? 1u : 0u)) : -1;
  t.t_low[row] = lo_row;
  t.t_high[row] = hi_row;
  if (nlo > 0)
  {
    t.t_rid[lo_row] = rid * 2 + 1;  // IndexBuilder.cs:99
    t.t_low[lo_row] = -1;
    t.t_high[lo_row] = -1;
    if (nlo == 1)
    {
      t.t_dim[lo_row] = -1;  // leaf, IndexBuilder.cs:81-82; its Id is written by k_scatter
      t.t_mid[lo_row] = 0.0f;
    }
    else
    {
      nx.start[a0] = p0;
      nx.count[a0] = nlo;
      nx.rid[a0] = rid * 2 + 1;
      nx.row[a0] = (u32)lo_row;
      if (nlo >= VI_BIG) big_list_next[atomicAdd(&counters[0], 1u)] = a0;
    }
  }
  if (nhi > 0)
  {
    t.t_rid[hi_row] = rid * 2 + 2;  // IndexBuilder.cs:104
    t.t_low[hi_row] = -1;
    t.t_high[hi_row] = -1;
    if (nhi == 1)
    {
      t.t_dim[hi_row] = -1;
      t.t_mid[hi_row] = 0.0f;
    }
    else
    {
      const u32 a1 = a0 + (nlo >= 2 ? 1u : 0u);
      nx.start[a1] = p0 + (nlo >= 2 ? nlo : 0u);
      nx.count[a1] = nhi;
      nx.rid[a1] = rid * 2 + 2;
      nx.row[a1] = (u32)hi_row;
      if (nhi >= VI_BIG) big_list_next[atomicAdd(&counters[0], 1u)] = a1;
    }
  }
}

__global__ void __launch_bounds__(256)
k_scatter(SegLevel sg, const u32* __restrict__ seg_of, const u32* __restrict__ perm, const i64* __restrict__ pid,
          u32 A, u32 R, const u32* __restrict__ fbits, const u32* __restrict__ wpre, const u32* __restrict__ seg_nlo,
          const u32* __restrict__ seg_hbase, const u32* __restrict__ c_rows, const u64* __restrict__ c_actpos,
          u32 row_base_next, u32* __restrict__ perm_n, i64* __restrict__ pid_n, u32* __restrict__ seg_of_n,
          i64* __restrict__ t_id, int* __restrict__ t_src, const u32* __restrict__ counters)
{
  const u32 p = blockIdx.x * 256u + threadIdx.x;
  if (p >= A || counters[1]) return;
  const u32 s = seg_of[p];
  const u32 S = sg.start[s], n = sg.count[s], nlo = seg_nlo[s], nhi = n - nlo;
  const u32 w = fbits[p >> 5];
  const bool hi = (w >> (p & 31)) & 1u;
  const u32 hb = wpre[p >> 5] + __popc(w & ((1u << (p & 31)) - 1u)) - seg_hbase[s];
  const u32 r0 = row_base_next + c_rows[s];
  const u32 a0 = (u32)(c_actpos[s] >> 32), p0 = (u32)c_actpos[s];
  const u32 r = perm[p];
  const i64 id = pid[p];
  if (!hi)
  {
    if (nlo >= 2)
    {
      const u32 dst = p0 + (p - S) - hb;
      perm_n[dst] = r;
      pid_n[dst] = id;
      seg_of_n[dst] = a0;
    }
    else
    {
      t_id[r0] = id;  // the single low point is a leaf: RangeValue.Id = its id
      t_src[r0] = (int)r;
    }
  }
  else
  {
    if (nhi >= 2)
    {
      const u32 dst = p0 + (nlo >= 2 ? nlo : 0u) + hb;
      perm_n[dst] = r;
      pid_n[dst] = id;
      seg_of_n[dst] = a0 + (nlo >= 2 ? 1u : 0u);
    }
    else
    {
      t_id[r0 + (nlo > 0 ? 1u : 0u)] = id;
      t_src[r0 + (nlo > 0 ? 1u : 0u)] = (int)r;
    }
  }
}

__global__ void k_totals(const u32* c_rows, const u64* c_actpos, u32 R, const u32* counters, const u32* chunk_first,
                         u32 chunk_bound, LevelTotals* out)
{
  out->rows = c_rows[R];
  out->segs = (u32)(c_actpos[R] >> 32);
  out->pos = (u32)c_actpos[R];
  out->nbig = counters[0];
  out->chunks = chunk_first ? chunk_first[chunk_bound] : 0u;
  out->pad[0] = counters[1];
}

__global__ void k_pack_nodes(const int* __restrict__ t_dim, const float* __restrict__ t_mid, const i64* __restrict__ t_id,
                             const int* __restrict__ t_low, const int* __restrict__ t_high, int4* __restrict__ node, u32 n)
{
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int d = t_dim[i];
  int4 v;
  v.x = d;
  v.y = __float_as_int(t_mid[i]);
  if (d < 0)
  {
    const u64 id = (u64)t_id[i];  // leaf: carry the TextID in the child slots
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

// =============================================================================================================
// host driver
// =============================================================================================================
template <typename T>
static cudaError_t dalloc(T** p, size_t count)
{
  return cudaMalloc((void**)p, count * sizeof(T) + 256);
}

void vi_free_workspace(vi_ctx* ctx)
{
  for (int i = 0; i < 2; ++i)
  {
    cudaFree(ctx->perm[i]); ctx->perm[i] = nullptr;
    cudaFree(ctx->pid[i]); ctx->pid[i] = nullptr;
    cudaFree(ctx->seg_of[i]); ctx->seg_of[i] = nullptr;
    cudaFree(ctx->big_list[i]); ctx->big_list[i] = nullptr;
    SegLevel& s = ctx->seg[i];
    cudaFree(s.start); cudaFree(s.count); cudaFree(s.rid); cudaFree(s.row); cudaFree(s.dim); cudaFree(s.mid);
    cudaFree(s.pivot);
    s = SegLevel();
  }
  cudaFree(ctx->chunk_first); ctx->chunk_first = nullptr;
  cudaFree(ctx->fbits); ctx->fbits = nullptr;
  cudaFree(ctx->wpre); ctx->wpre = nullptr;
  cudaFree(ctx->seg_nlo); ctx->seg_nlo = nullptr;
  cudaFree(ctx->seg_hbase); ctx->seg_hbase = nullptr;
  cudaFree(ctx->c_rows); ctx->c_rows = nullptr;
  cudaFree(ctx->c_actpos); ctx->c_actpos = nullptr;
  cudaFree(ctx->scan_tmp); ctx->scan_tmp = nullptr;
  cudaFree(ctx->gacc); ctx->gacc = nullptr;
  cudaFree(ctx->gstats); ctx->gstats = nullptr;
  cudaFree(ctx->d_absmax); ctx->d_absmax = nullptr;
  if (ctx->totals) cudaFreeHost(ctx->totals);
  ctx->totals = nullptr;
  ctx->ws_n = 0;
}

void vi_free_table(vi_ctx* ctx)
{
  cudaFree(ctx->t_rid); ctx->t_rid = nullptr;
  cudaFree(ctx->t_dim); ctx->t_dim = nullptr;
  cudaFree(ctx->t_mid); ctx->t_mid = nullptr;
  cudaFree(ctx->t_id); ctx->t_id = nullptr;
  cudaFree(ctx->t_low); ctx->t_low = nullptr;
  cudaFree(ctx->t_high); ctx->t_high = nullptr;
  cudaFree(ctx->t_node); ctx->t_node = nullptr;
  cudaFree(ctx->t_src); ctx->t_src = nullptr;
  ctx->t_cap = 0;
  ctx->t_rows = 0;
  ctx->built = false;
}

static int alloc_workspace(vi_ctx* ctx)
{
  const int64_t n = ctx->n;
  if (ctx->ws_n >= n && ctx->perm[0]) return VI_OK;
  vi_free_workspace(ctx);
  const size_t N = (size_t)n;
  const size_t maxseg = N / 2 + 2;
  const size_t maxbig = N / VI_BIG + 2;
  const size_t words = N / 32 + 4;
  for (int i = 0; i < 2; ++i)
  {
    VI_CUDA_TRY(dalloc(&ctx->perm[i], N));
    VI_CUDA_TRY(dalloc(&ctx->pid[i], N));
    VI_CUDA_TRY(dalloc(&ctx->seg_of[i], N));
    VI_CUDA_TRY(dalloc(&ctx->big_list[i], maxbig));
    SegLevel& s = ctx->seg[i];
    VI_CUDA_TRY(dalloc(&s.start, maxseg));
    VI_CUDA_TRY(dalloc(&s.count, maxseg));
    VI_CUDA_TRY(dalloc(&s.rid, maxseg));
    VI_CUDA_TRY(dalloc(&s.row, maxseg));
    VI_CUDA_TRY(dalloc(&s.dim, maxseg));
    VI_CUDA_TRY(dalloc(&s.mid, maxseg));
    VI_CUDA_TRY(dalloc(&s.pivot, maxseg));
  }
  VI_CUDA_TRY(dalloc(&ctx->chunk_first, maxbig + 1));
  VI_CUDA_TRY(dalloc(&ctx->fbits, words));
  VI_CUDA_TRY(dalloc(&ctx->wpre, words + 1));
  VI_CUDA_TRY(dalloc(&ctx->seg_nlo, maxseg));
  VI_CUDA_TRY(dalloc(&ctx->seg_hbase, maxseg));
  VI_CUDA_TRY(dalloc(&ctx->c_rows, maxseg + 1));
  VI_CUDA_TRY(dalloc(&ctx->c_actpos, maxseg + 1));
  VI_CUDA_TRY(dalloc((u64**)&ctx->scan_tmp, maxseg / SCAN_TILE + words / SCAN_TILE + 64));
  VI_CUDA_TRY(dalloc(&ctx->gacc, maxbig * ((size_t)ctx->ld * 4 + 2)));
  VI_CUDA_TRY(dalloc(&ctx->gstats, maxbig * (size_t)ctx->dims));
  VI_CUDA_TRY(dalloc(&ctx->d_absmax, (size_t)4));
  VI_CUDA_TRY(cudaMallocHost((void**)&ctx->totals, sizeof(LevelTotals)));
  ctx->ws_n = n;
  return VI_OK;
}

static int alloc_table(vi_ctx* ctx)
{
  // 2n-1 rows when no child is ever empty; an empty child (all points on one side) adds a one-child row.
  const int64_t cap = 2 * ctx->n + ctx->n / 8 + 1024;
  if (ctx->t_cap >= cap && ctx->t_rid) return VI_OK;
  vi_free_table(ctx);
  if (cap >= (int64_t)0x7fffffff) return ctx->fail(VI_ERR_CAPACITY, "too many points for 32-bit row indexes");
  VI_CUDA_TRY(dalloc(&ctx->t_rid, (size_t)cap));
  VI_CUDA_TRY(dalloc(&ctx->t_dim, (size_t)cap));
  VI_CUDA_TRY(dalloc(&ctx->t_mid, (size_t)cap));
  VI_CUDA_TRY(dalloc(&ctx->t_id, (size_t)cap));
  VI_CUDA_TRY(dalloc(&ctx->t_low, (size_t)cap));
  VI_CUDA_TRY(dalloc(&ctx->t_high, (size_t)cap));
  VI_CUDA_TRY(dalloc(&ctx->t_node, (size_t)cap));
  VI_CUDA_TRY(dalloc(&ctx->t_src, (size_t)cap));
  ctx->t_cap = cap;
  return VI_OK;
}

// ---- kernel dispatch on the dimension count -----------------------------------------------------------------
struct Q30Shape
{
  int ts, ch;
};
static Q30Shape q30_shape(int ld)
{
  const int c4 = ld / 4;
  if (c4 <= 8) return {8, 1};
  if (c4 <= 16) return {8, 2};
  if (c4 <= 24) return {8, 3};
  if (c4 <= 32) return {8, 4};
  if (c4 <= 64) return {32, 2};
  if (c4 <= 96) return {32, 3};
  if (c4 <= 128) return {32, 4};
  return {32, 6};  // wider rows take several column passes
}

#define Q30_DISPATCH(ts, ch, CALL)                  \
  do                                                \
  {                                                 \
    if (ts == 8 && ch == 1) { CALL(8, 1); }         \
    else if (ts == 8 && ch == 2) { CALL(8, 2); }    \
    else if (ts == 8 && ch == 3) { CALL(8, 3); }    \
    else if (ts == 8 && ch == 4) { CALL(8, 4); }    \
    else if (ts == 32 && ch == 2) { CALL(32, 2); }  \
    else if (ts == 32 && ch == 3) { CALL(32, 3); }  \
    else if (ts == 32 && ch == 4) { CALL(32, 4); }  \
    else { CALL(32, 6); }                           \
  } while (0)

static int exact_chx(int dims)
{
  if (dims <= 32) return 1;
  if (dims <= 64) return 2;
  if (dims <= 96) return 3;
  if (dims <= 128) return 4;
  return 8;
}

int vi_build_impl(vi_ctx* ctx, int mode)
{
  const int64_t n64 = ctx->n;
  ctx->built = false;
  ctx->levels.clear();
  ctx->info = vi_build_info();
  ctx->info.mode = mode;
  if (n64 >= (int64_t)0x7fffffff) return ctx->fail(VI_ERR_CAPACITY, "more than 2^31-2 points per context");
  if (n64 == 0)
  {
    // IndexBuilder.cs:70-73: an empty range emits no row
    ctx->t_rows = 0;
    ctx->built = true;
    return VI_OK;
  }
  int rc = alloc_table(ctx);
  if (rc != VI_OK) return rc;
  cudaStream_t st = ctx->stream;
  const u32 n = (u32)n64;
  const int ld = ctx->ld, dims = ctx->dims;
  int64_t launches = 0;

  cudaEvent_t ev_begin, ev_end;
  VI_CUDA_TRY(cudaEventCreate(&ev_begin));
  VI_CUDA_TRY(cudaEventCreate(&ev_end));
  std::vector<cudaEvent_t> lev_ev;
  auto new_event = [&]() -> cudaEvent_t
  {
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    lev_ev.push_back(e);
    return e;
  };
  auto cleanup = [&]()
  {
    for (cudaEvent_t e : lev_ev) cudaEventDestroy(e);
    cudaEventDestroy(ev_begin);
    cudaEventDestroy(ev_end);
  };

  VI_CUDA_TRY(cudaEventRecord(ev_begin, st));
  if (n == 1)
  {
    k_single_point<<<1, 1, 0, st>>>(ctx->ids, ctx->t_rid, ctx->t_dim, ctx->t_mid, ctx->t_id, ctx->t_low, ctx->t_high,
                                    ctx->t_src);
    k_pack_nodes<<<1, 32, 0, st>>>(ctx->t_dim, ctx->t_mid, ctx->t_id, ctx->t_low, ctx->t_high, ctx->t_node, 1);
    VI_CUDA_TRY(cudaEventRecord(ev_end, st));
    VI_CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, ev_begin, ev_end);
    cleanup();
    ctx->t_rows = 1;
    ctx->built = true;
    ctx->info.ranges = 1;
    ctx->info.levels = 1;
    ctx->info.kernel_launches = 2;
    ctx->info.build_ms = ms;
    return VI_OK;
  }
  rc = alloc_workspace(ctx);
  if (rc != VI_OK) { cleanup(); return rc; }

  // fast mode: quantisation exponent E with max|x| < 2^E
  int qe = 0;
  float qk = 1.0f;
  double qinv = 1.0;
  if (mode == VI_MODE_FAST)
  {
    VI_CUDA_TRY(cudaMemsetAsync(ctx->d_absmax, 0, 4, st));
    k_absmax<<<VI_NUM_SMS * 8, 256, 0, st>>>(reinterpret_cast<const float4*>(ctx->rows), (size_t)n * ld / 4,
                                             (u32*)ctx->d_absmax);
    ++launches;
    float amax = 0.f;
    VI_CUDA_TRY(cudaMemcpyAsync(&amax, ctx->d_absmax, 4, cudaMemcpyDeviceToHost, st));
    VI_CUDA_TRY(cudaStreamSynchronize(st));
    if (isinf(amax)) qe = 128;
    else if (amax > 0.f) (void)frexpf(amax, &qe);
    if (qe < -96) qe = -96;
    if (qe > 128) qe = 128;
    qk = ldexpf(1.0f, 30 - qe);
    qinv = ldexp(1.0, qe - 30);
    ctx->info.q30_exponent = qe;
  }

  TableOut tout{ctx->t_rid, ctx->t_dim, ctx->t_mid, ctx->t_id, ctx->t_low, ctx->t_high};
  StatsOut sout{ctx->t_dim, ctx->t_mid, ctx->t_id};

  k_init_level0<<<(n + 255) / 256, 256, 0, st>>>(ctx->perm[0], ctx->pid[0], ctx->ids, ctx->seg_of[0], n, ctx->seg[0],
                                                 ctx->big_list[0], ctx->t_rid, ctx->t_low, ctx->t_high);
  ++launches;

  u32 A = n, R = 1;
  u32 nbig = n >= VI_BIG ? 1u : 0u;
  u32 chunks = nbig ? (n + VI_CHUNK - 1) / VI_CHUNK : 0u;
  if (nbig)
  {
    const u32 h[2] = {0u, chunks};
    VI_CUDA_TRY(cudaMemcpyAsync(ctx->chunk_first, h, sizeof(h), cudaMemcpyHostToDevice, st));
  }
  u32 row_base = 0, nrows = 1;
  int cur = 0;
  int level = 0;
  const Q30Shape shp = q30_shape(ld);
  const int chx = exact_chx(dims);
  const size_t gstride = (size_t)ld * 4 + 2;

  while (R > 0)
  {
    if (level >= VI_MAX_DEPTH)
    {
      cleanup();
      return ctx->fail(VI_ERR_OVERFLOW, "rangeId overflow: a range at depth 62 still holds more than one point "
                                        "(IndexBuilder.cs:99 checked(rangeId * 2 + 1))");
    }
    const int nxt = cur ^ 1;
    const int mx = (level & 1) == 0;  // root max = true, children !max (IndexBuilder.cs:33,128-129)
    SegLevel& sg = ctx->seg[cur];
    cudaEvent_t e0 = new_event();

    // ---- statistics + split choice -----------------------------------------------------------------------
    if (mode == VI_MODE_FAST)
    {
      if (nbig)
      {
        VI_CUDA_TRY(cudaMemsetAsync(ctx->gacc, 0, (size_t)nbig * gstride * sizeof(u64), st));
#define CALL_BIG(TS, CH)                                                                                              \
  k_stats_big_q30<TS, CH><<<chunks, 256, 0, st>>>(sg, ctx->big_list[cur], ctx->chunk_first, nbig, ctx->perm[cur],       \
                                                  ctx->pid[cur], ctx->rows, ld, dims, qk, ctx->gacc)
        Q30_DISPATCH(shp.ts, shp.ch, CALL_BIG);
#undef CALL_BIG
        if (ctx->world > 1 && ctx->allreduce)
        {
          const int arc = ctx->allreduce(ctx->allreduce_user, ctx->gacc, (int64_t)((size_t)nbig * gstride));
          if (arc != 0) { cleanup(); return ctx->fail(VI_ERR_CUDA, "all-reduce callback failed"); }
        }
        k_finalize_big_q30<<<(nbig * 32 + 255) / 256, 256, 0, st>>>(sg, ctx->big_list[cur], nbig, ctx->gacc, ld, dims,
                                                                    qinv, mx, sout, ctx->rows, ctx->perm[cur]);
        launches += 2;
      }
      if (R > nbig)
      {
#define CALL_SMALL(TS, CH)                                                                                           \
  k_stats_small_q30<TS, CH><<<(u32)(((u64)R * TS + 255) / 256), 256, 0, st>>>(sg, R, ctx->perm[cur], ctx->pid[cur],    \
                                                                              ctx->rows, ld, dims, qk, qinv, mx, sout)
        Q30_DISPATCH(shp.ts, shp.ch, CALL_SMALL);
#undef CALL_SMALL
        ++launches;
      }
    }
    else
    {
      if (nbig)
      {
        const u32 nblk = (u32)((dims + 31) / 32);
        k_stats_big_exact<<<nbig * nblk, 32, 0, st>>>(sg, ctx->big_list[cur], nblk, ctx->perm[cur], ctx->rows, ld, dims,
                                                      ctx->gstats);
        k_finalize_big_exact<<<(nbig * 32 + 255) / 256, 256, 0, st>>>(sg, ctx->big_list[cur], nbig, ctx->gstats,
                                                                      ctx->pid[cur], dims, mx, sout);
        launches += 2;
      }
      if (R > nbig)
      {
        const u32 grid = (u32)(((u64)R * 32 + 255) / 256);
        switch (chx)
        {
          case 1: k_stats_small_exact<1><<<grid, 256, 0, st>>>(sg, R, ctx->perm[cur], ctx->pid[cur], ctx->rows, ld, dims, mx, sout); break;
          case 2: k_stats_small_exact<2><<<grid, 256, 0, st>>>(sg, R, ctx->perm[cur], ctx->pid[cur], ctx->rows, ld, dims, mx, sout); break;
          case 3: k_stats_small_exact<3><<<grid, 256, 0, st>>>(sg, R, ctx->perm[cur], ctx->pid[cur], ctx->rows, ld, dims, mx, sout); break;
          case 4: k_stats_small_exact<4><<<grid, 256, 0, st>>>(sg, R, ctx->perm[cur], ctx->pid[cur], ctx->rows, ld, dims, mx, sout); break;
          default: k_stats_small_exact<8><<<grid, 256, 0, st>>>(sg, R, ctx->perm[cur], ctx->pid[cur], ctx->rows, ld, dims, mx, sout); break;
        }
        ++launches;
      }
    }
    cudaEvent_t e1 = new_event();

    // ---- stable partition ----------------------------------------------------------------------------------
    const u32 W = (A + 31) / 32;
    k_flags<<<W / 8 + 1, 256, 0, st>>>(sg, ctx->seg_of[cur], ctx->perm[cur], ctx->pid[cur], ctx->rows, ld, A, ctx->fbits,
                                       ctx->wpre);
    ++launches;
    scan_exclusive<u32>(ctx, ctx->wpre, W + 1, launches);
    k_seg_children<<<(R + 255) / 256, 256, 0, st>>>(sg, R, ctx->wpre, ctx->fbits, ctx->seg_nlo, ctx->seg_hbase,
                                                    ctx->c_rows, ctx->c_actpos);
    ++launches;
    scan_exclusive<u32>(ctx, ctx->c_rows, R, launches);
    scan_exclusive<u64>(ctx, ctx->c_actpos, R, launches);
    VI_CUDA_TRY(cudaMemsetAsync(ctx->counters, 0, 8, st));
    const u32 row_base_next = row_base + nrows;
    k_emit_children<<<(R + 255) / 256, 256, 0, st>>>(sg, R, ctx->seg_nlo, ctx->c_rows, ctx->c_actpos, ctx->seg[nxt],
                                                     row_base_next, (u32)ctx->t_cap, tout, ctx->big_list[nxt],
                                                     ctx->counters);
    k_scatter<<<(A + 255) / 256, 256, 0, st>>>(sg, ctx->seg_of[cur], ctx->perm[cur], ctx->pid[cur], A, R, ctx->fbits,
                                               ctx->wpre, ctx->seg_nlo, ctx->seg_hbase, ctx->c_rows, ctx->c_actpos,
                                               row_base_next, ctx->perm[nxt], ctx->pid[nxt], ctx->seg_of[nxt], ctx->t_id,
                                               ctx->t_src, ctx->counters);
    launches += 2;
    const u32 big_bound = A / VI_BIG + 1;
    u32* chunk_arr = nullptr;
    if (mode == VI_MODE_FAST)
    {
      chunk_arr = ctx->chunk_first;
      k_big_chunks<<<(big_bound + 255) / 256, 256, 0, st>>>(ctx->seg[nxt].count, ctx->big_list[nxt], ctx->counters,
                                                            chunk_arr, big_bound);
      ++launches;
      scan_exclusive<u32>(ctx, chunk_arr, big_bound, launches);
    }
    k_totals<<<1, 1, 0, st>>>(ctx->c_rows, ctx->c_actpos, R, ctx->counters, chunk_arr, big_bound, ctx->totals);
    ++launches;
    cudaEvent_t e2 = new_event();
    VI_CUDA_TRY(cudaStreamSynchronize(st));
    const LevelTotals tt = *ctx->totals;
    if (tt.pad[0])
    {
      cleanup();
      return ctx->fail(VI_ERR_CAPACITY, "range table capacity exceeded (degenerate input: too many one-child ranges)");
    }
    vi_level_info li{};
    li.level = level;
    li.ranges = R;
    li.points = A;
    li.rows_emitted = nrows;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    li.stats_ms = ms;
    cudaEventElapsedTime(&ms, e1, e2);
    li.partition_ms = ms;
    ctx->levels.push_back(li);
    ctx->info.point_visits += A;

    row_base = row_base_next;
    nrows = tt.rows;
    R = tt.segs;
    A = tt.pos;
    nbig = tt.nbig;
    chunks = tt.chunks;
    cur = nxt;
    ++level;
  }
  // last level's rows are all leaves (or none)
  if (nrows > 0)
  {
    vi_level_info li{};
    li.level = level;
    li.rows_emitted = nrows;
    ctx->levels.push_back(li);
  }
  const u32 total_rows = row_base + nrows;
  k_pack_nodes<<<(total_rows + 255) / 256, 256, 0, st>>>(ctx->t_dim, ctx->t_mid, ctx->t_id, ctx->t_low, ctx->t_high,
                                                         ctx->t_node, total_rows);
  ++launches;
  VI_CUDA_TRY(cudaEventRecord(ev_end, st));
  VI_CUDA_TRY(cudaStreamSynchronize(st));
  VI_CUDA_TRY(cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, ev_begin, ev_end);
  cleanup();
  ctx->t_rows = total_rows;
  ctx->built = true;
  ctx->info.ranges = total_rows;
  ctx->info.levels = (int32_t)ctx->levels.size();
  ctx->info.kernel_launches = launches;
  ctx->info.build_ms = ms;
  return VI_OK;
}
