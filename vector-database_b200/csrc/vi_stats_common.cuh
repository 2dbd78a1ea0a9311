// vi_stats_common.cuh -- pieces shared by the exact and fast statistics kernels.
#pragma once
#include "vi_common.cuh"

struct StatsOut
{
  int* t_dim;
  float* t_mid;
  i64* t_id;
};

// RangeValue { Dimension, Mid, Id } of a split range (IndexBuilder.cs:83-88) + the copy the partition pass reads
// The `mx` argument of the statistics kernels: bit 0 = this level takes the max-variance dimension; VI_MODE_SQL adds
// bit 1 = a fallback range whose chosen Stdev2N is 0 gets a null Dimension / Mid (DDL.sql:193-194) and bit 2 = this is
// dbo.BuildIndex's root, which sends Value = Mean to the high child whatever the id (DDL.sql:104).
constexpr int VI_MX_MAX = 1, VI_MX_SQL = 2, VI_MX_ROOT_HIGH = 4;

__device__ __forceinline__ void write_split(const SegLevel& sg, const StatsOut& o, u32 s, int dim, float mid, i64 pivot,
                                            bool null_dim = false, bool root_high = false)
{
  const u32 row = sg.row[s];
  o.t_dim[row] = null_dim ? VI_DIM_NULL : dim;
  o.t_mid[row] = null_dim ? __int_as_float(0x7fc00000) : mid;
  o.t_id[row] = pivot;
  // the partition pass always splits on (dim, mid, pivot): with Stdev = 0 every value equals mid and the ids decide
  sg.dim[s] = dim;
  sg.mid[s] = mid;
  sg.pivot[s] = (root_high && !null_dim) ? (i64)INT64_MIN : pivot;  // id > INT64_MIN: ties go high
}

// (long)(IdN / Count), Int128 division truncating toward zero (IndexBuilder.cs:87).
// IdN = shi * 2^32 + slo  (slo = sum of low 32-bit halves, shi = sum of arithmetic-shifted high halves).
__device__ __forceinline__ i64 mean_id(u64 slo, i64 shi, u32 n)
{
  const i128 idn = ((i128)shi << 32) + (i128)slo;
  if (idn >= (i128)INT64_MIN && idn <= (i128)INT64_MAX) return (i64)idn / (i64)n;
  return (i64)(idn / (i128)n);
}

// Sum of ids of a range over `W` cooperating lanes (lane index gl, mask gmask), split in 32-bit halves so that two
// 64-bit accumulators hold the Int128 sum exactly (Stats.IdN, IndexBuilder.cs:170,194).
template <int W>
__device__ __forceinline__ i64 range_mean_id(const i64* __restrict__ pid, u32 n, int gl, u32 gmask)
{
  u64 slo = 0;
  i64 shi = 0;
  for (u32 j = gl; j < n; j += W)
  {
    const i64 id = pid[j];
    slo += (u32)id;
    shi += (id >> 32);
  }
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1)
  {
    slo += __shfl_xor_sync(gmask, slo, o);
    shi += __shfl_xor_sync(gmask, shi, o);
  }
  return mean_id(slo, shi, n);
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
  int idx;  // INT_MAX = none
};

// MaxBy replaces only on strictly greater (IndexBuilder.cs:77-79): the lowest index wins ties
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

// Float32 statistics of one range computed by a team of TS lanes, rows in position order: the fast mode's fallback
// for a poorly resolved range.  Lane tl owns the float4 column chunks c0 + k*TS + tl.  pp = perm + range start.
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
