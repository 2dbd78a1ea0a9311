// vi_stats_exact.cuh -- exact-mode statistics: the literal float32 recurrence of IndexBuilder.cs:159-197, one chain
// per (range, dimension) in stable position order, so the range table is bit-identical to the reference's.
//
// The recurrence is a serial dependency (mean_k depends on mean_{k-1}); its latency per point is what bounds the
// top levels.  The divide (value - pa) / count dominates that latency, so it is taken off the critical path:
// count is known in advance, r = RN(1/count) is computed by another lane ahead of time, and the quotient is
// q0 = d*r followed by two Markstein corrections q <- fma(fma(-q, c, d), r, q).  With r correctly rounded the first
// correction makes q faithful and the second makes it RN(d/c) (Markstein 1990; Muller et al., Handbook of
// Floating-Point Arithmetic, ch. 4) provided d - q*c is exact, which holds for 2^-100 <= |d| <= 2^100 and
// 1 <= c <= 2^31; anything else (zero, tiny, huge, Inf, NaN) takes __fdiv_rn.  vi_debug_divcheck compares the two
// on the device (tests/test_gpu_parity.py::test_fast_division_is_correctly_rounded).
//
//   n <  t_big   one warp per range, a lane runs CHX chains (dims lane, lane+32, ...)   k_stats_small_exact
//   n >= t_big   one warp per (range, 32 dims): one chain per thread over a cp.async shared-memory ring, 32-step
//                groups run speculatively with a 4-deep dependency chain and a check (k_stats_big_exact)
//                + k_finalize_big_exact for the arg-max and the id sum
#pragma once
#include "vi_stats_common.cuh"

// RN(d / c) for 2^-100 <= |d| <= 2^100 or d == +-0 (the quotient of a zero keeps the zero's sign since c > 0)
__device__ __forceinline__ float div_core(float d, float c, float r)
{
  const float q0 = __fmul_rn(d, r);
  const float e0 = __fmaf_rn(-q0, c, d);
  const float q1 = __fmaf_rn(e0, r, q0);
  const float e1 = __fmaf_rn(-q1, c, d);
  const float q2 = __fmaf_rn(e1, r, q1);
  return d == 0.f ? d : q2;
}

// operands div_core does not cover: tiny non-zero, huge, Inf, NaN
__device__ __forceinline__ bool div_needs_ieee(float d)
{
  const float ad = fabsf(d);
  return !(ad <= 0x1p100f) || (ad < 0x1p-100f && ad != 0.f);
}

__device__ __forceinline__ float div_by_count(float d, float c, float r)
{
  return div_needs_ieee(d) ? __fdiv_rn(d, c) : div_core(d, c, r);
}

// same results as welford_step (vi_stats_common.cuh), with r = RN(1/c) supplied
__device__ __forceinline__ void welford_step_r(float& mean, float& q, float value, float c, float r)
{
  const float d1 = __fsub_rn(value, mean);
  const float a = __fadd_rn(mean, div_by_count(d1, c, r));
  q = __fadd_rn(q, __fmul_rn(d1, __fsub_rn(value, a)));
  mean = a;
}

__device__ __forceinline__ ExBest ex_reduce(ExBest b) { return ex_reduce_w<32>(b, 0xffffffffu); }

template <int CHX>
__global__ void __launch_bounds__(256)
k_stats_small_exact(SegLevel sg, u32 R, u32 nmax, const u32* __restrict__ perm, const i64* __restrict__ pid,
                    const float* __restrict__ rows, int ld, int dims, int mx, StatsOut out)
{
  const u32 s = (blockIdx.x * 256u + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (s >= R) return;
  const u32 n = sg.count[s];
  if (n >= nmax) return;
  const u32 S = sg.start[s];
  const u32* pp = perm + S;
  ExBest best;
  best.key = 0.f;
  best.mean = 0.f;
  best.idx = 0x7fffffff;
  for (int c0 = 0; c0 < dims; c0 += 32 * CHX)
  {
    float mean[CHX], q[CHX];
    {
      const float* rp = rows + (size_t)pp[0] * ld;
#pragma unroll
      for (int k = 0; k < CHX; ++k)
      {
        const int c = c0 + k * 32 + lane;
        mean[k] = (c < dims) ? ldg_f_stream(rp + c) : 0.f;  // InitStats, IndexBuilder.cs:159-173
        q[k] = 0.f;
      }
    }
    u32 mine = (lane < n) ? pp[lane] : 0u;
    for (u32 jb = 0; jb < n; jb += 32)
    {
      const u32 nxt = (jb + 32 + lane < n) ? pp[jb + 32 + lane] : 0u;
      const u32 m = min(32u, n - jb);
      const float cmine = (float)(jb + lane + 1u);  // (float)(Count + 1), IndexBuilder.cs:185-186
      const float rmine = __frcp_rn(cmine);
#pragma unroll 4
      for (u32 jj = (jb == 0 ? 1u : 0u); jj < m; ++jj)
      {
        const u32 r = __shfl_sync(0xffffffffu, mine, jj);
        const float cnt = __shfl_sync(0xffffffffu, cmine, jj);
        const float rc = __shfl_sync(0xffffffffu, rmine, jj);
        const float* rp = rows + (size_t)r * ld;
        float v[CHX];
#pragma unroll
        for (int k = 0; k < CHX; ++k)
        {
          const int c = c0 + k * 32 + lane;
          v[k] = (c < dims) ? ldg_f_stream(rp + c) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < CHX; ++k) welford_step_r(mean[k], q[k], v[k], cnt, rc);
      }
      mine = nxt;
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
  const i64 pivot = range_mean_id<32>(pid + S, n, lane, 0xffffffffu);
  if (lane == 0) write_split(sg, out, s, best.idx, best.mean, pivot);
}

constexpr int EXU = 32;   // rows per prefetch group
constexpr int EXNG_DEFAULT = 6;  // groups in flight (192 rows ahead of the recurrence)

// Speculative step of the big-range kernel: 4 dependent operations per point instead of 7.
//   R = r_hi + r_lo ~ 1/c to 48 bits (r_hi = RN(1/c), r_lo = RN-ish(1/c - r_hi));
//   q0 = RN(d*r_hi + RN(d*r_lo)) is RN(d/c) unless d/c lies within ~2^-47 (relative) of a rounding boundary.
// Off the dependency chain, one Markstein correction q1 = RN(q0 + (d - q0*c)*r_hi) checks it: q1 == q0 implies
// |d/c - q0| <= (1/2 + 2^-25) ulp, and since d/c stays at least 1/(2C) ulp away from a rounding boundary (C = the
// 24-bit significand of c), q0 is then the correctly rounded quotient unless C is all ones -- groups holding such a
// count, operands outside [2^-60, 2^60] (not 0) and failed checks (about 2^-23 of the steps) redo their 32 steps
// with div_by_count from the saved state.
struct CountRcp  // per row of a group: r_hi, r_lo, c = (float)(Count + 1)
{
  float r_hi, r_lo, c, pad;
};

__device__ __forceinline__ CountRcp count_rcp(u32 count_plus_1)
{
  CountRcp t;
  t.c = (float)count_plus_1;  // (float)(Count + 1), IndexBuilder.cs:185-186
  t.r_hi = __frcp_rn(t.c);
  t.r_lo = __fmul_rn(__fmaf_rn(-t.r_hi, t.c, 1.f), t.r_hi);
  t.pad = 0.f;
  return t;
}

__device__ __forceinline__ bool count_all_ones(float c) { return (__float_as_uint(c) & 0x7fffffu) == 0x7fffffu; }

struct SpecCheck
{
  bool neq;         // some q1 != q0
  u32 umax, umin1;  // max |d| bits, min (|d| bits - 1) (a zero wraps to 0xffffffff and never lowers the minimum)
  __device__ __forceinline__ void reset() { neq = false; umax = 0u; umin1 = 0xffffffffu; }
  __device__ __forceinline__ bool bad() const
  {
    return neq || umax > 0x5d800000u /* 2^60 */ || umin1 < 0x21800000u - 1u /* 2^-60 */;
  }
};

__device__ __forceinline__ void welford_step_spec(float& mean, float& q, float value, const CountRcp& k, SpecCheck& chk)
{
  const float d = __fsub_rn(value, mean);
  const float q0 = __fmaf_rn(d, k.r_hi, __fmul_rn(d, k.r_lo));
  const float a = __fadd_rn(mean, q0);
  const float q1 = __fmaf_rn(__fmaf_rn(-q0, k.c, d), k.r_hi, q0);
  chk.neq |= !(q1 == q0);
  const u32 ub = __float_as_uint(d) & 0x7fffffffu;
  chk.umax = max(chk.umax, ub);
  chk.umin1 = min(chk.umin1, ub - 1u);
  q = __fadd_rn(q, __fmul_rn(d, __fsub_rn(value, a)));
  mean = a;
}

// One warp per (range, 32 dimensions), one chain per lane.  Rows are prefetched with cp.async into a per-warp
// shared-memory ring (EXNG groups of EXU rows) -- register prefetch would not do: a warp has 6 scoreboard slots, and
// waiting for the oldest of 64 outstanding loads also waits for the youngest.
//   VEC:  lane u copies the 128-byte slice of row u with eight 16-byte cp.async (chunk k lands at chunk k ^ (u & 7):
//         conflict-free for the copies and for the reads); needs ld % 4 == 0 and a 16-byte aligned base.
//   !VEC: a lane copies its own 4 bytes of every row.
template <bool VEC, int EXNG>
__global__ void __launch_bounds__(32)
k_stats_big_exact(SegLevel sg, const u32* __restrict__ big_list, u32 nblk, const u32* __restrict__ perm,
                  const float* __restrict__ rows, int ld, int dims, float2* __restrict__ gstats)
{
  const u32 slot = blockIdx.x / nblk;
  const int col0 = (int)(blockIdx.x % nblk) * 32;
  const int lane = threadIdx.x;
  const int col = col0 + lane;
  const u32 s = big_list[slot];
  const u32 S = sg.start[s], n = sg.count[s];
  const bool act = col < dims;
  const u32* pp = perm + S;

  __shared__ __align__(16) float ring[EXNG][EXU * 32];
  __shared__ __align__(16) CountRcp tab[2][EXU];
  __shared__ __align__(16) float first[32];  // this block's 32 dimensions of the range's first row
  __shared__ u32 pring[EXNG][EXU];           // slot g % EXNG: row indexes of group g + EXNG (they travel with group g)
  // word offset of this lane's dimension inside row u of a group
  auto ring_off = [&](int u) -> int
  { return VEC ? u * 32 + ((((lane >> 2) ^ (u & 7)) << 2) | (lane & 3)) : u * 32 + lane; };

  // No global load may stay pending across the loop: ptxas has 6 scoreboards, a load shares one with the cp.async
  // group counter or with the loop-carried state, and the wait it plants at the loop top then drains everything.
  // So the first row and the row indexes go through cp.async as well.
  {
    const float* src = rows + (size_t)pp[0] * ld + col0;
    if (VEC)
    {
      if (lane < 8 && col0 + 4 * lane + 4 <= ld)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((u32)__cvta_generic_to_shared(&first[lane * 4])), "l"(src + 4 * lane)
                     : "memory");
    }
    else if (act)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((u32)__cvta_generic_to_shared(&first[lane])), "l"(src + lane) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  auto issue = [&](u32 g, u32 pv)  // group g = rows 1 + g*EXU .. ; always commits (uniform group counting)
  {
    const u32 j0 = 1u + g * EXU;
    {
      const u32 jn = j0 + EXNG * EXU + lane;  // row index of group g + EXNG
      if (jn < n)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((u32)__cvta_generic_to_shared(&pring[g % EXNG][lane])), "l"(pp + jn)
                     : "memory");
    }
    if (VEC)
    {
      if (j0 + lane < n)
      {
        const float* src = rows + (size_t)pv * ld + col0;
        const u32 dst = (u32)__cvta_generic_to_shared(&ring[g % EXNG][lane * 32]);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (col0 + 4 * k + 4 <= ld)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (u32)((k ^ (lane & 7)) << 4)), "l"(src + 4 * k)
                         : "memory");
      }
    }
    else
    {
      const float* base = rows + (act ? col : 0);
      const u32 dst = (u32)__cvta_generic_to_shared(&ring[g % EXNG][lane]);
#pragma unroll
      for (int u = 0; u < EXU; ++u)
      {
        const u32 r = __shfl_sync(0xffffffffu, pv, u);
        if (j0 + u < n)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + u * 128), "l"(base + (size_t)r * ld) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const u32 ngroups = (n - 1 + EXU - 1) / EXU;
#pragma unroll
  for (int g = 0; g < EXNG; ++g)
  {
    const u32 j = 1u + g * EXU + lane;
    issue(g, (j < n) ? pp[j] : 0u);
  }
  asm volatile("cp.async.wait_group %0;" ::"n"(EXNG) : "memory");
  __syncwarp();
  float mean = first[lane], q = 0.f;  // InitStats, IndexBuilder.cs:159-173
  for (u32 g = 0; g < ngroups; ++g)
  {
    const u32 j0 = 1u + g * EXU;
    const CountRcp mine = count_rcp(j0 + lane + 1u);
    tab[g & 1][lane] = mine;
    const bool ones = __any_sync(0xffffffffu, count_all_ones(mine.c));
    asm volatile("cp.async.wait_group %0;" ::"n"(EXNG - 1) : "memory");
    __syncwarp();  // the group's rows (copied by other lanes when VEC) and tab[] are visible
    float v[EXU];
#pragma unroll
    for (int u = 0; u < EXU; ++u) v[u] = ring[g % EXNG][ring_off(u)];
    const u32 pnext = (j0 + EXNG * EXU + lane < n) ? pring[g % EXNG][lane] : 0u;
    __syncwarp();  // every lane has drained the slot
    issue(g + EXNG, pnext);
    const CountRcp* tb = tab[g & 1];
    const float m0 = mean, q0 = q;
    bool redo = true;
    if (j0 + EXU <= n && !ones)
    {
      SpecCheck chk;
      chk.reset();
#pragma unroll
      for (int u = 0; u < EXU; ++u) welford_step_spec(mean, q, v[u], tb[u], chk);
      redo = __any_sync(0xffffffffu, act && chk.bad());
    }
    if (redo)
    {
      mean = m0;
      q = q0;
#pragma unroll  // (static indexes keep v in registers)
      for (int u = 0; u < EXU; ++u)
        if (j0 + u < n) welford_step_r(mean, q, v[u], tb[u].c, tb[u].r_hi);
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (act) gstats[(size_t)slot * dims + col] = make_float2(mean, q);
}

// Id sums of the big ranges (Stats.IdN, IndexBuilder.cs:170,194), `per` CTAs per range: integer sums are order-free,
// so they need not ride on the serial chains (one warp summing 10^7 ids costs more than a tenth of the level).
__global__ void __launch_bounds__(256)
k_idsum_big(SegLevel sg, const u32* __restrict__ big_list, u32 per, const i64* __restrict__ pid, u64* __restrict__ idacc)
{
  const u32 slot = blockIdx.x / per, b = blockIdx.x % per;
  const u32 s = big_list[slot];
  const u32 S = sg.start[s], n = sg.count[s];
  u64 slo = 0;
  i64 shi = 0;
  for (u32 j = b * 256u + threadIdx.x; j < n; j += per * 256u)
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
  __shared__ u64 w_lo[8];
  __shared__ i64 w_hi[8];
  if ((threadIdx.x & 31) == 0)
  {
    w_lo[threadIdx.x >> 5] = slo;
    w_hi[threadIdx.x >> 5] = shi;
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    for (int w = 1; w < 8; ++w)
    {
      slo += w_lo[w];
      shi += w_hi[w];
    }
    if (slo | (u64)shi)
    {
      atomicAdd(&idacc[2 * slot], slo);
      atomicAdd(&idacc[2 * slot + 1], (u64)shi);
    }
  }
}

__global__ void __launch_bounds__(256)
k_finalize_big_exact(SegLevel sg, const u32* __restrict__ big_list, u32 nbig, const float2* __restrict__ gstats,
                     const u64* __restrict__ idacc, int dims, int mx, StatsOut out)
{
  const u32 warp = (blockIdx.x * 256u + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= nbig) return;
  const u32 s = big_list[warp];
  const u32 n = sg.count[s];
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
  if (lane == 0) write_split(sg, out, s, best.idx, best.mean, mean_id(idacc[2 * warp], (i64)idacc[2 * warp + 1], n));
}

// debug: compares div_by_count with __fdiv_rn on pseudo-random operands; returns the mismatch count
__global__ void k_divcheck(u64 seed, u64 per_thread, unsigned long long* mismatches)
{
  u64 x = seed ^ ((u64)(blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull);
  unsigned long long bad = 0;
  for (u64 i = 0; i < per_thread; ++i)
  {
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;  // xorshift64
    const u32 cbits = (u32)(x >> 40);          // 24-bit counts, plus some larger ones
    u32 ci = (i & 7) == 0 ? (u32)(x >> 33) : cbits;
    if (ci < 2) ci = 2;
    const float c = (float)ci;
    float d;
    switch ((x >> 8) & 3)
    {
      case 0: d = __uint_as_float((u32)x); break;                                   // any bit pattern
      case 1: d = __uint_as_float(((u32)x & 0x807fffffu) | (((u32)(x >> 32) % 60 + 97) << 23)); break;  // 2^-30..2^29
      case 2: d = __uint_as_float(((u32)x & 0x807fffffu) | (((u32)(x >> 32) % 16 + 120) << 23)); break; // near 1
      default: d = (float)(int)(x >> 20) * 0x1p-20f; break;
    }
    if ((i & 1023) == 5) d = 0.f;
    if ((i & 1023) == 6) d = -0.f;
    const float a = div_by_count(d, c, __frcp_rn(c));
    const float b = __fdiv_rn(d, c);
    if (__float_as_uint(a) != __float_as_uint(b) && !(isnan(a) && isnan(b))) ++bad;
    // the speculative step of k_stats_big_exact: whenever its check passes, mean + q0 must equal mean + d / c
    const CountRcp k = count_rcp(ci);
    if (!count_all_ones(k.c))
    {
      const float mean0 = __uint_as_float((u32)(x >> 13) & 0xbfffffffu);  // any finite-ish mean; value = mean + d
      float mean = mean0, q = 0.f, mean2 = mean0, q2 = 0.f;
      const float value = __fadd_rn(mean0, d);
      SpecCheck chk;
      chk.reset();
      welford_step_spec(mean, q, value, k, chk);
      welford_step(mean2, q2, value, k.c);
      if (!chk.bad() && (__float_as_uint(mean) != __float_as_uint(mean2) || __float_as_uint(q) != __float_as_uint(q2)))
        ++bad;
    }
  }
  if (bad) atomicAdd(mismatches, bad);
}
