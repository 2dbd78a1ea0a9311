// vi_subtree.cuh -- the bottom of the tree, fast mode.  A range with at most T (<= 32) points leaves the level loop
// (k_children / k_scatter put it on the sub-tree list) and ONE WARP builds its whole sub-tree in shared memory: the
// rows are read from HBM once (cp.async, 16 B per lane) instead of once per remaining level, and quantised once,
// xi = rint(x * 2^(26-E)), in place into an int32 copy.  The sub-tree is processed level by level; the warp's four
// 8-lane teams take four of the level's nodes at a time (a lane owns CH int4 columns; 64-bit keys).
//
// Same arithmetic as the level kernels (vi_stats_fast.cuh): exact integer sums S1, S2 of xi, key K = n*S2 - S1^2,
// arg-max / arg-min by depth parity with lowest-index ties, Mid = float((double)S1/n * 2^(E-26)), float32 Welford
// fallback (from the raw rows) for a poorly resolved range, Id = trunc(sum id / n), stable partition with the
// reference's predicate (IndexBuilder.cs:115).  The predicate is evaluated on the quantised value where that decides
// it: with M = Mid * 2^(26-E) and Mi = floor(M), xi >= Mi + 2 implies value > Mid and xi <= Mi - 1 implies
// value < Mid (|value * 2^(26-E) - xi| <= 1/2); the two values in between, NaN-carrying sub-trees and Welford ranges
// read the raw float.
//
// Rows: the sub-tree root's row exists already (its parent's level made it).  A sub-tree of n points has exactly
// 2n-2 further rows when every split has two non-empty sides; they go to a dense block at
// row_base + 2*sub_start - 2*sub_index in closed form (sub_children below).  A one-sided split adds a row: it is
// taken from an overflow area behind all blocks (atomic counter), so the table stays dense.
//
// (A CTA-per-sub-tree variant for ranges of up to 512 points was built and measured in round 2: correct, not faster;
// profiles/r2_subtree_cta_experiment.md.)
#pragma once
#include "vi_partition.cuh"
#include "vi_stats_exact.cuh"
#include "vi_stats_fast.cuh"

constexpr int SUB_WARPS = 8;          // warps per CTA, one sub-tree each at a time
constexpr int SUB_TMAX = 32;          // most points of a sub-tree (one point per lane when loading)
constexpr int SUB_NODES = 16;         // nodes with >= 2 points on one level of a sub-tree of <= 32 points

struct SubNode
{
  i64 rid;
  u32 row;
  unsigned short start, count;
};
static_assert(sizeof(SubNode) == 16, "SubNode");

// dynamic shared memory per warp: T quantised rows, ids, row indexes, two order lists, two node lists
__host__ __device__ inline size_t sub_smem_bytes_per_warp(int T, int ld)
{
  return (size_t)T * ((size_t)ld * 4 + 8 + 4 + 4) + 2 * SUB_NODES * 16;
}

// The kernel's dynamic shared memory.  SubCtx keeps BYTE OFFSETS into it, not pointers: a pointer that travels through
// a struct loses its address space and every access through it becomes a generic load (measured: 3x the latency of
// the whole kernel); an address formed from s_dyn itself stays a shared-memory access.
extern __shared__ int4 s_dyn[];

struct SubCtx
{
  u32 o_x, o_ids, o_perm, o_ord[2], o_nodes[2];  // byte offsets of this warp's x[T][XS], ids[T], perm[T], ord[2][T], nodes[2][16]
  const float* rows;
  int ld, dims, C4, XS;
  float qk;
  double qinv;
  bool exact_all;          // the sub-tree holds a NaN: every predicate reads the raw float
};

__device__ __forceinline__ char* sub_base() { return reinterpret_cast<char*>(s_dyn); }
__device__ __forceinline__ int4* sub_x(const SubCtx& c) { return reinterpret_cast<int4*>(sub_base() + c.o_x); }  // [T][XS] quantised rows
__device__ __forceinline__ i64* sub_ids(const SubCtx& c) { return reinterpret_cast<i64*>(sub_base() + c.o_ids); }
__device__ __forceinline__ u32* sub_perm(const SubCtx& c) { return reinterpret_cast<u32*>(sub_base() + c.o_perm); }
__device__ __forceinline__ unsigned short* sub_ord(const SubCtx& c, int b)
{
  return reinterpret_cast<unsigned short*>(sub_base() + c.o_ord[b]);
}

struct SubSplit
{
  int dim;
  float mid;
  i64 pivot;
  int tlo, thi;
  int exact;
};

__device__ __forceinline__ SubSplit sub_make_split(const SubCtx& c, int dim, float mid, i64 pivot, bool exact)
{
  SubSplit sp;
  sp.dim = dim;
  sp.mid = mid;
  sp.pivot = pivot;
  const float M = __fmul_rn(mid, c.qk);
  const int Mi = __float2int_rd(M);
  sp.tlo = Mi - 1;
  sp.thi = Mi + 2;
  sp.exact = (exact || c.exact_all || !(fabsf(M) < 1.0e9f)) ? 1 : 0;  // NaN / Inf Mid (Welford ranges only): raw floats
  return sp;
}

// IndexBuilder.cs:115 for point `pt` of the sub-tree
__device__ __forceinline__ bool sub_hi(const SubCtx& c, const SubSplit& sp, u32 pt)
{
  const int xi = reinterpret_cast<const int*>(sub_x(c))[pt * (u32)c.XS * 4u + (u32)sp.dim];
  if (!sp.exact)
  {
    if (xi >= sp.thi) return true;
    if (xi <= sp.tlo) return false;
  }
  const float v = c.rows[(size_t)sub_perm(c)[pt] * c.ld + sp.dim];
  return v > sp.mid || (v == sp.mid && sub_ids(c)[pt] > sp.pivot);
}

__device__ __forceinline__ void iacc(i64& s1, u64& s2, int xi)
{
  asm("mad.wide.s32 %0, %1, 1, %0;" : "+l"(s1) : "r"(xi));
  asm("mad.wide.s32 %0, %1, %1, %0;" : "+l"(s2) : "r"(xi));
}

__device__ __forceinline__ void iacc4(i64* s1, u64* s2, const int4& x)
{
  iacc(s1[0], s2[0], x.x);
  iacc(s1[1], s2[1], x.y);
  iacc(s1[2], s2[2], x.z);
  iacc(s1[3], s2[3], x.w);
}

// The reference's float32 recurrence over the node's points in order, from the raw rows (global memory; the shared
// copy is quantised), by a group of G lanes: the fallback of a poorly resolved range.  Lane g of the group takes
// dims g, g+G, ... in blocks of DB dims; the raw values of PB points x DB dims are fetched with independent loads
// before the dependent recurrence runs over them (one memory round trip per block instead of one per step).
template <int G>
__device__ __noinline__ ExBest sub_welford(const SubCtx& c, const unsigned short* ord, u32 m, int g, u32 gmask, bool mx)
{
  constexpr int DB = 4, PB = 8;
  ExBest eb;
  eb.key = 0.f;
  eb.mean = 0.f;
  eb.idx = 0x7fffffff;
  for (int d0 = g; d0 < c.dims; d0 += G * DB)
  {
    float mean[DB], q[DB];
#pragma unroll
    for (int k = 0; k < DB; ++k) { mean[k] = 0.f; q[k] = 0.f; }
    for (u32 i0 = 0; i0 < m; i0 += PB)
    {
      float v[PB][DB];
#pragma unroll
      for (int p = 0; p < PB; ++p)
      {
        const u32 i = i0 + p;
        const float* rp = c.rows + (size_t)sub_perm(c)[ord[i < m ? i : 0]] * c.ld;
#pragma unroll
        for (int k = 0; k < DB; ++k)
        {
          const int d = d0 + k * G;
          v[p][k] = (i < m && d < c.dims) ? rp[d] : 0.f;
        }
      }
#pragma unroll
      for (int p = 0; p < PB; ++p)
      {
        const u32 i = i0 + p;
        if (i < m)
        {
#pragma unroll
          for (int k = 0; k < DB; ++k)
          {
            if (i == 0) { mean[k] = v[p][k]; q[k] = 0.f; }  // InitStats, IndexBuilder.cs:159-173
            else welford_step(mean[k], q[k], v[p][k], (float)(i + 1u));
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < DB; ++k)
    {
      const int d = d0 + k * G;
      if (d < c.dims)
      {
        const float key = mx ? q[k] : -q[k];
        if (ex_better(key, d, eb.key, eb.idx))
        {
          eb.key = key;
          eb.mean = mean[k];
          eb.idx = d;
        }
      }
    }
  }
  return ex_reduce_w<G>(eb, gmask);
}

// ---- children of a split node ------------------------------------------------------------------------------------------
// Row numbering inside a sub-tree of n points whose block starts at row B: a two-sided split at boundary b (the first
// position of the high child in the sub-tree's order, 1 <= b <= n-1) gives its low child row B + 2b - 2 and its high
// child row B + 2b - 1.  Every boundary belongs to exactly one two-sided split, so the 2n-2 rows are used exactly
// once, without any scan or counter.  The only child of a one-sided split takes a row from the overflow area.
struct SubEnv
{
  TableOut t;
  int* t_src;
  u32 B;              // first row of this sub-tree's block
  u32 overflow_base;
  u32 t_cap;
  u32 root_depth;     // depth of the sub-tree's root: a node's order list is ord[(depth - root_depth) & 1]
  int sql;            // VI_MODE_SQL: min variance at depth 1 only, null Dimension / Mid when Stdev = 0
  u32* counters;      // global: [0] overflow rows used, [1] error, [2] float32 fallbacks, [3] work cursor
  u32* lvlp;          // shared: points / ranges per depth (accounting)
  u32* lvlr;
};

__device__ __forceinline__ u32 sub_depth(i64 rid) { return 63u - (u32)__clzll((long long)(rid + 1)); }

// Writes the rows of P's children (and P's links); returns through lo/hi the children that are ranges themselves
// (count >= 2), with count 0 otherwise.  One thread per node calls this after the partition is visible.
__device__ __forceinline__ void sub_children(const SubCtx& c, const SubEnv& e, const SubNode& P, u32 nlo,
                                             const unsigned short* nord, SubNode& lo, SubNode& hi)
{
  const u32 nhi = (u32)P.count - nlo;
  const u32 b = (u32)P.start + nlo;
  lo.count = 0;
  hi.count = 0;
  int lo_row = -1, hi_row = -1;
  if (nlo > 0 && nhi > 0)
  {
    lo_row = (int)(e.B + 2u * b - 2u);
    hi_row = lo_row + 1;
  }
  else
  {
    const u32 r = e.overflow_base + atomicAdd(&e.counters[0], 1u);
    if (r >= e.t_cap)
    {
      e.counters[1] = 1u;
      e.t.t_low[P.row] = -1;
      e.t.t_high[P.row] = -1;
      return;
    }
    if (nlo > 0) lo_row = (int)r; else hi_row = (int)r;
  }
  e.t.t_low[P.row] = lo_row;
  e.t.t_high[P.row] = hi_row;
  if (nlo > 0)
  {
    e.t.t_rid[lo_row] = P.rid * 2 + 1;  // IndexBuilder.cs:99
    e.t.t_low[lo_row] = -1;
    e.t.t_high[lo_row] = -1;
    if (nlo == 1)
    {
      const unsigned short p1 = nord[P.start];
      e.t.t_dim[lo_row] = -1;  // leaf (IndexBuilder.cs:81-82)
      e.t.t_mid[lo_row] = 0.f;
      e.t.t_id[lo_row] = sub_ids(c)[p1];
      e.t_src[lo_row] = (int)sub_perm(c)[p1];
    }
    else
    {
      lo.rid = P.rid * 2 + 1;
      lo.row = (u32)lo_row;
      lo.start = P.start;
      lo.count = (unsigned short)nlo;
    }
  }
  if (nhi > 0)
  {
    e.t.t_rid[hi_row] = P.rid * 2 + 2;  // IndexBuilder.cs:104
    e.t.t_low[hi_row] = -1;
    e.t.t_high[hi_row] = -1;
    if (nhi == 1)
    {
      const unsigned short p1 = nord[b];
      e.t.t_dim[hi_row] = -1;
      e.t.t_mid[hi_row] = 0.f;
      e.t.t_id[hi_row] = sub_ids(c)[p1];
      e.t_src[hi_row] = (int)sub_perm(c)[p1];
    }
    else
    {
      hi.rid = P.rid * 2 + 2;
      hi.row = (u32)hi_row;
      hi.start = (unsigned short)b;
      hi.count = (unsigned short)nhi;
    }
  }
}

// Int128-exact (long)(IdN / Count) with a short path for the common case (non-negative sum below 2^32)
__device__ __forceinline__ i64 sub_mean_id(u64 slo, i64 shi, u32 n)
{
  if (shi == 0 && slo < 0x100000000ull) return (i64)((u32)slo / n);
  return mean_id(slo, shi, n);
}

// ---- 2 .. 32 points: one 8-lane team ---------------------------------------------------------------------------------
// Split of node nd (its points are ord[0..m)); partitions into dst[0..m); returns the low child's size.  All lanes of
// the team return the same value.  The node's row (Dimension, Mid, Id) is written by lane 0.
template <int CH, bool FULL>
__device__ __forceinline__ u32 sub_team_split(const SubCtx& c, const SubEnv& env, const SubNode& nd, const unsigned short* ord,
                                              unsigned short* dst, bool mx, int tl, u32 tmask, int tshift)
{
  const u32 m = nd.count;
  const int C4 = FULL ? 8 * CH : c.C4;
  u64 bkey = 0;
  int bs1 = 0, bidx = 0x7fffffff;
  for (int c0 = 0; c0 < C4; c0 += 8 * CH)
  {
    i64 s1[CH * 4];
    u64 s2[CH * 4];
#pragma unroll
    for (int i = 0; i < CH * 4; ++i) { s1[i] = 0; s2[i] = 0; }
    for (u32 i = 0; i < m; ++i)
    {
      const int4* rp = sub_x(c) + (u32)ord[i] * (u32)c.XS + (u32)(c0 + tl);
#pragma unroll
      for (int k = 0; k < CH; ++k)
        if (FULL || c0 + k * 8 + tl < C4) iacc4(s1 + k * 4, s2 + k * 4, rp[k * 8]);
    }
    // keys in increasing dimension order: a strict comparison keeps the lowest index among equal keys
#pragma unroll
    for (int k = 0; k < CH; ++k)
#pragma unroll
      for (int e = 0; e < 4; ++e)
      {
        const int d = (c0 + k * 8 + tl) * 4 + e;
        if (FULL || d < c.dims)
        {
          const int a = (int)s1[k * 4 + e];  // |S1| <= 32 * (2^26 - 4) < 2^31
          const u64 key = (u64)m * s2[k * 4 + e] - (u64)((i64)a * (i64)a);  // m <= 32: below 2^63, exact
          const bool take = bidx == 0x7fffffff || (mx ? key > bkey : key < bkey);
          if (take)
          {
            bkey = key;
            bs1 = a;
            bidx = d;
          }
        }
      }
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1)
  {
    const u64 k2 = __shfl_xor_sync(tmask, bkey, o);
    const int a2 = __shfl_xor_sync(tmask, bs1, o);
    const int i2 = __shfl_xor_sync(tmask, bidx, o);
    const bool take = i2 != 0x7fffffff && (bidx == 0x7fffffff || (k2 != bkey ? (mx ? k2 > bkey : k2 < bkey) : i2 < bidx));
    if (take)
    {
      bkey = k2;
      bs1 = a2;
      bidx = i2;
    }
  }
  int dim = bidx;
  float mid = qfx_mid((i64)bs1, m, c.qinv);
  bool welford = false, null_dim = false;
  if (bkey < (((u64)m * (u64)m) << (2 * VI_QFX_MIN_RES_BITS)))  // poorly resolved (team-uniform): float32 statistics
  {
    if (tl == 0) atomicAdd(&env.counters[2], 1u);
    const ExBest eb = sub_welford<8>(c, ord, m, tl, tmask, mx);
    dim = eb.idx;
    mid = eb.mean;
    welford = true;
    null_dim = env.sql != 0 && eb.key == 0.f;  // Stdev = 0 (DDL.sql:193-194)
  }
  // pivot id and stable partition; lane tl looks after points tl, tl+8, ... (one round for m <= 8, else four)
  u32 nlo;
  i64 pivot;
  if (m <= 8)
  {
    const bool have = (u32)tl < m;
    const unsigned short pt = have ? ord[tl] : (unsigned short)0;
    const i64 id = have ? sub_ids(c)[pt] : 0;
    u64 slo = (u64)(u32)id;
    i64 shi = id >> 32;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1)
    {
      slo += __shfl_xor_sync(tmask, slo, o);
      shi += __shfl_xor_sync(tmask, shi, o);
    }
    pivot = sub_mean_id(slo, shi, m);
    const SubSplit sp = sub_make_split(c, dim, mid, pivot, welford);
    const bool hi = have && sub_hi(c, sp, pt);
    const u32 hib = (__ballot_sync(tmask, hi) >> tshift) & 0xffu;
    const u32 below = (1u << tl) - 1u;
    nlo = m - __popc(hib);
    if (have) dst[hi ? nlo + __popc(hib & below) : (u32)tl - __popc(hib & below)] = pt;
  }
  else
  {
    unsigned short pt[4];
    u64 slo = 0;
    i64 shi = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
    {
      const u32 i = (u32)(j * 8 + tl);
      pt[j] = i < m ? ord[i] : (unsigned short)0;
      const i64 id = i < m ? sub_ids(c)[pt[j]] : 0;
      slo += (u64)(u32)id;
      shi += id >> 32;
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1)
    {
      slo += __shfl_xor_sync(tmask, slo, o);
      shi += __shfl_xor_sync(tmask, shi, o);
    }
    pivot = sub_mean_id(slo, shi, m);
    const SubSplit sp = sub_make_split(c, dim, mid, pivot, welford);
    u32 hib[4];
    bool hi4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
    {
      hi4[j] = (u32)(j * 8 + tl) < m && sub_hi(c, sp, pt[j]);
      hib[j] = (__ballot_sync(tmask, hi4[j]) >> tshift) & 0xffu;
    }
    nlo = m - (__popc(hib[0]) + __popc(hib[1]) + __popc(hib[2]) + __popc(hib[3]));
    const u32 below = (1u << tl) - 1u;
    u32 hacc = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
    {
      const u32 i = (u32)(j * 8 + tl);
      const u32 hb = hacc + __popc(hib[j] & below);  // hi points before point i
      if (i < m) dst[hi4[j] ? nlo + hb : i - hb] = pt[j];
      hacc += __popc(hib[j]);
    }
  }
  if (tl == 0)
  {
    env.t.t_dim[nd.row] = null_dim ? VI_DIM_NULL : dim;
    env.t.t_mid[nd.row] = null_dim ? __int_as_float(0x7fc00000) : mid;
    env.t.t_id[nd.row] = pivot;
  }
  return nlo;
}

// ---- the kernel --------------------------------------------------------------------------------------------------------
// counters: [0] overflow rows used, [1] error (1 = table capacity, 2 = depth overflow), [2] float32 fallbacks,
//           [3] work cursor (next sub-tree to take)
template <int CH, bool FULL>
__global__ void __launch_bounds__(SUB_WARPS * 32, 2)
k_subtree_fast(SubList sl, u32 nsub, const u32* __restrict__ g_sub_perm, const i64* __restrict__ g_sub_pid,
               const float* __restrict__ rows, int ld, int dims, float qk, double qinv, TableOut t, int* __restrict__ t_src,
               u32 row_base, u32 overflow_base, u32 t_cap, u32* __restrict__ counters,
               unsigned long long* __restrict__ lvl_points, unsigned long long* __restrict__ lvl_ranges, int T, int sql,
               u32 k0)
{
  __shared__ u32 s_lvlp[64], s_lvlr[64];
  __shared__ u32 s_ncnt[SUB_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tl = lane & 7, team = lane >> 3;
  const int tshift = lane & 24;
  const u32 tmask = 0xffu << tshift;

  SubCtx c;
  c.ld = ld;
  c.dims = dims;
  c.C4 = FULL ? 8 * CH : (ld >> 2);
  c.XS = c.C4;
  c.qk = qk;
  c.qinv = qinv;
  c.rows = rows;
  c.exact_all = false;
  {
    u32 o = (u32)warp * (u32)sub_smem_bytes_per_warp(T, ld);
    c.o_x = o;
    o += (u32)T * (u32)c.XS * 16u;
    c.o_ids = o;
    o += (u32)T * 8u;
    c.o_nodes[0] = o;
    o += SUB_NODES * 16u;
    c.o_nodes[1] = o;
    o += SUB_NODES * 16u;
    c.o_perm = o;
    o += (u32)T * 4u;
    c.o_ord[0] = o;
    o += (u32)T * 2u;
    c.o_ord[1] = o;
  }
  if (threadIdx.x < 64) { s_lvlp[threadIdx.x] = 0; s_lvlr[threadIdx.x] = 0; }
  if (lane == 0) s_ncnt[warp] = 0;
  __syncthreads();
  SubEnv env;
  env.t = t;
  env.t_src = t_src;
  env.overflow_base = overflow_base;
  env.t_cap = t_cap;
  env.counters = counters;
  env.sql = sql;
  env.lvlp = s_lvlp;
  env.lvlr = s_lvlr;

  for (;;)
  {
    u32 k = 0;
    if (lane == 0) k = k0 + atomicAdd(&counters[3], 1u);  // this launch takes the sub-trees [k0, nsub)
    k = __shfl_sync(0xffffffffu, k, 0);
    if (k >= nsub) break;
    const u32 S = sl.start[k], n = sl.count[k];
    // ---- load the sub-tree's points: ids / row indexes one per lane, rows with 16-byte cp.async -------------------------
    u32 myrow = 0;
    if ((u32)lane < n)
    {
      myrow = g_sub_perm[S + lane];
      sub_perm(c)[lane] = myrow;
      sub_ids(c)[lane] = g_sub_pid[S + lane];
      sub_ord(c, 0)[lane] = (unsigned short)lane;
    }
    for (u32 j = 0; j < n; ++j)
    {
      const u32 r = __shfl_sync(0xffffffffu, myrow, j);
      const float4* src = reinterpret_cast<const float4*>(rows + (size_t)r * ld);
      for (int col = lane; col < c.C4; col += 32)
      {
        const u32 dsta = (u32)__cvta_generic_to_shared(sub_x(c) + j * (u32)c.XS + (u32)col);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dsta), "l"(src + col) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    env.B = row_base + 2u * S - 2u * k;
    env.root_depth = sl.depth[k];
    if (lane == 0)
    {
      SubNode root;
      root.rid = sl.rid[k];
      root.row = sl.row[k];
      root.start = 0;
      root.count = (unsigned short)n;
      reinterpret_cast<SubNode*>(sub_base() + c.o_nodes[0])[0] = root;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    // quantise in place: every lane converts exactly what it copied (its own cp.async data)
    {
      bool nan = false;
      for (u32 j = 0; j < n; ++j)
        for (int col = lane; col < c.C4; col += 32)
        {
          int4* px = sub_x(c) + j * (u32)c.XS + (u32)col;
          const float4 v = *reinterpret_cast<const float4*>(px);
          nan = nan || (v.x != v.x) || (v.y != v.y) || (v.z != v.z) || (v.w != v.w);
          int4 qv;
          qv.x = __float2int_rn(__fmul_rn(v.x, qk));
          qv.y = __float2int_rn(__fmul_rn(v.y, qk));
          qv.z = __float2int_rn(__fmul_rn(v.z, qk));
          qv.w = __float2int_rn(__fmul_rn(v.w, qk));
          *px = qv;
        }
      c.exact_all = __any_sync(0xffffffffu, nan) != 0;
    }
    __syncwarp();

    u32 ncur = 1;
    int cur = 0;
    u32 depth = env.root_depth;
    while (ncur > 0)
    {
      if (depth >= (u32)VI_MAX_DEPTH)
      {
        if (lane == 0) counters[1] = 2u;  // splitting a depth-62 range: rangeId overflow (IndexBuilder.cs:99)
        break;
      }
      const bool mx = sql ? depth != 1u : (depth & 1u) == 0u;  // DDL.sql:151,155 / IndexBuilder.cs:128-129
      const int ob = (int)((depth - env.root_depth) & 1u);
      const SubNode* nodes = reinterpret_cast<const SubNode*>(sub_base() + c.o_nodes[cur]);
      SubNode* nnodes = reinterpret_cast<SubNode*>(sub_base() + c.o_nodes[cur ^ 1]);
      for (u32 base = 0; base < ncur; base += 4)
      {
        const u32 ni = base + (u32)team;
        if (ni < ncur)
        {
          const SubNode nd = nodes[ni];
          const u32 nlo = sub_team_split<CH, FULL>(c, env, nd, sub_ord(c, ob) + nd.start, sub_ord(c, ob ^ 1) + nd.start, mx, tl,
                                                   tmask, tshift);
          __syncwarp(tmask);  // the partition is visible to lane 0 of the team
          if (tl == 0)
          {
            SubNode lo, hi2;
            sub_children(c, env, nd, nlo, sub_ord(c, ob ^ 1), lo, hi2);
            if (lo.count) nnodes[atomicAdd(&s_ncnt[warp], 1u)] = lo;
            if (hi2.count) nnodes[atomicAdd(&s_ncnt[warp], 1u)] = hi2;
            atomicAdd(&s_lvlp[depth & 63u], (u32)nd.count);
            atomicAdd(&s_lvlr[depth & 63u], 1u);
          }
        }
        __syncwarp();
      }
      ncur = s_ncnt[warp];
      __syncwarp();
      if (lane == 0) s_ncnt[warp] = 0;
      cur ^= 1;
      ++depth;
      __syncwarp();
    }
    __syncwarp();
  }
  __syncthreads();
  if (threadIdx.x < 64)
  {
    if (s_lvlp[threadIdx.x]) atomicAdd(&lvl_points[threadIdx.x], (unsigned long long)s_lvlp[threadIdx.x]);
    if (s_lvlr[threadIdx.x]) atomicAdd(&lvl_ranges[threadIdx.x], (unsigned long long)s_lvlr[threadIdx.x]);
  }
}
