// vi_subtree.cuh -- the bottom of the tree, fast mode.  A range with at most T (<= 512) points leaves the level loop
// (k_children / k_scatter put it on the sub-tree list) and ONE CTA builds its whole sub-tree in shared memory: the rows
// are read from HBM once -- instead of once per remaining level (about 12 levels of a 10M x 96 build) -- and quantised
// once, xi = rint(x * 2^(26-E)), into a shared int32 copy (cp.async of the raw rows, converted in place).
//
// The sub-tree is processed level by level.  Nodes of a level are handled by size class:
//   > 128 points   the whole CTA, one node at a time: a warp owns int4 columns, its lanes own points (shuffle sums)
//   33 .. 128      one warp per node: a lane owns an int4 column and walks the node's points
//   2 .. 32        one 8-lane team per node: a lane owns CH int4 columns (64-bit keys)
// then one pass (a thread per node) numbers the children's rows, writes leaf rows and the next level's node list.
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
// row_base + 2*sub_start - 2*sub_index, breadth-first inside the sub-tree.  A one-sided split adds a row: it is
// taken from an overflow area behind all blocks (atomic counter), so the table stays dense.
#pragma once
#include "vi_partition.cuh"
#include "vi_stats_exact.cuh"
#include "vi_stats_fast.cuh"

#ifndef VI_SUB_NT
#define VI_SUB_NT 512
#endif
constexpr int SUB_NT = VI_SUB_NT;     // threads per CTA
constexpr int SUB_NW = SUB_NT / 32;
constexpr int SUB_TEAMS = SUB_NT / 8;
constexpr int SUB_TMAX = SUB_NT;      // most points of a sub-tree (16-bit indexes; one thread per point of a CTA-wide node)
constexpr int SUB_WARP_MAX = 128;     // largest node one warp handles alone
#ifndef VI_SUB_TEAM_MAX
#define VI_SUB_TEAM_MAX 8
#endif
constexpr int SUB_TEAM_MAX = VI_SUB_TEAM_MAX;  // largest node an 8-lane team handles (64-bit keys; <= 32)
static_assert(SUB_TMAX <= SUB_NT, "the CTA-wide partition gives every point of a node its own thread");

struct SubNode
{
  i64 rid;
  u32 row;
  unsigned short start, count;
};
static_assert(sizeof(SubNode) == 16, "SubNode");

// dynamic shared memory per point: XS int4 of row (XS = ld/4 + 1: the pad spreads rows over the banks for the
// column-wise reads of the CTA-wide class), id, one node-queue entry, row index, two order entries
__host__ __device__ inline size_t sub_smem_bytes(int T, int ld) { return (size_t)T * ((size_t)(ld / 4 + 1) * 16 + 32); }

// The kernel's dynamic shared memory.  SubCtx keeps BYTE OFFSETS into it, not pointers: a pointer that travels through
// a struct loses its address space and every access through it becomes a generic load (measured: 3x the latency of
// the whole kernel); an address formed from s_dyn itself stays a shared-memory access.
extern __shared__ int4 s_dyn[];

struct SubCtx
{
  u32 o_ids, o_queue, o_perm, o_ord[2];  // byte offsets of ids[T], queue[T], perm[T], ord[2][T]; the rows start at 0
  const float* rows;
  int ld, dims, C4, XS;
  float qk;
  double qinv;
  bool exact_all;          // the sub-tree holds a NaN: every predicate reads the raw float
};

__device__ __forceinline__ char* sub_base() { return reinterpret_cast<char*>(s_dyn); }
__device__ __forceinline__ int4* sub_x(const SubCtx&) { return s_dyn; }                     // [T][XS] quantised rows
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
  bool welford = false;
  if (bkey < (((u64)m * (u64)m) << (2 * VI_QFX_MIN_RES_BITS)))  // poorly resolved (team-uniform): float32 statistics
  {
    if (tl == 0) atomicAdd(&env.counters[2], 1u);
    const ExBest eb = sub_welford<8>(c, ord, m, tl, tmask, mx);
    dim = eb.idx;
    mid = eb.mean;
    welford = true;
  }
  // pivot id and stable partition; lane tl looks after points tl, tl+8, ... (one round for m <= 8, else four)
  u32 nlo;
  i64 pivot;
  if (SUB_TEAM_MAX <= 8 || m <= 8)
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
    env.t.t_dim[nd.row] = dim;
    env.t.t_mid[nd.row] = mid;
    env.t.t_id[nd.row] = pivot;
  }
  return nlo;
}

// ---- 33 .. 128 points: one warp, a lane owns int4 columns ------------------------------------------------------------
__device__ __forceinline__ u32 sub_warp_split(const SubCtx& c, const SubEnv& env, const SubNode& nd, const unsigned short* ord,
                                              unsigned short* dst, bool mx, int lane)
{
  const u32 m = nd.count;
  QfxBest best;
  best.key.hi = 0;
  best.key.lo = 0;
  best.s1 = 0;
  best.idx = 0x7fffffff;
  for (int c0 = 0; c0 < c.C4; c0 += 32)
  {
    const int col = c0 + lane;
    i64 s1[4] = {0, 0, 0, 0};
    u64 s2[4] = {0, 0, 0, 0};
    if (col < c.C4)
    {
      const int4* xp = sub_x(c) + col;
#pragma unroll 4
      for (u32 i = 0; i < m; ++i) iacc4(s1, s2, xp[(u32)ord[i] * (u32)c.XS]);
#pragma unroll
      for (int e = 0; e < 4; ++e)
      {
        const int d = col * 4 + e;
        if (d < c.dims)
        {
          const Key128 key = qfx_key(m, s1[e], s2[e], 0ull);
          if (qfx_better(mx, key, d, best.key, best.idx))
          {
            best.key = key;
            best.s1 = s1[e];
            best.idx = d;
          }
        }
      }
    }
  }
  best = qfx_reduce<32>(best, mx, 0xffffffffu);
  int dim = best.idx;
  float mid = qfx_mid(best.s1, m, c.qinv);
  bool welford = false;
  if (key_lt(best.key, qfx_threshold(m)))
  {
    if (lane == 0) atomicAdd(&env.counters[2], 1u);
    const ExBest eb = sub_welford<32>(c, ord, m, lane, 0xffffffffu, mx);
    dim = eb.idx;
    mid = eb.mean;
    welford = true;
  }
  u32 nlo;
  i64 pivot;
  if (m <= 32)
  {
    const bool have = (u32)lane < m;
    const unsigned short pt = have ? ord[lane] : (unsigned short)0;
    const i64 id = have ? sub_ids(c)[pt] : 0;
    u64 slo = (u64)(u32)id;
    i64 shi = id >> 32;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
      slo += __shfl_xor_sync(0xffffffffu, slo, o);
      shi += __shfl_xor_sync(0xffffffffu, shi, o);
    }
    pivot = sub_mean_id(slo, shi, m);
    const SubSplit sp = sub_make_split(c, dim, mid, pivot, welford);
    const bool hi = have && sub_hi(c, sp, pt);
    const u32 hib = __ballot_sync(0xffffffffu, hi);
    const u32 hb = __popc(hib & ((1u << lane) - 1u));
    nlo = m - __popc(hib);
    if (have) dst[hi ? nlo + hb : (u32)lane - hb] = pt;
  }
  else
  {
    constexpr int ROUNDS = SUB_WARP_MAX / 32;
    unsigned short pt[ROUNDS];
    u64 slo = 0;
    i64 shi = 0;
#pragma unroll
    for (int j = 0; j < ROUNDS; ++j)
    {
      const u32 i = (u32)(j * 32 + lane);
      pt[j] = i < m ? ord[i] : (unsigned short)0;
      const i64 id = i < m ? sub_ids(c)[pt[j]] : 0;
      slo += (u64)(u32)id;
      shi += id >> 32;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
      slo += __shfl_xor_sync(0xffffffffu, slo, o);
      shi += __shfl_xor_sync(0xffffffffu, shi, o);
    }
    pivot = sub_mean_id(slo, shi, m);
    const SubSplit sp = sub_make_split(c, dim, mid, pivot, welford);
    u32 hib[ROUNDS];
    bool hi4[ROUNDS];
    u32 nhi = 0;
#pragma unroll
    for (int j = 0; j < ROUNDS; ++j)
    {
      hi4[j] = (u32)(j * 32 + lane) < m && sub_hi(c, sp, pt[j]);
      hib[j] = __ballot_sync(0xffffffffu, hi4[j]);
      nhi += __popc(hib[j]);
    }
    nlo = m - nhi;
    const u32 below = (1u << lane) - 1u;
    u32 hacc = 0;
#pragma unroll
    for (int j = 0; j < ROUNDS; ++j)
    {
      const u32 i = (u32)(j * 32 + lane);
      const u32 hb = hacc + __popc(hib[j] & below);
      if (i < m) dst[hi4[j] ? nlo + hb : i - hb] = pt[j];
      hacc += __popc(hib[j]);
    }
  }
  if (lane == 0)
  {
    env.t.t_dim[nd.row] = dim;
    env.t.t_mid[nd.row] = mid;
    env.t.t_id[nd.row] = pivot;
  }
  return nlo;
}

// ---- the kernel --------------------------------------------------------------------------------------------------------
// Phase 1 (CTA-synchronous, level by level): nodes of more than 32 points.  Their children of 2..32 points go to the
// node queue.  Phase 2 (no CTA barrier): every 8-lane team takes the next queue position, waits until a node has been
// published there, splits it and appends its children (2..32 points) to the queue -- breadth-first order without level
// barriers; `pending` (nodes published and not yet finished) reaching 0 ends the phase.  A sub-tree of n points has
// fewer than n nodes, so the queue never wraps.
struct SubLists  // static shared memory
{
  SubNode coop[2][4];    // nodes of more than 128 points (at most 3 per level of a sub-tree of <= 512 points)
  SubNode warpn[2][SUB_TMAX / (SUB_TEAM_MAX + 1) + 1];  // nodes of SUB_TEAM_MAX+1..128 points
  u32 n_coop[2], n_warp[2];
  u32 q_tail, q_head, pending;
};

// sum over the 32 lanes of eight 64-bit values per lane in 10 shuffles: afterwards lane l holds the total of value
// number l >> 2 (halving exchange over lane bits 4, 3, 2, then a plain butterfly over bits 1, 0)
__device__ __forceinline__ u64 sub_reduce8(const u64 (&v)[8], int lane)
{
  const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
  u64 w[4], z[2], y;
#pragma unroll
  for (int i = 0; i < 4; ++i)
  {
    const u64 send = b4 ? v[i] : v[i + 4];
    const u64 keep = b4 ? v[i + 4] : v[i];
    w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
  {
    const u64 send = b3 ? w[i] : w[i + 2];
    const u64 keep = b3 ? w[i + 2] : w[i];
    z[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  {
    const u64 send = b2 ? z[0] : z[1];
    const u64 keep = b2 ? z[1] : z[0];
    y = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  y += __shfl_xor_sync(0xffffffffu, y, 2);
  y += __shfl_xor_sync(0xffffffffu, y, 1);
  return y;
}

__device__ unsigned long long g_sub_dbg[8];

template <int CH, bool FULL>
__global__ void __launch_bounds__(SUB_NT, 512 / SUB_NT)
k_subtree_cta(SubList sl, u32 nsub, const u32* __restrict__ g_sub_perm, const i64* __restrict__ sub_pid,
              const float* __restrict__ rows, int ld, int dims, float qk, double qinv, TableOut t, int* __restrict__ t_src,
              u32 row_base, u32 overflow_base, u32 t_cap, u32* __restrict__ counters,
              unsigned long long* __restrict__ lvl_points, unsigned long long* __restrict__ lvl_ranges, int T)
{
  __shared__ QfxBest s_cbest[SUB_NW];
  __shared__ u64 s_idlo[SUB_NW];
  __shared__ i64 s_idhi[SUB_NW];
  __shared__ u32 s_wcnt[SUB_NW];
  __shared__ SubSplit s_split;
  __shared__ SubLists s_l;
  __shared__ u32 s_lvlp[64], s_lvlr[64];
  __shared__ u32 s_k, s_nan;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tl = lane & 7;
  const int tshift = lane & 24;
  const u32 tmask = 0xffu << tshift;

  SubCtx c;
  c.ld = ld;
  c.dims = dims;
  c.C4 = FULL ? 8 * CH : (ld >> 2);
  c.XS = c.C4 + 1;
  c.qk = qk;
  c.qinv = qinv;
  c.rows = rows;
  c.exact_all = false;
  {
    u32 o = (u32)T * (u32)c.XS * 16u;
    c.o_ids = o;
    o += (u32)T * 8u;
    c.o_queue = o;  // [T] nodes of 2..32 points, in the order they were published
    o += (u32)T * 16u;
    c.o_perm = o;
    o += (u32)T * 4u;
    c.o_ord[0] = o;
    o += (u32)T * 2u;
    c.o_ord[1] = o;
  }
#define s_queue (reinterpret_cast<SubNode*>(reinterpret_cast<char*>(s_dyn) + c.o_queue))
  if (threadIdx.x < 64) { s_lvlp[threadIdx.x] = 0; s_lvlr[threadIdx.x] = 0; }
  SubEnv env;
  env.t = t;
  env.t_src = t_src;
  env.overflow_base = overflow_base;
  env.t_cap = t_cap;
  env.counters = counters;
  env.lvlp = s_lvlp;
  env.lvlr = s_lvlr;

  // publishes a node of 2..32 points: fields first, the count (what consumers poll) last
  auto enqueue = [&](const SubNode& ch)
  {
    const u32 slot = atomicAdd(&s_l.q_tail, 1u);
    SubNode* q = s_queue + slot;
    q->rid = ch.rid;
    q->row = ch.row;
    q->start = ch.start;
    __threadfence_block();
    *reinterpret_cast<volatile unsigned short*>(&q->count) = ch.count;
  };
  // hands a child that is a range to the class that will split it (one thread)
  auto dispatch = [&](const SubNode& ch, int nxt)
  {
    if (ch.count == 0) return;
    if (ch.count > SUB_WARP_MAX) s_l.coop[nxt][atomicAdd(&s_l.n_coop[nxt], 1u)] = ch;
    else if (ch.count > SUB_TEAM_MAX) s_l.warpn[nxt][atomicAdd(&s_l.n_warp[nxt], 1u)] = ch;
    else enqueue(ch);
  };

  for (;;)
  {
    __syncthreads();  // the previous sub-tree is finished with shared memory (and s_k)
    if (threadIdx.x == 0)
    {
      s_k = atomicAdd(&counters[3], 1u);
      s_nan = 0;
      s_l.n_coop[0] = s_l.n_coop[1] = s_l.n_warp[0] = s_l.n_warp[1] = 0;
      s_l.q_tail = 0;
      s_l.q_head = 0;
      s_l.pending = 0;
    }
    __syncthreads();
    const u32 k = s_k;
    if (k >= nsub) break;
    const long long tk0 = clock64();
    const u32 S = sl.start[k], n = sl.count[k];
    // ---- load: ids, row indexes, empty queue; then the raw rows (cp.async, 16 B per thread) --------------------------------
    for (u32 j = threadIdx.x; j < n; j += SUB_NT)
    {
      sub_perm(c)[j] = g_sub_perm[S + j];
      sub_ids(c)[j] = sub_pid[S + j];
      sub_ord(c, 0)[j] = (unsigned short)j;
      s_queue[j].count = 0;
    }
    __syncthreads();
    const u32 n4 = n * (u32)c.C4;
    for (u32 q = threadIdx.x; q < n4; q += SUB_NT)
    {
      const u32 j = q / (u32)c.C4, col = q - j * (u32)c.C4;
      const float4* src = reinterpret_cast<const float4*>(rows + (size_t)sub_perm(c)[j] * ld) + col;
      const u32 dsta = (u32)__cvta_generic_to_shared(sub_x(c) + j * (u32)c.XS + col);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dsta), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    // quantise in place: every thread converts exactly what it copied (its own cp.async data: no barrier needed)
    {
      bool nan = false;
      for (u32 q = threadIdx.x; q < n4; q += SUB_NT)
      {
        const u32 j = q / (u32)c.C4, col = q - j * (u32)c.C4;
        int4* px = sub_x(c) + j * (u32)c.XS + col;
        const float4 v = *reinterpret_cast<const float4*>(px);
        nan = nan || (v.x != v.x) || (v.y != v.y) || (v.z != v.z) || (v.w != v.w);
        int4 qv;
        qv.x = __float2int_rn(__fmul_rn(v.x, qk));
        qv.y = __float2int_rn(__fmul_rn(v.y, qk));
        qv.z = __float2int_rn(__fmul_rn(v.z, qk));
        qv.w = __float2int_rn(__fmul_rn(v.w, qk));
        *px = qv;
      }
      if (nan) s_nan = 1u;
    }
    env.B = row_base + 2u * S - 2u * k;
    env.root_depth = sl.depth[k];
    if (threadIdx.x == 0)
    {
      SubNode root;
      root.rid = sl.rid[k];
      root.row = sl.row[k];
      root.start = 0;
      root.count = (unsigned short)n;
      if (n > (u32)SUB_WARP_MAX) { s_l.coop[0][0] = root; s_l.n_coop[0] = 1; }
      else if (n > (u32)SUB_TEAM_MAX) { s_l.warpn[0][0] = root; s_l.n_warp[0] = 1; }
      else enqueue(root);
    }
    __syncthreads();
    c.exact_all = s_nan != 0u;
    const long long tk1 = clock64();

    // ---- phase 1: nodes of more than 32 points, level by level ----------------------------------------------------------
    int cur = 0;
    u32 depth = env.root_depth;
    bool failed = false;
    for (;;)
    {
      const u32 n_coop = s_l.n_coop[cur], n_warp = s_l.n_warp[cur];
      if (n_coop + n_warp == 0) break;
      if (depth >= (u32)VI_MAX_DEPTH)
      {
        if (threadIdx.x == 0) counters[1] = 2u;  // splitting a depth-62 range: rangeId overflow (IndexBuilder.cs:99)
        failed = true;
        break;
      }
      const bool mx = (depth & 1u) == 0u;
      const int ob = (int)((depth - env.root_depth) & 1u);
      // CTA-wide class: one node at a time
      const long long tc0 = clock64();
      for (u32 ni = 0; ni < n_coop; ++ni)
      {
        const SubNode nd = s_l.coop[cur][ni];
        const u32 s0 = nd.start, m = nd.count;
        const unsigned short* ord = sub_ord(c, ob) + s0;
        // statistics: this warp's columns, lanes take points lane, lane + 32, ...
        QfxBest best;
        best.key.hi = 0;
        best.key.lo = 0;
        best.s1 = 0;
        best.idx = 0x7fffffff;
        for (int col = warp; col < c.C4; col += SUB_NW)
        {
          i64 s1[4] = {0, 0, 0, 0};
          u64 s2[4] = {0, 0, 0, 0};
          const int4* xp = sub_x(c) + col;
          for (u32 i = lane; i < m; i += 32) iacc4(s1, s2, xp[(u32)ord[i] * (u32)c.XS]);
          const u64 v8[8] = {(u64)s1[0], s2[0], (u64)s1[1], s2[1], (u64)s1[2], s2[2], (u64)s1[3], s2[3]};
          const u64 y = sub_reduce8(v8, lane);             // lane l: value l >> 2  (even: S1 of dim l >> 3, odd: S2)
          const u64 y2 = __shfl_xor_sync(0xffffffffu, y, 4);  // the other one of the pair
          const int d = col * 4 + (lane >> 3);
          if ((lane & 4) == 0 && d < dims)
          {
            const Key128 key = qfx_key(m, (i64)y, y2, 0ull);
            if (qfx_better(mx, key, d, best.key, best.idx))
            {
              best.key = key;
              best.s1 = (i64)y;
              best.idx = d;
            }
          }
        }
        // the warp's candidates sit in lanes with bit 2 clear, one dimension per value of lane >> 3
#pragma unroll
        for (int o = 16; o >= 8; o >>= 1)
        {
          QfxBest t2;
          t2.key.hi = __shfl_xor_sync(0xffffffffu, best.key.hi, o);
          t2.key.lo = __shfl_xor_sync(0xffffffffu, best.key.lo, o);
          t2.s1 = __shfl_xor_sync(0xffffffffu, best.s1, o);
          t2.idx = __shfl_xor_sync(0xffffffffu, best.idx, o);
          if (qfx_better(mx, t2.key, t2.idx, best.key, best.idx)) best = t2;
        }
        // id sum over the node's points
        u64 slo = 0;
        i64 shi = 0;
        for (u32 i = threadIdx.x; i < m; i += SUB_NT)
        {
          const i64 id = sub_ids(c)[ord[i]];
          slo += (u64)(u32)id;
          shi += id >> 32;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
          slo += __shfl_xor_sync(0xffffffffu, slo, o);
          shi += __shfl_xor_sync(0xffffffffu, shi, o);
        }
        if (lane == 0)
        {
          s_cbest[warp] = best;
          s_idlo[warp] = slo;
          s_idhi[warp] = shi;
        }
        __syncthreads();
        if (warp == 0)
        {
          QfxBest b2;
          b2.key.hi = 0;
          b2.key.lo = 0;
          b2.s1 = 0;
          b2.idx = 0x7fffffff;
          u64 l2 = 0;
          i64 h2 = 0;
          if (lane < SUB_NW)
          {
            b2 = s_cbest[lane];
            l2 = s_idlo[lane];
            h2 = s_idhi[lane];
          }
          b2 = qfx_reduce<32>(b2, mx, 0xffffffffu);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1)
          {
            l2 += __shfl_xor_sync(0xffffffffu, l2, o);
            h2 += __shfl_xor_sync(0xffffffffu, h2, o);
          }
          int dim = b2.idx;
          float mid = qfx_mid(b2.s1, m, qinv);
          bool welford = false;
          if (key_lt(b2.key, qfx_threshold(m)))
          {
            if (lane == 0) atomicAdd(&counters[2], 1u);
            const ExBest eb = sub_welford<32>(c, ord, m, lane, 0xffffffffu, mx);
            dim = eb.idx;
            mid = eb.mean;
            welford = true;
          }
          if (lane == 0)
          {
            const i64 pivot = sub_mean_id(l2, h2, m);
            s_split = sub_make_split(c, dim, mid, pivot, welford);
            t.t_dim[nd.row] = dim;
            t.t_mid[nd.row] = mid;
            t.t_id[nd.row] = pivot;
          }
        }
        __syncthreads();
        // stable partition: thread i looks after point i of the node (m <= SUB_TMAX <= SUB_NT)
        const SubSplit sp = s_split;
        const bool have = threadIdx.x < m;
        const unsigned short pt = have ? ord[threadIdx.x] : (unsigned short)0;
        const bool hi = have && sub_hi(c, sp, pt);
        const u32 hb = __ballot_sync(0xffffffffu, hi);
        if (lane == 0) s_wcnt[warp] = __popc(hb);
        __syncthreads();
        u32 nhi = 0, hi_before_w = 0;
#pragma unroll
        for (int w = 0; w < SUB_NW; ++w)
        {
          const u32 x = s_wcnt[w];
          if (w < warp) hi_before_w += x;
          nhi += x;
        }
        const u32 nlo = m - nhi;
        unsigned short* dst = sub_ord(c, ob ^ 1) + s0;
        if (have)
        {
          const u32 hbefore = hi_before_w + __popc(hb & ((1u << lane) - 1u));
          dst[hi ? nlo + hbefore : threadIdx.x - hbefore] = pt;
        }
        __syncthreads();  // the partition is visible; s_cbest / s_wcnt / s_split are free again
        if (threadIdx.x == 0)
        {
          SubNode lo, hi2;
          sub_children(c, env, nd, nlo, sub_ord(c, ob ^ 1), lo, hi2);
          dispatch(lo, cur ^ 1);
          dispatch(hi2, cur ^ 1);
          atomicAdd(&s_lvlp[depth & 63u], m);
          atomicAdd(&s_lvlr[depth & 63u], 1u);
        }
      }
      // warp class
      const long long tc1 = clock64();
      for (u32 ni = warp; ni < n_warp; ni += SUB_NW)
      {
        const SubNode nd = s_l.warpn[cur][ni];
        const u32 nlo = sub_warp_split(c, env, nd, sub_ord(c, ob) + nd.start, sub_ord(c, ob ^ 1) + nd.start, mx, lane);
        __syncwarp();
        if (lane == 0)
        {
          SubNode lo, hi2;
          sub_children(c, env, nd, nlo, sub_ord(c, ob ^ 1), lo, hi2);
          dispatch(lo, cur ^ 1);
          dispatch(hi2, cur ^ 1);
          atomicAdd(&s_lvlp[depth & 63u], (u32)nd.count);
          atomicAdd(&s_lvlr[depth & 63u], 1u);
        }
      }
      __syncthreads();
      if (threadIdx.x == 0)
      {
        const long long tc2 = clock64();
        atomicAdd(&g_sub_dbg[4], (unsigned long long)(tc1 - tc0));
        atomicAdd(&g_sub_dbg[5], (unsigned long long)(tc2 - tc1));
        atomicAdd(&g_sub_dbg[6], (unsigned long long)n_coop);
        atomicAdd(&g_sub_dbg[7], (unsigned long long)n_warp);
      }
      if (threadIdx.x == 0) { s_l.n_coop[cur] = 0; s_l.n_warp[cur] = 0; }  // this level's lists are the level after next's
      cur ^= 1;
      ++depth;
      __syncthreads();
    }
    if (threadIdx.x == 0) s_l.pending = failed ? 0u : s_l.q_tail;
    __syncthreads();
    const long long tk2 = clock64();
    // ---- phase 2: nodes of 2..32 points from the queue ---------------------------------------------------------------------
    // The four teams of a warp stay converged: every trip of the loop starts with a warp-wide vote, then each team
    // splits one node if one is ready at its queue position.
    {
      bool done = false, have_pos = false;
      u32 pos = 0, spins = 0;
      while (__any_sync(0xffffffffu, !done))
      {
        if (done) continue;
        if (!have_pos)
        {
          if (tl == 0) pos = atomicAdd(&s_l.q_head, 1u);
          pos = __shfl_sync(tmask, pos, tshift);
          have_pos = true;
        }
        // (a sub-tree of n points publishes fewer than n nodes: positions from n on never fill, their owners only wait
        // for the end; entries [0, n) were cleared when the sub-tree was loaded)
        const u32 cnt = pos < n ? (u32)*reinterpret_cast<volatile unsigned short*>(&s_queue[pos].count) : 0u;
        if (cnt == 0)
        {
          if (*reinterpret_cast<volatile u32*>(&s_l.pending) == 0u) done = true;
          else if (++spins > (1u << 24))
          {
            if (tl == 0) counters[1] = 3u;  // must not happen: a published node was never finished (reported as an error)
            done = true;
          }
          continue;
        }
        spins = 0;
        __threadfence_block();
        SubNode nd = s_queue[pos];
        nd.count = (unsigned short)cnt;
        have_pos = false;
        const u32 d = sub_depth(nd.rid);
        u32 nch = 0;
        if (d >= (u32)VI_MAX_DEPTH)
        {
          if (tl == 0) counters[1] = 2u;  // rangeId overflow (IndexBuilder.cs:99)
        }
        else
        {
          const int ob = (int)((d - env.root_depth) & 1u);
          const u32 nlo = sub_team_split<CH, FULL>(c, env, nd, sub_ord(c, ob) + nd.start, sub_ord(c, ob ^ 1) + nd.start, (d & 1u) == 0u,
                                                   tl, tmask, tshift);
          __syncwarp(tmask);  // the partition is visible to lane 0
          if (tl == 0)
          {
            SubNode lo, hi2;
            sub_children(c, env, nd, nlo, sub_ord(c, ob ^ 1), lo, hi2);
            if (lo.count) { enqueue(lo); ++nch; }
            if (hi2.count) { enqueue(hi2); ++nch; }
            atomicAdd(&s_lvlp[d & 63u], (u32)nd.count);
            atomicAdd(&s_lvlr[d & 63u], 1u);
          }
        }
        if (tl == 0)
        {
          __threadfence_block();
          atomicAdd(&s_l.pending, nch - 1u);  // children first, then this node leaves
        }
      }
    }
    if (threadIdx.x == 0)
    {
      const long long tk3 = clock64();
      atomicAdd(&g_sub_dbg[0], (unsigned long long)(tk1 - tk0));
      atomicAdd(&g_sub_dbg[1], (unsigned long long)(tk2 - tk1));
      atomicAdd(&g_sub_dbg[2], (unsigned long long)(tk3 - tk2));
      atomicAdd(&g_sub_dbg[3], 1ull);
    }
  }
  __syncthreads();
  if (threadIdx.x < 64)
  {
    if (s_lvlp[threadIdx.x]) atomicAdd(&lvl_points[threadIdx.x], (unsigned long long)s_lvlp[threadIdx.x]);
    if (s_lvlr[threadIdx.x]) atomicAdd(&lvl_ranges[threadIdx.x], (unsigned long long)s_lvlr[threadIdx.x]);
  }
}
#undef s_queue
