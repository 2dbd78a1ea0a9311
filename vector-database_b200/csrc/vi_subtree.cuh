// vi_subtree.cuh -- the bottom of the tree, fast mode: a range with at most `rows_max` (<= 32) points leaves the
// level loop (k_emit_children / k_scatter put it on the sub-tree list) and ONE WARP builds its whole sub-tree in
// shared memory: the rows are read from HBM once (cp.async, 16 B per lane) instead of once per remaining level.
// The sub-tree is processed level by level; the warp's four 8-lane teams take four of the level's nodes at a time
// (a level of a sub-tree with <= 32 points has <= 16 nodes with >= 2 points).
//
// Same arithmetic as the level kernels (vi_stats_fast.cuh): exact integer sums of xi = rint(x * 2^(26-E)), key
// K = n*S2 - S1^2 (fits 64 bits for n <= 32), arg-max / arg-min by depth parity with lowest-index ties,
// Mid = float((double)S1/n * 2^(E-26)), float32 Welford fallback for a poorly resolved range, Id = trunc(sum id / n),
// stable partition with the reference's predicate (IndexBuilder.cs:115).  A team lane owns CH float4 column chunks.
//
// Rows: the sub-tree root's row exists already (its parent's level made it).  A sub-tree of n points has exactly
// 2n-2 further rows when every split has two non-empty sides; they go to a dense block at
// row_base + 2*sub_start - 2*sub_index, breadth-first inside the sub-tree.  A one-sided split adds a row: it is
// taken from an overflow area behind all blocks (atomic counter), so the table stays dense.
#pragma once
#include "vi_partition.cuh"
#include "vi_stats_exact.cuh"
#include "vi_stats_fast.cuh"

// RN(1/c) for the counts a sub-tree can see (host constant folding is IEEE division): the float32 fallback divides
// through div_by_count (vi_stats_exact.cuh) instead of __fdiv_rn
__constant__ float c_rcp32[33] = {
    0.f,        1.f,        1.f / 2.f,  1.f / 3.f,  1.f / 4.f,  1.f / 5.f,  1.f / 6.f,  1.f / 7.f,  1.f / 8.f,
    1.f / 9.f,  1.f / 10.f, 1.f / 11.f, 1.f / 12.f, 1.f / 13.f, 1.f / 14.f, 1.f / 15.f, 1.f / 16.f, 1.f / 17.f,
    1.f / 18.f, 1.f / 19.f, 1.f / 20.f, 1.f / 21.f, 1.f / 22.f, 1.f / 23.f, 1.f / 24.f, 1.f / 25.f, 1.f / 26.f,
    1.f / 27.f, 1.f / 28.f, 1.f / 29.f, 1.f / 30.f, 1.f / 31.f, 1.f / 32.f};

constexpr int SUB_WARPS = 8;
constexpr int SUB_NODES = 16;  // nodes with >= 2 points on one level of a sub-tree of <= 32 points

struct SubNode
{
  i64 rid;
  u32 row;
  unsigned char start, count, pad0, pad1;
};

struct SubBest
{
  u64 key;
  i64 s1;
  int idx;
};

__device__ __forceinline__ bool sub_better(bool mx, u64 k, int i, u64 bk, int bi)
{
  if (i == 0x7fffffff) return false;
  if (bi == 0x7fffffff) return true;
  if (k != bk) return mx ? (k > bk) : (k < bk);
  return i < bi;
}

// counters: [0] overflow rows used, [1] error (1 = table capacity, 2 = depth overflow)
// FULL: rows are exactly 8*CH float4 wide and every column is a real dimension (no guards)
template <int CH, int MINB, bool FULL>
__global__ void __launch_bounds__(SUB_WARPS * 32, MINB)
k_subtree_fast(SubList sl, u32 nsub, const u32* __restrict__ sub_perm, const i64* __restrict__ sub_pid,
               const float* __restrict__ rows, int ld, int dims, float qk, double qinv, TableOut t,
               int* __restrict__ t_src, u32 row_base, u32 overflow_base, u32 t_cap, u32* __restrict__ counters,
               unsigned long long* __restrict__ lvl_points, unsigned long long* __restrict__ lvl_ranges, int rows_max)
{
  extern __shared__ float4 s_rows4[];  // [SUB_WARPS][rows_max][ld]
  __shared__ i64 s_ids[SUB_WARPS][32];
  __shared__ u32 s_perm[SUB_WARPS][32];
  __shared__ unsigned char s_lp[SUB_WARPS][32];
  __shared__ SubNode s_nodes[SUB_WARPS][2][SUB_NODES];
  __shared__ unsigned long long s_lvlp[64], s_lvlr[64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tl = lane & 7, team = lane >> 3;
  const u32 tmask = 0xffu << (team * 8);
  float* wrows = reinterpret_cast<float*>(s_rows4) + (size_t)warp * rows_max * ld;
  if (threadIdx.x < 64) { s_lvlp[threadIdx.x] = 0; s_lvlr[threadIdx.x] = 0; }
  __syncthreads();
  const int C4 = FULL ? 8 * CH : (ld >> 2);
  if (FULL) { ld = 32 * CH; dims = 32 * CH; }

  for (u32 k = blockIdx.x * SUB_WARPS + warp; k < nsub; k += gridDim.x * SUB_WARPS)
  {
    const u32 S = sl.start[k], n = sl.count[k];
    // ---- load the sub-tree's points: ids / row indexes one per lane, rows with 16-byte cp.async ----------------
    u32 myrow = 0;
    if ((u32)lane < n)
    {
      myrow = sub_perm[S + lane];
      s_perm[warp][lane] = myrow;
      s_ids[warp][lane] = sub_pid[S + lane];
      s_lp[warp][lane] = (unsigned char)lane;
    }
    for (u32 j = 0; j < n; ++j)
    {
      const u32 r = __shfl_sync(0xffffffffu, myrow, j);
      const float4* src = reinterpret_cast<const float4*>(rows + (size_t)r * ld);
      for (int c = lane; c < C4; c += 32)
      {
        const u32 dst = (u32)__cvta_generic_to_shared(wrows + (size_t)j * ld + c * 4);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + c) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    u32 next_row = row_base + 2u * S - 2u * k;  // warp-uniform
    const u32 block_end = next_row + 2u * n - 2u;
    u32 depth = sl.depth[k];
    if (lane == 0)
    {
      SubNode root;
      root.rid = sl.rid[k];
      root.row = sl.row[k];
      root.start = 0;
      root.count = (unsigned char)n;
      root.pad0 = root.pad1 = 0;
      s_nodes[warp][0][0] = root;
    }
    u32 ncur = 1;
    int cur = 0;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();

    while (ncur > 0)
    {
      if (depth >= (u32)VI_MAX_DEPTH)
      {
        if (lane == 0) counters[1] = 2u;  // splitting a depth-62 range: rangeId overflow (IndexBuilder.cs:99)
        break;
      }
      const bool mx = (depth & 1u) == 0u;
      u32 nnext = 0, lvl_pts = 0;
      for (u32 base = 0; base < ncur; base += 4)
      {
        const bool act = base + team < ncur;
        SubNode nd;
        nd.rid = 0; nd.row = 0; nd.start = 0; nd.count = 0;
        if (act) nd = s_nodes[warp][cur][base + team];
        const u32 s0 = nd.start, m = nd.count;
        // ---- statistics: the team's lanes own CH float4 column chunks; points in stable order -------------------
        i64 s1[CH * 4];
        u64 s2[CH * 4];
#pragma unroll
        for (int i = 0; i < CH * 4; ++i) { s1[i] = 0; s2[i] = 0; }
        for (u32 i = 0; i < m; ++i)
        {
          const float4* rp = reinterpret_cast<const float4*>(wrows + (size_t)s_lp[warp][s0 + i] * ld);
#pragma unroll
          for (int c = 0; c < CH; ++c)
          {
            const int col = c * 8 + tl;
            if (FULL || col < C4)
            {
              const float4 x = rp[col];
              qfx_acc4(s1 + c * 4, s2 + c * 4, x, qk);
            }
          }
        }
        const u64 thr = ((u64)m * (u64)m) << (2 * VI_QFX_MIN_RES_BITS);
        SubBest best;
        best.key = 0;
        best.s1 = 0;
        best.idx = 0x7fffffff;
#pragma unroll
        for (int c = 0; c < CH; ++c)
#pragma unroll
          for (int e = 0; e < 4; ++e)
          {
            const int d = (c * 8 + tl) * 4 + e;
            if ((FULL || d < dims) && m > 0)
            {
              const i64 a = s1[c * 4 + e];
              const u64 key = (u64)m * s2[c * 4 + e] - (u64)(a * a);  // m <= 32: below 2^63, exact
              if (sub_better(mx, key, d, best.key, best.idx))
              {
                best.key = key;
                best.s1 = a;
                best.idx = d;
              }
            }
          }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1)
        {
          SubBest b2;
          b2.key = __shfl_xor_sync(0xffffffffu, best.key, o);
          b2.s1 = __shfl_xor_sync(0xffffffffu, best.s1, o);
          b2.idx = __shfl_xor_sync(0xffffffffu, best.idx, o);
          if (sub_better(mx, b2.key, b2.idx, best.key, best.idx)) best = b2;
        }
        int dim = best.idx;
        float mid = m > 0 ? qfx_mid(best.s1, m, qinv) : 0.f;
        const bool unresolved = act && best.key < thr;  // chosen dimension poorly resolved (team-uniform)
        if (unresolved)
        {
          if (tl == 0) atomicAdd(&counters[2], 1u);  // fallback count (diagnostics)
          // poorly resolved: the reference's float32 recurrence over the same points in the same order
          ExBest eb;
          eb.key = 0.f;
          eb.mean = 0.f;
          eb.idx = 0x7fffffff;
          for (int d = tl; d < dims; d += 8)
          {
            float mean = wrows[(size_t)s_lp[warp][s0] * ld + d], q = 0.f;
            for (u32 i = 1; i < m; ++i)
              welford_step_r(mean, q, wrows[(size_t)s_lp[warp][s0 + i] * ld + d], (float)(i + 1u), c_rcp32[i + 1u]);
            const float key = mx ? q : -q;
            if (ex_better(key, d, eb.key, eb.idx))
            {
              eb.key = key;
              eb.mean = mean;
              eb.idx = d;
            }
          }
          eb = ex_reduce_w<8>(eb, tmask);
          dim = eb.idx;
          mid = eb.mean;
        }
        // ---- pivot id and stable partition of the node's slice of s_lp (lane handles points tl, tl+8, ...) ----------
        unsigned char pt[4];
        i64 pid4[4];
        u64 slo = 0;
        i64 shi = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
          const u32 i = (u32)(j * 8 + tl);
          pt[j] = i < m ? s_lp[warp][s0 + i] : 0;
          pid4[j] = i < m ? s_ids[warp][pt[j]] : 0;
          slo += (u64)(u32)pid4[j];
          shi += pid4[j] >> 32;
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1)
        {
          slo += __shfl_xor_sync(0xffffffffu, slo, o);
          shi += __shfl_xor_sync(0xffffffffu, shi, o);
        }
        const i64 pivot = m > 0 ? mean_id(slo, shi, m) : 0;
        u32 hib[4], lob[4];
        bool hi4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
          const u32 i = (u32)(j * 8 + tl);
          const bool have = i < m;
          const float v = have ? wrows[(size_t)pt[j] * ld + dim] : 0.f;
          hi4[j] = have && (v > mid || (v == mid && pid4[j] > pivot));  // IndexBuilder.cs:115
          hib[j] = (__ballot_sync(0xffffffffu, hi4[j]) >> (team * 8)) & 0xffu;
          lob[j] = (__ballot_sync(0xffffffffu, have && !hi4[j]) >> (team * 8)) & 0xffu;
        }
        const u32 nhi = __popc(hib[0]) + __popc(hib[1]) + __popc(hib[2]) + __popc(hib[3]);
        const u32 nlo = m - nhi;
        __syncwarp();
        {
          const u32 below = (1u << tl) - 1u;
          u32 hacc = 0, lacc = 0;
#pragma unroll
          for (int j = 0; j < 4; ++j)
          {
            const u32 i = (u32)(j * 8 + tl);
            if (i < m)
              s_lp[warp][hi4[j] ? s0 + nlo + hacc + __popc(hib[j] & below) : s0 + lacc + __popc(lob[j] & below)] = pt[j];
            hacc += __popc(hib[j]);
            lacc += __popc(lob[j]);
          }
        }
        __syncwarp();
        // ---- rows: each team needs (nlo>0)+(nhi>0) rows and pushes its children with >= 2 points --------------------
        const u32 need = act ? (u32)(nlo > 0) + (u32)(nhi > 0) : 0u;
        const u32 push = act ? (u32)(nlo > 1) + (u32)(nhi > 1) : 0u;
        u32 need_before = 0, push_before = 0, need_all = 0, push_all = 0;
#pragma unroll
        for (int tt = 0; tt < 4; ++tt)
        {
          const u32 nn = __shfl_sync(0xffffffffu, need, tt * 8);
          const u32 pp = __shfl_sync(0xffffffffu, push, tt * 8);
          if (tt < team) { need_before += nn; push_before += pp; }
          need_all += nn;
          push_all += pp;
        }
        if (act && tl == 0)
        {
          t.t_dim[nd.row] = dim;
          t.t_mid[nd.row] = mid;
          t.t_id[nd.row] = pivot;
          int child_row[2] = {-1, -1};
          const u32 cstart[2] = {s0, s0 + nlo}, ccount[2] = {nlo, nhi};
          u32 ri = next_row + need_before, pi = nnext + push_before;
#pragma unroll
          for (int side = 0; side < 2; ++side)
          {
            if (ccount[side] == 0) continue;  // empty range: no row (IndexBuilder.cs:70-73)
            const u32 slot = ri++;
            const u32 r = slot < block_end ? slot : overflow_base + atomicAdd(&counters[0], 1u);
            if (r >= t_cap) { counters[1] = 1u; continue; }
            child_row[side] = (int)r;
            t.t_rid[r] = nd.rid * 2 + 1 + side;  // IndexBuilder.cs:99,104
            t.t_low[r] = -1;
            t.t_high[r] = -1;
            if (ccount[side] == 1)
            {
              const unsigned char p1 = s_lp[warp][cstart[side]];
              t.t_dim[r] = -1;  // leaf (IndexBuilder.cs:81-82)
              t.t_mid[r] = 0.f;
              t.t_id[r] = s_ids[warp][p1];
              t_src[r] = (int)s_perm[warp][p1];
            }
            else if (pi < (u32)SUB_NODES)
            {
              SubNode ch;
              ch.rid = nd.rid * 2 + 1 + side;
              ch.row = r;
              ch.start = (unsigned char)cstart[side];
              ch.count = (unsigned char)ccount[side];
              ch.pad0 = ch.pad1 = 0;
              s_nodes[warp][cur ^ 1][pi++] = ch;
            }
            else
              counters[1] = 1u;
          }
          t.t_low[nd.row] = child_row[0];
          t.t_high[nd.row] = child_row[1];
        }
        next_row += need_all;
        nnext += push_all;
        lvl_pts += __shfl_sync(0xffffffffu, m, 0) + __shfl_sync(0xffffffffu, m, 8) + __shfl_sync(0xffffffffu, m, 16) +
                   __shfl_sync(0xffffffffu, m, 24);
        __syncwarp();
      }
      if (lane == 0)
      {
        atomicAdd(&s_lvlp[depth & 63u], (unsigned long long)lvl_pts);
        atomicAdd(&s_lvlr[depth & 63u], (unsigned long long)ncur);
      }
      // next_row may have run past the block when one-sided splits used overflow rows: keep it monotone
      ncur = nnext;
      cur ^= 1;
      ++depth;
      __syncwarp();
    }
    __syncwarp();
  }
  __syncthreads();
  if (threadIdx.x < 64)
  {
    if (s_lvlp[threadIdx.x]) atomicAdd(&lvl_points[threadIdx.x], s_lvlp[threadIdx.x]);
    if (s_lvlr[threadIdx.x]) atomicAdd(&lvl_ranges[threadIdx.x], s_lvlr[threadIdx.x]);
  }
}
