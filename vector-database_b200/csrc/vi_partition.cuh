// vi_partition.cuh -- stable segmented partition of every open range at once (IndexBuilder.cs:99-129).
//
//   k_flags         hi(p) = value > Mid || (value == Mid && id > Id)  (IndexBuilder.cs:115), one bit per position
//   scan            hi-count prefix per 32-position word
//   k_seg_children  per range: low/high sizes -> child rows, next-level ranges, next-level positions
//   scans           child row indexes, next-level range indexes and position offsets
//   k_emit_children rows and next-level range descriptors of the children (2r+1, 2r+2; an empty child gets no row)
//   k_scatter       stable scatter of (row index, id) into the compacted next-level position space; a single-point
//                   child is a leaf: its row gets Id = that point's id and the point leaves the position space
#pragma once
#include "vi_common.cuh"

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
    const float v = ldg_f_gather(rows + (size_t)perm[p] * ld + dim);
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
               u32* seg_hbase, u32* c_rows, u64* c_actpos, u64* c_sub, u32 t_sub)
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
  // a child with one point is a leaf, with 2..t_sub points it goes to the sub-tree list, otherwise it stays a range
  const bool lo_sub = nlo >= 2 && nlo <= t_sub, hi_sub = nhi >= 2 && nhi <= t_sub;
  const bool lo_act = nlo > 1 && !lo_sub, hi_act = nhi > 1 && !hi_sub;
  c_actpos[s] = ((u64)((u32)lo_act + (u32)hi_act) << 32) | (u64)((lo_act ? nlo : 0u) + (hi_act ? nhi : 0u));
  c_sub[s] = ((u64)((u32)lo_sub + (u32)hi_sub) << 32) | (u64)((lo_sub ? nlo : 0u) + (hi_sub ? nhi : 0u));
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

// counters: [0] next-level big count, [1] error flag, [4] min next-level range size, [5] max next-level range size
__global__ void __launch_bounds__(256)
k_emit_children(SegLevel sg, u32 R, const u32* __restrict__ seg_nlo, const u32* __restrict__ c_rows,
                const u64* __restrict__ c_actpos, SegLevel nx, u32 row_base_next, u32 t_cap, TableOut t,
                u32* big_list_next, u32 big_thr, u32* counters, const u64* __restrict__ c_sub, u32 t_sub,
                u32 sub_cnt_base, u32 sub_pos_base, u32 child_depth, u32* sub_start, u32* sub_count, i64* sub_rid,
                u32* sub_row, u32* sub_depth, u32* bl_parent_next, u32* bl_sib_next, int sibling)
{
  const u32 s = blockIdx.x * 256u + threadIdx.x;
  if ((u64)row_base_next + c_rows[R] > (u64)t_cap)
  {
    if (s == 0) counters[1] = 1;
    return;
  }
  u32 cmin = 0xffffffffu, cmax = 0u;
  if (s < R)
  {
    const u32 n = sg.count[s], nlo = seg_nlo[s], nhi = n - nlo;
    const i64 rid = sg.rid[s];
    const u32 row = sg.row[s];
    const u32 r0 = row_base_next + c_rows[s];
    const u32 a0 = (u32)(c_actpos[s] >> 32), p0 = (u32)c_actpos[s];
    const u32 b0 = sub_cnt_base + (u32)(c_sub[s] >> 32), q0 = sub_pos_base + (u32)c_sub[s];
    const bool lo_sub = nlo >= 2 && nlo <= t_sub, hi_sub = nhi >= 2 && nhi <= t_sub;
    const bool lo_act = nlo > 1 && !lo_sub;
    // Sibling derivation (fast mode): when both children are big and the parent's integer sums are in the previous
    // level's gacc, only the smaller child is summed; the other one's sums are parent - sibling (exact integers).
    // The pair takes two consecutive big-list slots.
    const bool pair = sibling && nlo >= big_thr && nhi >= big_thr && sg.bslot[s] != 0xffffffffu;
    const u32 pair_base = pair ? atomicAdd(&counters[0], 2u) : 0u;
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
      else if (lo_sub)
      {
        sub_start[b0] = q0;
        sub_count[b0] = nlo;
        sub_rid[b0] = rid * 2 + 1;
        sub_row[b0] = (u32)lo_row;
        sub_depth[b0] = child_depth;
      }
      else
      {
        nx.start[a0] = p0;
        nx.count[a0] = nlo;
        nx.rid[a0] = rid * 2 + 1;
        nx.row[a0] = (u32)lo_row;
        u32 slot = 0xffffffffu;
        if (pair)
        {
          slot = pair_base;
          const bool derived = nlo > nhi;  // the larger one is derived; ties: low is summed
          bl_parent_next[slot] = derived ? sg.bslot[s] : 0xffffffffu;
          bl_sib_next[slot] = pair_base + 1u;
        }
        else if (nlo >= big_thr)
        {
          slot = atomicAdd(&counters[0], 1u);
          bl_parent_next[slot] = 0xffffffffu;
          bl_sib_next[slot] = 0xffffffffu;
        }
        if (slot != 0xffffffffu) big_list_next[slot] = a0;
        nx.bslot[a0] = slot;
        cmin = min(cmin, nlo);
        cmax = max(cmax, nlo);
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
      else if (hi_sub)
      {
        const u32 b1 = b0 + (lo_sub ? 1u : 0u);
        sub_start[b1] = q0 + (lo_sub ? nlo : 0u);
        sub_count[b1] = nhi;
        sub_rid[b1] = rid * 2 + 2;
        sub_row[b1] = (u32)hi_row;
        sub_depth[b1] = child_depth;
      }
      else
      {
        const u32 a1 = a0 + (lo_act ? 1u : 0u);
        nx.start[a1] = p0 + (lo_act ? nlo : 0u);
        nx.count[a1] = nhi;
        nx.rid[a1] = rid * 2 + 2;
        nx.row[a1] = (u32)hi_row;
        u32 slot = 0xffffffffu;
        if (pair)
        {
          slot = pair_base + 1u;
          const bool derived = nlo <= nhi;
          bl_parent_next[slot] = derived ? sg.bslot[s] : 0xffffffffu;
          bl_sib_next[slot] = pair_base;
        }
        else if (nhi >= big_thr)
        {
          slot = atomicAdd(&counters[0], 1u);
          bl_parent_next[slot] = 0xffffffffu;
          bl_sib_next[slot] = 0xffffffffu;
        }
        if (slot != 0xffffffffu) big_list_next[slot] = a1;
        nx.bslot[a1] = slot;
        cmin = min(cmin, nhi);
        cmax = max(cmax, nhi);
      }
    }
  }
  cmin = __reduce_min_sync(0xffffffffu, cmin);
  cmax = __reduce_max_sync(0xffffffffu, cmax);
  if ((threadIdx.x & 31) == 0 && cmax > 0)
  {
    atomicMin(&counters[4], cmin);
    atomicMax(&counters[5], cmax);
  }
}

__global__ void __launch_bounds__(256)
k_scatter(SegLevel sg, const u32* __restrict__ seg_of, const u32* __restrict__ perm, const i64* __restrict__ pid,
          u32 A, const u32* __restrict__ fbits, const u32* __restrict__ wpre, const u32* __restrict__ seg_nlo,
          const u32* __restrict__ seg_hbase, const u32* __restrict__ c_rows, const u64* __restrict__ c_actpos,
          u32 row_base_next, u32* __restrict__ perm_n, i64* __restrict__ pid_n, u32* __restrict__ seg_of_n,
          i64* __restrict__ t_id, int* __restrict__ t_src, const u32* __restrict__ counters,
          const u64* __restrict__ c_sub, u32 t_sub, u32 sub_pos_base, u32* __restrict__ sub_perm,
          i64* __restrict__ sub_pid)
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
  const u32 q0 = sub_pos_base + (u32)c_sub[s];
  const bool lo_sub = nlo >= 2 && nlo <= t_sub, hi_sub = nhi >= 2 && nhi <= t_sub;
  const bool lo_act = nlo > 1 && !lo_sub;
  const u32 r = perm[p];
  const i64 id = pid[p];
  if (!hi)
  {
    if (lo_sub)
    {
      const u32 dst = q0 + (p - S) - hb;
      sub_perm[dst] = r;
      sub_pid[dst] = id;
    }
    else if (nlo >= 2)
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
    if (hi_sub)
    {
      const u32 dst = q0 + (lo_sub ? nlo : 0u) + hb;
      sub_perm[dst] = r;
      sub_pid[dst] = id;
    }
    else if (nhi >= 2)
    {
      const u32 dst = p0 + (lo_act ? nlo : 0u) + hb;
      perm_n[dst] = r;
      pid_n[dst] = id;
      seg_of_n[dst] = a0 + (lo_act ? 1u : 0u);
    }
    else
    {
      t_id[r0 + (nlo > 0 ? 1u : 0u)] = id;
      t_src[r0 + (nlo > 0 ? 1u : 0u)] = (int)r;
    }
  }
}

// chunk count per big-list slot (a range whose sums are derived from its parent and sibling needs none)
// counters[6] += points of the derived ranges (accounting: their rows are not read by the statistics pass)
__global__ void k_big_chunks(const u32* __restrict__ count, const u32* __restrict__ big_list, u32* counters, u32* chunks,
                             u32 bound, const u32* __restrict__ bl_parent)
{
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= bound) return;
  const u32 nbig = counters[0];
  const bool summed = i < nbig && (bl_parent == nullptr || bl_parent[i] == 0xffffffffu);
  chunks[i] = summed ? (count[big_list[i]] + VI_CHUNK - 1) / VI_CHUNK : 0u;
  if (i < nbig && !summed) atomicAdd(&counters[6], count[big_list[i]]);
}

// bslot of a level that did not come out of k_emit_children (the root, the roots of a rank's forest)
__global__ void k_init_bslot(u32* __restrict__ bslot, u32 R, const u32* __restrict__ big_list, u32 nbig, u32* __restrict__ bl_parent,
                             u32* __restrict__ bl_sib)
{
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nbig)
  {
    bslot[big_list[i]] = i;
    bl_parent[i] = 0xffffffffu;
    bl_sib[i] = 0xffffffffu;
  }
}

__global__ void k_totals(const u32* c_rows, const u64* c_actpos, const u64* c_sub, u32 R, const u32* counters,
                         const u32* chunk_first, u32 chunk_bound, LevelTotals* out)
{
  out->subs = (u32)(c_sub[R] >> 32);
  out->subpos = (u32)c_sub[R];
  out->rows = c_rows[R];
  out->segs = (u32)(c_actpos[R] >> 32);
  out->pos = (u32)c_actpos[R];
  out->nbig = counters[0];
  out->chunks = chunk_first ? chunk_first[chunk_bound] : 0u;
  out->err = counters[1];
  out->minseg = counters[4];
  out->maxseg = counters[5];
  out->derived = counters[6];
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
