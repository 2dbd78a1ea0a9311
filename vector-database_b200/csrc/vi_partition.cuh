// vi_partition.cuh -- stable segmented partition of every open range at once (IndexBuilder.cs:99-129), three kernels
// per level, all sizes read from the device-resident level record (LevelDev), grids launched on host-side bounds:
//
//   k_flags     hi(p) = value > Mid || (value == Mid && id > Id)  (IndexBuilder.cs:115): one bit per position, the
//               hi-count prefix of every flag word inside its 2048-position tile, and the tile totals; the last CTA
//               to finish scans the tile totals (threadfence + ticket; no spinning, no second launch)
//   k_children  per range: low/high sizes -> what becomes of the two children (leaf row, sub-tree list entry,
//               next-level range, big range, derived big range) and an 8-component exclusive scan of those counts
//               over all ranges (tile-local here, tile prefixes by the last CTA), which also yields the next level's
//               record: sizes, first free row, sub-tree list cursor, chunk count
//   k_scatter   stable scatter of (row index, id, range index) into the compacted next-level position space or the
//               sub-tree position space; a single-point child is a leaf: its row gets Id = that point's id.  The
//               thread of a range's first position also emits the range's child rows (2r+1, 2r+2; an empty child
//               gets no row, IndexBuilder.cs:70-73) and next-level descriptors.
#pragma once
#include "vi_common.cuh"
#include "vi_scan.cuh"

constexpr u32 VI_NOSLOT = 0xffffffffu;

struct TableOut
{
  i64* t_rid;
  int* t_dim;
  float* t_mid;
  i64* t_id;
  int* t_low;
  int* t_high;
};

struct SubList  // device arrays, one entry per sub-tree root
{
  u32* start;  // first position in the sub-tree position space (sub_perm / sub_pid)
  u32* count;
  i64* rid;
  u32* row;
  u32* depth;
};

// ---- k_flags ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_flags(const LevelDev* __restrict__ lvp, u32* __restrict__ ticket, SegLevel sg, const u32* __restrict__ seg_of,
        const u32* __restrict__ perm, const i64* __restrict__ pid, const float* __restrict__ rows, int ld,
        u32* __restrict__ fbits, u32* __restrict__ wloc, u32* __restrict__ ftile)
{
  __shared__ u32 s_bits[FL_WORDS];
  __shared__ u32 s_last;
  const u32 A = lvp->A;
  if (lvp->R == 0) return;
  const u32 ntiles = A / FL_TILE + 1;  // covers position A itself: hi_before(A) needs its word
  const u32 tile = blockIdx.x;
  if (tile >= ntiles) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 base = tile * FL_TILE + threadIdx.x;
  u32 seg[FL_ITEMS], row[FL_ITEMS];
#pragma unroll
  for (int i = 0; i < FL_ITEMS; ++i)
  {
    const u32 p = base + i * 256;
    seg[i] = p < A ? seg_of[p] : 0u;
    row[i] = p < A ? perm[p] : 0u;
  }
  float v[FL_ITEMS], mid[FL_ITEMS];
#pragma unroll
  for (int i = 0; i < FL_ITEMS; ++i)
  {
    const u32 p = base + i * 256;
    v[i] = 0.f;
    mid[i] = 0.f;
    if (p < A)
    {
      mid[i] = sg.mid[seg[i]];
      v[i] = ldg_f_gather(rows + (size_t)row[i] * ld + sg.dim[seg[i]]);
    }
  }
#pragma unroll
  for (int i = 0; i < FL_ITEMS; ++i)
  {
    const u32 p = base + i * 256;
    bool hi = false;
    if (p < A) hi = v[i] > mid[i] || (v[i] == mid[i] && pid[p] > sg.pivot[seg[i]]);
    const u32 b = __ballot_sync(0xffffffffu, hi);
    if (lane == 0) s_bits[i * 8 + warp] = b;  // position = tile*2048 + i*256 + warp*32 + lane
  }
  __syncthreads();
  if (warp == 0)
  {
    const u32 b0 = s_bits[2 * lane], b1 = s_bits[2 * lane + 1];
    const u32 c0 = __popc(b0), c1 = __popc(b1);
    const u32 incl = warp_inclusive_scan(c0 + c1);
    const u32 excl = incl - (c0 + c1);
    const u32 w = tile * FL_WORDS + 2 * lane;
    *reinterpret_cast<uint2*>(fbits + w) = make_uint2(b0, b1);
    *reinterpret_cast<uint2*>(wloc + w) = make_uint2(excl, excl + c0);
    if (lane == 31)
    {
      ftile[tile] = incl;
      __threadfence();
      s_last = (atomicAdd(ticket, 1u) == ntiles - 1u) ? 1u : 0u;
    }
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  cta_scan_inplace<u32, 256>(ftile, ntiles);
}

// ---- what becomes of the two children of a range ---------------------------------------------------------------------
struct Children
{
  u32 nlo, nhi;
  bool lo_sub, hi_sub;          // 2..t_sub points: finished by the sub-tree kernel
  bool lo_act, hi_act;          // more: a range of the next level
  bool lo_big, hi_big;          // >= t_slot points: the range gets a slot of the level's big list (its integer sums
                                //   can be kept in gacc); of those, ranges of >= t_big points are summed in chunks
  bool pair;                    // both slotted and the parent's integer sums are kept: only the smaller one is summed
  bool lo_derived, hi_derived;  // the larger one's sums are parent - sibling (ties: low is summed)
  u32 lo_chunks, hi_chunks;     // chunk CTAs of the statistics kernel
};

__device__ __forceinline__ u32 chunks_of(u32 n) { return (n + VI_CHUNK - 1) / VI_CHUNK; }

__device__ __forceinline__ Children classify_children(u32 n, u32 nlo, u32 t_sub, u32 t_slot, u32 t_big, bool can_pair)
{
  Children c;
  c.nlo = nlo;
  c.nhi = n - nlo;
  c.lo_sub = c.nlo >= 2 && c.nlo <= t_sub;
  c.hi_sub = c.nhi >= 2 && c.nhi <= t_sub;
  c.lo_act = c.nlo > 1 && !c.lo_sub;
  c.hi_act = c.nhi > 1 && !c.hi_sub;
  c.lo_big = c.lo_act && c.nlo >= t_slot;
  c.hi_big = c.hi_act && c.nhi >= t_slot;
  c.pair = can_pair && c.lo_big && c.hi_big;
  c.lo_derived = c.pair && c.nlo > c.nhi;
  c.hi_derived = c.pair && c.nlo <= c.nhi;
  c.lo_chunks = (c.lo_big && !c.lo_derived && c.nlo >= t_big) ? chunks_of(c.nlo) : 0u;
  c.hi_chunks = (c.hi_big && !c.hi_derived && c.nhi >= t_big) ? chunks_of(c.nhi) : 0u;
  return c;
}

__device__ __forceinline__ ChildAgg child_counts(const Children& c)
{
  ChildAgg a;
  a.rows = (u32)(c.nlo > 0) + (u32)(c.nhi > 0);
  a.act_cnt = (u32)c.lo_act + (u32)c.hi_act;
  a.act_pos = (c.lo_act ? c.nlo : 0u) + (c.hi_act ? c.nhi : 0u);
  a.sub_cnt = (u32)c.lo_sub + (u32)c.hi_sub;
  a.sub_pos = (c.lo_sub ? c.nlo : 0u) + (c.hi_sub ? c.nhi : 0u);
  a.big_cnt = (u32)c.lo_big + (u32)c.hi_big;
  a.chunks = c.lo_chunks + c.hi_chunks;
  a.derived = (c.lo_derived ? c.nlo : 0u) + (c.hi_derived ? c.nhi : 0u);
  return a;
}

__device__ __forceinline__ ChildAgg agg_zero()
{
  ChildAgg a;
  a.rows = a.act_cnt = a.act_pos = a.sub_cnt = a.sub_pos = a.big_cnt = a.chunks = a.derived = 0u;
  return a;
}

__device__ __forceinline__ ChildAgg agg_add(const ChildAgg& x, const ChildAgg& y)
{
  ChildAgg a;
  a.rows = x.rows + y.rows;
  a.act_cnt = x.act_cnt + y.act_cnt;
  a.act_pos = x.act_pos + y.act_pos;
  a.sub_cnt = x.sub_cnt + y.sub_cnt;
  a.sub_pos = x.sub_pos + y.sub_pos;
  a.big_cnt = x.big_cnt + y.big_cnt;
  a.chunks = x.chunks + y.chunks;
  a.derived = x.derived + y.derived;
  return a;
}

__device__ __forceinline__ ChildAgg agg_sub(const ChildAgg& x, const ChildAgg& y)
{
  ChildAgg a;
  a.rows = x.rows - y.rows;
  a.act_cnt = x.act_cnt - y.act_cnt;
  a.act_pos = x.act_pos - y.act_pos;
  a.sub_cnt = x.sub_cnt - y.sub_cnt;
  a.sub_pos = x.sub_pos - y.sub_pos;
  a.big_cnt = x.big_cnt - y.big_cnt;
  a.chunks = x.chunks - y.chunks;
  a.derived = x.derived - y.derived;
  return a;
}

__device__ __forceinline__ ChildAgg agg_shfl_up(const ChildAgg& x, int d)
{
  ChildAgg a;
  a.rows = __shfl_up_sync(0xffffffffu, x.rows, d);
  a.act_cnt = __shfl_up_sync(0xffffffffu, x.act_cnt, d);
  a.act_pos = __shfl_up_sync(0xffffffffu, x.act_pos, d);
  a.sub_cnt = __shfl_up_sync(0xffffffffu, x.sub_cnt, d);
  a.sub_pos = __shfl_up_sync(0xffffffffu, x.sub_pos, d);
  a.big_cnt = __shfl_up_sync(0xffffffffu, x.big_cnt, d);
  a.chunks = __shfl_up_sync(0xffffffffu, x.chunks, d);
  a.derived = __shfl_up_sync(0xffffffffu, x.derived, d);
  return a;
}

__device__ __forceinline__ ChildAgg agg_warp_inclusive(ChildAgg v)
{
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1)
  {
    const ChildAgg t = agg_shfl_up(v, d);
    if (lane >= d) v = agg_add(v, t);
  }
  return v;
}

__device__ __forceinline__ ChildAgg agg_load_cg(const ChildAgg* p)
{
  const uint4 a = __ldcg(reinterpret_cast<const uint4*>(p));
  const uint4 b = __ldcg(reinterpret_cast<const uint4*>(p) + 1);
  ChildAgg r;
  r.rows = a.x; r.act_cnt = a.y; r.act_pos = a.z; r.sub_cnt = a.w;
  r.sub_pos = b.x; r.big_cnt = b.y; r.chunks = b.z; r.derived = b.w;
  return r;
}

__device__ __forceinline__ ChildAgg agg_load(const ChildAgg* p)
{
  const uint4 a = *reinterpret_cast<const uint4*>(p);
  const uint4 b = *(reinterpret_cast<const uint4*>(p) + 1);
  ChildAgg r;
  r.rows = a.x; r.act_cnt = a.y; r.act_pos = a.z; r.sub_cnt = a.w;
  r.sub_pos = b.x; r.big_cnt = b.y; r.chunks = b.z; r.derived = b.w;
  return r;
}

__device__ __forceinline__ void agg_store(ChildAgg* p, const ChildAgg& v)
{
  *reinterpret_cast<uint4*>(p) = make_uint4(v.rows, v.act_cnt, v.act_pos, v.sub_cnt);
  *(reinterpret_cast<uint4*>(p) + 1) = make_uint4(v.sub_pos, v.big_cnt, v.chunks, v.derived);
}

__device__ __forceinline__ void lv_store(LevelDev* dst, const LevelDev& v)
{
  const uint4* s = reinterpret_cast<const uint4*>(&v);
  uint4* d = reinterpret_cast<uint4*>(dst);
  d[0] = s[0];
  d[1] = s[1];
  d[2] = s[2];
  d[3] = s[3];
}

// CTA-wide exclusive scan of one ChildAgg per thread (256 threads); returns the thread's exclusive prefix and the total
__device__ __forceinline__ ChildAgg agg_cta_exclusive(const ChildAgg& mine, ChildAgg* s_w /*[8]*/, ChildAgg& total)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const ChildAgg incl = agg_warp_inclusive(mine);
  if (lane == 31) s_w[warp] = incl;
  __syncthreads();
  ChildAgg woff = agg_zero();
  total = agg_zero();
#pragma unroll
  for (int w = 0; w < 8; ++w)
  {
    const ChildAgg y = s_w[w];
    if (w < warp) woff = agg_add(woff, y);
    total = agg_add(total, y);
  }
  __syncthreads();  // s_w may be reused by the caller's next round
  return agg_add(woff, agg_sub(incl, mine));
}

constexpr int CH_ITEMS = 4;
constexpr int CH_TILE = 256 * CH_ITEMS;  // ranges per CTA of k_children
constexpr int CH_TILE_LOG2 = 10;
static_assert((1 << CH_TILE_LOG2) == CH_TILE, "tile size");

// ---- k_children ------------------------------------------------------------------------------------------------------
// nxt / h_nxt: the next level's record on the device and its pinned host copy.
__global__ void __launch_bounds__(256)
k_children(const LevelDev* __restrict__ cur, u32* __restrict__ ticket, LevelDev* __restrict__ nxt, LevelDev* __restrict__ h_nxt,
           SegLevel sg, FlagScan fs, u32* __restrict__ seg_nlo, u32* __restrict__ seg_hbase, ChildAgg* __restrict__ c_pre,
           ChildAgg* __restrict__ ctile, uint2* __restrict__ ctile_mm, u32 t_sub, u32 t_slot, u32 t_big, int sibling,
           u32 t_cap, u32* __restrict__ chunk_first_next)
{
  __shared__ ChildAgg s_w[8];
  __shared__ u32 s_min[8], s_max[8];
  __shared__ u32 s_last;
  const u32 R = cur->R;
  if (R == 0)
  {
    // nothing is open: the level is a no-op (the host runs a level or two ahead of what it knows); forward the record
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
      LevelDev n = *cur;
      n.A = n.R = n.nbig = n.chunks = n.derived = n.rows = 0u;
      n.minseg = 0xffffffffu;
      n.maxseg = 0u;
      n.ticket[0] = n.ticket[1] = 0u;
      lv_store(nxt, n);
      lv_store(h_nxt, n);
    }
    return;
  }
  const u32 ntiles = (R + CH_TILE - 1) / CH_TILE;
  const u32 tile = blockIdx.x;
  if (tile >= ntiles) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 s0 = tile * CH_TILE + threadIdx.x * CH_ITEMS;
  ChildAgg ex[CH_ITEMS];
  ChildAgg run = agg_zero();
  u32 mn = 0xffffffffu, mx = 0u;
#pragma unroll
  for (int j = 0; j < CH_ITEMS; ++j)
  {
    const u32 s = s0 + j;
    ex[j] = run;
    if (s < R)
    {
      const u32 S = sg.start[s], n = sg.count[s];
      const u32 hb = hi_before(fs, S);
      const u32 nhi = hi_before(fs, S + n) - hb;
      const u32 nlo = n - nhi;
      seg_nlo[s] = nlo;
      seg_hbase[s] = hb;
      const Children c = classify_children(n, nlo, t_sub, t_slot, t_big, sibling && sg.bslot[s] != VI_NOSLOT);
      run = agg_add(run, child_counts(c));
      if (c.lo_act) { mn = min(mn, nlo); mx = max(mx, nlo); }
      if (c.hi_act) { mn = min(mn, nhi); mx = max(mx, nhi); }
    }
  }
  ChildAgg total;
  const ChildAgg off = agg_cta_exclusive(run, s_w, total);
#pragma unroll
  for (int j = 0; j < CH_ITEMS; ++j)
    if (s0 + j < R) agg_store(c_pre + s0 + j, agg_add(ex[j], off));
  mn = __reduce_min_sync(0xffffffffu, mn);
  mx = __reduce_max_sync(0xffffffffu, mx);
  if (lane == 0) { s_min[warp] = mn; s_max[warp] = mx; }
  __syncthreads();
  if (threadIdx.x == 0)
  {
#pragma unroll
    for (int w = 1; w < 8; ++w) { mn = min(mn, s_min[w]); mx = max(mx, s_max[w]); }
    agg_store(ctile + tile, total);
    ctile_mm[tile] = make_uint2(mn, mx);
    __threadfence();
    s_last = (atomicAdd(ticket, 1u) == ntiles - 1u) ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last) return;
  // ---- last CTA: tile aggregates -> exclusive tile prefixes, totals -> the next level's record ---------------------
  __threadfence();
  ChildAgg carry = agg_zero();
  u32 gmn = 0xffffffffu, gmx = 0u;
  for (u32 base = 0; base < ntiles; base += 256)
  {
    const u32 t = base + threadIdx.x;
    ChildAgg a = agg_zero();
    if (t < ntiles)
    {
      a = agg_load_cg(ctile + t);
      const uint2 mm = __ldcg(ctile_mm + t);
      gmn = min(gmn, mm.x);
      gmx = max(gmx, mm.y);
    }
    ChildAgg tot;
    const ChildAgg exl = agg_cta_exclusive(a, s_w, tot);
    if (t < ntiles) agg_store(ctile + t, agg_add(carry, exl));
    carry = agg_add(carry, tot);  // identical in every thread
  }
  gmn = __reduce_min_sync(0xffffffffu, gmn);
  gmx = __reduce_max_sync(0xffffffffu, gmx);
  if (lane == 0) { s_min[warp] = gmn; s_max[warp] = gmx; }
  __syncthreads();
  if (threadIdx.x == 0)
  {
#pragma unroll
    for (int w = 1; w < 8; ++w) { gmn = min(gmn, s_min[w]); gmx = max(gmx, s_max[w]); }
    LevelDev n;
    const bool err = cur->err != 0u || (u64)cur->row_next + (u64)carry.rows > (u64)t_cap;
    n.A = err ? 0u : carry.act_pos;
    n.R = err ? 0u : carry.act_cnt;
    n.nbig = err ? 0u : carry.big_cnt;
    n.chunks = err ? 0u : carry.chunks;
    n.minseg = gmn;
    n.maxseg = gmx;
    n.derived = carry.derived;
    n.row_next = cur->row_next + (err ? 0u : carry.rows);
    n.sub_cnt = cur->sub_cnt + (err ? 0u : carry.sub_cnt);
    n.sub_pos = cur->sub_pos + (err ? 0u : carry.sub_pos);
    n.err = err ? 1u : 0u;
    n.rows = err ? 0u : carry.rows;
    n.ticket[0] = n.ticket[1] = 0u;
    n.pad[0] = n.pad[1] = 0u;
    lv_store(nxt, n);
    lv_store(h_nxt, n);
    chunk_first_next[n.nbig] = n.chunks;
  }
}

// per-range child sizes only (shared phase of the multi-rank build: the bookkeeping there follows global sizes)
__global__ void __launch_bounds__(256)
k_seg_nlo(const LevelDev* __restrict__ lvp, SegLevel sg, FlagScan fs, u32* __restrict__ seg_nlo, u32* __restrict__ seg_hbase)
{
  const u32 s = blockIdx.x * 256u + threadIdx.x;
  if (s >= lvp->R) return;
  const u32 S = sg.start[s], n = sg.count[s];
  const u32 hb = hi_before(fs, S);
  seg_nlo[s] = n - (hi_before(fs, S + n) - hb);
  seg_hbase[s] = hb;
}

// ---- k_scatter -------------------------------------------------------------------------------------------------------
struct NextLevel  // what k_scatter writes for the next level
{
  SegLevel seg;
  u32* perm;
  i64* pid;
  u32* seg_of;
  u32* big_list;
  u32* bl_parent;
  u32* bl_sib;
  u32* chunk_first;
};

__global__ void __launch_bounds__(256)
k_scatter(const LevelDev* __restrict__ cur, const LevelDev* __restrict__ nxt, SegLevel sg, const u32* __restrict__ seg_of,
          const u32* __restrict__ perm, const i64* __restrict__ pid, FlagScan fs, const u32* __restrict__ seg_nlo,
          const u32* __restrict__ seg_hbase, const ChildAgg* __restrict__ c_pre, const ChildAgg* __restrict__ ctile,
          u32 t_sub, u32 t_slot, u32 t_big, int sibling, u32 child_depth, NextLevel nx, TableOut t, int* __restrict__ t_src,
          SubList sub,
          u32* __restrict__ sub_perm, i64* __restrict__ sub_pid)
{
  const u32 p = blockIdx.x * 256u + threadIdx.x;
  if (p >= cur->A || cur->R == 0u || nxt->err) return;
  const u32 s = seg_of[p];
  const u32 S = sg.start[s], n = sg.count[s];
  const u32 parent_slot = sibling ? sg.bslot[s] : VI_NOSLOT;
  const Children c = classify_children(n, seg_nlo[s], t_sub, t_slot, t_big, parent_slot != VI_NOSLOT);
  const ChildAgg P = agg_add(agg_load(c_pre + s), agg_load(ctile + (s >> CH_TILE_LOG2)));
  const u32 r0 = cur->row_next + P.rows;      // first child row
  const u32 a0 = P.act_cnt, p0 = P.act_pos;   // next-level range index / first position of the first staying child
  const u32 b0 = cur->sub_cnt + P.sub_cnt, q0 = cur->sub_pos + P.sub_pos;  // sub-tree list entry / position
  const u32 nlo = c.nlo, nhi = c.nhi;
  const u32 w = fs.fbits[p >> 5];
  const bool hi = (w >> (p & 31)) & 1u;
  const u32 hb = hi_before(fs, p) - seg_hbase[s];  // hi points of this range before p
  const u32 r = perm[p];
  const i64 id = pid[p];
  if (!hi)
  {
    const u32 rank = (p - S) - hb;
    if (c.lo_sub)
    {
      sub_perm[q0 + rank] = r;
      sub_pid[q0 + rank] = id;
    }
    else if (c.lo_act)
    {
      nx.perm[p0 + rank] = r;
      nx.pid[p0 + rank] = id;
      nx.seg_of[p0 + rank] = a0;
    }
    else
    {
      t.t_id[r0] = id;  // the single low point is a leaf: RangeValue.Id = its id (IndexBuilder.cs:81-82)
      t_src[r0] = (int)r;
    }
  }
  else
  {
    if (c.hi_sub)
    {
      const u32 dst = q0 + (c.lo_sub ? nlo : 0u) + hb;
      sub_perm[dst] = r;
      sub_pid[dst] = id;
    }
    else if (c.hi_act)
    {
      const u32 dst = p0 + (c.lo_act ? nlo : 0u) + hb;
      nx.perm[dst] = r;
      nx.pid[dst] = id;
      nx.seg_of[dst] = a0 + (c.lo_act ? 1u : 0u);
    }
    else
    {
      const u32 lr = r0 + (nlo > 0 ? 1u : 0u);
      t.t_id[lr] = id;
      t_src[lr] = (int)r;
    }
  }
  if (p != S) return;
  // ---- the range's first position also writes its child rows and next-level descriptors ----------------------------
  const i64 rid = sg.rid[s];
  const u32 row = sg.row[s];
  const int lo_row = nlo > 0 ? (int)r0 : -1;
  const int hi_row = nhi > 0 ? (int)(r0 + (nlo > 0 ? 1u : 0u)) : -1;
  t.t_low[row] = lo_row;
  t.t_high[row] = hi_row;
  const u32 lo_slot = c.lo_big ? P.big_cnt : VI_NOSLOT;
  const u32 hi_slot = c.hi_big ? P.big_cnt + (c.lo_big ? 1u : 0u) : VI_NOSLOT;
  const u32 lo_chunks = c.lo_chunks;
  if (nlo > 0)
  {
    t.t_rid[lo_row] = rid * 2 + 1;  // IndexBuilder.cs:99
    t.t_low[lo_row] = -1;
    t.t_high[lo_row] = -1;
    if (nlo == 1)
    {
      t.t_dim[lo_row] = -1;  // leaf, IndexBuilder.cs:81-82; its Id is written by the point's own thread above
      t.t_mid[lo_row] = 0.0f;
    }
    else if (c.lo_sub)
    {
      sub.start[b0] = q0;
      sub.count[b0] = nlo;
      sub.rid[b0] = rid * 2 + 1;
      sub.row[b0] = (u32)lo_row;
      sub.depth[b0] = child_depth;
    }
    else
    {
      nx.seg.start[a0] = p0;
      nx.seg.count[a0] = nlo;
      nx.seg.rid[a0] = rid * 2 + 1;
      nx.seg.row[a0] = (u32)lo_row;
      nx.seg.bslot[a0] = lo_slot;
      if (c.lo_big)
      {
        nx.big_list[lo_slot] = a0;
        nx.bl_parent[lo_slot] = c.lo_derived ? parent_slot : VI_NOSLOT;
        nx.bl_sib[lo_slot] = c.pair ? hi_slot : VI_NOSLOT;
        nx.chunk_first[lo_slot] = P.chunks;
      }
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
    else if (c.hi_sub)
    {
      const u32 b1 = b0 + (c.lo_sub ? 1u : 0u);
      sub.start[b1] = q0 + (c.lo_sub ? nlo : 0u);
      sub.count[b1] = nhi;
      sub.rid[b1] = rid * 2 + 2;
      sub.row[b1] = (u32)hi_row;
      sub.depth[b1] = child_depth;
    }
    else
    {
      const u32 a1 = a0 + (c.lo_act ? 1u : 0u);
      nx.seg.start[a1] = p0 + (c.lo_act ? nlo : 0u);
      nx.seg.count[a1] = nhi;
      nx.seg.rid[a1] = rid * 2 + 2;
      nx.seg.row[a1] = (u32)hi_row;
      nx.seg.bslot[a1] = hi_slot;
      if (c.hi_big)
      {
        nx.big_list[hi_slot] = a1;
        nx.bl_parent[hi_slot] = c.hi_derived ? parent_slot : VI_NOSLOT;
        nx.bl_sib[hi_slot] = c.pair ? lo_slot : VI_NOSLOT;
        nx.chunk_first[hi_slot] = P.chunks + lo_chunks;
      }
    }
  }
}

// bslot of a level that did not come out of k_scatter (the root, the roots of a rank's forest)
__global__ void k_init_bslot(u32* __restrict__ bslot, u32 R, const u32* __restrict__ big_list, u32 nbig, u32* __restrict__ bl_parent,
                             u32* __restrict__ bl_sib)
{
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nbig)
  {
    bslot[big_list[i]] = i;
    bl_parent[i] = VI_NOSLOT;
    bl_sib[i] = VI_NOSLOT;
  }
}

__global__ void k_pack_nodes(const int* __restrict__ t_dim, const float* __restrict__ t_mid, const i64* __restrict__ t_id,
                             const int* __restrict__ t_low, const int* __restrict__ t_high, int4* __restrict__ node, u32 n)
{
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int d = t_dim[i];
  int4 v;
  v.x = d == VI_DIM_NULL ? VI_NODE_BOTH : d;  // Dimension = null (VI_MODE_SQL): the walk follows both children
  v.y = __float_as_int(t_mid[i]);
  if (d < 0 && d != VI_DIM_NULL)
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
