// vi_sharded.cuh -- kernels of the multi-rank build (one process per GPU; protocol in vi_build.cu / DESIGN.md).
//
// Shared phase: every rank holds its local slice of every range of the top levels.  The per-range bookkeeping
// (child rows, next-level ranges, destination offsets) follows the all-reduced GLOBAL child sizes and is done by one
// thread on the device (k_sh_next; there are at most 2^L ranges), identically on every rank, so the host enqueues the
// whole phase -- kernels and NCCL collectives on one stream -- without a single synchronisation.
// Ownership exchange: k_pack_rows gathers the local rows of each range into the all-to-all send buffer;
// k_forest_init lays the received pieces out per owned range in source-rank order (= the global stable order).
#pragma once
#include "vi_common.cuh"
#include "vi_partition.cuh"

constexpr u32 VI_NONE = 0xffffffffu;
constexpr int VI_SH_MAXR = 128;     // ranges of a shared level: 2^L with L = ceil(log2 world) + 1, world <= 64
constexpr int VI_SH_MAXROWS = 1024; // table rows made by the shared levels

// The ranges of one shared level (device; the final one is copied to the host).  Identical on every rank except for
// the local slices (start, lcount).
struct ShLevel
{
  u32 R;          // ranges with >= 2 points (globally)
  u32 T;          // shared table rows so far
  u32 err_level;  // first level at which a range was too tightly clustered for the integer statistics (else VI_NONE)
  u32 pad;
  u32 start[VI_SH_MAXR];   // local slice
  u32 lcount[VI_SH_MAXR];
  u32 row[VI_SH_MAXR];
  u32 pad2[VI_SH_MAXR];
  u64 gcount[VI_SH_MAXR];  // points over all ranks
  i64 rid[VI_SH_MAXR];
};

struct ShScatter  // per range s of the current shared level (device arrays of VI_SH_MAXR entries)
{
  u32* lo_dst;   // first next-level position of the low child, VI_NONE if it does not stay a range
  u32* hi_dst;
  u32* lo_seg;   // next-level range index of the low / high child
  u32* hi_seg;
  int* lo_leaf;  // table row of a low / high child that is a (global) single point, else -1
  int* hi_leaf;
};

// level 0 of the shared phase (after k_init_level0): the root over all ranks
__global__ void k_sh_begin(ShLevel* __restrict__ sh, LevelDev* __restrict__ lv0, u32 nloc, u64 nglobal, u32* __restrict__ chunk_first,
                           SegLevel sg, u32* __restrict__ big_list, TableOut t)
{
  int* t_dim = t.t_dim;
  float* t_mid = t.t_mid;
  sg.start[0] = 0;  // (k_init_level0 writes the same when this rank holds points; a rank with an empty shard relies on this)
  sg.count[0] = nloc;
  sg.rid[0] = 0;
  sg.row[0] = 0;
  big_list[0] = 0;
  t.t_rid[0] = 0;
  t.t_low[0] = -1;
  t.t_high[0] = -1;
  sh->R = nglobal >= 2 ? 1u : 0u;
  sh->T = 1;
  sh->err_level = VI_NONE;
  sh->start[0] = 0;
  sh->lcount[0] = nloc;
  sh->row[0] = 0;
  sh->gcount[0] = nglobal;
  sh->rid[0] = 0;
  LevelDev v;
  v.A = nglobal >= 2 ? nloc : 0u;
  v.R = sh->R;
  v.nbig = sh->R;
  v.chunks = nglobal >= 2 ? chunks_of(nloc) : 0u;
  v.minseg = v.maxseg = nloc;
  v.derived = 0;
  v.row_next = 1;
  v.sub_cnt = v.sub_pos = v.err = v.rows = 0;
  v.ticket[0] = v.ticket[1] = 0;
  v.pad[0] = v.pad[1] = 0;
  lv_store(lv0, v);
  chunk_first[0] = 0;
  chunk_first[1] = v.chunks;
  if (nglobal == 1)
  {
    t_dim[0] = -1;  // IndexBuilder.cs:81-82: the only point is a leaf; its Id arrives with the leaf-id all-reduce
    t_mid[0] = 0.f;
  }
}

// local child sizes of the level's ranges as u64 words [2*Rb] (zero beyond R): the input of the size all-reduce
__global__ void k_sh_counts(const ShLevel* __restrict__ sh, const u32* __restrict__ seg_nlo, u32 Rb, u64* __restrict__ cnt_loc,
                            u64* __restrict__ cnt_glb)
{
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Rb) return;
  u64 lo = 0, hi = 0;
  if (i < sh->R)
  {
    const u32 lc = sh->lcount[i];
    lo = lc ? seg_nlo[i] : 0u;
    hi = lc - lo;
  }
  cnt_loc[2 * i] = lo;
  cnt_loc[2 * i + 1] = hi;
  cnt_glb[2 * i] = lo;
  cnt_glb[2 * i + 1] = hi;
}

// One thread: children of the level's ranges from the GLOBAL sizes (same order as the single-rank builder: by range,
// low child before high child): rows, leaves, next-level ranges, local destinations, the next level's records.
__global__ void k_sh_next(const ShLevel* __restrict__ cur, ShLevel* __restrict__ nxt, const u64* __restrict__ cnt_loc,
                          const u64* __restrict__ cnt_glb, u32 level, u32* __restrict__ stat_err, ShScatter sc, SegLevel nseg,
                          u32* __restrict__ big_list_n, u32* __restrict__ chunk_first_n, LevelDev* __restrict__ lv_n, TableOut t)
{
  const u32 R = cur->R;
  u32 T = cur->T, npos = 0, nR = 0, chunks = 0, rows0 = cur->T;
  u32 err = cur->err_level;
  if (*stat_err)
  {
    err = min(err, level);
    *stat_err = 0;
  }
  for (u32 i = 0; i < R; ++i)
  {
    sc.lo_dst[i] = VI_NONE; sc.hi_dst[i] = VI_NONE;
    sc.lo_seg[i] = 0; sc.hi_seg[i] = 0;
    sc.lo_leaf[i] = -1; sc.hi_leaf[i] = -1;
    const i64 rid = cur->rid[i];
    const u32 prow = cur->row[i];
    int link[2] = {-1, -1};
    for (int side = 0; side < 2; ++side)
    {
      const u64 g = cnt_glb[2 * i + side];
      const u32 l = (u32)cnt_loc[2 * i + side];
      if (g == 0) continue;  // empty range: no row (IndexBuilder.cs:70-73)
      if (T >= (u32)VI_SH_MAXROWS || nR >= (u32)VI_SH_MAXR) { err = min(err, level); continue; }
      const u32 row = T++;
      link[side] = (int)row;
      t.t_rid[row] = rid * 2 + 1 + side;
      t.t_low[row] = -1;
      t.t_high[row] = -1;
      if (g == 1)
      {
        t.t_dim[row] = -1;  // leaf; Id = the point's id, written by the rank that holds it (k_scatter_shared)
        t.t_mid[row] = 0.f;
        (side ? sc.hi_leaf : sc.lo_leaf)[i] = (int)row;
      }
      else
      {
        (side ? sc.hi_dst : sc.lo_dst)[i] = npos;
        (side ? sc.hi_seg : sc.lo_seg)[i] = nR;
        nxt->start[nR] = npos;
        nxt->lcount[nR] = l;
        nxt->gcount[nR] = g;
        nxt->rid[nR] = rid * 2 + 1 + side;
        nxt->row[nR] = row;
        nseg.start[nR] = npos;
        nseg.count[nR] = l;
        nseg.rid[nR] = rid * 2 + 1 + side;
        nseg.row[nR] = row;
        big_list_n[nR] = nR;
        chunk_first_n[nR] = chunks;
        chunks += chunks_of(l);
        npos += l;
        ++nR;
      }
    }
    t.t_low[prow] = link[0];
    t.t_high[prow] = link[1];
  }
  chunk_first_n[nR] = chunks;
  nxt->R = nR;
  nxt->T = T;
  nxt->err_level = err;
  LevelDev v;
  v.A = npos;
  v.R = nR;
  v.nbig = nR;
  v.chunks = chunks;
  v.minseg = 0;
  v.maxseg = 0;
  v.derived = 0;
  v.row_next = T;
  v.sub_cnt = v.sub_pos = v.err = 0;
  v.rows = T - rows0;
  v.ticket[0] = v.ticket[1] = 0;
  v.pad[0] = v.pad[1] = 0;
  lv_store(lv_n, v);
}

// ids of the shared leaf rows (all-reduced: the holder wrote, everybody else contributed 0)
__global__ void k_sh_leaf_ids(const ShLevel* __restrict__ sh, const u64* __restrict__ leaf_ids, const int* __restrict__ t_dim,
                              i64* __restrict__ t_id)
{
  const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < sh->T && t_dim[r] == -1) t_id[r] = (i64)leaf_ids[r];
}

// root rows of level-L ranges owned by another rank are placeholders here
__global__ void k_sh_placeholders(const u32* __restrict__ rows, u32 n, int* __restrict__ t_dim)
{
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) t_dim[rows[i]] = -2;
}

__global__ void __launch_bounds__(256)
k_scatter_shared(const LevelDev* __restrict__ lvp, SegLevel sg, const u32* __restrict__ seg_of, const u32* __restrict__ perm,
                 const i64* __restrict__ pid, FlagScan fs, const u32* __restrict__ seg_hbase, ShScatter sc,
                 u32* __restrict__ perm_n, i64* __restrict__ pid_n, u32* __restrict__ seg_of_n, u64* __restrict__ leaf_ids,
                 int* __restrict__ t_src)
{
  const u32 p = blockIdx.x * 256u + threadIdx.x;
  if (p >= lvp->A) return;
  const u32 s = seg_of[p];
  const u32 S = sg.start[s];
  const u32 w = fs.fbits[p >> 5];
  const bool hi = (w >> (p & 31)) & 1u;
  const u32 hb = hi_before(fs, p) - seg_hbase[s];
  const u32 rank = hi ? hb : (p - S) - hb;
  const u32 dstb = hi ? sc.hi_dst[s] : sc.lo_dst[s];
  const u32 r = perm[p];
  const i64 id = pid[p];
  if (dstb != VI_NONE)
  {
    perm_n[dstb + rank] = r;
    pid_n[dstb + rank] = id;
    seg_of_n[dstb + rank] = hi ? sc.hi_seg[s] : sc.lo_seg[s];
  }
  else
  {
    const int leaf = hi ? sc.hi_leaf[s] : sc.lo_leaf[s];
    if (leaf >= 0)
    {
      leaf_ids[leaf] = (u64)id;  // only the rank that holds the point writes; the host all-reduces the array
      t_src[leaf] = (int)r;
    }
  }
}

// send buffer row of position p: send_base[range] + (p - start[range]); one thread per float4
__global__ void __launch_bounds__(256)
k_pack_rows(SegLevel sg, const u32* __restrict__ seg_of, const u32* __restrict__ perm, const i64* __restrict__ pid,
            const float* __restrict__ rows, int ld, u32 A, const u32* __restrict__ send_base,
            float* __restrict__ send_rows, i64* __restrict__ send_ids)
{
  const int C4 = ld >> 2;
  const size_t idx = (size_t)blockIdx.x * 256u + threadIdx.x;
  const u32 p = (u32)(idx / C4);
  const int c = (int)(idx % C4);
  if (p >= A) return;
  const u32 s = seg_of[p];
  const u32 dst = send_base[s] + (p - sg.start[s]);
  const float4 v = reinterpret_cast<const float4*>(rows + (size_t)perm[p] * ld)[c];
  reinterpret_cast<float4*>(send_rows + (size_t)dst * ld)[c] = v;
  if (c == 0) send_ids[dst] = pid[p];
}

// Received pieces (sorted by destination position): position p of the forest takes row piece_src + (p - piece_dst)
// of the receive buffer and belongs to range piece_seg.
__global__ void __launch_bounds__(256)
k_forest_init(u32 A, u32 npieces, const u32* __restrict__ piece_dst, const u32* __restrict__ piece_src,
              const u32* __restrict__ piece_seg, const i64* __restrict__ ids, u32* __restrict__ perm,
              i64* __restrict__ pid, u32* __restrict__ seg_of)
{
  const u32 p = blockIdx.x * 256u + threadIdx.x;
  if (p >= A) return;
  u32 lo = 0, hi = npieces;
  while (hi - lo > 1)
  {
    const u32 m = (lo + hi) >> 1;
    if (piece_dst[m] <= p) lo = m; else hi = m;
  }
  const u32 r = piece_src[lo] + (p - piece_dst[lo]);
  perm[p] = r;
  pid[p] = ids[r];
  seg_of[p] = piece_seg[lo];
}

// ---- table replication (query-sharded search against a replicated table) ---------------------------------------
// Every row becomes 4 u64 words [rid][dim | mid][id][low | high]; a rank writes the rows it is responsible for into
// its slice of a zero-filled buffer (links already remapped to global row indexes), one sum-all-reduce concatenates.
__global__ void __launch_bounds__(256)
k_pack_table(const i64* __restrict__ t_rid, const int* __restrict__ t_dim, const float* __restrict__ t_mid,
             const i64* __restrict__ t_id, const int* __restrict__ t_low, const int* __restrict__ t_high, u32 rows,
             u32 shared, u32 own_offset, int contribute_shared, u64* __restrict__ out)
{
  const u32 i = blockIdx.x * 256u + threadIdx.x;
  if (i >= rows) return;
  const int dim = t_dim[i];
  u32 g;  // global row index
  if (i < shared)
  {
    // replicated rows come from rank 0 only; a level-L root row comes from its owner (placeholders have dim -2)
    const bool root_owned = t_low[i] >= (int)shared || t_high[i] >= (int)shared;
    if (dim == -2 || (!contribute_shared && !root_owned)) return;
    g = i;
  }
  else
    g = own_offset + (i - shared);
  int lo = t_low[i], hi = t_high[i];
  if (lo >= (int)shared) lo = (int)(own_offset + ((u32)lo - shared));
  if (hi >= (int)shared) hi = (int)(own_offset + ((u32)hi - shared));
  u64* o = out + (size_t)g * 4;
  o[0] = (u64)t_rid[i];
  o[1] = ((u64)(u32)__float_as_int(t_mid[i]) << 32) | (u64)(u32)dim;
  o[2] = (u64)t_id[i];
  o[3] = ((u64)(u32)hi << 32) | (u64)(u32)lo;
}

__global__ void __launch_bounds__(256)
k_unpack_table(const u64* __restrict__ in, u32 rows, i64* __restrict__ t_rid, int* __restrict__ t_dim,
               float* __restrict__ t_mid, i64* __restrict__ t_id, int* __restrict__ t_low, int* __restrict__ t_high,
               int* __restrict__ t_src)
{
  const u32 i = blockIdx.x * 256u + threadIdx.x;
  if (i >= rows) return;
  const u64* o = in + (size_t)i * 4;
  t_rid[i] = (i64)o[0];
  t_dim[i] = (int)(u32)o[1];
  t_mid[i] = __int_as_float((int)(u32)(o[1] >> 32));
  t_id[i] = (i64)o[2];
  t_low[i] = (int)(u32)o[3];
  t_high[i] = (int)(u32)(o[3] >> 32);
  t_src[i] = -1;  // the vectors stay with their owners: no candidate verification on a replicated table
}
