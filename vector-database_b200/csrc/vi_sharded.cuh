// vi_sharded.cuh -- kernels of the multi-rank build (one process per GPU; protocol in vi_build.cu / DESIGN.md).
//
// Shared phase: every rank holds its local slice of every range of the top levels; the per-range bookkeeping
// (child rows, next-level ranges, destination offsets) is done on the host from all-reduced counts because there
// are at most 2^L ranges; the device only scatters.
// Ownership exchange: k_pack_rows gathers the local rows of each range into the all-to-all send buffer;
// k_forest_init lays the received pieces out per owned range in source-rank order (= the global stable order).
#pragma once
#include "vi_common.cuh"

constexpr u32 VI_NONE = 0xffffffffu;

struct ShScatter  // per range s of the current shared level (device arrays)
{
  const u32* lo_dst;   // first next-level position of the low child, VI_NONE if it does not stay a range
  const u32* hi_dst;
  const u32* lo_seg;   // next-level range index of the low / high child
  const u32* hi_seg;
  const int* lo_leaf;  // table row of a low / high child that is a (global) single point, else -1
  const int* hi_leaf;
};

__global__ void __launch_bounds__(256)
k_scatter_shared(SegLevel sg, const u32* __restrict__ seg_of, const u32* __restrict__ perm, const i64* __restrict__ pid,
                 u32 A, FlagScan fs, const u32* __restrict__ seg_hbase, ShScatter sc, u32* __restrict__ perm_n, i64* __restrict__ pid_n, u32* __restrict__ seg_of_n,
                 u64* __restrict__ leaf_ids, int* __restrict__ t_src)
{
  const u32 p = blockIdx.x * 256u + threadIdx.x;
  if (p >= A) return;
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
