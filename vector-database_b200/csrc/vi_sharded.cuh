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
                 u32 A, const u32* __restrict__ fbits, const u32* __restrict__ wpre, const u32* __restrict__ seg_hbase,
                 ShScatter sc, u32* __restrict__ perm_n, i64* __restrict__ pid_n, u32* __restrict__ seg_of_n,
                 u64* __restrict__ leaf_ids, int* __restrict__ t_src)
{
  const u32 p = blockIdx.x * 256u + threadIdx.x;
  if (p >= A) return;
  const u32 s = seg_of[p];
  const u32 S = sg.start[s];
  const u32 w = fbits[p >> 5];
  const bool hi = (w >> (p & 31)) & 1u;
  const u32 hb = wpre[p >> 5] + __popc(w & ((1u << (p & 31)) - 1u)) - seg_hbase[s];
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
