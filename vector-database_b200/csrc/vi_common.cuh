// vi_common.cuh -- shared types and device helpers of libvi_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "vi_b200.h"

typedef uint32_t u32;
typedef unsigned long long u64;  // the type CUDA's 64-bit atomics are declared on
typedef int64_t i64;
typedef __int128 i128;
typedef unsigned __int128 u128;

// Ranges with at least t_big points are "big": their statistics are computed by many CTAs (fast mode) or by one
// thread per (range, dimension) chain (exact mode); smaller ranges are owned by one team/warp (vi_build.cu).
constexpr u32 VI_MIN_BIG = 512;  // smallest allowed "big range" threshold (sizes the big-range work space)
// Rows of a big range handled by one CTA of the fast-mode statistics kernel.
constexpr u32 VI_CHUNK = 4096;
constexpr int VI_NUM_SMS = 148;
constexpr int VI_MAX_DEPTH = 62;  // IndexBuilder.cs:99,104: splitting a depth-62 range overflows rangeId

#define VI_CUDA_TRY(expr)                                                                 \
  do                                                                                      \
  {                                                                                       \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) return ctx->fail_cuda(_e, #expr, __FILE__, __LINE__);          \
  } while (0)

// One tree level's open ranges ("segments" of the position space). All device arrays.
struct SegLevel
{
  u32* start = nullptr;   // first position
  u32* count = nullptr;   // points (>= 2: leaves never become segments)
  i64* rid = nullptr;     // RangeID
  u32* row = nullptr;     // row index in the range table
  int* dim = nullptr;     // split choice, written by the statistics pass
  float* mid = nullptr;
  i64* pivot = nullptr;
  u32* bslot = nullptr;   // fast mode: index of the range in the level's big list (its gacc slot), or 0xffffffff
};

struct LevelTotals  // pinned host, written by the device at the end of every level
{
  u32 rows;      // table rows created for the next level (non-empty children)
  u32 segs;      // next-level segments (children with >= 2 points)
  u32 pos;       // next-level active positions
  u32 nbig;      // next-level big segments
  u32 chunks;    // next-level fast-mode chunk count
  u32 err;       // range table capacity exceeded
  u32 minseg;    // smallest / largest next-level range
  u32 maxseg;
  u32 subs;      // children handed to the sub-tree kernel this level, and their points
  u32 subpos;
  u32 derived;   // next-level points in ranges whose sums are derived (parent - sibling), fast mode
};

struct vi_ctx
{
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;

  // points
  int32_t dims = 0;
  int32_t ld = 0;  // row stride in floats (dims rounded up to 4)
  int64_t capacity = 0;
  int64_t n = 0;
  float* rows = nullptr;
  i64* ids = nullptr;

  // build work space (allocated by vi_build for n points)
  int64_t ws_n = 0;
  u32* perm[2] = {nullptr, nullptr};
  i64* pid[2] = {nullptr, nullptr};
  u32* seg_of[2] = {nullptr, nullptr};
  SegLevel seg[2];
  u32* big_list[2] = {nullptr, nullptr};    // segment indexes of big segments
  u32* chunk_first = nullptr;               // exclusive scan of chunks per big slot (+ total)
  u32* fbits = nullptr;                     // hi flags, 1 bit per position
  u32* wpre = nullptr;                      // exclusive popcount prefix per flag word (+ total)
  u32* seg_nlo = nullptr;                   // per segment: low child size
  u32* seg_hbase = nullptr;                 // per segment: hi flags before its first position
  u32* c_rows = nullptr;                    // per segment child row count (scanned in place, + total)
  u64* c_actpos = nullptr;                  // per segment (active children << 32 | active positions) (+ total)
  u64* c_sub = nullptr;                     // per segment (sub-tree children << 32 | their points) (+ total)
  u32* sub_perm = nullptr;                  // sub-tree position space: row index / id per point
  i64* sub_pid = nullptr;
  u32* sub_start = nullptr;                 // sub-tree list
  u32* sub_count = nullptr;
  i64* sub_rid = nullptr;
  u32* sub_row = nullptr;
  u32* sub_depth = nullptr;
  u64* sub_stats = nullptr;                 // [64] points + [64] ranges per depth, then the kernel's 2 counters
  void* scan_tmp = nullptr;                 // block sums for scans
  u64* gacc_prev = nullptr;                 // fast mode: the previous level's gacc (sibling derivation)
  u32* bl_parent[2] = {nullptr, nullptr};   // per big-list slot: the parent's slot in gacc_prev if the range's sums are
  u32* bl_sib[2] = {nullptr, nullptr};      //   derived as parent - sibling (bl_sib = the sibling's slot), else 0xffffffff
  u64* gacc = nullptr;                      // fast mode: per big slot [dims][4] + [2] id sums
  float2* gstats = nullptr;                 // exact mode: per big slot [dims] (mean, q)
  u32* counters = nullptr;                  // device counters (nbig_next, ...)
  LevelTotals* totals = nullptr;            // pinned host
  float* d_absmax = nullptr;

  // table
  int64_t t_cap = 0;
  int64_t t_rows = 0;
  i64* t_rid = nullptr;
  int* t_dim = nullptr;
  float* t_mid = nullptr;
  i64* t_id = nullptr;
  int* t_low = nullptr;
  int* t_high = nullptr;
  int4* t_node = nullptr;  // packed traversal rows (dim, mid, low|id.lo, high|id.hi)
  int* t_src = nullptr;    // leaf rows: row index of the point in `rows` (for candidate verification)
  bool built = false;

  vi_build_info info{};
  std::vector<vi_level_info> levels;

  // search scratch (grown on demand)
  float* q_buf = nullptr;   int64_t q_cap = 0;      // staged queries (floats)
  i64* off_buf = nullptr;   int64_t off_cap = 0;    // offsets
  i64* ids_buf = nullptr;   int64_t ids_cap = 0;    // candidate ids
  i64* off2_buf = nullptr;  int64_t off2_cap = 0;   // verified offsets
  i64* ids2_buf = nullptr;  int64_t ids2_cap = 0;   // verified ids
  int* search_src = nullptr; int64_t src_cap = 0;   // candidate source rows (only while verifying)
  u32* verify_keep = nullptr; int64_t keep_cap = 0;

  // collective
  int rank = 0, world = 1;
  vi_allreduce_u64_fn allreduce = nullptr;
  vi_alltoallv_fn alltoallv = nullptr;
  void* coll_user = nullptr;
  int64_t shared_rows = 0;      // rows of the replicated top levels (multi-rank build), else 0
  bool replicated = false;      // vi_table_replicate done: every rank holds the whole table
  float* own_rows = nullptr;    // multi-rank build: rows / ids of the ranges this rank owns (replace rows/ids as the
  i64* own_ids = nullptr;       // data the table's t_src refers to)
  int64_t own_n = 0, own_cap = 0;
  float* send_rows = nullptr;   // all-to-all send buffers (kept across builds)
  i64* send_ids = nullptr;
  int64_t send_cap = 0;

  int fail(int code, const std::string& msg)
  {
    err = msg;
    return code;
  }
  int fail_cuda(cudaError_t e, const char* what, const char* file, int line)
  {
    err = std::string("CUDA error: ") + cudaGetErrorString(e) + " at " + what + " (" + file + ":" + std::to_string(line) + ")";
    cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? VI_ERR_OOM : VI_ERR_CUDA;
  }
};

// ---- device helpers -----------------------------------------------------------------------------------------

// float.CompareTo (Comparer<float>.Default inside MaxBy, IndexBuilder.cs:77-79): NaN lowest, -0 == +0.
__device__ __forceinline__ int cmp_float_dotnet(float a, float b)
{
  if (a < b) return -1;
  if (a > b) return 1;
  if (a == b) return 0;
  if (isnan(a)) return isnan(b) ? 0 : -1;
  return 1;
}

__device__ __forceinline__ float4 ldg_f4_stream(const float4* p)
{
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ float ldg_f_stream(const float* p)
{
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}

// one float of a row for the partition gather: ask L2 for the smallest fill it offers (a plain load of 4 bytes was
// measured to move a whole 128-byte line from HBM: 1.36 GB per level at 10M points, profiles/r1_ncu_final.md)
__device__ __forceinline__ float ldg_f_gather(const float* p)
{
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::64B.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}

// hi flags before position x: word prefix + bits below x in its word
__device__ __forceinline__ u32 hi_before(const u32* __restrict__ wpre, const u32* __restrict__ fbits, u32 x)
{
  u32 w = x >> 5, b = x & 31;
  return wpre[w] + __popc(fbits[w] & ((1u << b) - 1u));
}

// build entry points implemented in vi_build.cu / vi_search.cu
int vi_table_replicate_impl(vi_ctx* ctx);
int vi_debug_divcheck_impl(vi_ctx* ctx, uint64_t seed, int64_t samples, int64_t* mismatches);
int vi_build_impl(vi_ctx* ctx, int mode);
int vi_search_impl(vi_ctx* ctx, const float* d_queries, int64_t nq, float proximity, i64* d_offsets, i64* d_ids,
                   int64_t cap, int64_t* total, int64_t* visits, bool have_offsets);
int vi_verify_impl(vi_ctx* ctx, const float* d_queries, int64_t nq, float distance, const i64* d_offsets_in,
                   const i64* d_ids_in, int64_t total_in, i64* d_offsets_out, i64* d_ids_out, int64_t* total_out);
void vi_free_workspace(vi_ctx* ctx);
void vi_free_table(vi_ctx* ctx);
