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
constexpr u32 VI_MIN_BIG = 64;   // smallest allowed "big range" threshold (sizes the big-range work space)
// Rows of a big range handled by one CTA of the fast-mode statistics kernel.
constexpr u32 VI_CHUNK = 4096;
constexpr int VI_NUM_SMS = 148;
// packed traversal row: x = Dimension, or < 0 for a leaf, or VI_NODE_BOTH for an internal row with Dimension = null
constexpr int VI_NODE_BOTH = 0x7fffffff;
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

// State of one tree level, resident on the device: every kernel of a level reads its sizes from here and the last
// kernel of the partition pass writes the next level's record (and a copy into pinned host memory), so the level
// loop never has to wait for the host to learn a size before it can launch (vi_build.cu run_levels).
struct LevelDev
{
  u32 A;          // active positions (points in open ranges)
  u32 R;          // open ranges (segments, each >= 2 points)
  u32 nbig;       // of them, big ones (fast mode: chunked statistics; exact mode: one chain per (range, dim))
  u32 chunks;     // fast mode: chunk CTAs of the statistics kernel
  u32 minseg;     // smallest / largest open range
  u32 maxseg;
  u32 derived;    // fast mode: points in big ranges whose sums are derived as parent - sibling
  u32 row_next;   // first free table row
  u32 sub_cnt;    // sub-tree list so far: entries and points
  u32 sub_pos;
  u32 err;        // 1 = range table capacity exceeded
  u32 rows;       // table rows created by the partition pass that produced this level (accounting)
  u32 ticket[2];  // last-block tickets of the level's two scans (k_flags, k_children)
  u32 pad[2];
};
static_assert(sizeof(LevelDev) == 64, "LevelDev is copied as 4 x uint4");
constexpr int VI_LV_N = VI_MAX_DEPTH + 4;

// One tile aggregate / prefix of k_children's scan over the level's ranges (vi_partition.cuh)
struct ChildAgg
{
  u32 rows;     // table rows of the children (non-empty ones)
  u32 act_cnt;  // children that stay ranges of the next level
  u32 act_pos;  // their points
  u32 sub_cnt;  // children handed to the sub-tree kernel
  u32 sub_pos;  // their points
  u32 big_cnt;  // big children (next level's big list slots)
  u32 chunks;   // chunk CTAs of the big children that are summed
  u32 derived;  // points of the big children that are derived
};
static_assert(sizeof(ChildAgg) == 32, "ChildAgg is moved as 2 x uint4");

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
  u32* chunk_first[2] = {nullptr, nullptr}; // exclusive scan of chunks per big slot (+ total), per level parity
  u32* fbits = nullptr;                     // hi flags, 1 bit per position
  u32* wloc = nullptr;                      // hi flags before each flag word inside its 2048-position tile
  u32* ftile = nullptr;                     // per flag tile: hi count, then (last block) exclusive prefix (+ total)
  u32* seg_nlo = nullptr;                   // per segment: low child size
  u32* seg_hbase = nullptr;                 // per segment: hi flags before its first position
  ChildAgg* c_pre = nullptr;                // per segment: exclusive prefix of its children's counts inside its tile
  ChildAgg* ctile = nullptr;                // per 1024-segment tile: aggregate, then exclusive prefix
  uint2* ctile_mm = nullptr;                // per tile: (min, max) size of the children that stay ranges
  LevelDev* lv = nullptr;                   // device level records [VI_LV_N]
  LevelDev* h_lv = nullptr;                 // pinned host copies, written by the device
  u32* sub_perm = nullptr;                  // sub-tree position space: row index / id per point
  i64* sub_pid = nullptr;
  u32* sub_start = nullptr;                 // sub-tree list
  u32* sub_count = nullptr;
  i64* sub_rid = nullptr;
  u32* sub_row = nullptr;
  u32* sub_depth = nullptr;
  u64* sub_stats = nullptr;                 // [64] points + [64] ranges per depth, then the kernel's 2 counters
  void* scan_tmp = nullptr;                 // block sums for scans (multi-rank shared phase)
  u64* gacc_prev = nullptr;                 // fast mode: the previous level's gacc (sibling derivation)
  u32* bl_parent[2] = {nullptr, nullptr};   // per big-list slot: the parent's slot in gacc_prev if the range's sums are
  u32* bl_sib[2] = {nullptr, nullptr};      //   derived as parent - sibling (bl_sib = the sibling's slot), else 0xffffffff
  u64* gacc = nullptr;                      // fast mode: per big slot [dims][4] + [2] id sums
  float2* gstats = nullptr;                 // exact mode: per big slot [dims] (mean, q)
  u32* counters = nullptr;                  // device counters (nbig_next, ...)
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
  const float* src_rows = nullptr;  // the row store the table's t_src indexes (rows, or own_rows after a multi-rank
                                    // build); null when the table has no vectors behind it (imported / replicated)

  // vi_build_copy: host destinations of the range table; finished row blocks are copied out on copy_stream while the
  // build's last kernel is still running (null = a plain vi_build)
  struct CopyOut
  {
    i64* rid = nullptr;
    int* dim = nullptr;
    float* mid = nullptr;
    i64* id = nullptr;
    int64_t cap = 0;
    int64_t copied_lo = 0, copied_hi = 0;  // rows [copied_lo, copied_hi) are already on their way
    bool active = false;
  } out;
  cudaStream_t copy_stream = nullptr;

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
  int64_t pending_nq = -1, pending_total = 0;        // vi_search_begin ... vi_search_fetch
  float pending_prox = 0.f;
  // warp-per-query traversal (vi_search.cu): candidate pool of the single-pass form, per-query chunk heads, per-warp
  // stack spill area, control words
  i64* sw_pool = nullptr;    int64_t sw_pool_cap = 0;   // slots (chunks of SW_CHUNK: [link][63 ids])
  i64* sw_head = nullptr;    int64_t sw_head_cap = 0;
  u32* sw_spill = nullptr;   int64_t sw_spill_cap = 0;
  int search_path = 0;                                  // of the last count pass: 0 = thread per query, 1 = warp per query
  bool search_want_src = false;                         // the caller will fill with source rows (verify / top-k): no pool
  bool sw_pool_valid = false;                           // the pool holds the candidates of (sw_q, sw_off, sw_nq, sw_prox)
  bool sw_pool_overflow = false;
  const float* sw_q = nullptr; const i64* sw_off = nullptr; int64_t sw_nq = 0; float sw_prox = 0.f;

  // collective
  int rank = 0, world = 1;
  vi_allreduce_u64_fn allreduce = nullptr;
  vi_alltoallv_fn alltoallv = nullptr;
  void* coll_user = nullptr;
  void* nccl = nullptr;         // ncclComm_t when the library owns the communicator (vi_comm_init)
  int64_t coll_calls[3] = {0, 0, 0};  // all-reduce, all-to-all, all-gather: calls and bytes sent (vi_comm_stats)
  int64_t coll_bytes[3] = {0, 0, 0};
  void* sh_dev = nullptr;       // multi-rank shared phase: device records (vi_sharded.cuh) and their pinned host copy
  void* sh_host = nullptr;
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

// ---- hi-flag prefix: tile prefix + word prefix inside the tile + bits below x in its word ------------------------
constexpr int FL_ITEMS = 8;                    // positions per thread of k_flags
constexpr int FL_TILE = 256 * FL_ITEMS;        // positions per CTA
constexpr int FL_WORDS = FL_TILE / 32;         // flag words per tile
constexpr int FL_WORDS_LOG2 = 6;
static_assert((1 << FL_WORDS_LOG2) == FL_WORDS, "tile size");

struct FlagScan
{
  const u32* fbits;
  const u32* wloc;
  const u32* ftile;
};

// hi flags among positions [0, x); valid for x <= A (k_flags covers position A itself)
__device__ __forceinline__ u32 hi_before(const FlagScan& f, u32 x)
{
  const u32 w = x >> 5, b = x & 31;
  return f.ftile[w >> FL_WORDS_LOG2] + f.wloc[w] + __popc(f.fbits[w] & ((1u << b) - 1u));
}

// build entry points implemented in vi_build.cu / vi_search.cu
int vi_table_replicate_impl(vi_ctx* ctx);
// collectives (vi_comm.cu): NCCL on ctx->stream when the library owns a communicator, else the host callbacks
bool vi_coll_in_stream(const vi_ctx* ctx);
int vi_coll_allreduce_u64(vi_ctx* ctx, void* d_buf, int64_t count);
int vi_coll_alltoallv(vi_ctx* ctx, const void* d_send, const int64_t* send_bytes, void* d_recv, const int64_t* recv_bytes);
int vi_coll_allgather(vi_ctx* ctx, const void* d_send, void* d_recv, int64_t bytes);
int vi_coll_allgatherv_inplace(vi_ctx* ctx, void* d_buf, const int64_t* offset, const int64_t* bytes);
void vi_comm_release(vi_ctx* ctx);
int vi_debug_divcheck_impl(vi_ctx* ctx, uint64_t seed, int64_t samples, int64_t* mismatches);
int vi_build_impl(vi_ctx* ctx, int mode);
int vi_search_impl(vi_ctx* ctx, const float* d_queries, int64_t nq, float proximity, i64* d_offsets, i64* d_ids,
                   int64_t cap, int64_t* total, int64_t* visits, bool have_offsets);
int vi_verify_impl(vi_ctx* ctx, const float* d_queries, int64_t nq, float distance, const i64* d_offsets_in,
                   const i64* d_ids_in, int64_t total_in, i64* d_offsets_out, i64* d_ids_out, int64_t* total_out);
void vi_free_workspace(vi_ctx* ctx);
void vi_free_table(vi_ctx* ctx);
