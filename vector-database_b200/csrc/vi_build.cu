// vi_build.cu -- level-synchronous split-tree builder (replaces the sequential walker IndexBuilder.Build,
// VectorIndex/IndexBuilder.cs:23-157).  One pass per tree level over every open range at once:
//
//   statistics + split choice  IndexBuilder.cs:55-88   vi_stats_fast.cuh (fast mode) / vi_stats_exact.cuh (exact mode)
//   stable partition           IndexBuilder.cs:99-129  vi_partition.cuh
//
// Points never move: a range is a contiguous slice [start, start+count) of a position space; perm[] maps a
// position to its row, pid[] carries the point id along.  Leaves (count == 1) get their row written when their
// parent is partitioned and drop out of the position space (it is compacted every level).
#include <math.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>

#include "vi_common.cuh"
#include "vi_partition.cuh"
#include "vi_scan.cuh"
#include "vi_sharded.cuh"
#include "vi_stats_exact_px.cuh"
#include "vi_stats_fast.cuh"
#include "vi_subtree.cuh"

// =============================================================================================================
// level 0 set-up
// =============================================================================================================
__global__ void k_init_level0(u32* perm, i64* pid, const i64* ids, u32* seg_of, u32 n, SegLevel sg, u32* big_list,
                              i64* t_rid, int* t_low, int* t_high)
{
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
  {
    perm[i] = i;
    pid[i] = ids[i];
    seg_of[i] = 0;
  }
  if (i == 0)
  {
    sg.start[0] = 0;
    sg.count[0] = n;
    sg.rid[0] = 0;
    sg.row[0] = 0;
    big_list[0] = 0;
    t_rid[0] = 0;
    t_low[0] = -1;
    t_high[0] = -1;
  }
}

__global__ void k_single_point(const i64* ids, i64* t_rid, int* t_dim, float* t_mid, i64* t_id, int* t_low, int* t_high,
                               int* t_src)
{
  // IndexBuilder.cs:81-82: count == 1 -> Dimension -1, Mid default 0, Id = the point id
  t_rid[0] = 0;
  t_dim[0] = -1;
  t_mid[0] = 0.0f;
  t_id[0] = ids[0];
  t_low[0] = -1;
  t_high[0] = -1;
  t_src[0] = 0;
}

__global__ void k_absmax(const float4* __restrict__ rows, size_t n4, u32* out)
{
  float m = 0.0f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
  {
    float4 v = ldg_f4_stream(rows + i);
    float a = fabsf(v.x), b = fabsf(v.y), c = fabsf(v.z), d = fabsf(v.w);
    if (a > m) m = a;  // NaN never wins
    if (b > m) m = b;
    if (c > m) m = c;
    if (d > m) m = d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));  // non-negative floats order like their bits
}

// =============================================================================================================
// memory
// =============================================================================================================
template <typename T>
static cudaError_t dalloc(T** p, size_t count)
{
  return cudaMalloc((void**)p, count * sizeof(T) + 256);
}

void vi_free_workspace(vi_ctx* ctx)
{
  for (int i = 0; i < 2; ++i)
  {
    cudaFree(ctx->perm[i]); ctx->perm[i] = nullptr;
    cudaFree(ctx->pid[i]); ctx->pid[i] = nullptr;
    cudaFree(ctx->seg_of[i]); ctx->seg_of[i] = nullptr;
    cudaFree(ctx->big_list[i]); ctx->big_list[i] = nullptr;
    SegLevel& s = ctx->seg[i];
    cudaFree(s.start); cudaFree(s.count); cudaFree(s.rid); cudaFree(s.row); cudaFree(s.dim); cudaFree(s.mid);
    cudaFree(s.pivot); cudaFree(s.bslot);
    s = SegLevel();
    cudaFree(ctx->bl_parent[i]); ctx->bl_parent[i] = nullptr;
    cudaFree(ctx->bl_sib[i]); ctx->bl_sib[i] = nullptr;
    cudaFree(ctx->chunk_first[i]); ctx->chunk_first[i] = nullptr;
  }
  cudaFree(ctx->fbits); ctx->fbits = nullptr;
  cudaFree(ctx->wloc); ctx->wloc = nullptr;
  cudaFree(ctx->ftile); ctx->ftile = nullptr;
  cudaFree(ctx->seg_nlo); ctx->seg_nlo = nullptr;
  cudaFree(ctx->seg_hbase); ctx->seg_hbase = nullptr;
  cudaFree(ctx->c_pre); ctx->c_pre = nullptr;
  cudaFree(ctx->ctile); ctx->ctile = nullptr;
  cudaFree(ctx->ctile_mm); ctx->ctile_mm = nullptr;
  cudaFree(ctx->lv); ctx->lv = nullptr;
  cudaFree(ctx->scan_tmp); ctx->scan_tmp = nullptr;
  cudaFree(ctx->sub_perm); ctx->sub_perm = nullptr;
  cudaFree(ctx->sub_pid); ctx->sub_pid = nullptr;
  cudaFree(ctx->sub_start); ctx->sub_start = nullptr;
  cudaFree(ctx->sub_count); ctx->sub_count = nullptr;
  cudaFree(ctx->sub_rid); ctx->sub_rid = nullptr;
  cudaFree(ctx->sub_row); ctx->sub_row = nullptr;
  cudaFree(ctx->sub_depth); ctx->sub_depth = nullptr;
  cudaFree(ctx->sub_stats); ctx->sub_stats = nullptr;
  cudaFree(ctx->gacc); ctx->gacc = nullptr;
  cudaFree(ctx->gacc_prev); ctx->gacc_prev = nullptr;
  cudaFree(ctx->gstats); ctx->gstats = nullptr;
  cudaFree(ctx->d_absmax); ctx->d_absmax = nullptr;
  if (ctx->h_lv) cudaFreeHost(ctx->h_lv);
  ctx->h_lv = nullptr;
  ctx->ws_n = 0;
}

void vi_free_table(vi_ctx* ctx)
{
  cudaFree(ctx->t_rid); ctx->t_rid = nullptr;
  cudaFree(ctx->t_dim); ctx->t_dim = nullptr;
  cudaFree(ctx->t_mid); ctx->t_mid = nullptr;
  cudaFree(ctx->t_id); ctx->t_id = nullptr;
  cudaFree(ctx->t_low); ctx->t_low = nullptr;
  cudaFree(ctx->t_high); ctx->t_high = nullptr;
  cudaFree(ctx->t_node); ctx->t_node = nullptr;
  cudaFree(ctx->t_src); ctx->t_src = nullptr;
  ctx->t_cap = 0;
  ctx->t_rows = 0;
  ctx->built = false;
}

static int alloc_workspace(vi_ctx* ctx, int64_t n)
{
  if (ctx->ws_n >= n && ctx->perm[0]) return VI_OK;
  vi_free_workspace(ctx);
  const size_t N = (size_t)n;
  // lower bounds: the multi-rank build uses these arrays as scratch for its (small) host-built tables
  const size_t maxseg = std::max<size_t>(N / 2 + 2, 65536);
  const size_t maxbig = std::max<size_t>(N / VI_MIN_BIG + 2, 256);
  const size_t words = (N / FL_TILE + 2) * FL_WORDS;  // whole flag tiles, position N included
  for (int i = 0; i < 2; ++i)
  {
    VI_CUDA_TRY(dalloc(&ctx->perm[i], N));
    VI_CUDA_TRY(dalloc(&ctx->pid[i], N));
    VI_CUDA_TRY(dalloc(&ctx->seg_of[i], N));
    VI_CUDA_TRY(dalloc(&ctx->big_list[i], maxbig));
    SegLevel& s = ctx->seg[i];
    VI_CUDA_TRY(dalloc(&s.start, maxseg));
    VI_CUDA_TRY(dalloc(&s.count, maxseg));
    VI_CUDA_TRY(dalloc(&s.rid, maxseg));
    VI_CUDA_TRY(dalloc(&s.row, maxseg));
    VI_CUDA_TRY(dalloc(&s.dim, maxseg));
    VI_CUDA_TRY(dalloc(&s.mid, maxseg));
    VI_CUDA_TRY(dalloc(&s.pivot, maxseg));
    VI_CUDA_TRY(dalloc(&s.bslot, maxseg));
    VI_CUDA_TRY(dalloc(&ctx->bl_parent[i], maxbig));
    VI_CUDA_TRY(dalloc(&ctx->bl_sib[i], maxbig));
    VI_CUDA_TRY(dalloc(&ctx->chunk_first[i], maxbig + 1));
  }
  VI_CUDA_TRY(dalloc(&ctx->fbits, words));
  VI_CUDA_TRY(dalloc(&ctx->wloc, words));
  VI_CUDA_TRY(dalloc(&ctx->ftile, N / FL_TILE + 8));
  VI_CUDA_TRY(dalloc(&ctx->seg_nlo, maxseg));
  VI_CUDA_TRY(dalloc(&ctx->seg_hbase, maxseg));
  VI_CUDA_TRY(dalloc(&ctx->c_pre, maxseg));
  VI_CUDA_TRY(dalloc(&ctx->ctile, maxseg / CH_TILE + 8));
  VI_CUDA_TRY(dalloc(&ctx->ctile_mm, maxseg / CH_TILE + 8));
  VI_CUDA_TRY(dalloc(&ctx->lv, (size_t)VI_LV_N));
  VI_CUDA_TRY(dalloc(&ctx->sub_perm, N));
  VI_CUDA_TRY(dalloc(&ctx->sub_pid, N));
  VI_CUDA_TRY(dalloc(&ctx->sub_start, maxseg));
  VI_CUDA_TRY(dalloc(&ctx->sub_count, maxseg));
  VI_CUDA_TRY(dalloc(&ctx->sub_rid, maxseg));
  VI_CUDA_TRY(dalloc(&ctx->sub_row, maxseg));
  VI_CUDA_TRY(dalloc(&ctx->sub_depth, maxseg));
  VI_CUDA_TRY(dalloc(&ctx->sub_stats, (size_t)160));
  VI_CUDA_TRY(dalloc((u64**)&ctx->scan_tmp, maxseg / SCAN_TILE + words / SCAN_TILE + 64));
  VI_CUDA_TRY(dalloc(&ctx->gacc, maxbig * ((size_t)ctx->ld * 3 + 3)));
  VI_CUDA_TRY(dalloc(&ctx->gacc_prev, maxbig * ((size_t)ctx->ld * 3 + 3)));
  VI_CUDA_TRY(dalloc(&ctx->gstats, maxbig * (size_t)ctx->dims));
  VI_CUDA_TRY(dalloc(&ctx->d_absmax, (size_t)4));
  VI_CUDA_TRY(cudaMallocHost((void**)&ctx->h_lv, sizeof(LevelDev) * VI_LV_N));
  ctx->ws_n = n;
  return VI_OK;
}

static int alloc_table(vi_ctx* ctx, int64_t n)
{
  // 2n-1 rows when no child is ever empty; an empty child (all points on one side) adds a one-child row.
  const int64_t cap = 2 * n + n / 8 + 1024;
  if (ctx->t_cap >= cap && ctx->t_rid) return VI_OK;
  vi_free_table(ctx);
  if (cap >= (int64_t)0x7fffffff) return ctx->fail(VI_ERR_CAPACITY, "too many points for 32-bit row indexes");
  VI_CUDA_TRY(dalloc(&ctx->t_rid, (size_t)cap));
  VI_CUDA_TRY(dalloc(&ctx->t_dim, (size_t)cap));
  VI_CUDA_TRY(dalloc(&ctx->t_mid, (size_t)cap));
  VI_CUDA_TRY(dalloc(&ctx->t_id, (size_t)cap));
  VI_CUDA_TRY(dalloc(&ctx->t_low, (size_t)cap));
  VI_CUDA_TRY(dalloc(&ctx->t_high, (size_t)cap));
  VI_CUDA_TRY(dalloc(&ctx->t_node, (size_t)cap));
  VI_CUDA_TRY(dalloc(&ctx->t_src, (size_t)cap));
  ctx->t_cap = cap;
  return VI_OK;
}

// a table of `rows` rows for vi_ranges_load (vi_table.cu)
int vi_alloc_table_rows(vi_ctx* ctx, int64_t rows) { return alloc_table(ctx, rows / 2 + 1); }

// grows the table to hold `rows_needed` rows, keeping the first `keep` rows (multi-rank build: the shared top rows
// exist before a rank learns how many points it will own)
template <typename T>
static cudaError_t regrow(T** p, size_t newcount, size_t keep, cudaStream_t st)
{
  T* q = nullptr;
  cudaError_t e = cudaMalloc((void**)&q, newcount * sizeof(T) + 256);
  if (e != cudaSuccess) return e;
  if (keep) e = cudaMemcpyAsync(q, *p, keep * sizeof(T), cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(*p);
  *p = q;
  return e;
}

static int grow_table(vi_ctx* ctx, int64_t rows_needed, int64_t keep)
{
  if (ctx->t_cap >= rows_needed) return VI_OK;
  if (rows_needed >= (int64_t)0x7fffffff) return ctx->fail(VI_ERR_CAPACITY, "too many points for 32-bit row indexes");
  const size_t c = (size_t)rows_needed, k = (size_t)keep;
  cudaStream_t st = ctx->stream;
  VI_CUDA_TRY(regrow(&ctx->t_rid, c, k, st));
  VI_CUDA_TRY(regrow(&ctx->t_dim, c, k, st));
  VI_CUDA_TRY(regrow(&ctx->t_mid, c, k, st));
  VI_CUDA_TRY(regrow(&ctx->t_id, c, k, st));
  VI_CUDA_TRY(regrow(&ctx->t_low, c, k, st));
  VI_CUDA_TRY(regrow(&ctx->t_high, c, k, st));
  VI_CUDA_TRY(regrow(&ctx->t_node, c, k, st));
  VI_CUDA_TRY(regrow(&ctx->t_src, c, k, st));
  ctx->t_cap = rows_needed;
  return VI_OK;
}

// =============================================================================================================
// kernel dispatch on the row width
// =============================================================================================================
static u32 env_u32(const char* name, u32 dflt, u32 lo, u32 hi)
{
  const char* v = getenv(name);
  if (!v || !*v) return dflt;
  long x = strtol(v, nullptr, 10);
  if (x < (long)lo) x = lo;
  if (x > (long)hi) x = hi;
  return (u32)x;
}

struct FastShape
{
  int ts, ch;
  bool full;
};
static FastShape fast_shape(int ld, bool warp_rows)
{
  const int c4 = ld / 4;
  FastShape s;
  if (c4 <= 8) s = {8, 1, false};
  else if (c4 <= 16) s = {8, 2, false};
  else if (c4 <= 24) s = {8, 3, false};
  else if (c4 <= 32) s = {8, 4, false};
  else if (c4 <= 64) s = {32, 2, false};
  else if (c4 <= 96) s = {32, 3, false};
  else if (c4 <= 128) s = {32, 4, false};
  else s = {32, 6, false};  // wider rows take several column passes
  // experiment knob: rows up to 128 floats wide with one float4 per lane of a whole warp (lanes >= c4 idle)
  if (c4 <= 32 && warp_rows) s = {32, 1, false};
  s.full = (c4 == s.ts * s.ch);
  return s;
}

// CALL(TS, CH, FULL)
#define FAST_DISPATCH2(ts, ch, full, CALL)                         \
  do                                                               \
  {                                                                \
    if (full) { CALL(ts, ch, true); } else { CALL(ts, ch, false); } \
  } while (0)
#define FAST_DISPATCH(shp, CALL)                                                   \
  do                                                                               \
  {                                                                                \
    if (shp.ts == 32 && shp.ch == 1) FAST_DISPATCH2(32, 1, shp.full, CALL);        \
    else if (shp.ts == 8 && shp.ch == 1) FAST_DISPATCH2(8, 1, shp.full, CALL);     \
    else if (shp.ts == 8 && shp.ch == 2) FAST_DISPATCH2(8, 2, shp.full, CALL);     \
    else if (shp.ts == 8 && shp.ch == 3) FAST_DISPATCH2(8, 3, shp.full, CALL);     \
    else if (shp.ts == 8 && shp.ch == 4) FAST_DISPATCH2(8, 4, shp.full, CALL);     \
    else if (shp.ts == 32 && shp.ch == 2) FAST_DISPATCH2(32, 2, shp.full, CALL);   \
    else if (shp.ts == 32 && shp.ch == 3) FAST_DISPATCH2(32, 3, shp.full, CALL);   \
    else if (shp.ts == 32 && shp.ch == 4) FAST_DISPATCH2(32, 4, shp.full, CALL);   \
    else FAST_DISPATCH2(32, 6, shp.full, CALL);                                    \
  } while (0)

static int exact_chx(int dims)
{
  if (dims <= 32) return 1;
  if (dims <= 64) return 2;
  if (dims <= 96) return 3;
  if (dims <= 128) return 4;
  return 8;
}

int vi_debug_divcheck_impl(vi_ctx* ctx, uint64_t seed, int64_t samples, int64_t* mismatches)
{
  unsigned long long* d = (unsigned long long*)ctx->counters + 2;  // counters u32[4..5]
  VI_CUDA_TRY(cudaMemsetAsync(d, 0, 8, ctx->stream));
  const int blocks = VI_NUM_SMS * 8, threads = 256;
  const u64 per_thread = (u64)(samples / ((int64_t)blocks * threads)) + 1;
  k_divcheck<<<blocks, threads, 0, ctx->stream>>>(seed, per_thread, d);
  unsigned long long h = 0;
  VI_CUDA_TRY(cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, ctx->stream));
  VI_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  VI_CUDA_TRY(cudaGetLastError());
  *mismatches = (int64_t)h;
  return VI_OK;
}

// =============================================================================================================
// the build
// =============================================================================================================
struct BuildEnv
{
  int mode;     // statistics: VI_MODE_EXACT or VI_MODE_FAST (VI_MODE_SQL builds with the fast mode's kernels and sql = 1)
  int sql;      // dbo.BuildIndex's rules (DDL.sql:44-202): level_mx()
  u32 t_team, t_big, big_unroll;
  u32 t_slot;   // ranges of >= t_slot points get a big-list slot (their sums can be kept / derived); t_slot <= t_big
  u32 ring;     // warp-per-range kernel: rows through the bulk-async shared-memory ring (vi_stats_fast.cuh RING)
  u32 sibling;  // fast mode: sum only the smaller child of a big pair, derive the other from the parent (1 = on)
  u32 t_sub;    // ranges of 2..t_sub points are finished by the sub-tree kernel (0 = off)
  u32 sub_minb;
  u32 lag;      // levels the host may run ahead of the last level record it has read back (1 = none)
  FastShape shp;      // team / warp-per-range kernels
  FastShape shp_big;  // chunk kernel
  int chx;
  float qk;
  double qinv;
  size_t gstride;
  int64_t launches = 0;
  std::vector<cudaEvent_t> ev;
};

// Host-side copy of a level record (exact once the device's copy has been read back).
struct LevelState
{
  u32 A, R, nbig, chunks, minseg, maxseg;
  u32 derived;   // of A, the points in ranges whose sums are derived rather than summed (fast mode)
  u32 row_next;  // first free table row
  u32 sub_cnt, sub_pos;  // sub-tree list: entries and points so far
  int cur;       // ping-pong index of the level's buffers
  int level;     // depth of the ranges in seg[cur]
};

static cudaEvent_t env_event(vi_ctx* ctx, BuildEnv& env)
{
  cudaEvent_t e;
  cudaEventCreate(&e);
  cudaEventRecord(e, ctx->stream);
  env.ev.push_back(e);
  return e;
}

static void env_cleanup(BuildEnv& env)
{
  for (cudaEvent_t e : env.ev) cudaEventDestroy(e);
  env.ev.clear();
}

// the `mx` argument of a level's statistics kernels (vi_stats_common.cuh): IndexBuilder alternates max / min from the
// root (IndexBuilder.cs:33,128-129); dbo.BuildIndex takes the minimum at depth 1 only (DDL.sql:113,151,155) and its
// root sends ties high (DDL.sql:104)
static int level_mx(const BuildEnv& env, int level)
{
  if (!env.sql) return (level & 1) == 0 ? VI_MX_MAX : 0;
  return (level != 1 ? VI_MX_MAX : 0) | VI_MX_SQL | (level == 0 ? VI_MX_ROOT_HIGH : 0);
}

static void env_init(vi_ctx* ctx, BuildEnv& env, int mode)
{
  env.sql = mode == VI_MODE_SQL ? 1 : 0;
  if (env.sql) mode = VI_MODE_FAST;
  env.mode = mode;
  // range-size classes (see vi_stats_fast.cuh / vi_stats_exact.cuh); defaults from scripts/sweep.py on 10M x 96
  // (profiles/r1_sweep.txt); the variables exist for such sweeps
  env.t_team = env_u32("VI_B200_T_TEAM", 32, 2, VI_MAX_ROWS_PER_LANE);
  // (rows wider than 512 floats: a 128-row range is already 400 KB, enough for a CTA; profiles/r2_sweep_768.txt)
  const u32 t_big_fast = env_u32("VI_B200_T_BIG", ctx->ld > 512 ? 128 : 512, VI_MIN_BIG, VI_MAX_ROWS_PER_LANE);
  const u32 t_big_exact = env_u32("VI_B200_T_BIG_EXACT", 512, VI_MIN_BIG, 1u << 30);
  env.t_big = mode == VI_MODE_FAST ? t_big_fast : t_big_exact;
  env.big_unroll = env_u32("VI_B200_BIG_UNROLL", 8, 0, 8);  // 0 = cp.async ring
  env.sibling = env_u32("VI_B200_SIBLING", 1, 0, 1);
  env.ring = env_u32("VI_B200_RING", 0, 0, 1);
  // sibling derivation reaches below the chunked class: a warp-per-range range of >= t_slot points keeps its sums when
  // its children may pair up, and the larger child of such a pair is derived instead of summed
  env.t_slot = (mode == VI_MODE_FAST && env.sibling) ? std::min(env.t_big, env_u32("VI_B200_T_SLOT", 128, VI_MIN_BIG, 1u << 30))
                                                     : env.t_big;
  // sub-tree kernel (fast mode, vi_subtree.cuh): a range of up to t_sub points is finished by one warp in shared
  // memory; as many rows as fit 12.5 KB per warp (8 warps per CTA, 2 CTAs per SM), at most 32 (one point per lane)
  {
    u32 sub_rows = std::min<u32>((u32)SUB_TMAX, (u32)(12800 / (ctx->ld * 4 + 16)));
    if (sub_rows < 4 || mode != VI_MODE_FAST) sub_rows = 0;  // wide rows stay on the level path
    env.t_sub = std::min(env_u32("VI_B200_T_SUB", sub_rows, 0, (u32)SUB_TMAX), sub_rows);
  }
  env.sub_minb = 2;
  // The host enqueues level l+1 before it has read level l's record back (the kernels take their sizes from the
  // device-resident record, the grids from bounds derived from the last record the host knows).  The exact mode's
  // kernels are launched on exact sizes (its top levels are latency-bound chains anyway): no run-ahead there.
  env.lag = mode == VI_MODE_FAST ? env_u32("VI_B200_LAG", 2, 1, 4) : 1u;
  // Rows up to 128 floats wide: the chunk kernel reads a row with ONE float4 per lane of a whole warp (lanes beyond
  // the row idle): 16 accumulator registers per lane instead of 48, 63 registers, unroll 8 -> 0.72 ms per 10M x 96
  // level instead of 0.90 (profiles/r1_sweep.txt).  The small-range kernels keep 8-lane teams (bit 1 switches them).
  const u32 wr = env_u32("VI_B200_WARP_ROWS", 1, 0, 3);
  env.shp = fast_shape(ctx->ld, (wr & 2) != 0);
  env.shp_big = fast_shape(ctx->ld, (wr & 1) != 0);
  // rows wider than 128 floats: the warp-per-range kernel with 6 float4 per lane holds 48 64-bit accumulators (255
  // registers, 8 warps per SM: latency-bound); fewer columns per pass and several passes over the range's rows keep
  // more warps resident (the passes read disjoint columns: no byte is read twice)
  if (ctx->ld / 4 > 128)
  {
    const u32 wch = env_u32("VI_B200_WIDE_CH", 1, 1, 6);  // 1M x 768: 18.1 -> 15.6 ms per build (profiles/r2_sweep_768.txt)
    const int ch = wch >= 6 ? 6 : wch >= 4 ? 4 : wch >= 3 ? 3 : wch >= 2 ? 2 : 1;
    env.shp = {32, ch, (ctx->ld / 4) == 32 * ch};
  }
  env.chx = exact_chx(ctx->dims);
  env.qk = 1.0f;
  env.qinv = 1.0;
  env.gstride = (size_t)ctx->ld * 3 + 3;
}

// +-Inf has no fixed-point image (xi would saturate and the 64-bit partial sums of xi^2 overflow): the fast / SQL modes
// refuse it, as their CPU statement does; VI_MODE_EXACT takes any float
static const char* const kInfMessage =
    "fast / SQL mode: the data contains +-Inf, which has no fixed-point image; use VI_MODE_EXACT";

static void set_q_exponent(vi_ctx* ctx, BuildEnv& env, float amax)
{
  int qe = 0;
  if (amax > 0.f) (void)frexpf(amax, &qe);
  if (qe < -96) qe = -96;
  if (qe > 128) qe = 128;
  env.qk = ldexpf(1.0f, VI_QBITS - qe);
  env.qinv = ldexp(1.0, qe - VI_QBITS);
  ctx->info.q_exponent = qe;
}

// local max |x| over n rows of `rows` (NaN ignored)
static int local_absmax(vi_ctx* ctx, const float* rows, int64_t n, BuildEnv& env, float* out)
{
  cudaStream_t st = ctx->stream;
  VI_CUDA_TRY(cudaMemsetAsync(ctx->d_absmax, 0, 4, st));
  if (n > 0)
  {
    k_absmax<<<VI_NUM_SMS * 8, 256, 0, st>>>(reinterpret_cast<const float4*>(rows), (size_t)n * ctx->ld / 4,
                                             (u32*)ctx->d_absmax);
    ++env.launches;
  }
  VI_CUDA_TRY(cudaMemcpyAsync(out, ctx->d_absmax, 4, cudaMemcpyDeviceToHost, st));
  VI_CUDA_TRY(cudaStreamSynchronize(st));
  return VI_OK;
}

// launches the fast-mode chunk kernel over the ranges in big_list[cur]; `chunks_bound` >= the level's chunk count
// (the kernel reads the count from the level record `lvp` and surplus CTAs exit)
static void launch_big_fast(vi_ctx* ctx, BuildEnv& env, const LevelDev* lvp, const float* rows, int cur, u32 chunks_bound,
                            int mx, int allow_whole, u64* gacc, u32 keep_thr, const u32* bl_sib)
{
  cudaStream_t st = ctx->stream;
  SegLevel& sg = ctx->seg[cur];
  StatsOut sout{ctx->t_dim, ctx->t_mid, ctx->t_id};
  const int ld = ctx->ld, dims = ctx->dims;
  const u32 chunks = chunks_bound;
#define CALL_BIG_ARGS                                                                                              \
  lvp, sg, ctx->big_list[cur], ctx->chunk_first[cur], ctx->perm[cur], ctx->pid[cur], rows, ld, dims, env.qk, env.qinv, mx, \
      sout, gacc, allow_whole, keep_thr, bl_sib
#define CALL_BIG(TS, CH, FULL)                                                                                     \
  if (env.big_unroll == 0)                                                                                         \
  {                                                                                                                \
    const size_t ring = (size_t)BIG_NST * CH * 256 * sizeof(float4);                                               \
    cudaFuncSetAttribute(k_stats_big_fast<TS, CH, FULL, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring); \
    k_stats_big_fast<TS, CH, FULL, 0><<<chunks, 256, ring, st>>>(CALL_BIG_ARGS);                                   \
  }                                                                                                                \
  else if (env.big_unroll >= 8)                                                                                    \
    k_stats_big_fast<TS, CH, FULL, 8><<<chunks, 256, 0, st>>>(CALL_BIG_ARGS);                                      \
  else if (env.big_unroll >= 4)                                                                                    \
    k_stats_big_fast<TS, CH, FULL, 4><<<chunks, 256, 0, st>>>(CALL_BIG_ARGS);                                      \
  else                                                                                                             \
    k_stats_big_fast<TS, CH, FULL, 2><<<chunks, 256, 0, st>>>(CALL_BIG_ARGS)
  FAST_DISPATCH(env.shp_big, CALL_BIG);
#undef CALL_BIG
#undef CALL_BIG_ARGS
  ++env.launches;
}

// vi_build_copy: rows [r0, r1) of the table are final once the work enqueued on the build stream so far has run; send
// them to the caller's buffers on the copy stream.  Blocks are handed over in increasing row order.
static int copy_out_rows(vi_ctx* ctx, BuildEnv& env, int64_t r0, int64_t r1)
{
  vi_ctx::CopyOut& o = ctx->out;
  if (!o.active) return VI_OK;
  r1 = std::min(r1, o.cap);  // what does not fit is reported by vi_build_copy at the end
  if (r1 <= r0) return VI_OK;
  if (!ctx->copy_stream) VI_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  cudaEvent_t ev = env_event(ctx, env);
  VI_CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, ev, 0));
  const size_t n = (size_t)(r1 - r0);
  cudaStream_t cs = ctx->copy_stream;
  if (o.rid) VI_CUDA_TRY(cudaMemcpyAsync(o.rid + r0, ctx->t_rid + r0, n * 8, cudaMemcpyDeviceToHost, cs));
  if (o.dim) VI_CUDA_TRY(cudaMemcpyAsync(o.dim + r0, ctx->t_dim + r0, n * 4, cudaMemcpyDeviceToHost, cs));
  if (o.mid) VI_CUDA_TRY(cudaMemcpyAsync(o.mid + r0, ctx->t_mid + r0, n * 4, cudaMemcpyDeviceToHost, cs));
  if (o.id) VI_CUDA_TRY(cudaMemcpyAsync(o.id + r0, ctx->t_id + r0, n * 8, cudaMemcpyDeviceToHost, cs));
  if (o.copied_hi == o.copied_lo) o.copied_lo = r0;
  o.copied_hi = r1;
  return VI_OK;
}

// Finishes every range on the sub-tree list (vi_subtree.cuh); advances s.row_next past their rows.
static int run_subtrees(vi_ctx* ctx, BuildEnv& env, LevelState& s, const float* rows)
{
  if (s.sub_cnt == 0) return VI_OK;
  cudaStream_t st = ctx->stream;
  const int ld = ctx->ld, dims = ctx->dims;
  const int rows_max = (int)env.t_sub;
  const u32 row_base = s.row_next;
  const u32 overflow_base = row_base + 2u * s.sub_pos - 2u * s.sub_cnt;
  unsigned long long* lvlp = (unsigned long long*)ctx->sub_stats;  // [64] points, [64] ranges, then 2 u32 counters
  unsigned long long* lvlr = lvlp + 64;
  u32* cnt = (u32*)(lvlr + 64);
  VI_CUDA_TRY(cudaMemsetAsync(ctx->sub_stats, 0, 160 * 8, st));
  TableOut tout{ctx->t_rid, ctx->t_dim, ctx->t_mid, ctx->t_id, ctx->t_low, ctx->t_high};
  SubList sl{ctx->sub_start, ctx->sub_count, ctx->sub_rid, ctx->sub_row, ctx->sub_depth};
  const size_t smem = (size_t)SUB_WARPS * sub_smem_bytes_per_warp(rows_max, ld);
  // vi_build_copy: the list is taken in a few slices, and the dense row block of a finished slice (closed-form row
  // numbers: sub-tree k starts at row_base + 2 * sub_start[k] - 2 * k) goes to the host while the next slice runs
  const int nslice = (ctx->out.active && s.sub_cnt >= 8192u) ? (int)env_u32("VI_B200_COPY_SLICES", 8, 1, 8) : 1;
  u32 kb[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  u32 sb[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  kb[nslice] = s.sub_cnt;
  sb[nslice] = s.sub_pos;
  for (int i = 1; i < nslice; ++i)
  {
    kb[i] = (u32)((u64)s.sub_cnt * i / nslice);
    VI_CUDA_TRY(cudaMemcpyAsync(&sb[i], ctx->sub_start + kb[i], 4, cudaMemcpyDeviceToHost, st));
  }
  if (nslice > 1) VI_CUDA_TRY(cudaStreamSynchronize(st));
  cudaEvent_t e0 = env_event(ctx, env);
#define CALL_SUB(CH, FULL)                                                                                                \
  do                                                                                                                      \
  {                                                                                                                       \
    VI_CUDA_TRY(cudaFuncSetAttribute(k_subtree_fast<CH, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
    k_subtree_fast<CH, FULL><<<grid, SUB_WARPS * 32, smem, st>>>(sl, k1, ctx->sub_perm, ctx->sub_pid, rows, ld, dims,       \
                                                                 env.qk, env.qinv, tout, ctx->t_src, row_base,            \
                                                                 overflow_base, (u32)ctx->t_cap, cnt, lvlp, lvlr,         \
                                                                 rows_max, env.sql, k0);                                  \
  } while (0)
  const int ch = std::min(4, (ld / 4 + 7) / 8);  // int4 columns per team lane and pass
  const bool sub_full = ld == 32 * ch && dims == ld;
  for (int i = 0; i < nslice; ++i)
  {
    const u32 k0 = kb[i], k1 = nslice == 1 ? s.sub_cnt : kb[i + 1];
    if (k1 <= k0) continue;
    const u32 grid = std::min<u32>((k1 - k0 + SUB_WARPS - 1) / SUB_WARPS, (u32)VI_NUM_SMS * 2u);  // persistent warps, work cursor
    if (i > 0) VI_CUDA_TRY(cudaMemsetAsync(cnt + 3, 0, 4, st));  // the work cursor
    if (ch <= 1) { if (sub_full) CALL_SUB(1, true); else CALL_SUB(1, false); }
    else if (ch == 2) { if (sub_full) CALL_SUB(2, true); else CALL_SUB(2, false); }
    else if (ch == 3) { if (sub_full) CALL_SUB(3, true); else CALL_SUB(3, false); }
    else { if (sub_full) CALL_SUB(4, true); else CALL_SUB(4, false); }
    ++env.launches;
    if (nslice > 1)
    {
      const int64_t r0 = (int64_t)row_base + 2 * (int64_t)sb[i] - 2 * (int64_t)k0;
      const int64_t r1 = i + 1 == nslice ? (int64_t)overflow_base : (int64_t)row_base + 2 * (int64_t)sb[i + 1] - 2 * (int64_t)k1;
      const int rc = copy_out_rows(ctx, env, r0, r1);
      if (rc != VI_OK) return rc;
    }
  }
#undef CALL_SUB
  cudaEvent_t e1 = env_event(ctx, env);
  unsigned long long h[130];
  VI_CUDA_TRY(cudaMemcpyAsync(h, ctx->sub_stats, sizeof(h), cudaMemcpyDeviceToHost, st));
  VI_CUDA_TRY(cudaStreamSynchronize(st));
  VI_CUDA_TRY(cudaGetLastError());
  const u32 overflow = (u32)(h[128] & 0xffffffffu), err = (u32)(h[128] >> 32);
  if (getenv("VI_B200_TRACE"))
  {
    fprintf(stderr, "[vi_b200] sub-trees %u, float32 fallbacks %u, overflow rows %u\n", s.sub_cnt, (u32)(h[129] & 0xffffffffu),
            overflow);
  }
  if (err == 1) return ctx->fail(VI_ERR_CAPACITY, "range table capacity exceeded (degenerate input: too many one-child ranges)");
  if (err == 2)
    return ctx->fail(VI_ERR_OVERFLOW, "rangeId overflow: a range at depth 62 still holds more than one point "
                                      "(IndexBuilder.cs:99 checked(rangeId * 2 + 1))");
  if (err != 0) return ctx->fail(VI_ERR_STATE, "sub-tree kernel: internal error (node queue stalled)");
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  ctx->info.subtree_ms += ms;
  ctx->info.subtree_ranges += s.sub_cnt;
  // per-depth accounting: the sub-trees' ranges belong to levels like everyone else's
  for (int d = 0; d < 64; ++d)
  {
    if (!h[d] && !h[64 + d]) continue;
    vi_level_info* li = nullptr;
    for (auto& l : ctx->levels)
      if (l.level == d) li = &l;
    if (!li)
    {
      vi_level_info n{};
      n.level = d;
      ctx->levels.push_back(n);
      li = &ctx->levels.back();
    }
    li->points += (int64_t)h[d];
    li->ranges += (int64_t)h[64 + d];
    li->in_subtrees += (int64_t)h[d];
    ctx->info.point_visits += (int64_t)h[d];
  }
  s.row_next = overflow_base + overflow;
  s.sub_cnt = 0;
  s.sub_pos = 0;
  return VI_OK;
}

static LevelDev lv_from_state(const LevelState& s)
{
  LevelDev v{};
  v.A = s.A;
  v.R = s.R;
  v.nbig = s.nbig;
  v.chunks = s.chunks;
  v.minseg = s.minseg;
  v.maxseg = s.maxseg;
  v.derived = s.derived;
  v.row_next = s.row_next;
  v.sub_cnt = s.sub_cnt;
  v.sub_pos = s.sub_pos;
  return v;
}

// Puts the first level record of a level loop on the device (all later ones are written by k_children).
static int upload_level(vi_ctx* ctx, const LevelState& s)
{
  cudaStream_t st = ctx->stream;
  VI_CUDA_TRY(cudaMemsetAsync(ctx->lv, 0, sizeof(LevelDev) * VI_LV_N, st));  // tickets of every level start at 0
  ctx->h_lv[s.level] = lv_from_state(s);  // pinned staging: stays untouched until the copy has run (end of the build)
  VI_CUDA_TRY(cudaMemcpyAsync(ctx->lv + s.level, ctx->h_lv + s.level, sizeof(LevelDev), cudaMemcpyHostToDevice, st));
  return VI_OK;
}

// One level's launches.  `b` holds bounds on the level's sizes (exact when lag == 1); the kernels read the real sizes
// from lv[level].
struct LevelEvents
{
  cudaEvent_t e0, e1, e2;
};

static int enqueue_level(vi_ctx* ctx, BuildEnv& env, const LevelState& b, int level, int cur, bool exact_sizes,
                         bool first_level, const float* rows, u64* gacc_cur, u64* gacc_prev, LevelEvents& ev)
{
  cudaStream_t st = ctx->stream;
  const int ld = ctx->ld, dims = ctx->dims;
  const int mode = env.mode;
  const int nxt = cur ^ 1;
  const int mx = level_mx(env, level);
  SegLevel& sg = ctx->seg[cur];
  TableOut tout{ctx->t_rid, ctx->t_dim, ctx->t_mid, ctx->t_id, ctx->t_low, ctx->t_high};
  StatsOut sout{ctx->t_dim, ctx->t_mid, ctx->t_id};
  const u32 t_team = env.t_team, t_big = env.t_big;
  const FastShape shp = env.shp;
  LevelDev* lvp = ctx->lv + level;
  ev.e0 = env_event(ctx, env);

  // ---- statistics + split choice ---------------------------------------------------------------------------
  if (mode == VI_MODE_FAST)
  {
    const u32 t_slot = env.t_slot;
    const bool chunked = b.nbig && b.maxseg >= t_big;
    if (chunked)
    {
      // gacc is accumulated with atomics only by ranges of several chunks (or column passes); every other record is
      // written whole by whoever keeps it
      const int single_pass0 = (ld / 4) <= env.shp_big.ts * env.shp_big.ch;
      if (b.maxseg > VI_CHUNK || !single_pass0)
        VI_CUDA_TRY(cudaMemsetAsync(gacc_cur, 0, (size_t)b.nbig * env.gstride * sizeof(u64), st));
      launch_big_fast(ctx, env, lvp, rows, cur, b.chunks, mx, 1, gacc_cur, env.sibling ? 2 * t_slot : 0xffffffffu,
                      env.sibling ? ctx->bl_sib[cur] : nullptr);
    }
    // warp-per-range class (teams of a warp share one range); for TS == 32 it also covers the team class.
    // Below the first level of a loop every range has more than t_sub points (smaller children went to the sub-tree
    // list); the first level's ranges (the root, the roots of a rank's forest) can have any size.
    const u32 floor_n = first_level ? 2u : env.t_sub + 1u;
    const u32 wlo = shp.ts == 32 ? 2u : t_team;
    const bool may_warp = std::max(wlo, floor_n) < t_big && b.maxseg >= wlo && (!exact_sizes || b.minseg < t_big);
    if (may_warp)
    {
      u64* wg = (env.sibling && b.nbig) ? gacc_cur : nullptr;
      const u32 wgrid = (u32)(((u64)b.R * 32 + 255) / 256);
      if (env.ring && shp.ts == 8 && shp.full)
      {
        // rows through the bulk-async shared-memory ring (vi_stats_fast.cuh RING)
        const size_t rsm = 8 * ring_smem_bytes_per_warp(ld);
#define CALL_RING(CH)                                                                                                  \
  do                                                                                                                   \
  {                                                                                                                    \
    VI_CUDA_TRY(cudaFuncSetAttribute(k_stats_small_fast<8, CH, true, true, true>,                                      \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsm));                          \
    k_stats_small_fast<8, CH, true, true, true><<<wgrid, 256, rsm, st>>>(                                              \
        lvp, sg, wlo, t_big, ctx->perm[cur], ctx->pid[cur], rows, ld, dims, env.qk, env.qinv, mx, sout, wg,            \
        ctx->bl_parent[cur], ctx->bl_sib[cur], 2 * t_slot);                                                            \
  } while (0)
        if (shp.ch == 1) CALL_RING(1);
        else if (shp.ch == 2) CALL_RING(2);
        else if (shp.ch == 3) CALL_RING(3);
        else CALL_RING(4);
#undef CALL_RING
      }
      else
      {
#define CALL_WARP(TS, CH, FULL)                                                                              \
  k_stats_small_fast<TS, CH, FULL, true><<<wgrid, 256, 0, st>>>(                                              \
      lvp, sg, wlo, t_big, ctx->perm[cur], ctx->pid[cur], rows, ld, dims, env.qk, env.qinv, mx, sout, wg,      \
      ctx->bl_parent[cur], ctx->bl_sib[cur], 2 * t_slot)
        FAST_DISPATCH(shp, CALL_WARP);
#undef CALL_WARP
      }
      ++env.launches;
    }
    if (b.nbig)
    {
      // derived ranges (parent - sibling), ranges that span several chunks or column passes: split choice from gacc;
      // ranges that fit one chunk and one column pass were finished by their CTA, smaller ones by their warp
      const int single_pass = (ld / 4) <= env.shp_big.ts * env.shp_big.ch;
      if (!single_pass || b.maxseg > VI_CHUNK || env.sibling)
      {
        k_finalize_big_fast<<<(b.nbig * 32 + 255) / 256, 256, 0, st>>>(
            lvp, sg, ctx->big_list[cur], gacc_cur, gacc_prev, ld, dims, env.qinv, mx, sout, rows, ctx->perm[cur], single_pass,
            0, nullptr, env.sibling ? ctx->bl_parent[cur] : nullptr, ctx->bl_sib[cur], t_big);
        ++env.launches;
      }
    }
    const bool may_team = shp.ts < 32 && floor_n < t_team && (!exact_sizes || b.minseg < t_team);
    if (may_team)
    {
#define CALL_TEAM(TS, CH, FULL)                                                                              \
  k_stats_small_fast<TS, CH, FULL, false><<<(u32)(((u64)b.R * TS + 255) / 256), 256, 0, st>>>(                \
      lvp, sg, 2u, t_team, ctx->perm[cur], ctx->pid[cur], rows, ld, dims, env.qk, env.qinv, mx, sout, nullptr, nullptr, \
      nullptr, 0xffffffffu)
      FAST_DISPATCH(shp, CALL_TEAM);
#undef CALL_TEAM
      ++env.launches;
    }
  }
  else
  {
    // exact mode: `b` is exact (lag == 1)
    if (b.nbig)
    {
      const u32 nblk = (u32)((dims + 31) / 32);
      const bool vec_ok = ld % 4 == 0 && ((uintptr_t)rows & 15) == 0;
      // top levels (few chains in the whole GPU): the warp-specialised pipeline, vi_stats_exact_px.cuh
      const u32 px_max = env_u32("VI_B200_EX_PX", VI_NUM_SMS, 0, 1u << 20);
      if (vec_ok && b.nbig * nblk <= px_max)
      {
        // > 48 KB of dynamic shared memory: opt in (per device; a handful of launches per build)
        VI_CUDA_TRY(cudaFuncSetAttribute(k_stats_big_exact_px<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(PxShared<16>)));
        auto px_trace = [&](const char* what) -> int
        {
          if (!getenv("VI_B200_TRACE")) return VI_OK;
          unsigned long long h[8][4];
          VI_CUDA_TRY(cudaStreamSynchronize(st));
          VI_CUDA_TRY(cudaMemcpyFromSymbol(h, g_px_dbg, sizeof(h)));
          static const char* role[8] = {"chain", "verify0", "verify1", "variance", "-", "loader0", "loader1", "-"};
          for (int w = 0; w < 8; ++w)
            if (h[w][2])
              fprintf(stderr, "[vi_b200] px %s level %d %-8s waitA %5.1f%% waitB %5.1f%% other %5.1f%% of %.2f Mcycles\n", what,
                      level, role[w], 100.0 * h[w][0] / h[w][2], 100.0 * h[w][1] / h[w][2], 100.0 * h[w][3] / h[w][2],
                      h[w][2] * 1e-6);
          unsigned long long z[8][4] = {};
          VI_CUDA_TRY(cudaMemcpyToSymbol(g_px_dbg, z, sizeof(z)));
          return VI_OK;
        };
        k_stats_big_exact_px<16><<<b.nbig * nblk, PX_THREADS, sizeof(PxShared<16>), st>>>(
            sg, ctx->big_list[cur], nblk, ctx->perm[cur], rows, ld, dims, ctx->gstats);
        px_trace("full");
      }
      else
      {
        const bool vec = vec_ok && env_u32("VI_B200_EX_VEC", 1, 0, 1) != 0;
        const u32 ng = env_u32("VI_B200_EX_NG", EXNG_DEFAULT, 4, 12);
#define CALL_BIGEX(VEC, NG)                                                                                   \
  k_stats_big_exact<VEC, NG><<<b.nbig * nblk, 32, 0, st>>>(sg, ctx->big_list[cur], nblk, ctx->perm[cur], rows, ld, \
                                                           dims, ctx->gstats)
        if (vec) { if (ng <= 4) CALL_BIGEX(true, 4); else if (ng <= 6) CALL_BIGEX(true, 6); else CALL_BIGEX(true, 10); }
        else { if (ng <= 4) CALL_BIGEX(false, 4); else if (ng <= 6) CALL_BIGEX(false, 6); else CALL_BIGEX(false, 10); }
#undef CALL_BIGEX
      }
      {
        const u32 per = std::min(128u, std::max(1u, 2048u / b.nbig));
        VI_CUDA_TRY(cudaMemsetAsync(ctx->gacc, 0, (size_t)b.nbig * 16, st));
        k_idsum_big<<<b.nbig * per, 256, 0, st>>>(sg, ctx->big_list[cur], per, ctx->pid[cur], ctx->gacc);
      }
      k_finalize_big_exact<<<(b.nbig * 32 + 255) / 256, 256, 0, st>>>(sg, ctx->big_list[cur], b.nbig, ctx->gstats,
                                                                      ctx->gacc, dims, mx, sout);
      env.launches += 3;
    }
    if (b.minseg < t_big)
    {
      const u32 grid = (u32)(((u64)b.R * 32 + 255) / 256);
#define CALL_EX(CHX) \
  k_stats_small_exact<CHX><<<grid, 256, 0, st>>>(sg, b.R, t_big, ctx->perm[cur], ctx->pid[cur], rows, ld, dims, mx, sout)
      switch (env.chx)
      {
        case 1: CALL_EX(1); break;
        case 2: CALL_EX(2); break;
        case 3: CALL_EX(3); break;
        case 4: CALL_EX(4); break;
        default: CALL_EX(8); break;
      }
#undef CALL_EX
      ++env.launches;
    }
  }
  ev.e1 = env_event(ctx, env);

  // ---- stable partition: three launches ------------------------------------------------------------------------
  const FlagScan fs{ctx->fbits, ctx->wloc, ctx->ftile};
  const int sibling = (mode == VI_MODE_FAST && env.sibling) ? 1 : 0;
  k_flags<<<b.A / FL_TILE + 1, 256, 0, st>>>(lvp, &lvp->ticket[0], sg, ctx->seg_of[cur], ctx->perm[cur], ctx->pid[cur], rows, ld,
                                            ctx->fbits, ctx->wloc, ctx->ftile);
  k_children<<<(b.R + CH_TILE - 1) / CH_TILE, 256, 0, st>>>(lvp, &lvp->ticket[1], lvp + 1, ctx->h_lv + level + 1, sg, fs,
                                                           ctx->seg_nlo, ctx->seg_hbase, ctx->c_pre, ctx->ctile,
                                                           ctx->ctile_mm, env.t_sub, env.t_slot, t_big, sibling,
                                                           (u32)ctx->t_cap, ctx->chunk_first[nxt]);
  NextLevel nx{ctx->seg[nxt], ctx->perm[nxt], ctx->pid[nxt], ctx->seg_of[nxt], ctx->big_list[nxt], ctx->bl_parent[nxt],
               ctx->bl_sib[nxt], ctx->chunk_first[nxt]};
  SubList sl{ctx->sub_start, ctx->sub_count, ctx->sub_rid, ctx->sub_row, ctx->sub_depth};
  k_scatter<<<(b.A + 255) / 256, 256, 0, st>>>(lvp, lvp + 1, sg, ctx->seg_of[cur], ctx->perm[cur], ctx->pid[cur], fs,
                                               ctx->seg_nlo, ctx->seg_hbase, ctx->c_pre, ctx->ctile, env.t_sub, env.t_slot, t_big,
                                               sibling, (u32)level + 1u, nx, tout, ctx->t_src, sl, ctx->sub_perm, ctx->sub_pid);
  env.launches += 3;
  ev.e2 = env_event(ctx, env);
  return VI_OK;
}

// The level loop: processes seg[s.cur] (ranges of depth s.level) until no range with >= 2 points is left.
// `rows` is the row store perm[] indexes.  `s` is the exact state of the first level; on return it holds the
// final row / sub-tree cursors.
// stop_level >= 0: return before enqueuing that level, `s` = its exact state (the exact-mode multi-rank build filters the
// ranges there and calls again).
static int run_levels(vi_ctx* ctx, BuildEnv& env, LevelState& s, const float* rows, int stop_level = -1)
{
  cudaStream_t st = ctx->stream;
  const int mode = env.mode;
  const u32 t_big = env.t_big;
  int rc = upload_level(ctx, s);
  if (rc != VI_OK) return rc;
  // fast mode: the level's big-list slots and, per slot, whether its sums are derived (sibling derivation)
  u64 *gacc_cur = ctx->gacc, *gacc_prev = ctx->gacc_prev;
  if (mode == VI_MODE_FAST && s.R > 0)
  {
    VI_CUDA_TRY(cudaMemsetAsync(ctx->seg[s.cur].bslot, 0xff, (size_t)s.R * sizeof(u32), st));
    if (s.nbig)
      k_init_bslot<<<(s.nbig + 255) / 256, 256, 0, st>>>(ctx->seg[s.cur].bslot, s.R, ctx->big_list[s.cur], s.nbig,
                                                         ctx->bl_parent[s.cur], ctx->bl_sib[s.cur]);
  }
  const int first = s.level;
  std::vector<LevelEvents> evs;  // per enqueued level, index level - first
  LevelState known = s;          // exact state of level `known.level`
  int level = s.level;           // next level to enqueue
  int cur = s.cur;
  const int lag = (int)env.lag;

  // waits for the end of level known.level and reads the record of the next level the device wrote
  auto absorb = [&]() -> int
  {
    const LevelEvents& ev = evs[known.level - first];
    VI_CUDA_TRY(cudaEventSynchronize(ev.e2));
    const LevelDev d = ctx->h_lv[known.level + 1];
    if (d.err)
      return ctx->fail(VI_ERR_CAPACITY, "range table capacity exceeded (degenerate input: too many one-child ranges)");
    if (known.R > 0)
    {
      vi_level_info li{};
      li.level = known.level;
      li.ranges = known.R;
      li.points = known.A;
      li.rows_emitted = d.rows;
      li.derived_points = (int32_t)known.derived;
      float ms = 0;
      cudaEventElapsedTime(&ms, ev.e0, ev.e1);
      li.stats_ms = ms;
      cudaEventElapsedTime(&ms, ev.e1, ev.e2);
      li.partition_ms = ms;
      ctx->levels.push_back(li);
      ctx->info.point_visits += known.A;
    }
    known.A = d.A;
    known.R = d.R;
    known.nbig = d.nbig;
    known.chunks = d.chunks;
    known.minseg = d.minseg;
    known.maxseg = d.maxseg;
    known.derived = d.derived;
    known.row_next = d.row_next;
    known.sub_cnt = d.sub_cnt;
    known.sub_pos = d.sub_pos;
    ++known.level;
    return VI_OK;
  };

  for (;;)
  {
    while (known.level + (lag - 1) < level)
      if ((rc = absorb()) != VI_OK) return rc;
    if (known.R == 0) break;  // nothing open at known.level: every level already enqueued behind it is a no-op
    if (stop_level >= 0 && level >= stop_level) break;
    if (level >= VI_MAX_DEPTH)
    {
      // splitting a depth-62 range overflows rangeId: make sure such a range really exists before failing
      while (known.level < level)
        if ((rc = absorb()) != VI_OK) return rc;
      if (known.R == 0) break;
      return ctx->fail(VI_ERR_OVERFLOW, "rangeId overflow: a range at depth 62 still holds more than one point "
                                        "(IndexBuilder.cs:99 checked(rangeId * 2 + 1))");
    }
    // bounds on this level's sizes from the last record the host knows (`d` levels above)
    const int d = level - known.level;
    LevelState b = known;
    if (d > 0)
    {
      const u64 grow = 1ull << d;
      b.R = (u32)std::min<u64>((u64)known.R * grow, (u64)known.A / 2);
      b.nbig = (u32)std::min<u64>((u64)known.nbig * grow, (u64)known.A / env.t_slot);
      b.chunks = known.A / VI_CHUNK + std::min<u32>(b.nbig, known.A / t_big);
      b.minseg = 2;
    }
    LevelEvents ev;
    rc = enqueue_level(ctx, env, b, level, cur, d == 0, level == first, rows, gacc_cur, gacc_prev, ev);
    if (rc != VI_OK) return rc;
    evs.push_back(ev);
    std::swap(gacc_cur, gacc_prev);
    cur ^= 1;
    ++level;
  }
  while (known.level < level)
    if ((rc = absorb()) != VI_OK) return rc;
  if (stop_level >= 0 && known.level == stop_level && known.R > 0)
  {
    // stopped in front of an open level: hand its exact record back
    const int keep_cur = cur;
    s = known;
    s.cur = keep_cur;
    s.level = known.level;
    return VI_OK;
  }
  s.A = 0;
  s.R = 0;
  s.row_next = known.row_next;
  s.sub_cnt = known.sub_cnt;
  s.sub_pos = known.sub_pos;
  s.level = known.level;
  s.cur = cur;
  return run_subtrees(ctx, env, s, rows);
}

static int finish_table(vi_ctx* ctx, BuildEnv& env, u32 total_rows, cudaEvent_t ev_begin, const float* src_rows)
{
  ctx->src_rows = src_rows;  // the store the table's t_src indexes (candidate verification, top-k)
  cudaStream_t st = ctx->stream;
  if (total_rows > 0)
  {
    k_pack_nodes<<<(total_rows + 255) / 256, 256, 0, st>>>(ctx->t_dim, ctx->t_mid, ctx->t_id, ctx->t_low, ctx->t_high,
                                                           ctx->t_node, total_rows);
    ++env.launches;
  }
  cudaEvent_t ev_end = env_event(ctx, env);
  VI_CUDA_TRY(cudaStreamSynchronize(st));
  VI_CUDA_TRY(cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, ev_begin, ev_end);
  std::stable_sort(ctx->levels.begin(), ctx->levels.end(),
                   [](const vi_level_info& a, const vi_level_info& b) { return a.level < b.level; });
  ctx->t_rows = total_rows;
  ctx->built = true;
  ctx->info.ranges = total_rows;
  ctx->info.levels = (int32_t)ctx->levels.size();
  ctx->info.kernel_launches = env.launches;
  ctx->info.build_ms = ms;
  return VI_OK;
}

static int build_single(vi_ctx* ctx, BuildEnv& env)
{
  cudaStream_t st = ctx->stream;
  const u32 n = (u32)ctx->n;
  cudaEvent_t ev_begin = env_event(ctx, env);
  if (n == 1)
  {
    k_single_point<<<1, 1, 0, st>>>(ctx->ids, ctx->t_rid, ctx->t_dim, ctx->t_mid, ctx->t_id, ctx->t_low, ctx->t_high,
                                    ctx->t_src);
    ++env.launches;
    vi_level_info li{};
    li.rows_emitted = 1;
    ctx->levels.push_back(li);
    return finish_table(ctx, env, 1, ev_begin, ctx->rows);
  }
  int rc = alloc_workspace(ctx, ctx->n);
  if (rc != VI_OK) return rc;
  if (env.mode == VI_MODE_FAST)
  {
    float amax = 0.f;
    rc = local_absmax(ctx, ctx->rows, ctx->n, env, &amax);
    if (rc != VI_OK) return rc;
    if (isinf(amax)) return ctx->fail(VI_ERR_INVALID_ARG, kInfMessage);
    set_q_exponent(ctx, env, amax);
  }
  k_init_level0<<<(n + 255) / 256, 256, 0, st>>>(ctx->perm[0], ctx->pid[0], ctx->ids, ctx->seg_of[0], n, ctx->seg[0],
                                                 ctx->big_list[0], ctx->t_rid, ctx->t_low, ctx->t_high);
  ++env.launches;
  LevelState s{};
  s.A = n;
  s.R = 1;
  s.nbig = n >= env.t_big ? 1u : 0u;
  s.chunks = s.nbig ? (n + VI_CHUNK - 1) / VI_CHUNK : 0u;
  s.minseg = s.maxseg = n;
  s.row_next = 1;
  s.cur = 0;
  s.level = 0;
  if (s.nbig)
  {
    const u32 h[2] = {0u, s.chunks};
    VI_CUDA_TRY(cudaMemcpyAsync(ctx->chunk_first[0], h, sizeof(h), cudaMemcpyHostToDevice, st));
  }
  rc = run_levels(ctx, env, s, ctx->rows);
  if (rc != VI_OK) return rc;
  return finish_table(ctx, env, s.row_next, ev_begin, ctx->rows);
}

// =============================================================================================================
// multi-rank build (protocol: include/vi_b200.h vi_comm_init / vi_set_collective, DESIGN.md "Multi-GPU build")
// =============================================================================================================
constexpr size_t VI_SH_STAGE = 512 * 1024;  // small host-built tables of the ownership phase, one upload

struct ShDev  // device scratch of the shared phase
{
  ShLevel lvl[2];
  u64 cnt_loc[2 * VI_SH_MAXR];
  u64 cnt_glb[2 * VI_SH_MAXR];
  u32 sc[6][VI_SH_MAXR];
  u64 leaf_ids[VI_SH_MAXROWS];
  u64 v0[2 * 64];                    // sizes and max|x| of every rank
  u32 lc_all[64 * VI_SH_MAXR];       // all-gathered local sizes of the last shared level's ranges
  u32 stat_err;
  u32 pad[3];
  unsigned char tables[VI_SH_STAGE];
};

struct ShHost  // pinned
{
  ShLevel fin;
  u64 v0[2 * 64];
  u32 lc_all[64 * VI_SH_MAXR];
  unsigned char tables[VI_SH_STAGE];
};

__global__ void k_sh_pack0(const float* __restrict__ d_absmax, u32 nloc, int me, int G, u64* __restrict__ v0)
{
  const int i = threadIdx.x;
  if (i < 2 * G) v0[i] = 0;
  __syncthreads();
  if (i == 0)
  {
    v0[me] = nloc;
    v0[G + me] = (u64)__float_as_uint(*d_absmax);  // non-negative floats order like their bit patterns
  }
}

// a host blob of small tables that becomes one pinned -> device copy
struct Blob
{
  std::vector<unsigned char> bytes;
  template <typename T>
  size_t add(const std::vector<T>& v)
  {
    const size_t off = (bytes.size() + 15) & ~(size_t)15;
    bytes.resize(off + v.size() * sizeof(T) + 16);
    if (!v.empty()) memcpy(bytes.data() + off, v.data(), v.size() * sizeof(T));
    return off;
  }
};

constexpr int VI_RETRY = -1000;  // internal: rebuild with fewer shared levels

static int build_sharded_try(vi_ctx* ctx, BuildEnv& env, int Lcap, int* retry_level)
{
  cudaStream_t st = ctx->stream;
  const int G = ctx->world, me = ctx->rank;
  const int ld = ctx->ld, dims = ctx->dims;
  const u32 nloc = (u32)ctx->n;
  cudaEvent_t ev_begin = env_event(ctx, env);

  // 25 % slack: the points a rank owns after the exchange are rarely exactly its shard size
  int rc = alloc_workspace(ctx, std::max<int64_t>(ctx->n + ctx->n / 4, 4096));
  if (rc != VI_OK) return rc;
  if (!ctx->sh_dev)
  {
    VI_CUDA_TRY(cudaMalloc(&ctx->sh_dev, sizeof(ShDev)));
    VI_CUDA_TRY(cudaMallocHost(&ctx->sh_host, sizeof(ShHost)));
  }
  ShDev* sd = (ShDev*)ctx->sh_dev;
  ShHost* shh = (ShHost*)ctx->sh_host;

  // ---- global size and quantisation exponent: one small all-reduce, the phase's first host synchronisation -----------
  VI_CUDA_TRY(cudaMemsetAsync(ctx->d_absmax, 0, 4, st));
  if (nloc > 0)
  {
    k_absmax<<<VI_NUM_SMS * 8, 256, 0, st>>>(reinterpret_cast<const float4*>(ctx->rows), (size_t)nloc * ld / 4,
                                             (u32*)ctx->d_absmax);
    ++env.launches;
  }
  k_sh_pack0<<<1, 128, 0, st>>>(ctx->d_absmax, nloc, me, G, sd->v0);
  ++env.launches;
  if ((rc = vi_coll_allreduce_u64(ctx, sd->v0, 2 * G))) return rc;
  VI_CUDA_TRY(cudaMemcpyAsync(shh->v0, sd->v0, (size_t)2 * G * 8, cudaMemcpyDeviceToHost, st));
  VI_CUDA_TRY(cudaStreamSynchronize(st));
  u64 nglobal = 0;
  u32 amax_bits = 0;
  for (int g = 0; g < G; ++g)
  {
    nglobal += shh->v0[g];
    amax_bits = std::max(amax_bits, (u32)shh->v0[G + g]);
  }
  float amax = 0.f;
  memcpy(&amax, &amax_bits, 4);
  if (isinf(amax)) return ctx->fail(VI_ERR_INVALID_ARG, kInfMessage);  // the same on every rank (all-reduced maximum)
  set_q_exponent(ctx, env, amax);
  if (nglobal >= 0x7fffffffull) return ctx->fail(VI_ERR_CAPACITY, "more than 2^31-2 points");
  if (nglobal == 0) return finish_table(ctx, env, 0, ev_begin, nullptr);

  // ---- phase A: shared levels, enqueued without a host synchronisation ------------------------------------------------
  int L = 1;
  while ((1 << (L - 1)) < G) ++L;  // L = ceil(log2 G) + 1: about 2G ranges to balance over G owners
  L = std::min(L, Lcap);
  if (nloc > 0)
  {
    k_init_level0<<<(nloc + 255) / 256, 256, 0, st>>>(ctx->perm[0], ctx->pid[0], ctx->ids, ctx->seg_of[0], nloc,
                                                      ctx->seg[0], ctx->big_list[0], ctx->t_rid, ctx->t_low, ctx->t_high);
    ++env.launches;
  }
  VI_CUDA_TRY(cudaMemsetAsync(ctx->lv, 0, sizeof(LevelDev) * VI_LV_N, st));
  VI_CUDA_TRY(cudaMemsetAsync(sd->leaf_ids, 0, sizeof(sd->leaf_ids), st));
  VI_CUDA_TRY(cudaMemsetAsync(&sd->stat_err, 0, 16, st));
  TableOut tout{ctx->t_rid, ctx->t_dim, ctx->t_mid, ctx->t_id, ctx->t_low, ctx->t_high};
  StatsOut sout{ctx->t_dim, ctx->t_mid, ctx->t_id};
  k_sh_begin<<<1, 1, 0, st>>>(&sd->lvl[0], ctx->lv, nloc, nglobal, ctx->chunk_first[0], ctx->seg[0], ctx->big_list[0], tout);
  ++env.launches;
  if (nglobal == 1 && nloc == 1) VI_CUDA_TRY(cudaMemcpyAsync(sd->leaf_ids, ctx->ids, 8, cudaMemcpyDeviceToDevice, st));
  const FlagScan fs{ctx->fbits, ctx->wloc, ctx->ftile};
  ShScatter sc{sd->sc[0], sd->sc[1], sd->sc[2], sd->sc[3], (int*)sd->sc[4], (int*)sd->sc[5]};
  std::vector<LevelEvents> evs;
  int level = 0;
  for (; level < L; ++level)
  {
    if (level >= VI_MAX_DEPTH) return ctx->fail(VI_ERR_OVERFLOW, "rangeId overflow (IndexBuilder.cs:99)");
    const int cur = level & 1, nxt = cur ^ 1;
    const int mx = level_mx(env, level);
    const u32 Rb = (u32)std::min(1 << level, VI_SH_MAXR);  // bound on the level's ranges
    SegLevel& sg = ctx->seg[cur];
    LevelDev* lvp = ctx->lv + level;
    LevelEvents ev;
    ev.e0 = env_event(ctx, env);
    // local sums of every range -> gacc, one all-reduce, identical split on every rank
    VI_CUDA_TRY(cudaMemsetAsync(ctx->gacc, 0, (size_t)Rb * env.gstride * sizeof(u64), st));
    launch_big_fast(ctx, env, lvp, ctx->rows, cur, nloc / VI_CHUNK + Rb + 1, mx, 0, ctx->gacc, 0xffffffffu, nullptr);
    if ((rc = vi_coll_allreduce_u64(ctx, ctx->gacc, (int64_t)((size_t)Rb * env.gstride)))) return rc;
    k_finalize_big_fast<<<(Rb * 32 + 255) / 256, 256, 0, st>>>(lvp, sg, ctx->big_list[cur], ctx->gacc, nullptr, ld, dims,
                                                               env.qinv, mx, sout, ctx->rows, ctx->perm[cur], 0, 1,
                                                               &sd->stat_err, nullptr, nullptr, 0u);
    ev.e1 = env_event(ctx, env);
    // local partition flags and child sizes; global child sizes; children (identical on every rank); local scatter
    k_flags<<<nloc / FL_TILE + 1, 256, 0, st>>>(lvp, &lvp->ticket[0], sg, ctx->seg_of[cur], ctx->perm[cur], ctx->pid[cur],
                                                ctx->rows, ld, ctx->fbits, ctx->wloc, ctx->ftile);
    k_seg_nlo<<<(Rb + 255) / 256, 256, 0, st>>>(lvp, sg, fs, ctx->seg_nlo, ctx->seg_hbase);
    k_sh_counts<<<(Rb + 255) / 256, 256, 0, st>>>(&sd->lvl[cur], ctx->seg_nlo, Rb, sd->cnt_loc, sd->cnt_glb);
    if ((rc = vi_coll_allreduce_u64(ctx, sd->cnt_glb, 2 * (int64_t)Rb))) return rc;
    k_sh_next<<<1, 1, 0, st>>>(&sd->lvl[cur], &sd->lvl[nxt], sd->cnt_loc, sd->cnt_glb, (u32)level, &sd->stat_err, sc,
                               ctx->seg[nxt], ctx->big_list[nxt], ctx->chunk_first[nxt], lvp + 1, tout);
    k_scatter_shared<<<nloc / 256 + 1, 256, 0, st>>>(lvp, sg, ctx->seg_of[cur], ctx->perm[cur], ctx->pid[cur], fs,
                                                         ctx->seg_hbase, sc, ctx->perm[nxt], ctx->pid[nxt], ctx->seg_of[nxt],
                                                         sd->leaf_ids, ctx->t_src);
    env.launches += 6;
    ev.e2 = env_event(ctx, env);
    evs.push_back(ev);
  }
  const int cur = level & 1;  // buffers of the ranges of level L
  ShLevel* fin_d = &sd->lvl[cur];
  if ((rc = vi_coll_allreduce_u64(ctx, sd->leaf_ids, VI_SH_MAXROWS))) return rc;
  k_sh_leaf_ids<<<(VI_SH_MAXROWS + 255) / 256, 256, 0, st>>>(fin_d, sd->leaf_ids, ctx->t_dim, ctx->t_id);
  ++env.launches;
  if ((rc = vi_coll_allgather(ctx, fin_d->lcount, sd->lc_all, (int64_t)VI_SH_MAXR * 4))) return rc;
  VI_CUDA_TRY(cudaMemcpyAsync(&shh->fin, fin_d, sizeof(ShLevel), cudaMemcpyDeviceToHost, st));
  VI_CUDA_TRY(cudaMemcpyAsync(shh->lc_all, sd->lc_all, (size_t)G * VI_SH_MAXR * 4, cudaMemcpyDeviceToHost, st));
  VI_CUDA_TRY(cudaMemcpyAsync(ctx->h_lv, ctx->lv, sizeof(LevelDev) * (size_t)(level + 1), cudaMemcpyDeviceToHost, st));
  VI_CUDA_TRY(cudaStreamSynchronize(st));  // the phase's second host synchronisation
  VI_CUDA_TRY(cudaGetLastError());
  const ShLevel& fin = shh->fin;
  if (fin.err_level != VI_NONE)
  {
    // a range of the shared levels is too tightly clustered for the integer statistics: its float32 fallback needs all
    // its rows on one rank in global order, so the ranges are handed to their owners one level earlier
    *retry_level = (int)fin.err_level;
    return VI_RETRY;
  }
  for (int l = 0; l < level; ++l)
  {
    if (ctx->h_lv[l].R == 0) continue;
    vi_level_info li{};
    li.level = l;
    li.ranges = ctx->h_lv[l].R;
    li.points = ctx->h_lv[l].A;  // local points visited
    li.rows_emitted = ctx->h_lv[l + 1].rows;
    float ms = 0;
    cudaEventElapsedTime(&ms, evs[l].e0, evs[l].e1);
    li.stats_ms = ms;
    cudaEventElapsedTime(&ms, evs[l].e1, evs[l].e2);
    li.partition_ms = ms;
    ctx->levels.push_back(li);
    ctx->info.point_visits += li.points;
  }
  const u32 T = fin.T;
  ctx->shared_rows = T;

  // ---- phase B: ownership exchange --------------------------------------------------------------------------------------
  const u32 RL = fin.R;
  ctx->own_n = 0;
  LevelState s{};
  s.row_next = T;
  s.level = level;
  s.cur = 0;
  if (RL > 0)
  {
    // who holds how much of each range
    auto held = [&](int g, u32 i) -> u64 { return shh->lc_all[(size_t)g * VI_SH_MAXR + i]; };
    // owners: largest range first to the least loaded rank (ties: lower range index, lower rank)
    std::vector<u32> order(RL), owner(RL);
    for (u32 i = 0; i < RL; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](u32 a, u32 b) { return fin.gcount[a] > fin.gcount[b]; });
    std::vector<u64> load(G, 0);
    for (u32 i : order)
    {
      int best = 0;
      for (int g = 1; g < G; ++g)
        if (load[g] < load[best]) best = g;
      owner[i] = (u32)best;
      load[best] += fin.gcount[i];
    }
    // send layout: by destination rank, then range index, then local position
    std::vector<u32> send_base(RL, 0), placeholders;
    std::vector<int64_t> send_rows_n(G, 0), recv_rows_n(G, 0);
    u32 off = 0, A = 0;
    for (int d = 0; d < G; ++d)
      for (u32 i = 0; i < RL; ++i)
        if (owner[i] == (u32)d)
        {
          send_base[i] = off;
          off += fin.lcount[i];
          send_rows_n[d] += fin.lcount[i];
        }
    u64 nown = 0;
    for (u32 i = 0; i < RL; ++i)
    {
      A += fin.lcount[i];
      if (owner[i] != (u32)me) placeholders.push_back(fin.row[i]);  // root rows of ranges owned elsewhere: Dimension -2
    }
    for (int g = 0; g < G; ++g)
      for (u32 i = 0; i < RL; ++i)
        if (owner[i] == (u32)me)
        {
          recv_rows_n[g] += (int64_t)held(g, i);
          nown += held(g, i);
        }
    // forest of owned ranges, pieces in source-rank order (= global stable order)
    std::vector<u32> piece_dst, piece_src, piece_seg, f_start, f_count, f_row, f_big, f_cf;
    std::vector<i64> f_rid;
    std::vector<u32> src_off(G, 0);  // running offset inside each source's block of the receive buffer
    {
      u32 acc = 0;
      for (int g = 0; g < G; ++g) { src_off[g] = acc; acc += (u32)recv_rows_n[g]; }
    }
    u32 pos = 0, minseg = 0xffffffffu, maxseg = 0, chunks = 0;
    for (u32 i = 0; i < RL; ++i)
    {
      if (owner[i] != (u32)me) continue;
      const u32 fsi = (u32)f_start.size();
      f_start.push_back(pos);
      f_count.push_back((u32)fin.gcount[i]);
      f_rid.push_back(fin.rid[i]);
      f_row.push_back(fin.row[i]);
      for (int g = 0; g < G; ++g)
      {
        const u32 len = (u32)held(g, i);
        if (len == 0) continue;
        piece_dst.push_back(pos);
        piece_src.push_back(src_off[g]);
        piece_seg.push_back(fsi);
        pos += len;
        src_off[g] += len;
      }
      const u32 c = (u32)fin.gcount[i];
      minseg = std::min(minseg, c);
      maxseg = std::max(maxseg, c);
      if (c >= env.t_big)
      {
        f_big.push_back(fsi);
        f_cf.push_back(chunks);
        chunks += (c + VI_CHUNK - 1) / VI_CHUNK;
      }
    }
    f_cf.push_back(chunks);
    // every small table in one upload
    Blob blob;
    const size_t o_rid = blob.add(f_rid), o_sb = blob.add(send_base), o_ph = blob.add(placeholders), o_pd = blob.add(piece_dst),
                 o_ps = blob.add(piece_src), o_pg = blob.add(piece_seg), o_fs = blob.add(f_start), o_fc = blob.add(f_count),
                 o_fr = blob.add(f_row), o_fb = blob.add(f_big), o_cf = blob.add(f_cf);
    if (blob.bytes.size() > VI_SH_STAGE) return ctx->fail(VI_ERR_CAPACITY, "ownership tables too large");
    memcpy(shh->tables, blob.bytes.data(), blob.bytes.size());
    VI_CUDA_TRY(cudaMemcpyAsync(sd->tables, shh->tables, blob.bytes.size(), cudaMemcpyHostToDevice, st));
    auto dptr = [&](size_t o) { return sd->tables + o; };
    if (!placeholders.empty())
    {
      k_sh_placeholders<<<((u32)placeholders.size() + 255) / 256, 256, 0, st>>>((const u32*)dptr(o_ph), (u32)placeholders.size(),
                                                                                ctx->t_dim);
      ++env.launches;
    }
    // exchange buffers are kept across builds (cudaMalloc/cudaFree of GB-sized buffers costs milliseconds)
    if ((int64_t)nloc > ctx->send_cap)
    {
      cudaFree(ctx->send_rows); cudaFree(ctx->send_ids);
      ctx->send_rows = nullptr; ctx->send_ids = nullptr; ctx->send_cap = 0;
      const size_t c = (size_t)nloc + 1024;
      VI_CUDA_TRY(cudaMalloc((void**)&ctx->send_rows, c * ld * sizeof(float)));
      VI_CUDA_TRY(cudaMalloc((void**)&ctx->send_ids, c * sizeof(i64)));
      ctx->send_cap = (int64_t)c;
    }
    if ((int64_t)nown > ctx->own_cap)
    {
      cudaFree(ctx->own_rows); cudaFree(ctx->own_ids);
      ctx->own_rows = nullptr; ctx->own_ids = nullptr; ctx->own_cap = 0;
      const size_t c = (size_t)nown + (size_t)nown / 8 + 1024;
      VI_CUDA_TRY(cudaMalloc((void**)&ctx->own_rows, c * ld * sizeof(float)));
      VI_CUDA_TRY(cudaMalloc((void**)&ctx->own_ids, c * sizeof(i64)));
      ctx->own_cap = (int64_t)c;
    }
    if (A > 0)
    {
      // seg[cur] (written by the last k_sh_next) describes the local slices of the level-L ranges
      const size_t threads = (size_t)A * (ld / 4);
      k_pack_rows<<<(u32)((threads + 255) / 256), 256, 0, st>>>(ctx->seg[cur], ctx->seg_of[cur], ctx->perm[cur],
                                                                ctx->pid[cur], ctx->rows, ld, A, (const u32*)dptr(o_sb),
                                                                ctx->send_rows, ctx->send_ids);
      ++env.launches;
    }
    std::vector<int64_t> sb(G), rb(G);
    for (int g = 0; g < G; ++g) { sb[g] = send_rows_n[g] * ld * 4; rb[g] = recv_rows_n[g] * ld * 4; }
    if ((rc = vi_coll_alltoallv(ctx, ctx->send_rows, sb.data(), ctx->own_rows, rb.data()))) return rc;
    for (int g = 0; g < G; ++g) { sb[g] = send_rows_n[g] * 8; rb[g] = recv_rows_n[g] * 8; }
    if ((rc = vi_coll_alltoallv(ctx, ctx->send_ids, sb.data(), ctx->own_ids, rb.data()))) return rc;
    ctx->own_n = (int64_t)nown;

    // ---- phase C: the owned forest, ordinary level loop ------------------------------------------------------------------
    rc = alloc_workspace(ctx, std::max<int64_t>((int64_t)nown, 1024));
    if (rc != VI_OK) return rc;
    rc = grow_table(ctx, (int64_t)(2 * nown + nown / 8 + 1024 + T), T);
    if (rc != VI_OK) return rc;
    s.A = pos;
    s.R = (u32)f_start.size();
    s.nbig = (u32)f_big.size();
    s.chunks = chunks;
    s.minseg = minseg;
    s.maxseg = maxseg;
    if (s.R > 0)
    {
      SegLevel& f = ctx->seg[0];
      const size_t r4 = (size_t)s.R * 4;
      VI_CUDA_TRY(cudaMemcpyAsync(f.start, dptr(o_fs), r4, cudaMemcpyDeviceToDevice, st));
      VI_CUDA_TRY(cudaMemcpyAsync(f.count, dptr(o_fc), r4, cudaMemcpyDeviceToDevice, st));
      VI_CUDA_TRY(cudaMemcpyAsync(f.row, dptr(o_fr), r4, cudaMemcpyDeviceToDevice, st));
      VI_CUDA_TRY(cudaMemcpyAsync(f.rid, dptr(o_rid), (size_t)s.R * 8, cudaMemcpyDeviceToDevice, st));
      if (s.nbig) VI_CUDA_TRY(cudaMemcpyAsync(ctx->big_list[0], dptr(o_fb), (size_t)s.nbig * 4, cudaMemcpyDeviceToDevice, st));
      VI_CUDA_TRY(cudaMemcpyAsync(ctx->chunk_first[0], dptr(o_cf), ((size_t)s.nbig + 1) * 4, cudaMemcpyDeviceToDevice, st));
      const u32 np = (u32)piece_dst.size();
      k_forest_init<<<(s.A + 255) / 256, 256, 0, st>>>(s.A, np, (const u32*)dptr(o_pd), (const u32*)dptr(o_ps),
                                                       (const u32*)dptr(o_pg), ctx->own_ids, ctx->perm[0], ctx->pid[0],
                                                       ctx->seg_of[0]);
      ++env.launches;
      rc = run_levels(ctx, env, s, ctx->own_rows);
      if (rc != VI_OK) return rc;
    }
  }
  return finish_table(ctx, env, s.row_next, ev_begin, ctx->own_n > 0 ? ctx->own_rows : ctx->rows);
}

// ---- multi-rank build, exact mode --------------------------------------------------------------
// The literal recurrence is one sequential chain per (range, dimension) over the GLOBAL stable order: the top levels
// cannot be split over ranks.  So every rank first receives all points (one all-gather-v in rank order = global order),
// builds the levels 0 .. L-1, L = ceil(log2 G), redundantly -- the same deterministic kernels on the same data give the same
// bits everywhere, no communication -- and then keeps only the level-L ranges it owns (largest first to the least loaded
// rank, as the fast mode) and finishes their sub-trees alone.  What the ranks end up with has the fast mode's shape:
// rows [0, shared_rows) numbered while the ranges were common (a level-L range's row is filled in by its owner, a
// Dimension = -2 placeholder elsewhere), local rows behind them; vi_table_replicate works unchanged.  Gain: only the
// levels below L divide by G (levels 0-2 are 260 of the 342 ms of a 10M x 96 build).
__global__ void k_owned_filter(const SegLevel sg, const u32* __restrict__ seg_of, const u32* __restrict__ perm,
                               const i64* __restrict__ pid, u32 A, const u32* __restrict__ new_seg /* [R] or VI_NONE */,
                               const u32* __restrict__ shift /* [R] positions removed before the range */,
                               u32* __restrict__ perm2, i64* __restrict__ pid2, u32* __restrict__ seg_of2)
{
  const u32 p = blockIdx.x * 256u + threadIdx.x;
  if (p >= A) return;
  const u32 s = seg_of[p];
  const u32 ns = new_seg[s];
  if (ns == VI_NONE) return;
  const u32 q = p - shift[s];
  perm2[q] = perm[p];
  pid2[q] = pid[p];
  seg_of2[q] = ns;
}

static int build_sharded_exact(vi_ctx* ctx, BuildEnv& env)
{
  cudaStream_t st = ctx->stream;
  const int G = ctx->world, me = ctx->rank;
  const int ld = ctx->ld;
  cudaEvent_t ev_begin = env_event(ctx, env);
  if (!ctx->sh_dev)
  {
    VI_CUDA_TRY(cudaMalloc(&ctx->sh_dev, sizeof(ShDev)));
    VI_CUDA_TRY(cudaMallocHost(&ctx->sh_host, sizeof(ShHost)));
  }
  ShDev* sd = (ShDev*)ctx->sh_dev;
  ShHost* shh = (ShHost*)ctx->sh_host;
  // ---- every rank gets every point, in rank order ----------------------------------------------------------------------
  const u64 nloc = (u64)ctx->n;
  VI_CUDA_TRY(cudaMemsetAsync(sd->v0, 0, sizeof(sd->v0), st));
  VI_CUDA_TRY(cudaMemcpyAsync(sd->v0 + me, &nloc, 8, cudaMemcpyHostToDevice, st));
  int rc = vi_coll_allreduce_u64(ctx, sd->v0, G);
  if (rc != VI_OK) return rc;
  VI_CUDA_TRY(cudaMemcpyAsync(shh->v0, sd->v0, (size_t)G * 8, cudaMemcpyDeviceToHost, st));
  VI_CUDA_TRY(cudaStreamSynchronize(st));
  u64 N = 0, my_off = 0;
  std::vector<int64_t> off_r(G), len_r(G), off_i(G), len_i(G);
  for (int g = 0; g < G; ++g)
  {
    if (g == me) my_off = N;
    off_r[g] = (int64_t)N * ld * 4;
    len_r[g] = (int64_t)shh->v0[g] * ld * 4;
    off_i[g] = (int64_t)N * 8;
    len_i[g] = (int64_t)shh->v0[g] * 8;
    N += shh->v0[g];
  }
  if (N >= 0x7fffffffull) return ctx->fail(VI_ERR_CAPACITY, "more than 2^31-2 points");
  if (N == 0) return finish_table(ctx, env, 0, ev_begin, nullptr);
  if ((int64_t)N > ctx->own_cap)
  {
    cudaFree(ctx->own_rows); cudaFree(ctx->own_ids);
    ctx->own_rows = nullptr; ctx->own_ids = nullptr; ctx->own_cap = 0;
    VI_CUDA_TRY(cudaMalloc((void**)&ctx->own_rows, ((size_t)N + 64) * ld * sizeof(float)));
    VI_CUDA_TRY(cudaMalloc((void**)&ctx->own_ids, ((size_t)N + 64) * sizeof(i64)));
    ctx->own_cap = (int64_t)N;
  }
  if (nloc)
  {
    VI_CUDA_TRY(cudaMemcpyAsync(ctx->own_rows + my_off * ld, ctx->rows, (size_t)nloc * ld * 4, cudaMemcpyDeviceToDevice, st));
    VI_CUDA_TRY(cudaMemcpyAsync(ctx->own_ids + my_off, ctx->ids, (size_t)nloc * 8, cudaMemcpyDeviceToDevice, st));
  }
  if ((rc = vi_coll_allgatherv_inplace(ctx, ctx->own_rows, off_r.data(), len_r.data()))) return rc;
  if ((rc = vi_coll_allgatherv_inplace(ctx, ctx->own_ids, off_i.data(), len_i.data()))) return rc;
  ctx->own_n = (int64_t)N;
  const float* rows = ctx->own_rows;
  const u32 n = (u32)N;
  rc = grow_table(ctx, (int64_t)(2 * N + N / 8 + 1024), 0);
  if (rc != VI_OK) return rc;
  if (n == 1)
  {
    k_single_point<<<1, 1, 0, st>>>(ctx->own_ids, ctx->t_rid, ctx->t_dim, ctx->t_mid, ctx->t_id, ctx->t_low, ctx->t_high,
                                    ctx->t_src);
    ++env.launches;
    ctx->shared_rows = 1;
    return finish_table(ctx, env, 1, ev_begin, rows);
  }
  rc = alloc_workspace(ctx, (int64_t)N);
  if (rc != VI_OK) return rc;
  // ---- the common top: levels 0 .. L-1, the same on every rank ----------------------------------------------------------
  int L = 1;
  while ((1 << L) < G) ++L;  // stop in front of level L = ceil(log2 G): 2^L >= G ranges (mean splits are near halves)
  k_init_level0<<<(n + 255) / 256, 256, 0, st>>>(ctx->perm[0], ctx->pid[0], ctx->own_ids, ctx->seg_of[0], n, ctx->seg[0],
                                                 ctx->big_list[0], ctx->t_rid, ctx->t_low, ctx->t_high);
  ++env.launches;
  LevelState s{};
  s.A = n;
  s.R = 1;
  s.nbig = n >= env.t_big ? 1u : 0u;
  s.chunks = 0;
  s.minseg = n;
  s.maxseg = n;
  s.row_next = 1;
  s.cur = 0;
  s.level = 0;
  rc = run_levels(ctx, env, s, rows, L);
  if (rc != VI_OK) return rc;
  if (s.R == 0)
  {
    // the whole tree fitted in the common levels: every rank holds all of it
    ctx->shared_rows = s.row_next;
    return finish_table(ctx, env, s.row_next, ev_begin, rows);
  }
  // ---- ownership of the level-L ranges, then the rest alone --------------------------------------------------------
  const u32 R = s.R, T = s.row_next;
  if (R > (u32)VI_SH_MAXR * 2) return ctx->fail(VI_ERR_STATE, "too many ranges at the ownership level");
  std::vector<u32> cnt(R), start(R), rowi(R);
  std::vector<i64> rid(R);
  SegLevel& sg = ctx->seg[s.cur];
  VI_CUDA_TRY(cudaMemcpyAsync(cnt.data(), sg.count, (size_t)R * 4, cudaMemcpyDeviceToHost, st));
  VI_CUDA_TRY(cudaMemcpyAsync(start.data(), sg.start, (size_t)R * 4, cudaMemcpyDeviceToHost, st));
  VI_CUDA_TRY(cudaMemcpyAsync(rowi.data(), sg.row, (size_t)R * 4, cudaMemcpyDeviceToHost, st));
  VI_CUDA_TRY(cudaMemcpyAsync(rid.data(), sg.rid, (size_t)R * 8, cudaMemcpyDeviceToHost, st));
  VI_CUDA_TRY(cudaStreamSynchronize(st));
  std::vector<u32> order(R), owner(R);
  for (u32 i = 0; i < R; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](u32 a, u32 b) { return cnt[a] > cnt[b]; });
  std::vector<u64> load(G, 0);
  for (u32 i : order)
  {
    int best = 0;
    for (int g = 1; g < G; ++g)
      if (load[g] < load[best]) best = g;
    owner[i] = (u32)best;
    load[best] += cnt[i];
  }
  std::vector<u32> new_seg(R, VI_NONE), shift(R, 0), placeholders, f_start, f_count, f_row, f_big;
  std::vector<i64> f_rid;
  u32 removed = 0, pos = 0, minseg = 0xffffffffu, maxseg = 0;
  for (u32 i = 0; i < R; ++i)  // ranges are contiguous position slices in index order
  {
    if (owner[i] != (u32)me)
    {
      removed += cnt[i];
      placeholders.push_back(rowi[i]);
      continue;
    }
    new_seg[i] = (u32)f_start.size();
    shift[i] = removed;
    if (cnt[i] >= env.t_big) f_big.push_back((u32)f_start.size());
    f_start.push_back(pos);
    f_count.push_back(cnt[i]);
    f_row.push_back(rowi[i]);
    f_rid.push_back(rid[i]);
    pos += cnt[i];
    minseg = std::min(minseg, cnt[i]);
    maxseg = std::max(maxseg, cnt[i]);
  }
  Blob blob;
  const size_t o_ns = blob.add(new_seg), o_sh = blob.add(shift), o_ph = blob.add(placeholders), o_fs = blob.add(f_start),
               o_fc = blob.add(f_count), o_fr = blob.add(f_row), o_fb = blob.add(f_big), o_rid = blob.add(f_rid);
  if (blob.bytes.size() > VI_SH_STAGE) return ctx->fail(VI_ERR_CAPACITY, "ownership tables too large");
  memcpy(shh->tables, blob.bytes.data(), blob.bytes.size());
  VI_CUDA_TRY(cudaMemcpyAsync(sd->tables, shh->tables, blob.bytes.size(), cudaMemcpyHostToDevice, st));
  auto dptr = [&](size_t o) { return sd->tables + o; };
  if (!placeholders.empty())
  {
    k_sh_placeholders<<<((u32)placeholders.size() + 255) / 256, 256, 0, st>>>((const u32*)dptr(o_ph), (u32)placeholders.size(),
                                                                              ctx->t_dim);
    ++env.launches;
  }
  ctx->shared_rows = T;
  LevelState f{};
  f.row_next = T;
  f.level = s.level;
  f.cur = s.cur ^ 1;  // the filtered level goes to the other half of the ping-pong
  f.A = pos;
  f.R = (u32)f_start.size();
  f.nbig = (u32)f_big.size();
  f.chunks = 0;
  f.minseg = minseg;
  f.maxseg = maxseg;
  if (f.R > 0)
  {
    k_owned_filter<<<(s.A + 255) / 256, 256, 0, st>>>(sg, ctx->seg_of[s.cur], ctx->perm[s.cur], ctx->pid[s.cur], s.A,
                                                      (const u32*)dptr(o_ns), (const u32*)dptr(o_sh), ctx->perm[f.cur],
                                                      ctx->pid[f.cur], ctx->seg_of[f.cur]);
    ++env.launches;
    SegLevel& d = ctx->seg[f.cur];
    const size_t r4 = (size_t)f.R * 4;
    VI_CUDA_TRY(cudaMemcpyAsync(d.start, dptr(o_fs), r4, cudaMemcpyDeviceToDevice, st));
    VI_CUDA_TRY(cudaMemcpyAsync(d.count, dptr(o_fc), r4, cudaMemcpyDeviceToDevice, st));
    VI_CUDA_TRY(cudaMemcpyAsync(d.row, dptr(o_fr), r4, cudaMemcpyDeviceToDevice, st));
    VI_CUDA_TRY(cudaMemcpyAsync(d.rid, dptr(o_rid), (size_t)f.R * 8, cudaMemcpyDeviceToDevice, st));
    if (f.nbig) VI_CUDA_TRY(cudaMemcpyAsync(ctx->big_list[f.cur], dptr(o_fb), (size_t)f.nbig * 4, cudaMemcpyDeviceToDevice, st));
    rc = run_levels(ctx, env, f, rows);
    if (rc != VI_OK) return rc;
  }
  return finish_table(ctx, env, f.row_next, ev_begin, rows);
}

static int build_sharded(vi_ctx* ctx, BuildEnv& env)
{
  if (env.mode == VI_MODE_EXACT) return build_sharded_exact(ctx, env);
  int Lcap = 64;
  for (;;)
  {
    int retry_level = 0;
    const int rc = build_sharded_try(ctx, env, Lcap, &retry_level);
    if (rc != VI_RETRY) return rc;
    // every rank sees the same err_level (it comes from all-reduced statistics), so every rank retries alike
    Lcap = retry_level;
    ctx->levels.clear();
    ctx->info.point_visits = 0;
    ctx->info.shared_retry = 1;
  }
}

// Replicates the table of a multi-rank build on every rank (for query-sharded search): the rows of the shared levels
// meet in a small sum-all-reduce (a level-L root row comes from its owner), the owned sub-trees in one all-gather.
int vi_table_replicate_impl(vi_ctx* ctx)
{
  if (ctx->world <= 1 || ctx->replicated) return VI_OK;
  cudaStream_t st = ctx->stream;
  const int G = ctx->world, me = ctx->rank;
  const u32 T = (u32)ctx->shared_rows, rows = (u32)ctx->t_rows;
  if (!ctx->sh_dev) return ctx->fail(VI_ERR_STATE, "no multi-rank build to replicate");
  ShDev* sd = (ShDev*)ctx->sh_dev;
  ShHost* shh = (ShHost*)ctx->sh_host;
  // how many rows every rank owns
  const u64 mine = rows - T;
  VI_CUDA_TRY(cudaMemsetAsync(sd->v0, 0, sizeof(sd->v0), st));
  VI_CUDA_TRY(cudaMemcpyAsync(sd->v0 + me, &mine, 8, cudaMemcpyHostToDevice, st));
  int rc = vi_coll_allreduce_u64(ctx, sd->v0, G);
  if (rc != VI_OK) return rc;
  VI_CUDA_TRY(cudaMemcpyAsync(shh->v0, sd->v0, (size_t)G * 8, cudaMemcpyDeviceToHost, st));
  VI_CUDA_TRY(cudaStreamSynchronize(st));
  u64 total = T, my_off = T;
  std::vector<int64_t> off(G), len(G);
  for (int g = 0; g < G; ++g)
  {
    off[g] = (int64_t)total * 32;
    len[g] = (int64_t)shh->v0[g] * 32;
    if (g < me) my_off += shh->v0[g];
    total += shh->v0[g];
  }
  if (total >= 0x7fffffffull) return ctx->fail(VI_ERR_CAPACITY, "replicated table too large for 32-bit row indexes");
  u64* buf = nullptr;
  VI_CUDA_TRY(cudaMalloc((void**)&buf, (size_t)total * 32 + 256));
  cudaError_t e = cudaMemsetAsync(buf, 0, (size_t)T * 32, st);  // only the shared rows are summed
  if (e == cudaSuccess && rows > 0)
    k_pack_table<<<(rows + 255) / 256, 256, 0, st>>>(ctx->t_rid, ctx->t_dim, ctx->t_mid, ctx->t_id, ctx->t_low, ctx->t_high,
                                                     rows, T, (u32)my_off, me == 0 ? 1 : 0, buf);
  if (e != cudaSuccess) { cudaFree(buf); return ctx->fail_cuda(e, "pack table", __FILE__, __LINE__); }
  rc = vi_coll_allreduce_u64(ctx, buf, (int64_t)T * 4);
  if (rc == VI_OK) rc = vi_coll_allgatherv_inplace(ctx, buf, off.data(), len.data());
  if (rc != VI_OK) { cudaStreamSynchronize(st); cudaFree(buf); return rc; }
  rc = grow_table(ctx, (int64_t)total + 1024, 0);
  if (rc != VI_OK) { cudaFree(buf); return rc; }
  k_unpack_table<<<((u32)total + 255) / 256, 256, 0, st>>>(buf, (u32)total, ctx->t_rid, ctx->t_dim, ctx->t_mid, ctx->t_id,
                                                           ctx->t_low, ctx->t_high, ctx->t_src);
  k_pack_nodes<<<((u32)total + 255) / 256, 256, 0, st>>>(ctx->t_dim, ctx->t_mid, ctx->t_id, ctx->t_low, ctx->t_high,
                                                         ctx->t_node, (u32)total);
  e = cudaStreamSynchronize(st);
  cudaFree(buf);
  if (e != cudaSuccess) return ctx->fail_cuda(e, "unpack table", __FILE__, __LINE__);
  ctx->t_rows = (int64_t)total;
  ctx->shared_rows = (int64_t)total;
  ctx->replicated = true;
  ctx->src_rows = nullptr;  // the vectors stay with their owners
  ctx->info.ranges = (int64_t)total;
  return VI_OK;
}

int vi_build_impl(vi_ctx* ctx, int mode)
{
  ctx->replicated = false;
  const int64_t n64 = ctx->n;
  ctx->built = false;
  ctx->shared_rows = 0;
  ctx->levels.clear();
  ctx->info = vi_build_info();
  ctx->info.mode = mode;
  if (n64 >= (int64_t)0x7fffffff) return ctx->fail(VI_ERR_CAPACITY, "more than 2^31-2 points per context");
  BuildEnv env;
  env_init(ctx, env, mode);
  int rc;
  if (ctx->world > 1)
  {
    // sized for the local shard; grown once the rank knows how many points it owns (build_sharded phase C)
    rc = alloc_table(ctx, std::max<int64_t>(n64 + n64 / 4, 4096));
    if (rc == VI_OK) rc = build_sharded(ctx, env);
  }
  else if (n64 == 0)
  {
    // IndexBuilder.cs:70-73: an empty range emits no row
    ctx->t_rows = 0;
    ctx->built = true;
    rc = VI_OK;
  }
  else
  {
    rc = alloc_table(ctx, n64);
    if (rc == VI_OK) rc = build_single(ctx, env);
  }
  env_cleanup(env);
  return rc;
}
