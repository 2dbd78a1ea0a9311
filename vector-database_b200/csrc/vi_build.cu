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
  int mode;
  u32 t_team, t_big, big_unroll;
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

static void env_init(vi_ctx* ctx, BuildEnv& env, int mode)
{
  env.mode = mode;
  // range-size classes (see vi_stats_fast.cuh / vi_stats_exact.cuh); defaults from scripts/sweep.py on 10M x 96
  // (profiles/r1_sweep.txt); the variables exist for such sweeps
  env.t_team = env_u32("VI_B200_T_TEAM", 32, 2, VI_MAX_ROWS_PER_LANE);
  const u32 t_big_fast = env_u32("VI_B200_T_BIG", 512, VI_MIN_BIG, VI_MAX_ROWS_PER_LANE);  // raised to t_sub + 1 below
  const u32 t_big_exact = env_u32("VI_B200_T_BIG_EXACT", 512, VI_MIN_BIG, 1u << 30);
  env.t_big = mode == VI_MODE_FAST ? t_big_fast : t_big_exact;
  env.big_unroll = env_u32("VI_B200_BIG_UNROLL", 8, 0, 8);  // 0 = cp.async ring
  env.sibling = env_u32("VI_B200_SIBLING", 1, 0, 1);
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
  env.chx = exact_chx(ctx->dims);
  env.qk = 1.0f;
  env.qinv = 1.0;
  env.gstride = (size_t)ctx->ld * 3 + 3;
}

static void set_q_exponent(vi_ctx* ctx, BuildEnv& env, float amax)
{
  int qe = 0;
  if (isinf(amax)) qe = 128;
  else if (amax > 0.f) (void)frexpf(amax, &qe);
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
  const u32 grid = std::min<u32>((s.sub_cnt + SUB_WARPS - 1) / SUB_WARPS, (u32)VI_NUM_SMS * 2u);  // persistent warps, work cursor
  cudaEvent_t e0 = env_event(ctx, env);
#define CALL_SUB(CH, FULL)                                                                                                \
  do                                                                                                                      \
  {                                                                                                                       \
    VI_CUDA_TRY(cudaFuncSetAttribute(k_subtree_fast<CH, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
    k_subtree_fast<CH, FULL><<<grid, SUB_WARPS * 32, smem, st>>>(sl, s.sub_cnt, ctx->sub_perm, ctx->sub_pid, rows, ld, dims, \
                                                                 env.qk, env.qinv, tout, ctx->t_src, row_base,            \
                                                                 overflow_base, (u32)ctx->t_cap, cnt, lvlp, lvlr,         \
                                                                 rows_max);                                               \
  } while (0)
  const int ch = std::min(4, (ld / 4 + 7) / 8);  // int4 columns per team lane and pass
  const bool sub_full = ld == 32 * ch && dims == ld;
  if (ch <= 1) { if (sub_full) CALL_SUB(1, true); else CALL_SUB(1, false); }
  else if (ch == 2) { if (sub_full) CALL_SUB(2, true); else CALL_SUB(2, false); }
  else if (ch == 3) { if (sub_full) CALL_SUB(3, true); else CALL_SUB(3, false); }
  else { if (sub_full) CALL_SUB(4, true); else CALL_SUB(4, false); }
#undef CALL_SUB
  ++env.launches;
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
  const int mx = (level & 1) == 0;  // root max = true, children !max (IndexBuilder.cs:33,128-129)
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
    if (b.nbig)
    {
      VI_CUDA_TRY(cudaMemsetAsync(gacc_cur, 0, (size_t)b.nbig * env.gstride * sizeof(u64), st));
      launch_big_fast(ctx, env, lvp, rows, cur, b.chunks, mx, 1, gacc_cur, env.sibling ? 2 * t_big : 0xffffffffu,
                      env.sibling ? ctx->bl_sib[cur] : nullptr);
      // derived ranges (parent - sibling), ranges that span several chunks or column passes: split choice from gacc;
      // ranges that fit one chunk and one column pass were finished by their CTA
      const int single_pass = (ld / 4) <= env.shp_big.ts * env.shp_big.ch;
      if (!single_pass || b.maxseg > VI_CHUNK || env.sibling)
      {
        k_finalize_big_fast<<<(b.nbig * 32 + 255) / 256, 256, 0, st>>>(
            lvp, sg, ctx->big_list[cur], gacc_cur, gacc_prev, ld, dims, env.qinv, mx, sout, rows, ctx->perm[cur], single_pass,
            0, nullptr, env.sibling ? ctx->bl_parent[cur] : nullptr, ctx->bl_sib[cur]);
        ++env.launches;
      }
    }
    // warp-per-range class (teams of a warp share one range); for TS == 32 it also covers the team class.
    // Below the first level of a loop every range has more than t_sub points (smaller children went to the sub-tree
    // list); the first level's ranges (the root, the roots of a rank's forest) can have any size.
    const u32 floor_n = first_level ? 2u : env.t_sub + 1u;
    const u32 wlo = shp.ts == 32 ? 2u : t_team;
    const bool may_warp = std::max(wlo, floor_n) < t_big && b.maxseg >= wlo && (!exact_sizes || b.minseg < t_big);
    if (may_warp)
    {
#define CALL_WARP(TS, CH, FULL)                                                                              \
  k_stats_small_fast<TS, CH, FULL, true><<<(u32)(((u64)b.R * 32 + 255) / 256), 256, 0, st>>>(                 \
      lvp, sg, wlo, t_big, ctx->perm[cur], ctx->pid[cur], rows, ld, dims, env.qk, env.qinv, mx, sout)
      FAST_DISPATCH(shp, CALL_WARP);
#undef CALL_WARP
      ++env.launches;
    }
    const bool may_team = shp.ts < 32 && floor_n < t_team && (!exact_sizes || b.minseg < t_team);
    if (may_team)
    {
#define CALL_TEAM(TS, CH, FULL)                                                                              \
  k_stats_small_fast<TS, CH, FULL, false><<<(u32)(((u64)b.R * TS + 255) / 256), 256, 0, st>>>(                \
      lvp, sg, 2u, t_team, ctx->perm[cur], ctx->pid[cur], rows, ld, dims, env.qk, env.qinv, mx, sout)
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
                                                           ctx->ctile_mm, env.t_sub, t_big, sibling, (u32)ctx->t_cap,
                                                           ctx->chunk_first[nxt]);
  NextLevel nx{ctx->seg[nxt], ctx->perm[nxt], ctx->pid[nxt], ctx->seg_of[nxt], ctx->big_list[nxt], ctx->bl_parent[nxt],
               ctx->bl_sib[nxt], ctx->chunk_first[nxt]};
  SubList sl{ctx->sub_start, ctx->sub_count, ctx->sub_rid, ctx->sub_row, ctx->sub_depth};
  k_scatter<<<(b.A + 255) / 256, 256, 0, st>>>(lvp, lvp + 1, sg, ctx->seg_of[cur], ctx->perm[cur], ctx->pid[cur], fs,
                                               ctx->seg_nlo, ctx->seg_hbase, ctx->c_pre, ctx->ctile, env.t_sub, t_big, sibling,
                                               (u32)level + 1u, nx, tout, ctx->t_src, sl, ctx->sub_perm, ctx->sub_pid);
  env.launches += 3;
  ev.e2 = env_event(ctx, env);
  return VI_OK;
}

// The level loop: processes seg[s.cur] (ranges of depth s.level) until no range with >= 2 points is left.
// `rows` is the row store perm[] indexes.  `s` is the exact state of the first level; on return it holds the
// final row / sub-tree cursors.
static int run_levels(vi_ctx* ctx, BuildEnv& env, LevelState& s, const float* rows)
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
      b.nbig = (u32)std::min<u64>((u64)known.nbig * grow, (u64)known.A / t_big);
      b.chunks = known.A / VI_CHUNK + b.nbig;
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
// multi-rank build (protocol: include/vi_b200.h vi_set_collective, DESIGN.md "Multi-GPU build")
// =============================================================================================================
struct HostSeg
{
  u32 start, lcount;  // local slice
  u64 gcount;         // points over all ranks
  i64 rid;
  u32 row;
};

template <typename T>
static int upload(vi_ctx* ctx, T* dst, const std::vector<T>& src)
{
  if (src.empty()) return VI_OK;
  VI_CUDA_TRY(cudaMemcpyAsync(dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  VI_CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // src is a stack/heap temporary
  return VI_OK;
}

// all-reduce (sum) of a small host vector of u64 through the device scratch `gacc`
static int allreduce_host(vi_ctx* ctx, std::vector<u64>& v)
{
  if (v.empty()) return VI_OK;
  VI_CUDA_TRY(cudaMemcpyAsync(ctx->gacc, v.data(), v.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  VI_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  if (ctx->allreduce(ctx->coll_user, ctx->gacc, (int64_t)v.size()) != 0)
    return ctx->fail(VI_ERR_CUDA, "all-reduce callback failed");
  VI_CUDA_TRY(cudaMemcpyAsync(v.data(), ctx->gacc, v.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  VI_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return VI_OK;
}

static int build_sharded(vi_ctx* ctx, BuildEnv& env)
{
  cudaStream_t st = ctx->stream;
  const int G = ctx->world, me = ctx->rank;
  const int ld = ctx->ld, dims = ctx->dims;
  const u32 nloc = (u32)ctx->n;
  cudaEvent_t ev_begin = env_event(ctx, env);

  // ---- global size and quantisation exponent ----------------------------------------------------------------
  // 25 % slack: the points a rank owns after the exchange are rarely exactly its shard size
  int rc = alloc_workspace(ctx, std::max<int64_t>(ctx->n + ctx->n / 4, 4096));
  if (rc != VI_OK) return rc;
  float amax = 0.f;
  rc = local_absmax(ctx, ctx->rows, ctx->n, env, &amax);
  if (rc != VI_OK) return rc;
  std::vector<u64> v0((size_t)2 * G, 0);
  v0[me] = nloc;
  u32 amax_mine = 0;
  memcpy(&amax_mine, &amax, 4);
  v0[G + me] = amax_mine;  // non-negative floats order like their bit patterns
  rc = allreduce_host(ctx, v0);
  if (rc != VI_OK) return rc;
  u64 nglobal = 0;
  u32 amax_bits = 0;
  for (int g = 0; g < G; ++g)
  {
    nglobal += v0[g];
    amax_bits = std::max(amax_bits, (u32)v0[G + g]);
  }
  memcpy(&amax, &amax_bits, 4);
  set_q_exponent(ctx, env, amax);
  if (nglobal >= 0x7fffffffull) return ctx->fail(VI_ERR_CAPACITY, "more than 2^31-2 points");
  if (nglobal == 0) return finish_table(ctx, env, 0, ev_begin, nullptr);

  // ---- phase A: shared levels -------------------------------------------------------------------------------
  int L = 1;
  while ((1 << (L - 1)) < G) ++L;  // L = ceil(log2 G) + 1: about 2G ranges to balance over G owners
  std::vector<HostSeg> segs{{0u, nloc, nglobal, 0, 0u}};
  // shared rows are assembled on the host (a few dozen rows) and uploaded at the end of the phase
  std::vector<i64> h_rid{0};
  std::vector<int> h_low{-1}, h_high{-1}, h_leaf{nglobal == 1 ? 1 : 0};
  u32 T = 1;  // shared rows so far
  int cur = 0;
  int level = 0;
  if (nloc > 0)
  {
    k_init_level0<<<(nloc + 255) / 256, 256, 0, st>>>(ctx->perm[0], ctx->pid[0], ctx->ids, ctx->seg_of[0], nloc,
                                                      ctx->seg[0], ctx->big_list[0], ctx->t_rid, ctx->t_low, ctx->t_high);
    ++env.launches;
  }
  // scratch (the level loop's per-range prefix array is unused until phase C): ids of shared leaf rows (the holder
  // writes), then the small host-built tables of this phase
  u64* leaf_ids = reinterpret_cast<u64*>(ctx->c_pre);
  u32* scratch32 = reinterpret_cast<u32*>(ctx->c_pre) + 16384;
  VI_CUDA_TRY(cudaMemsetAsync(leaf_ids, 0, 4096 * sizeof(u64), st));
  VI_CUDA_TRY(cudaMemsetAsync(ctx->lv, 0, sizeof(LevelDev) * VI_LV_N, st));
  if (nglobal == 1)
  {
    if (nloc == 1)
      VI_CUDA_TRY(cudaMemcpyAsync(leaf_ids, ctx->ids, 8, cudaMemcpyDeviceToDevice, st));
    segs.clear();
  }
  StatsOut sout{ctx->t_dim, ctx->t_mid, ctx->t_id};
  u32* lvl_counters = ctx->counters + 16;
  const FlagScan fs{ctx->fbits, ctx->wloc, ctx->ftile};

  for (; level < L && !segs.empty(); ++level)
  {
    if (level >= VI_MAX_DEPTH) return ctx->fail(VI_ERR_OVERFLOW, "rangeId overflow (IndexBuilder.cs:99)");
    const int nxt = cur ^ 1;
    const int mx = (level & 1) == 0;
    const u32 R = (u32)segs.size();
    const u32 T_before = T;
    SegLevel& sg = ctx->seg[cur];
    cudaEvent_t e0 = env_event(ctx, env);
    // range list of this level (identical on every rank except for the local slices)
    std::vector<u32> h_start(R), h_count(R), h_row(R), h_big(R), h_cf(R + 1);
    std::vector<i64> h_srid(R);
    u32 A = 0, chunks = 0;
    for (u32 i = 0; i < R; ++i)
    {
      h_start[i] = segs[i].start;
      h_count[i] = segs[i].lcount;
      h_row[i] = segs[i].row;
      h_srid[i] = segs[i].rid;
      h_big[i] = i;
      h_cf[i] = chunks;
      chunks += (segs[i].lcount + VI_CHUNK - 1) / VI_CHUNK;
      A += segs[i].lcount;
    }
    h_cf[R] = chunks;
    LevelDev hl{};
    hl.A = A;
    hl.R = R;
    hl.nbig = R;
    hl.chunks = chunks;
    std::vector<LevelDev> h_lvrec{hl};
    LevelDev* lvp = ctx->lv + level;
    if ((rc = upload(ctx, sg.start, h_start)) || (rc = upload(ctx, sg.count, h_count)) || (rc = upload(ctx, sg.row, h_row)) ||
        (rc = upload(ctx, sg.rid, h_srid)) || (rc = upload(ctx, ctx->big_list[cur], h_big)) ||
        (rc = upload(ctx, ctx->chunk_first[cur], h_cf)) || (rc = upload(ctx, lvp, h_lvrec)))
      return rc;
    // local sums of every range -> gacc, one all-reduce, identical split on every rank
    VI_CUDA_TRY(cudaMemsetAsync(ctx->gacc, 0, (size_t)R * env.gstride * sizeof(u64), st));
    if (chunks > 0) launch_big_fast(ctx, env, lvp, ctx->rows, cur, chunks, mx, 0, ctx->gacc, 0xffffffffu, nullptr);
    VI_CUDA_TRY(cudaStreamSynchronize(st));
    if (ctx->allreduce(ctx->coll_user, ctx->gacc, (int64_t)((size_t)R * env.gstride)) != 0)
      return ctx->fail(VI_ERR_CUDA, "all-reduce callback failed");
    VI_CUDA_TRY(cudaMemsetAsync(lvl_counters, 0, 32, st));
    k_finalize_big_fast<<<(R * 32 + 255) / 256, 256, 0, st>>>(lvp, sg, ctx->big_list[cur], ctx->gacc, nullptr, ld, dims,
                                                              env.qinv, mx, sout, ctx->rows, ctx->perm[cur], 0, 1,
                                                              lvl_counters + 1, nullptr, nullptr);
    ++env.launches;
    cudaEvent_t e1 = env_event(ctx, env);
    // local partition flags and child sizes
    std::vector<u32> h_nlo(R, 0);
    if (A > 0)
    {
      k_flags<<<A / FL_TILE + 1, 256, 0, st>>>(lvp, &lvp->ticket[0], sg, ctx->seg_of[cur], ctx->perm[cur], ctx->pid[cur],
                                               ctx->rows, ld, ctx->fbits, ctx->wloc, ctx->ftile);
      k_seg_nlo<<<(R + 255) / 256, 256, 0, st>>>(sg, R, fs, ctx->seg_nlo, ctx->seg_hbase);
      env.launches += 2;
      VI_CUDA_TRY(cudaMemcpyAsync(h_nlo.data(), ctx->seg_nlo, R * 4, cudaMemcpyDeviceToHost, st));
    }
    u32 errflag = 0;
    VI_CUDA_TRY(cudaMemcpyAsync(&errflag, lvl_counters + 1, 4, cudaMemcpyDeviceToHost, st));
    VI_CUDA_TRY(cudaStreamSynchronize(st));
    if (errflag)
      return ctx->fail(VI_ERR_STATE, "multi-rank build: a range of the shared top levels is too tightly clustered for the "
                                     "fixed-point statistics (its float32 fallback needs all rows on one rank)");
    std::vector<u64> cnt((size_t)2 * R);
    for (u32 i = 0; i < R; ++i)
    {
      const u32 llo = segs[i].lcount ? h_nlo[i] : 0u;
      cnt[2 * i] = llo;
      cnt[2 * i + 1] = segs[i].lcount - llo;
    }
    std::vector<u64> gcnt = cnt;
    if ((rc = allreduce_host(ctx, gcnt))) return rc;
    // children: rows, leaves, next-level ranges, local destinations (same order as the single-rank builder:
    // by range, low child before high child)
    std::vector<HostSeg> next;
    std::vector<u32> lo_dst(R, VI_NONE), hi_dst(R, VI_NONE), lo_seg(R, 0), hi_seg(R, 0);
    std::vector<int> lo_leaf(R, -1), hi_leaf(R, -1);
    u32 npos = 0;
    for (u32 i = 0; i < R; ++i)
    {
      for (int side = 0; side < 2; ++side)
      {
        const u64 g = gcnt[2 * i + side];
        const u32 l = (u32)cnt[2 * i + side];
        if (g == 0) continue;  // empty range: no row (IndexBuilder.cs:70-73)
        const u32 row = T++;
        h_rid.push_back(segs[i].rid * 2 + 1 + side);
        h_low.push_back(-1);
        h_high.push_back(-1);
        h_leaf.push_back(g == 1 ? 1 : 0);
        (side ? h_high : h_low)[segs[i].row] = (int)row;
        if (g == 1)
          (side ? hi_leaf : lo_leaf)[i] = (int)row;
        else
        {
          (side ? hi_dst : lo_dst)[i] = npos;
          (side ? hi_seg : lo_seg)[i] = (u32)next.size();
          next.push_back({npos, l, g, segs[i].rid * 2 + 1 + side, row});
          npos += l;
        }
      }
    }
    if (T >= 4096) return ctx->fail(VI_ERR_CAPACITY, "too many shared rows");
    if (A > 0)
    {
      u32* d = scratch32;  // the six per-range arrays
      if ((rc = upload(ctx, d, lo_dst)) || (rc = upload(ctx, d + R, hi_dst)) || (rc = upload(ctx, d + 2 * R, lo_seg)) ||
          (rc = upload(ctx, d + 3 * R, hi_seg)) || (rc = upload(ctx, (int*)(d + 4 * R), lo_leaf)) ||
          (rc = upload(ctx, (int*)(d + 5 * R), hi_leaf)))
        return rc;
      ShScatter sc{d, d + R, d + 2 * R, d + 3 * R, (const int*)(d + 4 * R), (const int*)(d + 5 * R)};
      k_scatter_shared<<<(A + 255) / 256, 256, 0, st>>>(sg, ctx->seg_of[cur], ctx->perm[cur], ctx->pid[cur], A, fs,
                                                        ctx->seg_hbase, sc, ctx->perm[nxt], ctx->pid[nxt],
                                                        ctx->seg_of[nxt], leaf_ids, ctx->t_src);
      ++env.launches;
    }
    cudaEvent_t e2 = env_event(ctx, env);
    VI_CUDA_TRY(cudaStreamSynchronize(st));
    vi_level_info li{};
    li.level = level;
    li.ranges = R;
    li.points = A;  // local points visited
    li.rows_emitted = (int64_t)T - (int64_t)T_before;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    li.stats_ms = ms;
    cudaEventElapsedTime(&ms, e1, e2);
    li.partition_ms = ms;
    ctx->levels.push_back(li);
    ctx->info.point_visits += A;
    segs.swap(next);
    cur = nxt;
  }

  // shared rows: host-assembled columns + leaf ids (owner wrote, everyone sums)
  {
    std::vector<u64> lid(T, 0);
    VI_CUDA_TRY(cudaMemcpyAsync(lid.data(), leaf_ids, T * 8, cudaMemcpyDeviceToHost, st));
    VI_CUDA_TRY(cudaStreamSynchronize(st));
    if ((rc = allreduce_host(ctx, lid))) return rc;
    if ((rc = upload(ctx, ctx->t_rid, h_rid)) || (rc = upload(ctx, ctx->t_low, h_low)) || (rc = upload(ctx, ctx->t_high, h_high)))
      return rc;
    for (u32 r = 0; r < T; ++r)
      if (h_leaf[r])
      {
        const int dimv = -1;
        const float midv = 0.f;
        const i64 idv = (i64)lid[r];
        VI_CUDA_TRY(cudaMemcpyAsync(ctx->t_dim + r, &dimv, 4, cudaMemcpyHostToDevice, st));
        VI_CUDA_TRY(cudaMemcpyAsync(ctx->t_mid + r, &midv, 4, cudaMemcpyHostToDevice, st));
        VI_CUDA_TRY(cudaMemcpyAsync(ctx->t_id + r, &idv, 8, cudaMemcpyHostToDevice, st));
        VI_CUDA_TRY(cudaStreamSynchronize(st));
      }
  }
  ctx->shared_rows = T;

  // ---- phase B: ownership exchange --------------------------------------------------------------------------
  const u32 RL = (u32)segs.size();
  ctx->own_n = 0;
  LevelState s{};
  s.row_next = T;
  s.level = level;
  s.cur = 0;
  if (RL > 0)
  {
    // who holds how much of each range
    std::vector<u64> mat((size_t)G * RL, 0);
    for (u32 i = 0; i < RL; ++i) mat[(size_t)me * RL + i] = segs[i].lcount;
    if ((rc = allreduce_host(ctx, mat))) return rc;
    // owners: largest range first to the least loaded rank (ties: lower range index, lower rank)
    std::vector<u32> order(RL), owner(RL);
    for (u32 i = 0; i < RL; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](u32 a, u32 b) { return segs[a].gcount > segs[b].gcount; });
    std::vector<u64> load(G, 0);
    for (u32 i : order)
    {
      int best = 0;
      for (int g = 1; g < G; ++g)
        if (load[g] < load[best]) best = g;
      owner[i] = (u32)best;
      load[best] += segs[i].gcount;
    }
    // root rows of ranges owned elsewhere are placeholders here: Dimension -2
    for (u32 i = 0; i < RL; ++i)
      if (owner[i] != (u32)me)
      {
        const int dimv = -2;
        VI_CUDA_TRY(cudaMemcpyAsync(ctx->t_dim + segs[i].row, &dimv, 4, cudaMemcpyHostToDevice, st));
        VI_CUDA_TRY(cudaStreamSynchronize(st));
      }
    // send layout: by destination rank, then range index, then local position
    std::vector<u32> send_base(RL, 0);
    std::vector<int64_t> send_rows_n(G, 0), recv_rows_n(G, 0);
    u32 off = 0;
    for (int d = 0; d < G; ++d)
      for (u32 i = 0; i < RL; ++i)
        if (owner[i] == (u32)d)
        {
          send_base[i] = off;
          off += segs[i].lcount;
          send_rows_n[d] += segs[i].lcount;
        }
    u64 nown = 0;
    for (int g = 0; g < G; ++g)
      for (u32 i = 0; i < RL; ++i)
        if (owner[i] == (u32)me)
        {
          recv_rows_n[g] += (int64_t)mat[(size_t)g * RL + i];
          nown += mat[(size_t)g * RL + i];
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
    float* send_rows = ctx->send_rows;
    i64* send_ids = ctx->send_ids;
    u32 A = 0;
    for (u32 i = 0; i < RL; ++i) A += segs[i].lcount;
    if (A > 0)
    {
      // seg[cur] of the last shared level still describes local slices: refresh starts/counts for the pack kernel
      std::vector<u32> h_start(RL), h_count(RL);
      for (u32 i = 0; i < RL; ++i) { h_start[i] = segs[i].start; h_count[i] = segs[i].lcount; }
      if ((rc = upload(ctx, ctx->seg[cur].start, h_start)) || (rc = upload(ctx, ctx->seg[cur].count, h_count)) ||
          (rc = upload(ctx, scratch32, send_base)))
        return rc;
      const size_t threads = (size_t)A * (ld / 4);
      k_pack_rows<<<(u32)((threads + 255) / 256), 256, 0, st>>>(ctx->seg[cur], ctx->seg_of[cur], ctx->perm[cur],
                                                                ctx->pid[cur], ctx->rows, ld, A, scratch32, send_rows,
                                                                send_ids);
      ++env.launches;
    }
    VI_CUDA_TRY(cudaStreamSynchronize(st));
    std::vector<int64_t> sb(G), rb(G);
    for (int g = 0; g < G; ++g) { sb[g] = send_rows_n[g] * ld * 4; rb[g] = recv_rows_n[g] * ld * 4; }
    int cerr = ctx->alltoallv(ctx->coll_user, send_rows, sb.data(), ctx->own_rows, rb.data());
    for (int g = 0; g < G; ++g) { sb[g] = send_rows_n[g] * 8; rb[g] = recv_rows_n[g] * 8; }
    if (cerr == 0) cerr = ctx->alltoallv(ctx->coll_user, send_ids, sb.data(), ctx->own_ids, rb.data());
    if (cerr != 0) return ctx->fail(VI_ERR_CUDA, "all-to-all callback failed");
    ctx->own_n = (int64_t)nown;

    // ---- phase C: forest of owned ranges, pieces in source-rank order (= global stable order) --------------
    rc = alloc_workspace(ctx, std::max<int64_t>((int64_t)nown, 1024));
    if (rc != VI_OK) return rc;
    rc = grow_table(ctx, (int64_t)(2 * nown + nown / 8 + 1024 + T), T);
    if (rc != VI_OK) return rc;
    std::vector<u32> piece_dst, piece_src, piece_seg, f_start, f_count, f_row, f_big;
    std::vector<i64> f_rid;
    std::vector<u32> src_off(G, 0);  // running offset inside each source's block of the receive buffer
    {
      u32 acc = 0;
      for (int g = 0; g < G; ++g) { src_off[g] = acc; acc += (u32)recv_rows_n[g]; }
    }
    u32 pos = 0, minseg = 0xffffffffu, maxseg = 0, chunks = 0;
    std::vector<u32> f_cf;
    for (u32 i = 0; i < RL; ++i)
    {
      if (owner[i] != (u32)me) continue;
      const u32 fs = (u32)f_start.size();
      f_start.push_back(pos);
      f_count.push_back((u32)segs[i].gcount);
      f_rid.push_back(segs[i].rid);
      f_row.push_back(segs[i].row);
      for (int g = 0; g < G; ++g)
      {
        const u32 len = (u32)mat[(size_t)g * RL + i];
        if (len == 0) continue;
        piece_dst.push_back(pos);
        piece_src.push_back(src_off[g]);
        piece_seg.push_back(fs);
        pos += len;
        src_off[g] += len;
      }
      const u32 c = (u32)segs[i].gcount;
      minseg = std::min(minseg, c);
      maxseg = std::max(maxseg, c);
      if (c >= env.t_big)
      {
        f_big.push_back(fs);
        f_cf.push_back(chunks);
        chunks += (c + VI_CHUNK - 1) / VI_CHUNK;
      }
    }
    f_cf.push_back(chunks);
    s.A = pos;
    s.R = (u32)f_start.size();
    s.nbig = (u32)f_big.size();
    s.chunks = chunks;
    s.minseg = minseg;
    s.maxseg = maxseg;
    if (s.R > 0)
    {
      SegLevel& f = ctx->seg[0];
      u32* d = scratch32;
      const u32 np = (u32)piece_dst.size();
      if ((rc = upload(ctx, f.start, f_start)) || (rc = upload(ctx, f.count, f_count)) || (rc = upload(ctx, f.rid, f_rid)) ||
          (rc = upload(ctx, f.row, f_row)) || (rc = upload(ctx, ctx->big_list[0], f_big)) ||
          (rc = upload(ctx, ctx->chunk_first[0], f_cf)) || (rc = upload(ctx, d, piece_dst)) ||
          (rc = upload(ctx, d + np, piece_src)) || (rc = upload(ctx, d + 2 * np, piece_seg)))
        return rc;
      k_forest_init<<<(s.A + 255) / 256, 256, 0, st>>>(s.A, np, d, d + np, d + 2 * np, ctx->own_ids, ctx->perm[0],
                                                       ctx->pid[0], ctx->seg_of[0]);
      ++env.launches;
      rc = run_levels(ctx, env, s, ctx->own_rows);
      if (rc != VI_OK) return rc;
    }
  }
  return finish_table(ctx, env, s.row_next, ev_begin, ctx->own_n > 0 ? ctx->own_rows : ctx->rows);
}

// Replicates the table of a multi-rank build on every rank (for query-sharded search).
int vi_table_replicate_impl(vi_ctx* ctx)
{
  if (ctx->world <= 1 || ctx->replicated) return VI_OK;
  cudaStream_t st = ctx->stream;
  const int G = ctx->world, me = ctx->rank;
  const u32 T = (u32)ctx->shared_rows, rows = (u32)ctx->t_rows;
  std::vector<u64> own((size_t)G, 0);
  own[me] = rows - T;
  int rc = allreduce_host(ctx, own);
  if (rc != VI_OK) return rc;
  u64 total = T, my_off = T;
  for (int g = 0; g < G; ++g)
  {
    if (g < me) my_off += own[g];
    total += own[g];
  }
  if (total >= 0x7fffffffull) return ctx->fail(VI_ERR_CAPACITY, "replicated table too large for 32-bit row indexes");
  u64* buf = nullptr;
  VI_CUDA_TRY(cudaMalloc((void**)&buf, (size_t)total * 32 + 256));
  cudaError_t e = cudaMemsetAsync(buf, 0, (size_t)total * 32, st);
  if (e == cudaSuccess && rows > 0)
    k_pack_table<<<(rows + 255) / 256, 256, 0, st>>>(ctx->t_rid, ctx->t_dim, ctx->t_mid, ctx->t_id, ctx->t_low, ctx->t_high,
                                                     rows, T, (u32)my_off, me == 0 ? 1 : 0, buf);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { cudaFree(buf); return ctx->fail_cuda(e, "pack table", __FILE__, __LINE__); }
  if (ctx->allreduce(ctx->coll_user, buf, (int64_t)total * 4) != 0)
  {
    cudaFree(buf);
    return ctx->fail(VI_ERR_CUDA, "all-reduce callback failed");
  }
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
