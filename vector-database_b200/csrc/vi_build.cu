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

#include "vi_common.cuh"
#include "vi_partition.cuh"
#include "vi_scan.cuh"
#include "vi_stats_exact.cuh"
#include "vi_stats_fast.cuh"

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
    cudaFree(s.pivot);
    s = SegLevel();
  }
  cudaFree(ctx->chunk_first); ctx->chunk_first = nullptr;
  cudaFree(ctx->fbits); ctx->fbits = nullptr;
  cudaFree(ctx->wpre); ctx->wpre = nullptr;
  cudaFree(ctx->seg_nlo); ctx->seg_nlo = nullptr;
  cudaFree(ctx->seg_hbase); ctx->seg_hbase = nullptr;
  cudaFree(ctx->c_rows); ctx->c_rows = nullptr;
  cudaFree(ctx->c_actpos); ctx->c_actpos = nullptr;
  cudaFree(ctx->scan_tmp); ctx->scan_tmp = nullptr;
  cudaFree(ctx->gacc); ctx->gacc = nullptr;
  cudaFree(ctx->gstats); ctx->gstats = nullptr;
  cudaFree(ctx->d_absmax); ctx->d_absmax = nullptr;
  if (ctx->totals) cudaFreeHost(ctx->totals);
  ctx->totals = nullptr;
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

static int alloc_workspace(vi_ctx* ctx)
{
  const int64_t n = ctx->n;
  if (ctx->ws_n >= n && ctx->perm[0]) return VI_OK;
  vi_free_workspace(ctx);
  const size_t N = (size_t)n;
  const size_t maxseg = N / 2 + 2;
  const size_t maxbig = N / VI_MIN_BIG + 2;
  const size_t words = N / 32 + 4;
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
  }
  VI_CUDA_TRY(dalloc(&ctx->chunk_first, maxbig + 1));
  VI_CUDA_TRY(dalloc(&ctx->fbits, words));
  VI_CUDA_TRY(dalloc(&ctx->wpre, words + 1));
  VI_CUDA_TRY(dalloc(&ctx->seg_nlo, maxseg));
  VI_CUDA_TRY(dalloc(&ctx->seg_hbase, maxseg));
  VI_CUDA_TRY(dalloc(&ctx->c_rows, maxseg + 1));
  VI_CUDA_TRY(dalloc(&ctx->c_actpos, maxseg + 1));
  VI_CUDA_TRY(dalloc((u64**)&ctx->scan_tmp, maxseg / SCAN_TILE + words / SCAN_TILE + 64));
  VI_CUDA_TRY(dalloc(&ctx->gacc, maxbig * ((size_t)ctx->ld * 3 + 2)));
  VI_CUDA_TRY(dalloc(&ctx->gstats, maxbig * (size_t)ctx->dims));
  VI_CUDA_TRY(dalloc(&ctx->d_absmax, (size_t)4));
  VI_CUDA_TRY(cudaMallocHost((void**)&ctx->totals, sizeof(LevelTotals)));
  ctx->ws_n = n;
  return VI_OK;
}

static int alloc_table(vi_ctx* ctx)
{
  // 2n-1 rows when no child is ever empty; an empty child (all points on one side) adds a one-child row.
  const int64_t cap = 2 * ctx->n + ctx->n / 8 + 1024;
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

// =============================================================================================================
// kernel dispatch on the row width
// =============================================================================================================
struct FastShape
{
  int ts, ch;
  bool full;
};
static FastShape fast_shape(int ld)
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
    if (shp.ts == 8 && shp.ch == 1) FAST_DISPATCH2(8, 1, shp.full, CALL);          \
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

static u32 env_u32(const char* name, u32 dflt, u32 lo, u32 hi)
{
  const char* v = getenv(name);
  if (!v || !*v) return dflt;
  long x = strtol(v, nullptr, 10);
  if (x < (long)lo) x = lo;
  if (x > (long)hi) x = hi;
  return (u32)x;
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
int vi_build_impl(vi_ctx* ctx, int mode)
{
  const int64_t n64 = ctx->n;
  ctx->built = false;
  ctx->levels.clear();
  ctx->info = vi_build_info();
  ctx->info.mode = mode;
  if (n64 >= (int64_t)0x7fffffff) return ctx->fail(VI_ERR_CAPACITY, "more than 2^31-2 points per context");
  if (n64 == 0)
  {
    // IndexBuilder.cs:70-73: an empty range emits no row
    ctx->t_rows = 0;
    ctx->built = true;
    return VI_OK;
  }
  int rc = alloc_table(ctx);
  if (rc != VI_OK) return rc;
  cudaStream_t st = ctx->stream;
  const u32 n = (u32)n64;
  const int ld = ctx->ld, dims = ctx->dims;
  int64_t launches = 0;

  // range-size classes (see vi_stats_fast.cuh / vi_stats_exact.cuh); tunable for experiments
  // defaults from scripts/sweep.py on 10M x 96 (profiles/r1_sweep.txt)
  const u32 t_team = env_u32("VI_B200_T_TEAM", 32, 2, VI_MAX_ROWS_PER_LANE);
  const u32 t_big_fast = env_u32("VI_B200_T_BIG", 1024, VI_MIN_BIG, VI_MAX_ROWS_PER_LANE);
  const u32 t_big_exact = env_u32("VI_B200_T_BIG_EXACT", 512, VI_MIN_BIG, 1u << 30);
  const u32 t_big = mode == VI_MODE_FAST ? t_big_fast : t_big_exact;
  const u32 big_unroll = env_u32("VI_B200_BIG_UNROLL", 4, 1, 4);

  cudaEvent_t ev_begin, ev_end;
  VI_CUDA_TRY(cudaEventCreate(&ev_begin));
  VI_CUDA_TRY(cudaEventCreate(&ev_end));
  std::vector<cudaEvent_t> lev_ev;
  auto new_event = [&]() -> cudaEvent_t
  {
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    lev_ev.push_back(e);
    return e;
  };
  auto cleanup = [&]()
  {
    for (cudaEvent_t e : lev_ev) cudaEventDestroy(e);
    cudaEventDestroy(ev_begin);
    cudaEventDestroy(ev_end);
  };

  VI_CUDA_TRY(cudaEventRecord(ev_begin, st));
  if (n == 1)
  {
    k_single_point<<<1, 1, 0, st>>>(ctx->ids, ctx->t_rid, ctx->t_dim, ctx->t_mid, ctx->t_id, ctx->t_low, ctx->t_high,
                                    ctx->t_src);
    k_pack_nodes<<<1, 32, 0, st>>>(ctx->t_dim, ctx->t_mid, ctx->t_id, ctx->t_low, ctx->t_high, ctx->t_node, 1);
    VI_CUDA_TRY(cudaEventRecord(ev_end, st));
    VI_CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, ev_begin, ev_end);
    cleanup();
    ctx->t_rows = 1;
    ctx->built = true;
    ctx->info.ranges = 1;
    ctx->info.levels = 1;
    ctx->info.kernel_launches = 2;
    ctx->info.build_ms = ms;
    return VI_OK;
  }
  rc = alloc_workspace(ctx);
  if (rc != VI_OK) { cleanup(); return rc; }

  // fast mode: quantisation exponent E with max|x| < 2^E
  int qe = 0;
  float qk = 1.0f;
  double qinv = 1.0;
  if (mode == VI_MODE_FAST)
  {
    VI_CUDA_TRY(cudaMemsetAsync(ctx->d_absmax, 0, 4, st));
    k_absmax<<<VI_NUM_SMS * 8, 256, 0, st>>>(reinterpret_cast<const float4*>(ctx->rows), (size_t)n * ld / 4,
                                             (u32*)ctx->d_absmax);
    ++launches;
    float amax = 0.f;
    VI_CUDA_TRY(cudaMemcpyAsync(&amax, ctx->d_absmax, 4, cudaMemcpyDeviceToHost, st));
    VI_CUDA_TRY(cudaStreamSynchronize(st));
    if (isinf(amax)) qe = 128;
    else if (amax > 0.f) (void)frexpf(amax, &qe);
    if (qe < -96) qe = -96;
    if (qe > 128) qe = 128;
    qk = ldexpf(1.0f, VI_QBITS - qe);
    qinv = ldexp(1.0, qe - VI_QBITS);
    ctx->info.q_exponent = qe;
  }

  TableOut tout{ctx->t_rid, ctx->t_dim, ctx->t_mid, ctx->t_id, ctx->t_low, ctx->t_high};
  StatsOut sout{ctx->t_dim, ctx->t_mid, ctx->t_id};

  k_init_level0<<<(n + 255) / 256, 256, 0, st>>>(ctx->perm[0], ctx->pid[0], ctx->ids, ctx->seg_of[0], n, ctx->seg[0],
                                                 ctx->big_list[0], ctx->t_rid, ctx->t_low, ctx->t_high);
  ++launches;

  u32 A = n, R = 1;
  u32 nbig = n >= t_big ? 1u : 0u;
  u32 chunks = nbig ? (n + VI_CHUNK - 1) / VI_CHUNK : 0u;
  u32 minseg = n, maxseg = n;
  if (nbig)
  {
    const u32 h[2] = {0u, chunks};
    VI_CUDA_TRY(cudaMemcpyAsync(ctx->chunk_first, h, sizeof(h), cudaMemcpyHostToDevice, st));
  }
  u32 row_base = 0, nrows = 1;
  int cur = 0;
  int level = 0;
  const FastShape shp = fast_shape(ld);
  const int chx = exact_chx(dims);
  const size_t gstride = (size_t)ld * 3 + 2;

  while (R > 0)
  {
    if (level >= VI_MAX_DEPTH)
    {
      cleanup();
      return ctx->fail(VI_ERR_OVERFLOW, "rangeId overflow: a range at depth 62 still holds more than one point "
                                        "(IndexBuilder.cs:99 checked(rangeId * 2 + 1))");
    }
    const int nxt = cur ^ 1;
    const int mx = (level & 1) == 0;  // root max = true, children !max (IndexBuilder.cs:33,128-129)
    SegLevel& sg = ctx->seg[cur];
    cudaEvent_t e0 = new_event();

    // ---- statistics + split choice -----------------------------------------------------------------------
    if (mode == VI_MODE_FAST)
    {
      if (nbig)
      {
        VI_CUDA_TRY(cudaMemsetAsync(ctx->gacc, 0, (size_t)nbig * gstride * sizeof(u64), st));
#define CALL_BIG(TS, CH, FULL)                                                                                         \
  if (big_unroll >= 4)                                                                                                 \
    k_stats_big_fast<TS, CH, FULL, 4><<<chunks, 256, 0, st>>>(sg, ctx->big_list[cur], ctx->chunk_first, nbig,           \
                                                              ctx->perm[cur], ctx->pid[cur], ctx->rows, ld, dims, qk,   \
                                                              qinv, mx, sout, ctx->gacc);                              \
  else                                                                                                                 \
    k_stats_big_fast<TS, CH, FULL, 2><<<chunks, 256, 0, st>>>(sg, ctx->big_list[cur], ctx->chunk_first, nbig, ctx->perm[cur], \
                                                         ctx->pid[cur], ctx->rows, ld, dims, qk, qinv, mx, sout, ctx->gacc)
        FAST_DISPATCH(shp, CALL_BIG);
#undef CALL_BIG
        ++launches;
        // ranges that fit one chunk and one column pass were finished by their CTA
        const int single_pass = (ld / 4) <= shp.ts * shp.ch;
        if (!single_pass || maxseg > VI_CHUNK)
        {
          k_finalize_big_fast<<<(nbig * 32 + 255) / 256, 256, 0, st>>>(sg, ctx->big_list[cur], nbig, ctx->gacc, ld, dims,
                                                                       qinv, mx, sout, ctx->rows, ctx->perm[cur],
                                                                       single_pass);
          ++launches;
        }
      }
      // warp-per-range class (teams of a warp share one range); for TS == 32 it also covers the team class
      const u32 wlo = shp.ts == 32 ? 2u : t_team;
      if (minseg < t_big && maxseg >= wlo)
      {
#define CALL_WARP(TS, CH, FULL)                                                                                        \
  k_stats_small_fast<TS, CH, FULL, true><<<(u32)(((u64)R * 32 + 255) / 256), 256, 0, st>>>(                             \
      sg, R, wlo, t_big, ctx->perm[cur], ctx->pid[cur], ctx->rows, ld, dims, qk, qinv, mx, sout)
        FAST_DISPATCH(shp, CALL_WARP);
#undef CALL_WARP
        ++launches;
      }
      if (shp.ts < 32 && minseg < t_team)
      {
#define CALL_TEAM(TS, CH, FULL)                                                                                        \
  k_stats_small_fast<TS, CH, FULL, false><<<(u32)(((u64)R * TS + 255) / 256), 256, 0, st>>>(                            \
      sg, R, 2u, t_team, ctx->perm[cur], ctx->pid[cur], ctx->rows, ld, dims, qk, qinv, mx, sout)
        FAST_DISPATCH(shp, CALL_TEAM);
#undef CALL_TEAM
        ++launches;
      }
    }
    else
    {
      if (nbig)
      {
        const u32 nblk = (u32)((dims + 31) / 32);
        k_stats_big_exact<<<nbig * nblk, 32, 0, st>>>(sg, ctx->big_list[cur], nblk, ctx->perm[cur], ctx->rows, ld, dims,
                                                      ctx->gstats);
        k_finalize_big_exact<<<(nbig * 32 + 255) / 256, 256, 0, st>>>(sg, ctx->big_list[cur], nbig, ctx->gstats,
                                                                      ctx->pid[cur], dims, mx, sout);
        launches += 2;
      }
      if (minseg < t_big)
      {
        const u32 grid = (u32)(((u64)R * 32 + 255) / 256);
#define CALL_EX(CHX) \
  k_stats_small_exact<CHX><<<grid, 256, 0, st>>>(sg, R, t_big, ctx->perm[cur], ctx->pid[cur], ctx->rows, ld, dims, mx, sout)
        switch (chx)
        {
          case 1: CALL_EX(1); break;
          case 2: CALL_EX(2); break;
          case 3: CALL_EX(3); break;
          case 4: CALL_EX(4); break;
          default: CALL_EX(8); break;
        }
#undef CALL_EX
        ++launches;
      }
    }
    cudaEvent_t e1 = new_event();

    // ---- stable partition ----------------------------------------------------------------------------------
    const u32 W = (A + 31) / 32;
    k_flags<<<W / 8 + 1, 256, 0, st>>>(sg, ctx->seg_of[cur], ctx->perm[cur], ctx->pid[cur], ctx->rows, ld, A, ctx->fbits,
                                       ctx->wpre);
    ++launches;
    scan_exclusive<u32>(ctx, ctx->wpre, W + 1, launches);
    k_seg_children<<<(R + 255) / 256, 256, 0, st>>>(sg, R, ctx->wpre, ctx->fbits, ctx->seg_nlo, ctx->seg_hbase,
                                                    ctx->c_rows, ctx->c_actpos);
    ++launches;
    scan_exclusive<u32>(ctx, ctx->c_rows, R, launches);
    scan_exclusive<u64>(ctx, ctx->c_actpos, R, launches);
    {
      const u32 init[8] = {0u, 0u, 0u, 0u, 0xffffffffu, 0u, 0u, 0u};  // [0] nbig [1] err [4] minseg [5] maxseg
      VI_CUDA_TRY(cudaMemcpyAsync(ctx->counters + 16, init, sizeof(init), cudaMemcpyHostToDevice, st));
    }
    u32* lvl_counters = ctx->counters + 16;  // u32[16..23]; [2..3] search visits, [4..5] divcheck
    const u32 row_base_next = row_base + nrows;
    k_emit_children<<<(R + 255) / 256, 256, 0, st>>>(sg, R, ctx->seg_nlo, ctx->c_rows, ctx->c_actpos, ctx->seg[nxt],
                                                     row_base_next, (u32)ctx->t_cap, tout, ctx->big_list[nxt], t_big,
                                                     lvl_counters);
    k_scatter<<<(A + 255) / 256, 256, 0, st>>>(sg, ctx->seg_of[cur], ctx->perm[cur], ctx->pid[cur], A, ctx->fbits,
                                               ctx->wpre, ctx->seg_nlo, ctx->seg_hbase, ctx->c_rows, ctx->c_actpos,
                                               row_base_next, ctx->perm[nxt], ctx->pid[nxt], ctx->seg_of[nxt], ctx->t_id,
                                               ctx->t_src, lvl_counters);
    launches += 2;
    const u32 big_bound = A / t_big + 1;
    u32* chunk_arr = nullptr;
    if (mode == VI_MODE_FAST)
    {
      chunk_arr = ctx->chunk_first;
      k_big_chunks<<<(big_bound + 255) / 256, 256, 0, st>>>(ctx->seg[nxt].count, ctx->big_list[nxt], lvl_counters,
                                                            chunk_arr, big_bound);
      ++launches;
      scan_exclusive<u32>(ctx, chunk_arr, big_bound, launches);
    }
    k_totals<<<1, 1, 0, st>>>(ctx->c_rows, ctx->c_actpos, R, lvl_counters, chunk_arr, big_bound, ctx->totals);
    ++launches;
    cudaEvent_t e2 = new_event();
    VI_CUDA_TRY(cudaStreamSynchronize(st));
    const LevelTotals tt = *ctx->totals;
    if (tt.err)
    {
      cleanup();
      return ctx->fail(VI_ERR_CAPACITY, "range table capacity exceeded (degenerate input: too many one-child ranges)");
    }
    vi_level_info li{};
    li.level = level;
    li.ranges = R;
    li.points = A;
    li.rows_emitted = nrows;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    li.stats_ms = ms;
    cudaEventElapsedTime(&ms, e1, e2);
    li.partition_ms = ms;
    ctx->levels.push_back(li);
    ctx->info.point_visits += A;

    row_base = row_base_next;
    nrows = tt.rows;
    R = tt.segs;
    A = tt.pos;
    nbig = tt.nbig;
    chunks = tt.chunks;
    minseg = tt.minseg;
    maxseg = tt.maxseg;
    cur = nxt;
    ++level;
  }
  // last level's rows are all leaves (or none)
  if (nrows > 0)
  {
    vi_level_info li{};
    li.level = level;
    li.rows_emitted = nrows;
    ctx->levels.push_back(li);
  }
  const u32 total_rows = row_base + nrows;
  k_pack_nodes<<<(total_rows + 255) / 256, 256, 0, st>>>(ctx->t_dim, ctx->t_mid, ctx->t_id, ctx->t_low, ctx->t_high,
                                                         ctx->t_node, total_rows);
  ++launches;
  VI_CUDA_TRY(cudaEventRecord(ev_end, st));
  VI_CUDA_TRY(cudaStreamSynchronize(st));
  VI_CUDA_TRY(cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, ev_begin, ev_end);
  cleanup();
  ctx->t_rows = total_rows;
  ctx->built = true;
  ctx->info.ranges = total_rows;
  ctx->info.levels = (int32_t)ctx->levels.size();
  ctx->info.kernel_launches = launches;
  ctx->info.build_ms = ms;
  return VI_OK;
}
