// vi_search.cu -- batched proximity search over the device-resident range table.
//
// Replaces the recursive CTE dbo.Search (DDL.sql:234-295): from RangeID 0 follow LowRangeID iff
// Mid >= v[Dimension] - domain, HighRangeID iff Mid <= v[Dimension] + domain, emit TextID of every leaf reached.
// One thread walks one query with an explicit stack (depth <= 63 => at most 64 pending entries); rows are the
// packed 16-byte traversal rows built by k_pack_nodes.  Two passes (count, scan, fill) give a CSR result in the
// oracle's DFS order (low branch first).
//
// Candidate verification (the predicate half of Find, MemoryVectorIndex.cs:237-241,336-342): Euclidean distance
// with float32 accumulation in index order, as MemoryVectorIndexTests.cs:209-217, one thread per candidate so the
// summation order is the oracle's.
#include <algorithm>
#include <cstdlib>

#include "vi_common.cuh"

template <typename T>
__device__ __forceinline__ T warp_incl(T v)
{
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1)
  {
    T t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

// single-CTA exclusive scan of an int64 array in place, a[n] <- total (nq is at most a few million)
__global__ void __launch_bounds__(1024) k_scan_i64(i64* a, u32 n)
{
  __shared__ i64 wsum[32];
  __shared__ i64 carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int ITEMS = 8;
  for (u32 base = 0; base < n; base += 1024 * ITEMS)
  {
    const u32 i0 = base + threadIdx.x * ITEMS;
    i64 v[ITEMS];
    i64 run = 0;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k)
    {
      i64 t = (i0 + k < n) ? a[i0 + k] : 0;
      v[k] = run;
      run += t;
    }
    i64 incl = warp_incl(run);
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    i64 woff = 0, total = 0;
    for (int w = 0; w < 32; ++w)
    {
      i64 y = wsum[w];
      if (w < warp) woff += y;
      total += y;
    }
    const i64 off = carry_s + woff + incl - run;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k)
      if (i0 + k < n) a[i0 + k] = v[k] + off;
    __syncthreads();
    if (threadIdx.x == 0) carry_s += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) a[n] = carry_s;
}

// `stride`: the thread's query is i * stride (the sampling pass that picks the traversal kernel walks every
// stride-th query; the passes proper use stride 1)
template <bool FILL>
__global__ void __launch_bounds__(128)
k_search(const int4* __restrict__ node, const int* __restrict__ t_src, const float* __restrict__ queries, int ldq,
         u32 nq, u32 stride, float prox, i64* __restrict__ offsets, i64* __restrict__ ids_out, int* __restrict__ src_out,
         i64 cap, unsigned long long* __restrict__ visits)
{
  const u32 i = blockIdx.x * 128u + threadIdx.x;
  unsigned long long v = 0;
  if (i < nq)
  {
    const float* q = queries + (size_t)i * stride * ldq;
    u32 stack[64];
    int sp = 0;
    stack[sp++] = 0u;
    i64 cnt = 0;
    const i64 base = FILL ? offsets[i] : 0;
    while (sp > 0)
    {
      const u32 r = stack[--sp];
      const int4 nd = __ldg(node + r);
      ++v;
      if (nd.x < 0)
      {
        // leaf: TextID = RangeValue.Id (DDL.sql:295 "where TextID is not null")
        if (FILL && base + cnt < cap)
        {
          ids_out[base + cnt] = (i64)(((u64)(u32)nd.w << 32) | (u64)(u32)nd.z);
          if (src_out) src_out[base + cnt] = t_src[r];
        }
        ++cnt;
        continue;
      }
      const bool both = nd.x == VI_NODE_BOTH;  // `N.Dimension is null or ...` DDL.sql:275,290
      const float x = both ? 0.f : __ldg(q + nd.x);
      const float lo = __fsub_rn(x, prox);  // MinValue = value - @domain, DDL.sql:249
      const float hi = __fadd_rn(x, prox);  // MaxValue = value + @domain, DDL.sql:250
      const float mid = __int_as_float(nd.y);
      if ((both || mid <= hi) && nd.w >= 0) stack[sp++] = (u32)nd.w;  // DDL.sql:280-293 (pushed first, visited second)
      if ((both || mid >= lo) && nd.z >= 0) stack[sp++] = (u32)nd.z;  // DDL.sql:265-278
    }
    if (!FILL) offsets[i] = cnt;
  }
  if (!FILL && visits)
  {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(visits, v);
  }
}

// ---- warp per query: the per-warp frontier stack -------------------------------------------------------------------
// Queries that visit many rows (proximity > 0: ~2000 rows and ~600 candidates per query at 10M x 96, p = 0.01) are
// walked by a whole warp.  The pending rows of a query form ONE list in the oracle's DFS order (low branch first):
//   frontier (<= 32 rows, one per lane, in shared memory)  ++  stack (top first; shared memory, spilling to global).
// A round: every lane loads the packed row of its frontier entry (32 independent loads in flight per warp); the
// leaves at the head of the list are final and are written out together, in order (coalesced); every other entry is
// replaced in place by itself (a leaf waiting for its turn, flagged so that it is not loaded again until then) or by
// the children the query's box reaches (low, then high).  The first 32 entries of the new list stay the frontier, the
// rest goes back on the stack in reverse, and a frontier with free lanes is topped up from the stack.  The candidates
// therefore come out exactly in the order of the one-row-at-a-time walk (DDL.sql:246-295 visited low first), and each
// row is visited once.
//
// Three forms: COUNT (candidates per query), FILL (offsets known: ids, and source rows for verification, written in
// place) and POOL = count and keep: the ids go to chunks of a device pool (64 slots: a link to the next chunk and 63
// ids) taken with one atomic per chunk, so that one walk serves the two-call protocol -- k_search_gather then copies
// each query's chunks to its CSR slice.  A query that finds the pool full is only counted (head -2) and walked again by
// the FILL form.
constexpr int SW_WARPS = 8;
constexpr int SW_CHUNK = 64;
constexpr int SW_CHUNK_IDS = SW_CHUNK - 1;
constexpr u32 SW_LEAF = 0x80000000u;
constexpr int SW_SPILL = 4096;  // per-warp global spill entries (pending rows <= 64 per tree level: 64 * 63 + 64)
constexpr int SW_MODE_COUNT = 0, SW_MODE_POOL = 1, SW_MODE_FILL = 2;

struct SwCtl  // device control block of one launch
{
  unsigned long long visits;
  unsigned long long pool_cursor;
  u32 work_cursor;
  u32 err;  // bit 0: spill area exhausted, bit 1: some query found the pool full
};

template <int MODE>
__global__ void __launch_bounds__(SW_WARPS * 32)
k_search_warp(const int4* __restrict__ node, const int* __restrict__ t_src, const float* __restrict__ queries, int dims,
              int qpad, u32 nq, u32 stride, float prox, i64* __restrict__ offsets, i64* __restrict__ ids_out,
              int* __restrict__ src_out, i64 cap, i64* __restrict__ pool, u32 pool_cap, i64* __restrict__ head,
              int only_unwritten, u32* __restrict__ spill, int stack_cap, SwCtl* __restrict__ ctl)
{
  extern __shared__ u32 sw_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  u32* wbase = sw_smem + (size_t)warp * (qpad + 64 + stack_cap);
  float* qs = reinterpret_cast<float*>(wbase);
  u32* buf = wbase + qpad;   // [0, nf): the frontier; [0, 64): the new list of a round
  u32* stack = buf + 64;
  u32* gspill = spill + (size_t)(blockIdx.x * SW_WARPS + warp) * SW_SPILL;
  const int H = stack_cap >> 1;
  unsigned long long v = 0;
  for (;;)
  {
    u32 qi = 0;
    if (lane == 0) qi = atomicAdd(&ctl->work_cursor, 1u);
    qi = __shfl_sync(0xffffffffu, qi, 0);
    if (qi >= nq) break;
    if (MODE == SW_MODE_FILL && only_unwritten && head[qi] != -2) continue;
    __syncwarp();
    for (int d = lane; d < dims; d += 32) qs[d] = queries[(size_t)qi * stride * dims + d];
    if (lane == 0) buf[0] = 0u;  // RangeID 0
    __syncwarp();
    int nf = 1, sp = 0, gsp = 0;
    u32 cnt = 0;  // candidates so far (a table has fewer than 2^31 rows)
    const i64 obase = MODE == SW_MODE_FILL ? offsets[qi] : 0;
    // POOL: the chunk being filled (slot index of its link word), the one before it, the first one, and the number of
    // ids already in the current chunk
    u32 chunk = 0xffffffffu, prevc = 0xffffffffu, head_val = 0xffffffffu;
    int fill = SW_CHUNK_IDS;  // "full": the first emission takes a chunk
    bool dead = false;
    for (;;)
    {
      if (sp == 0 && gsp > 0)
      {
        gsp -= H;
        for (int i = lane; i < H; i += 32) stack[i] = __ldcg(gspill + gsp + i);
        sp = H;
        __syncwarp();
      }
      if (nf < 32 && sp > 0)
      {
        const int t = min(32 - nf, sp);
        if (lane < t) buf[nf + lane] = stack[sp - 1 - lane];
        sp -= t;
        nf += t;
        __syncwarp();
      }
      if (nf == 0) break;
      const bool valid = lane < nf;
      const u32 e = valid ? buf[lane] : 0u;
      bool leaf = valid && (e & SW_LEAF) != 0u;
      const u32 r = e & ~SW_LEAF;
      const u32 m_unknown = __ballot_sync(0xffffffffu, valid && !leaf);
      const int j0 = m_unknown ? __ffs(m_unknown) - 1 : nf;
      int4 nd = make_int4(0, 0, 0, 0);
      // rows not seen yet, and the flagged leaves that are certain to leave in this round (for their id)
      const bool loaded = valid && (!leaf || lane < j0);
      if (loaded) nd = __ldg(node + r);
      if (valid && !leaf)
      {
        ++v;
        leaf = nd.x < 0;
      }
      const u32 m_int = __ballot_sync(0xffffffffu, valid && !leaf);
      const int j = m_int ? __ffs(m_int) - 1 : nf;  // entries [0, j) are leaves at the head of the list: final
      // flagged leaves behind a row that has just turned out to be a leaf leave with it
      if (valid && lane < j && !loaded) nd = __ldg(node + r);
      if (j > 0)
      {
        if (MODE == SW_MODE_FILL)
        {
          const i64 o = obase + (i64)cnt + lane;
          if (lane < j && o < cap)
          {
            ids_out[o] = (i64)(((u64)(u32)nd.w << 32) | (u64)(u32)nd.z);
            if (src_out) src_out[o] = t_src[r];
          }
        }
        else if (MODE == SW_MODE_POOL)
        {
          if (!dead)
          {
            // lanes [0, room) still fit the current chunk, the others open the next one (j <= 32 < 63: at most one)
            const int room = SW_CHUNK_IDS - fill;
            if (j > room)
            {
              unsigned long long nw = 0;
              if (lane == 0) nw = atomicAdd(&ctl->pool_cursor, (unsigned long long)SW_CHUNK);
              nw = __shfl_sync(0xffffffffu, nw, 0);
              if (nw + SW_CHUNK > (unsigned long long)pool_cap) dead = true;
              else
              {
                if (lane == 0 && chunk != 0xffffffffu) pool[chunk] = (i64)nw;
                if (chunk == 0xffffffffu) head_val = (u32)nw;
                prevc = chunk;
                chunk = (u32)nw;
              }
            }
            if (!dead)
            {
              if (lane < j)
              {
                const u32 slot = lane < room ? (j > room ? prevc : chunk) + 1u + (u32)(fill + lane)
                                             : chunk + 1u + (u32)(lane - room);
                pool[slot] = (i64)(((u64)(u32)nd.w << 32) | (u64)(u32)nd.z);
              }
              fill = j > room ? j - room : fill + j;
            }
          }
        }
        cnt += (u32)j;
      }
      bool gl = false, gh = false;
      if (valid && !leaf)
      {
        const bool both = nd.x == VI_NODE_BOTH;  // `N.Dimension is null or ...` DDL.sql:275,290
        const float x = both ? 0.f : qs[nd.x];
        const float lo = __fsub_rn(x, prox);  // MinValue = value - @domain, DDL.sql:249
        const float hi = __fadd_rn(x, prox);  // MaxValue = value + @domain, DDL.sql:250
        const float mid = __int_as_float(nd.y);
        gh = (both || mid <= hi) && nd.w >= 0;  // DDL.sql:280-293
        gl = (both || mid >= lo) && nd.z >= 0;  // DDL.sql:265-278
      }
      // new list position of this lane's entries: leaves behind j keep one slot, rows expand to 0, 1 or 2
      const bool keep = valid && lane >= j;
      const bool one = keep && (leaf || gl || gh);
      const bool two = keep && !leaf && gl && gh;
      const int c = (int)one + (int)two;
      const u32 m1 = __ballot_sync(0xffffffffu, one), m2 = __ballot_sync(0xffffffffu, two);
      const u32 below = (1u << lane) - 1u;
      const int pos = __popc(m1 & below) + __popc(m2 & below);
      const int total = __popc(m1) + __popc(m2);
      __syncwarp();  // every lane has read its frontier entry
      if (c)
      {
        if (leaf) buf[pos] = r | SW_LEAF;
        else
        {
          if (gl) buf[pos] = (u32)nd.z;
          if (gh) buf[pos + (gl ? 1 : 0)] = (u32)nd.w;
        }
      }
      __syncwarp();
      nf = min(total, 32);
      const int over = total - nf;
      if (over > 0)
      {
        if (sp + over > stack_cap)
        {
          // shared stack full: its lower half goes to the warp's global area
          if (gsp + H > SW_SPILL)
          {
            if (lane == 0) atomicOr(&ctl->err, 1u);
            nf = 0;
            sp = 0;
            gsp = 0;
            break;
          }
          for (int i = lane; i < H; i += 32) __stcg(gspill + gsp + i, stack[i]);
          gsp += H;
          const int rest = sp - H;
          for (int b = 0; b < rest; b += 32)
          {
            const u32 t = (b + lane < rest) ? stack[H + b + lane] : 0u;
            __syncwarp();
            if (b + lane < rest) stack[b + lane] = t;
            __syncwarp();
          }
          sp = rest;
        }
        if (lane < over) stack[sp + over - 1 - lane] = buf[32 + lane];  // reversed: the earliest ends on top
        sp += over;
        __syncwarp();
      }
    }
    if (lane == 0)
    {
      if (MODE != SW_MODE_FILL) offsets[qi] = (i64)cnt;
      if (MODE == SW_MODE_POOL)
      {
        head[qi] = dead ? -2 : (head_val == 0xffffffffu ? -1 : (i64)head_val);
        if (dead) atomicOr(&ctl->err, 2u);
      }
    }
  }
  if (MODE != SW_MODE_FILL)
  {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && v) atomicAdd(&ctl->visits, v);
  }
}

// POOL form, second half: one warp per query copies the query's chunks to its CSR slice
__global__ void __launch_bounds__(256)
k_search_gather(const i64* __restrict__ pool, const i64* __restrict__ head, const i64* __restrict__ offsets, u32 nq,
                i64* __restrict__ ids_out, i64 cap)
{
  const int lane = threadIdx.x & 31;
  const u32 nwarps = gridDim.x * 8u;
  for (u32 q = blockIdx.x * 8u + (threadIdx.x >> 5); q < nq; q += nwarps)
  {
    i64 c = head[q];
    if (c < 0) continue;  // no candidates, or not written (the FILL form walks that query again)
    const i64 base = offsets[q], n = offsets[q + 1] - base;
    for (i64 k = 0; k < n; k += SW_CHUNK_IDS)
    {
      const i64 m = min((i64)SW_CHUNK_IDS, n - k);
      const i64 a = pool[c + lane];                             // slot 0 (lane 0): the link
      const i64 b = (lane + 32 <= m) ? pool[c + lane + 32] : 0;  // slots 32..63
      if (lane >= 1 && lane <= m && base + k + lane - 1 < cap) ids_out[base + k + lane - 1] = a;
      if (lane + 32 <= m && base + k + lane + 31 < cap) ids_out[base + k + lane + 31] = b;
      c = __shfl_sync(0xffffffffu, a, 0);
    }
  }
}

static int sw_grid(int64_t nq) { return (int)std::min<int64_t>((nq + SW_WARPS - 1) / SW_WARPS, (int64_t)VI_NUM_SMS * 8); }

static u32 sw_env(const char* name, u32 def, u32 lo, u32 hi)
{
  const char* s = getenv(name);
  if (!s || !*s) return def;
  const long v = strtol(s, nullptr, 10);
  return (u32)std::min<long>(std::max<long>(v, (long)lo), (long)hi);
}

template <typename T>
static cudaError_t sw_ensure(T** p, int64_t* cap, int64_t need)
{
  if (*p && *cap >= need) return cudaSuccess;
  cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  cudaError_t e = cudaMalloc((void**)p, (size_t)need * sizeof(T));
  if (e == cudaSuccess) *cap = need;
  return e;
}

// one launch of the warp kernel; ctl is zeroed first
static int sw_launch(vi_ctx* ctx, int mode, const float* d_queries, int64_t nq, u32 stride, float prox, i64* d_offsets,
                     i64* d_ids, int* d_src, int64_t cap, int only_unwritten)
{
  cudaStream_t st = ctx->stream;
  const int grid = sw_grid(nq);
  VI_CUDA_TRY(sw_ensure(&ctx->sw_spill, &ctx->sw_spill_cap, (int64_t)VI_NUM_SMS * 8 * SW_WARPS * SW_SPILL));
  const int stack_cap = (int)sw_env("VI_B200_SEARCH_STACK", 256, 64, 2048) & ~63;
  const int qpad = (ctx->dims + 31) & ~31;
  const size_t smem = (size_t)SW_WARPS * (qpad + 64 + stack_cap) * sizeof(u32);
  SwCtl* ctl = reinterpret_cast<SwCtl*>((unsigned long long*)ctx->counters + 8);
  VI_CUDA_TRY(cudaMemsetAsync(ctl, 0, sizeof(SwCtl), st));
#define SW_ARGS                                                                                                         \
  ctx->t_node, ctx->t_src, d_queries, ctx->dims, qpad, (u32)nq, stride, prox, d_offsets, d_ids, d_src, (i64)cap,           \
      ctx->sw_pool, (u32)std::min<int64_t>(ctx->sw_pool_cap, 0xffffff00ll), ctx->sw_head, only_unwritten, ctx->sw_spill,  \
      stack_cap, ctl
  if (mode == SW_MODE_COUNT)
  {
    VI_CUDA_TRY(cudaFuncSetAttribute(k_search_warp<SW_MODE_COUNT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_search_warp<SW_MODE_COUNT><<<grid, SW_WARPS * 32, smem, st>>>(SW_ARGS);
  }
  else if (mode == SW_MODE_POOL)
  {
    VI_CUDA_TRY(cudaFuncSetAttribute(k_search_warp<SW_MODE_POOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_search_warp<SW_MODE_POOL><<<grid, SW_WARPS * 32, smem, st>>>(SW_ARGS);
  }
  else
  {
    VI_CUDA_TRY(cudaFuncSetAttribute(k_search_warp<SW_MODE_FILL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_search_warp<SW_MODE_FILL><<<grid, SW_WARPS * 32, smem, st>>>(SW_ARGS);
  }
#undef SW_ARGS
  return VI_OK;
}

static int sw_check(vi_ctx* ctx, const SwCtl& h)
{
  if (h.err & 1u) return ctx->fail(VI_ERR_STATE, "search: a warp's pending-row stack outgrew its spill area");
  return VI_OK;
}

// have_offsets: d_offsets already holds the scanned counts of this very batch (skip the count pass)
//
// The count pass first walks a sample of the batch one thread per query; batches whose queries visit few rows (point
// lookups) stay on that kernel, the others go to the warp-per-query kernel.  Unless the caller wants source rows
// (ctx->search_want_src: verification, top-k), the warp count pass keeps the candidates in the pool, and the fill that
// follows (here when d_ids is given, else the next call with have_offsets) is a copy instead of a second walk.
int vi_search_impl(vi_ctx* ctx, const float* d_queries, int64_t nq, float proximity, i64* d_offsets, i64* d_ids,
                   int64_t cap, int64_t* total, int64_t* visits, bool have_offsets)
{
  cudaStream_t st = ctx->stream;
  *total = 0;
  if (visits) *visits = 0;
  if (!have_offsets) ctx->sw_pool_valid = false;
  if (nq == 0)
  {
    VI_CUDA_TRY(cudaMemsetAsync(d_offsets, 0, sizeof(i64), st));
    VI_CUDA_TRY(cudaStreamSynchronize(st));
    return VI_OK;
  }
  if (ctx->t_rows == 0)
  {
    VI_CUDA_TRY(cudaMemsetAsync(d_offsets, 0, sizeof(i64) * (size_t)(nq + 1), st));
    VI_CUDA_TRY(cudaStreamSynchronize(st));
    return VI_OK;
  }
  unsigned long long* d_vis = (unsigned long long*)ctx->counters + 1;  // counters[2..3]
  SwCtl* d_ctl = reinterpret_cast<SwCtl*>((unsigned long long*)ctx->counters + 8);
  const u32 grid = (u32)((nq + 127) / 128);
  i64 tot = 0;
  unsigned long long vis = 0;
  if (!have_offsets)
  {
    // ---- which kernel: visits per query on a sample ------------------------------------------------------------
    const u32 force = sw_env("VI_B200_SEARCH_PATH", 2, 0, 2);  // 0 thread, 1 warp, 2 by sample
    int path = (int)force;
    double cand_per_query = 0.0;
    // the sample is walked by the warp kernel (count form): a warp finishes a 2000-row query in ~100 rounds of
    // independent loads where a single thread needs 2000 dependent ones (3.9 ms for the sample alone, profiles/)
    const int64_t ns = std::min<int64_t>(nq, 256);
    if (force == 2 || force == 1)
    {
      int rc = sw_launch(ctx, SW_MODE_COUNT, d_queries, ns, (u32)(nq / ns), proximity, d_offsets, nullptr, nullptr, 0, 0);
      if (rc != VI_OK) return rc;
      k_scan_i64<<<1, 1024, 0, st>>>(d_offsets, (u32)ns);
      i64 stot = 0;
      SwCtl h{};
      VI_CUDA_TRY(cudaMemcpyAsync(&h, d_ctl, sizeof(SwCtl), cudaMemcpyDeviceToHost, st));
      VI_CUDA_TRY(cudaMemcpyAsync(&stot, d_offsets + ns, sizeof(i64), cudaMemcpyDeviceToHost, st));
      VI_CUDA_TRY(cudaStreamSynchronize(st));
      if ((rc = sw_check(ctx, h)) != VI_OK) return rc;
      cand_per_query = (double)stot / (double)ns;
      if (force == 2)
        path = ((double)h.visits / (double)ns >= (double)sw_env("VI_B200_SEARCH_WARP_VISITS", 96, 1, 1u << 30)) ? 1 : 0;
    }
    ctx->search_path = path;
    if (path == 0)
    {
      VI_CUDA_TRY(cudaMemsetAsync(d_vis, 0, 8, st));
      k_search<false><<<grid, 128, 0, st>>>(ctx->t_node, ctx->t_src, d_queries, ctx->dims, (u32)nq, 1u, proximity, d_offsets,
                                            nullptr, nullptr, 0, d_vis);
      k_scan_i64<<<1, 1024, 0, st>>>(d_offsets, (u32)nq);
      VI_CUDA_TRY(cudaMemcpyAsync(&vis, d_vis, 8, cudaMemcpyDeviceToHost, st));
    }
    else
    {
      bool pool = !ctx->search_want_src && sw_env("VI_B200_SEARCH_POOL", 1, 0, 1) != 0;
      if (pool)
      {
        // pool size from the sample: candidates per query + half a chunk of slack per query, 30 % head room
        const double est = (cand_per_query * 1.3 * SW_CHUNK / SW_CHUNK_IDS + SW_CHUNK) * (double)nq + 65536.0;
        int64_t need = (int64_t)std::min(est, 4.0e9);  // at most 32 GB (32-bit slot indexes); what does not fit is walked twice
        const u32 cap_env = sw_env("VI_B200_SEARCH_POOL_SLOTS", 0, 0, 0x7fffffffu);
        if (cap_env) need = cap_env;
        if (sw_ensure(&ctx->sw_pool, &ctx->sw_pool_cap, need) != cudaSuccess ||
            sw_ensure(&ctx->sw_head, &ctx->sw_head_cap, nq + 1) != cudaSuccess)
        {
          cudaGetLastError();
          pool = false;
        }
      }
      int rc = sw_launch(ctx, pool ? SW_MODE_POOL : SW_MODE_COUNT, d_queries, nq, 1u, proximity, d_offsets, nullptr, nullptr, 0, 0);
      if (rc != VI_OK) return rc;
      k_scan_i64<<<1, 1024, 0, st>>>(d_offsets, (u32)nq);
      SwCtl h{};
      VI_CUDA_TRY(cudaMemcpyAsync(&h, d_ctl, sizeof(SwCtl), cudaMemcpyDeviceToHost, st));
      VI_CUDA_TRY(cudaMemcpyAsync(&tot, d_offsets + nq, sizeof(i64), cudaMemcpyDeviceToHost, st));
      VI_CUDA_TRY(cudaStreamSynchronize(st));
      if ((rc = sw_check(ctx, h)) != VI_OK) return rc;
      vis = h.visits;
      if (pool)
      {
        ctx->sw_pool_valid = true;
        ctx->sw_pool_overflow = (h.err & 2u) != 0;
        ctx->sw_q = d_queries;
        ctx->sw_off = d_offsets;
        ctx->sw_nq = nq;
        ctx->sw_prox = proximity;
      }
    }
  }
  VI_CUDA_TRY(cudaMemcpyAsync(&tot, d_offsets + nq, sizeof(i64), cudaMemcpyDeviceToHost, st));
  VI_CUDA_TRY(cudaStreamSynchronize(st));
  *total = tot;
  if (visits && !have_offsets) *visits = (int64_t)vis;
  if (d_ids == nullptr) return VI_OK;
  if (cap < tot) return ctx->fail(VI_ERR_CAPACITY, "ids capacity smaller than the number of candidates");
  if (tot > 0)
  {
    if (ctx->search_path == 0)
      k_search<true><<<grid, 128, 0, st>>>(ctx->t_node, ctx->t_src, d_queries, ctx->dims, (u32)nq, 1u, proximity, d_offsets,
                                           d_ids, ctx->search_src, cap, nullptr);
    else
    {
      const bool from_pool = ctx->sw_pool_valid && ctx->search_src == nullptr && ctx->sw_q == d_queries &&
                             ctx->sw_off == d_offsets && ctx->sw_nq == nq && ctx->sw_prox == proximity;
      if (from_pool)
        k_search_gather<<<(u32)std::min<int64_t>((nq + 7) / 8, (int64_t)VI_NUM_SMS * 16), 256, 0, st>>>(
            ctx->sw_pool, ctx->sw_head, d_offsets, (u32)nq, d_ids, cap);
      if (!from_pool || ctx->sw_pool_overflow)
      {
        int rc = sw_launch(ctx, SW_MODE_FILL, d_queries, nq, 1u, proximity, d_offsets, d_ids, ctx->search_src, cap, from_pool ? 1 : 0);
        if (rc != VI_OK) return rc;
        SwCtl h{};
        VI_CUDA_TRY(cudaMemcpyAsync(&h, d_ctl, sizeof(SwCtl), cudaMemcpyDeviceToHost, st));
        VI_CUDA_TRY(cudaStreamSynchronize(st));
        if ((rc = sw_check(ctx, h)) != VI_OK) return rc;
      }
    }
    VI_CUDA_TRY(cudaStreamSynchronize(st));
  }
  VI_CUDA_TRY(cudaGetLastError());
  return VI_OK;
}

// ---- verification ------------------------------------------------------------------------------------------------
// one thread per candidate: float32 sum of squares in index order, sqrt, compare (MemoryVectorIndexTests.cs:209-217)
__global__ void __launch_bounds__(128)
k_verify_flags(const float* __restrict__ rows, int ld, int dims, const float* __restrict__ queries, int ldq,
               const i64* __restrict__ offsets, u32 nq, const int* __restrict__ src, i64 total, float distance,
               u32* __restrict__ keep)
{
  const i64 c = (i64)blockIdx.x * 128 + threadIdx.x;
  if (c >= total) return;
  // owning query: largest q with offsets[q] <= c
  u32 lo = 0, hi = nq;
  while (hi - lo > 1)
  {
    const u32 m = (lo + hi) >> 1;
    if (offsets[m] <= c) lo = m; else hi = m;
  }
  const float* a = rows + (size_t)src[c] * ld;
  const float* b = queries + (size_t)lo * ldq;
  float s = 0.f;
  for (int i = 0; i < dims; ++i)
  {
    const float t = __fsub_rn(a[i], b[i]);
    s = __fadd_rn(s, __fmul_rn(t, t));
  }
  keep[c] = (__fsqrt_rn(s) <= distance) ? 1u : 0u;
}

__global__ void __launch_bounds__(1024) k_scan_u32_single(u32* a, u32 n)
{
  __shared__ u32 wsum[32];
  __shared__ u32 carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int ITEMS = 8;
  for (u32 base = 0; base < n; base += 1024 * ITEMS)
  {
    const u32 i0 = base + threadIdx.x * ITEMS;
    u32 v[ITEMS];
    u32 run = 0;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k)
    {
      u32 t = (i0 + k < n) ? a[i0 + k] : 0;
      v[k] = run;
      run += t;
    }
    u32 incl = warp_incl(run);
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    u32 woff = 0, total = 0;
    for (int w = 0; w < 32; ++w)
    {
      u32 y = wsum[w];
      if (w < warp) woff += y;
      total += y;
    }
    const u32 off = carry_s + woff + incl - run;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k)
      if (i0 + k < n) a[i0 + k] = v[k] + off;
    __syncthreads();
    if (threadIdx.x == 0) carry_s += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) a[n] = carry_s;
}

__global__ void __launch_bounds__(256)
k_verify_compact(const i64* __restrict__ ids_in, const u32* __restrict__ keep_scan, i64 total, i64* __restrict__ ids_out,
                 i64 cap)
{
  const i64 c = (i64)blockIdx.x * 256 + threadIdx.x;
  if (c >= total) return;
  const u32 a = keep_scan[c], b = keep_scan[c + 1];
  if (b != a && (i64)a < cap) ids_out[a] = ids_in[c];
}

__global__ void __launch_bounds__(256)
k_verify_offsets(const i64* __restrict__ offsets_in, const u32* __restrict__ keep_scan, u32 nq, i64* __restrict__ offsets_out)
{
  const u32 q = blockIdx.x * 256u + threadIdx.x;
  if (q > nq) return;
  offsets_out[q] = (i64)keep_scan[offsets_in[q]];
}

// d_offsets_in/d_ids_in/src: result of a traversal with src rows; writes the filtered CSR.
int vi_verify_impl(vi_ctx* ctx, const float* d_queries, int64_t nq, float distance, const i64* d_offsets_in,
                   const i64* d_ids_in, int64_t total_in, i64* d_offsets_out, i64* d_ids_out, int64_t* total_out)
{
  cudaStream_t st = ctx->stream;
  *total_out = 0;
  if (total_in == 0)
  {
    VI_CUDA_TRY(cudaMemsetAsync(d_offsets_out, 0, sizeof(i64) * (size_t)(nq + 1), st));
    VI_CUDA_TRY(cudaStreamSynchronize(st));
    return VI_OK;
  }
  if (total_in >= (int64_t)0xfffffff0) return ctx->fail(VI_ERR_CAPACITY, "too many candidates to verify in one call");
  u32* keep = ctx->verify_keep;
  k_verify_flags<<<(u32)((total_in + 127) / 128), 128, 0, st>>>(ctx->src_rows, ctx->ld, ctx->dims, d_queries, ctx->dims,
                                                                d_offsets_in, (u32)nq, ctx->search_src, total_in, distance,
                                                                keep);
  k_scan_u32_single<<<1, 1024, 0, st>>>(keep, (u32)total_in);
  u32 kept = 0;
  VI_CUDA_TRY(cudaMemcpyAsync(&kept, keep + total_in, 4, cudaMemcpyDeviceToHost, st));
  k_verify_offsets<<<(u32)((nq + 1 + 255) / 256), 256, 0, st>>>(d_offsets_in, keep, (u32)nq, d_offsets_out);
  VI_CUDA_TRY(cudaStreamSynchronize(st));
  *total_out = kept;
  if (d_ids_out)
  {
    k_verify_compact<<<(u32)((total_in + 255) / 256), 256, 0, st>>>(d_ids_in, keep, total_in, d_ids_out, (i64)kept);
    VI_CUDA_TRY(cudaStreamSynchronize(st));
  }
  VI_CUDA_TRY(cudaGetLastError());
  return VI_OK;
}
