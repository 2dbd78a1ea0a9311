/*
 * vi_oracle_mt.c -- multi-threaded CPU baseline of the reference's builder.  TEST / BENCH INFRASTRUCTURE ONLY
 * (same rules as vi_oracle.c: only tests/ and bench.py's CPU legs may load it).
 *
 * The reference (IndexBuilder.cs) is strictly sequential.  This file keeps its arithmetic bit for bit -- the same
 * float32 recurrence per (range, dimension) in the same point order, the same split choice and stable partition --
 * and only uses the two kinds of independence the algorithm has:
 *   - the D per-dimension chains of one range are independent  -> big ranges: threads split the dimensions;
 *   - disjoint ranges are independent                           -> small ranges: threads take whole sub-trees.
 * tests/test_oracle_kat.py checks that its table equals vi_oracle.c's.  It exists so that `bench.py --impl reference`
 * can give the CPU "all the host threads it can use".
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef __int128 i128;

typedef struct
{
  int64_t rid, start, count;
  int max;
} range_t;

typedef struct
{
  int64_t* rid;
  int32_t* dim;
  float* mid;
  int64_t* id;
  int64_t n, cap;
} rows_t;

static int rows_push(rows_t* o, int64_t rid, int32_t dim, float mid, int64_t id)
{
  if (o->n == o->cap)
  {
    int64_t nc = o->cap ? o->cap * 2 : 4096;
    int64_t* a = (int64_t*)realloc(o->rid, (size_t)nc * 8);
    int32_t* b = (int32_t*)realloc(o->dim, (size_t)nc * 4);
    float* c = (float*)realloc(o->mid, (size_t)nc * 4);
    int64_t* e = (int64_t*)realloc(o->id, (size_t)nc * 8);
    if (a) o->rid = a;
    if (b) o->dim = b;
    if (c) o->mid = c;
    if (e) o->id = e;
    if (!a || !b || !c || !e) return 0;
    o->cap = nc;
  }
  o->rid[o->n] = rid;
  o->dim[o->n] = dim;
  o->mid[o->n] = mid;
  o->id[o->n] = id;
  ++o->n;
  return 1;
}

static inline int cmp_float_dotnet(float a, float b)
{
  if (a < b) return -1;
  if (a > b) return 1;
  if (a == b) return 0;
  if (isnan(a)) return isnan(b) ? 0 : -1;
  return 1;
}

/* IndexBuilder.cs:159-197 for the dimensions [d0, d1) of one range */
static void welford_dims(const float* rows, int64_t ld, const int64_t* p, int64_t count, int32_t d0, int32_t d1,
                         float* mean, float* q)
{
  const float* v0 = rows + p[0] * ld;
  for (int32_t i = d0; i < d1; ++i) { mean[i] = v0[i]; q[i] = 0.0f; }
  for (int64_t j = 1; j < count; ++j)
  {
    const float* v = rows + p[j] * ld;
    const float c = (float)(j + 1);
    for (int32_t i = d0; i < d1; ++i)
    {
      const float value = v[i], pa = mean[i];
      const float a = pa + (value - pa) / c;
      q[i] = q[i] + (value - pa) * (value - a);
      mean[i] = a;
    }
  }
}

/* split choice + row + stable partition of one range whose mean/q are known; returns 0 on allocation failure */
static int finish_range(const float* rows, int64_t ld, int32_t d, const int64_t* ids, int64_t* perm, int64_t* tmp,
                        range_t it, const float* mean, const float* q, rows_t* out, range_t* lo_out, range_t* hi_out,
                        int* overflow)
{
  const int64_t* p = perm + it.start;
  i128 idn = 0;
  for (int64_t j = 0; j < it.count; ++j) idn += (i128)ids[p[j]];
  lo_out->count = hi_out->count = 0;
  if (it.count == 1) return rows_push(out, it.rid, -1, 0.0f, (int64_t)idn); /* IndexBuilder.cs:81-82 */
  int32_t index = 0;
  float best = it.max ? q[0] : -q[0];
  for (int32_t i = 1; i < d; ++i) /* MaxBy: strictly greater replaces (IndexBuilder.cs:77-79) */
  {
    const float key = it.max ? q[i] : -q[i];
    if (cmp_float_dotnet(key, best) > 0) { best = key; index = i; }
  }
  const float mid = mean[index];
  const int64_t pivot = (int64_t)(idn / (i128)it.count);
  if (!rows_push(out, it.rid, index, mid, pivot)) return 0;
  if (it.rid > (INT64_MAX - 2) / 2) { *overflow = 1; return 1; } /* IndexBuilder.cs:99,104 */
  int64_t nlo = 0, nhi = 0;
  int64_t* lo = perm + it.start;
  int64_t* hi = tmp + it.start;
  for (int64_t j = 0; j < it.count; ++j) /* IndexBuilder.cs:111-124 */
  {
    const int64_t r = p[j];
    const float value = rows[r * ld + index];
    if (value > mid || (value == mid && ids[r] > pivot)) hi[nhi++] = r; else lo[nlo++] = r;
  }
  memcpy(perm + it.start + nlo, hi, (size_t)nhi * sizeof(int64_t));
  lo_out->rid = it.rid * 2 + 1; lo_out->start = it.start; lo_out->count = nlo; lo_out->max = !it.max;
  hi_out->rid = it.rid * 2 + 2; hi_out->start = it.start + nlo; hi_out->count = nhi; hi_out->max = !it.max;
  return 1;
}

static int cmp_range_desc(const void* a, const void* b)
{
  const int64_t x = ((const range_t*)a)->count, y = ((const range_t*)b)->count;
  return x < y ? 1 : (x > y ? -1 : 0);
}

/* literal build with `threads` threads; rows come out in no particular order */
int vio_build_mt(int64_t n, int32_t d, int64_t ld, const int64_t* ids, const float* rows, int64_t cap,
                 int64_t* out_range_id, int32_t* out_dim, float* out_mid, int64_t* out_id, int64_t* out_count,
                 int threads)
{
  *out_count = 0;
  if (n <= 0) return 0;
  if (threads < 1) threads = 1;
  omp_set_num_threads(threads);
  int64_t* perm = (int64_t*)malloc((size_t)n * 8);
  int64_t* tmp = (int64_t*)malloc((size_t)n * 8);
  float* mean = (float*)malloc((size_t)d * 4);
  float* q = (float*)malloc((size_t)d * 4);
  rows_t* outs = (rows_t*)calloc((size_t)threads + 1, sizeof(rows_t));
  const size_t dpad = ((size_t)d + 31) & ~(size_t)31; /* 128-byte multiples */
  float* priv = (float*)aligned_alloc(128, (size_t)threads * 2 * dpad * 4);
  int64_t tcap = 65536, ntask = 0, nopen = 1;
  range_t* tasks = (range_t*)malloc((size_t)tcap * sizeof(range_t));
  range_t* open_ = (range_t*)malloc((size_t)tcap * sizeof(range_t));
  range_t* next = (range_t*)malloc((size_t)tcap * sizeof(range_t));
  int rc = 0, overflow = 0, fail = 0;
  if (!perm || !tmp || !mean || !q || !outs || !tasks || !open_ || !next || !priv) { rc = -3; goto done; }
  for (int64_t i = 0; i < n; ++i) perm[i] = i;
  open_[0] = (range_t){0, 0, n, 1};
  /* top of the tree: ranges too big to hand to one thread -> threads split the dimensions */
  const int64_t big = n / (4 * (int64_t)threads) > 4096 ? n / (4 * (int64_t)threads) : 4096;
  while (nopen > 0 && !overflow && !fail)
  {
    int64_t nnext = 0;
    for (int64_t r = 0; r < nopen; ++r)
    {
      range_t it = open_[r];
      if (it.count == 0) continue;
      if (it.count < big || threads == 1)
      {
        if (ntask == tcap) { fail = 1; break; }
        tasks[ntask++] = it;
        continue;
      }
#pragma omp parallel
      {
        /* every thread keeps its chains in its own cache-line-aligned block (the shared mean[]/q[] would put the
         * slices of neighbouring threads on one line: false sharing on every step) and publishes them once */
        const int t = omp_get_thread_num(), T = omp_get_num_threads();
        const int32_t d0 = (int32_t)((int64_t)d * t / T), d1 = (int32_t)((int64_t)d * (t + 1) / T);
        if (d1 > d0)
        {
          float* pm = priv + (size_t)t * 2 * dpad;
          float* pq = pm + dpad;
          welford_dims(rows, ld, perm + it.start, it.count, d0, d1, pm, pq);
          memcpy(mean + d0, pm + d0, (size_t)(d1 - d0) * 4);
          memcpy(q + d0, pq + d0, (size_t)(d1 - d0) * 4);
        }
      }
      range_t lo, hi;
      if (!finish_range(rows, ld, d, ids, perm, tmp, it, mean, q, &outs[threads], &lo, &hi, &overflow)) { fail = 1; break; }
      if (nnext + 2 > tcap) { fail = 1; break; }
      if (lo.count) next[nnext++] = lo;
      if (hi.count) next[nnext++] = hi;
    }
    range_t* sw = open_; open_ = next; next = sw;
    nopen = nnext;
  }
  if (fail) { rc = -3; goto done; }
  /* below: whole sub-trees, largest first, one thread each */
  qsort(tasks, (size_t)ntask, sizeof(range_t), cmp_range_desc);
#pragma omp parallel
  {
    const int t = omp_get_thread_num();
    float* m2 = (float*)malloc((size_t)d * 4);
    float* q2 = (float*)malloc((size_t)d * 4);
    int64_t scap = 256, ssize = 0;
    range_t* st = (range_t*)malloc((size_t)scap * sizeof(range_t));
#pragma omp for schedule(dynamic, 1)
    for (int64_t k = 0; k < ntask; ++k)
    {
      if (!m2 || !q2 || !st) { fail = 1; continue; }
      st[0] = tasks[k];
      ssize = 1;
      while (ssize > 0 && !overflow && !fail)
      {
        range_t it = st[--ssize];
        if (it.count == 0) continue;
        welford_dims(rows, ld, perm + it.start, it.count, 0, d, m2, q2);
        range_t lo, hi;
        if (!finish_range(rows, ld, d, ids, perm, tmp, it, m2, q2, &outs[t], &lo, &hi, &overflow)) { fail = 1; break; }
        if (ssize + 2 > scap)
        {
          scap *= 2;
          range_t* ns = (range_t*)realloc(st, (size_t)scap * sizeof(range_t));
          if (!ns) { fail = 1; break; }
          st = ns;
        }
        if (lo.count) st[ssize++] = lo;
        if (hi.count) st[ssize++] = hi;
      }
    }
    free(m2); free(q2); free(st);
  }
  if (fail) { rc = -3; goto done; }
  if (overflow) { rc = -2; goto done; }
  {
    int64_t k = 0;
    for (int t = 0; t <= threads; ++t)
    {
      if (k + outs[t].n > cap) { rc = -1; goto done; }
      memcpy(out_range_id + k, outs[t].rid, (size_t)outs[t].n * 8);
      memcpy(out_dim + k, outs[t].dim, (size_t)outs[t].n * 4);
      memcpy(out_mid + k, outs[t].mid, (size_t)outs[t].n * 4);
      memcpy(out_id + k, outs[t].id, (size_t)outs[t].n * 8);
      k += outs[t].n;
    }
    *out_count = k;
  }
done:
  if (outs)
    for (int t = 0; t <= threads; ++t) { free(outs[t].rid); free(outs[t].dim); free(outs[t].mid); free(outs[t].id); }
  free(outs); free(perm); free(tmp); free(mean); free(q); free(tasks); free(open_); free(next); free(priv);
  return rc;
}
