/*
 * vi_oracle.c -- CPU ORACLE for the split-tree vector index.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is a plain-C restatement of the reference's algorithm for the hot path.  It exists so the
 * CUDA path can be checked against it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` leg may load it.  The product (vector-database_b200/) never links, imports or calls it.
 *
 * PARITY UNPINNED: the reference cannot run in this image (no .NET SDK, no SQL Server) and its own tests
 * hold no golden vector / known-answer test for IndexBuilder.Build or dbo.Search.
 * The oracle is therefore anchored on the reference's *source text*, cited line by line below, plus
 * hand-derivable known answers kept in tests/test_oracle_kat.py and an independent numpy restatement in
 * oracle/np_oracle.py.  The one executable expectation the reference's tests do state -- Find + the Euclidean predicate
 * returns exactly the plain scan's records on five fixed fixtures (MemoryVectorIndexTests.cs:10-113,161-204) -- is
 * replayed against this oracle (build + search + verify) in tests/test_reference_fixtures.py: it pins the
 * search-then-verify contract, not the rows of the range table.
 *
 * Reference files followed (paths relative to /root/reference):
 *   VectorIndex/IndexBuilder.cs:23-157   Build driver loop (DFS over ranges, explicit stack)
 *   VectorIndex/IndexBuilder.cs:159-173  InitStats  (first point of a range)
 *   VectorIndex/IndexBuilder.cs:175-197  UpdateStats (float32 Welford recurrence + Int128 id sum)
 *   VectorIndex/IndexBuilder.cs:77-88    split choice (MaxBy, first index wins) and RangeValue fields
 *   VectorIndex/IndexBuilder.cs:111-124  stable partition at Mid with the id tie-break
 *   VectorIndex/Stats.cs:6-27            accumulator field types (float Mean, float Stdev2N, long Count, Int128 IdN)
 *   VectorIndex/RangeValue.cs:6-22       output row (int Dimension, float Mid, long Id)
 *   DDL.sql:246-295                      dbo.Search traversal (vector +- domain from RangeID 0)
 *
 * Build flags (oracle/Makefile): -O2 -ffp-contract=off -fno-fast-math -mfpmath=sse : every float operation is a
 * single IEEE binary32 operation, exactly as RyuJIT x64 evaluates C# `float` arithmetic (no FMA contraction).
 *
 * Two build modes:
 *   mode 0  "literal": the reference's arithmetic, bit for bit.
 *   mode 1  "qfx":     the B200 fast mode's *specification* (order-independent fixed-point integer sums),
 *                      restated on the CPU so the GPU fast path can be checked bit-exactly against it.  It is
 *                      not the reference's arithmetic; tests report its divergence from mode 0.
 *   mode 2  "sql":     the rules of the T-SQL builder dbo.BuildIndex (DDL.sql:44-202) over the qfx statistics:
 *                      - max stdev at depth 0, MIN at depth 1, max at every depth >= 2: `iif(@level % 2 = 1, Stdev,
 *                        -Stdev) desc` with @level = 0, 1, 3, 7, ... (DDL.sql:113,151,155);
 *                      - the root sends Value = Mean to the HIGH child whatever the id (DDL.sql:104, `iif(Value <
 *                        Mean, 1, 2)`); deeper levels use Value < Mean / Value > Mean / ID <= avg(ID) (DDL.sql:161-167),
 *                        which is IndexBuilder's predicate;
 *                      - a range whose chosen dimension has Stdev = 0 keeps splitting by ID but its row has
 *                        Dimension = null, Mid = null (DDL.sql:193-194): here Dimension -3, Mid NaN, and dbo.Search
 *                        follows both children of such a row (DDL.sql:275,290 `N.Dimension is null or ...`);
 *                      - avg(ID) is the truncating bigint average, as IndexBuilder's Int128 one.
 *                      SQL Server's float(53) avg / stdev aggregate in an unspecified order, and `top 1 ... order by
 *                      Stdev` breaks ties arbitrarily: the statistics here are the qfx integer sums (order-free, Mean
 *                      within 2^-27 * 2^E of the exact average), ties go to the lowest dimension, and "Stdev = 0" is
 *                      decided where the qfx rules already hand a poorly resolved range to the float32 recurrence
 *                      (its Stdev2N == 0, i.e. all values of the chosen dimension are equal).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef __int128 i128;
typedef unsigned __int128 u128;

#define VIO_OK 0
#define VIO_ERR_CAPACITY -1 /* output arrays too small */
#define VIO_ERR_OVERFLOW -2 /* checked(rangeId*2+1/2) overflow, IndexBuilder.cs:99,104 */
#define VIO_ERR_NOMEM -3
#define VIO_ERR_ARG -4

/* qfx mode: a range is "resolved" when some dimension has n^2 * var >= n^2 * 2^(2*5) in quantised units */
#define VIO_QFX_MIN_RES_BITS 5
#define VIO_QBITS 26 /* fixed-point fraction bits: xi = rint(x * 2^(VIO_QBITS - E)), |xi| <= 2^26 */

/* float.CompareTo as used by Comparer<float>.Default inside Enumerable.MaxBy (IndexBuilder.cs:77-79):
 * NaN sorts below every number, NaN == NaN, -0 == +0. */
static inline int cmp_float_dotnet(float a, float b)
{
  if (a < b) return -1;
  if (a > b) return 1;
  if (a == b) return 0;
  if (isnan(a)) return isnan(b) ? 0 : -1;
  return 1;
}

typedef struct
{
  int64_t range_id;
  int64_t start;
  int64_t count;
  int max; /* IndexBuilder.cs:31 : (rangeId, store, max) */
} work_item;

typedef struct
{
  work_item* items;
  int64_t size, cap;
} work_stack;

static int push(work_stack* s, work_item it)
{
  if (s->size == s->cap)
  {
    int64_t ncap = s->cap ? s->cap * 2 : 1024;
    work_item* p = (work_item*)realloc(s->items, (size_t)ncap * sizeof(work_item));
    if (!p) return 0;
    s->items = p;
    s->cap = ncap;
  }
  s->items[s->size++] = it;
  return 1;
}

/* Quantisation exponent of the qfx mode: smallest E (clamped) with max|x| < 2^E, from frexp. */
int vio_qfx_exponent(const float* rows, int64_t n, int32_t d, int64_t ld)
{
  float amax = 0.0f;
  for (int64_t r = 0; r < n; ++r)
    for (int32_t i = 0; i < d; ++i)
    {
      float a = fabsf(rows[r * ld + i]);
      if (a > amax) amax = a; /* NaN never wins: comparisons with NaN are false */
    }
  int e = 0;
  if (isinf(amax)) return INT32_MAX; /* +-Inf has no fixed-point image: the qfx modes refuse it (vio_build: VIO_ERR_ARG) */
  if (amax > 0.0f) (void)frexpf(amax, &e);
  if (e < -96) e = -96;
  if (e > 128) e = 128;
  return e;
}

static inline int32_t qfx_quantise(float x, float k)
{
  /* GPU: __float2int_rn(x * k) : round-to-nearest-even, NaN -> 0, saturating. */
  float y = x * k;
  if (isnan(y)) return 0;
  if (y >= 2147483648.0f) return INT32_MAX;
  if (y <= -2147483648.0f) return INT32_MIN;
  return (int32_t)lrintf(y);
}

/*
 * Build the range table.
 *   ids[n], rows[n*ld] (row r at rows + r*ld, d used floats), mode 0 literal / 1 qfx.
 * Output rows are written in the reference's emission order (DFS, high child popped first:
 * IndexBuilder.cs:128-129 pushes low then high onto a Stack).  Returns VIO_OK and *out_count.
 */
/* vio_build_ex: the same walk started from an arbitrary range (root_rid, root_max = its depth is even) with a given
 * fixed-point exponent (qe_override != INT32_MIN): what one rank of a multi-rank build does for a sub-tree it owns. */
int vio_build_ex(int64_t n, int32_t d, int64_t ld, const int64_t* ids, const float* rows, int mode,
                 int64_t cap, int64_t* out_range_id, int32_t* out_dim, float* out_mid, int64_t* out_id,
                 int64_t* out_count, int64_t root_rid, int root_max, int32_t qe_override);

int vio_build(int64_t n, int32_t d, int64_t ld, const int64_t* ids, const float* rows, int mode,
              int64_t cap, int64_t* out_range_id, int32_t* out_dim, float* out_mid, int64_t* out_id,
              int64_t* out_count)
{
  return vio_build_ex(n, d, ld, ids, rows, mode, cap, out_range_id, out_dim, out_mid, out_id, out_count, 0, 1,
                      INT32_MIN);
}

int vio_build_ex(int64_t n, int32_t d, int64_t ld, const int64_t* ids, const float* rows, int mode,
                 int64_t cap, int64_t* out_range_id, int32_t* out_dim, float* out_mid, int64_t* out_id,
                 int64_t* out_count, int64_t root_rid, int root_max, int32_t qe_override)
{
  if (n < 0 || d <= 0 || ld < d || (mode != 0 && mode != 1 && mode != 2)) return VIO_ERR_ARG;
  const int sql = (mode == 2);
  *out_count = 0;
  if (n == 0) return VIO_OK; /* IndexBuilder.cs:70-73 : empty root emits nothing */

  int rc = VIO_OK;
  int64_t* perm = (int64_t*)malloc((size_t)n * sizeof(int64_t));
  int64_t* tmp = (int64_t*)malloc((size_t)n * sizeof(int64_t));
  float* mean = (float*)malloc((size_t)d * sizeof(float));
  float* q = (float*)malloc((size_t)d * sizeof(float));
  int64_t* s1 = (int64_t*)malloc((size_t)d * sizeof(int64_t));
  u128* s2 = (u128*)malloc((size_t)d * sizeof(u128));
  work_stack st = {0, 0, 0};
  if (!perm || !tmp || !mean || !q || !s1 || !s2) { rc = VIO_ERR_NOMEM; goto done; }
  for (int64_t i = 0; i < n; ++i) perm[i] = i;

  int qe = 0;
  float qk = 1.0f;
  double qinv = 1.0;
  if (mode != 0 && (n > 1 || qe_override != INT32_MIN))
  {
    qe = qe_override != INT32_MIN ? qe_override : vio_qfx_exponent(rows, n, d, ld);
    if (qe == INT32_MAX) { rc = VIO_ERR_ARG; goto done; } /* +-Inf in the data: only the literal mode takes it */
    qk = ldexpf(1.0f, VIO_QBITS - qe);
    qinv = ldexp(1.0, qe - VIO_QBITS);
  }

  work_item root = {root_rid, 0, n, root_max}; /* IndexBuilder.cs:33 : (0, points, true) */
  if (!push(&st, root)) { rc = VIO_ERR_NOMEM; goto done; }

  int64_t emitted = 0;
  while (st.size > 0)
  {
    work_item it = st.items[--st.size]; /* IndexBuilder.cs:37 */
    const int64_t* p = perm + it.start;
    int64_t count = it.count;
    if (count == 0) continue; /* IndexBuilder.cs:70-73 */

    i128 idn = 0;
    int32_t index = 0;
    float mid = 0.0f;

    int literal = (mode == 0);
    int depth = 0;
    for (uint64_t t = (uint64_t)it.range_id + 1; t > 1; t >>= 1) ++depth;
    if (sql) it.max = (depth != 1); /* DDL.sql:113 (root: Stdev desc), :151 with @level = 0, 1, 3, 7, ... (:155) */
    int null_dim = 0;
    if (mode != 0)
    {
      /* qfx specification (DESIGN.md "fast mode"): exact integer sums of xi = rint(x * 2^(26-E)). */
      for (int32_t i = 0; i < d; ++i) { s1[i] = 0; s2[i] = 0; }
      for (int64_t j = 0; j < count; ++j)
      {
        const float* v = rows + p[j] * ld;
        for (int32_t i = 0; i < d; ++i)
        {
          int64_t xi = qfx_quantise(v[i], qk);
          s1[i] += xi;
          s2[i] += (u128)(uint64_t)(xi * xi);
        }
        idn += (i128)ids[p[j]];
      }
      /* key K = n*S2 - S1^2 (exact, >= 0); even depth argmax, odd depth argmin, lowest index wins ties */
      u128 bestk = 0;
      u128 thr = ((u128)(uint64_t)count * (u128)(uint64_t)count) << (2 * VIO_QFX_MIN_RES_BITS);
      for (int32_t i = 0; i < d; ++i)
      {
        i128 a = (i128)s1[i];
        u128 k = (u128)(uint64_t)count * s2[i] - (u128)(a * a);
        if (i == 0 || (it.max ? (k > bestk) : (k < bestk))) { bestk = k; index = i; }
      }
      mid = (float)(((double)s1[index] / (double)count) * qinv);
      /* Poorly resolved range: the CHOSEN dimension spreads over fewer than 2^5 quantisation steps (stdev), so the
       * integer statistics cannot place Mid between its points; the range takes the reference's own float32
       * statistics instead (on max-variance levels this means no dimension is resolved). */
      if (bestk < thr) { literal = 1; idn = 0; index = 0; }
    }
    if (literal)
    {
      /* IndexBuilder.cs:159-173 InitStats */
      const float* v0 = rows + p[0] * ld;
      for (int32_t i = 0; i < d; ++i) { mean[i] = v0[i]; q[i] = 0.0f; }
      idn = (i128)ids[p[0]];
      /* IndexBuilder.cs:175-197 UpdateStats, binary32, one divide per (point, dim) */
      for (int64_t j = 1; j < count; ++j)
      {
        const float* v = rows + p[j] * ld;
        float c = (float)(j + 1); /* long -> float conversion of Count+1, IndexBuilder.cs:185-186 */
        for (int32_t i = 0; i < d; ++i)
        {
          float value = v[i];
          float pa = mean[i];
          float pq = q[i];
          float a = pa + (value - pa) / c;
          float qq = pq + (value - pa) * (value - a);
          mean[i] = a;
          q[i] = qq;
        }
        idn += (i128)ids[p[j]];
      }
      /* IndexBuilder.cs:77-79 MaxBy(max ? Stdev2N : -Stdev2N): strictly-greater replacement */
      float best = it.max ? q[0] : -q[0];
      for (int32_t i = 1; i < d; ++i)
      {
        float key = it.max ? q[i] : -q[i];
        if (cmp_float_dotnet(key, best) > 0) { best = key; index = i; }
      }
      mid = mean[index];
      if (sql && q[index] == 0.0f) null_dim = 1; /* Stdev = 0 (DDL.sql:193-194) */
    }

    if (emitted >= cap) { rc = VIO_ERR_CAPACITY; goto done; }
    out_range_id[emitted] = it.range_id;
    if (count == 1)
    {
      /* IndexBuilder.cs:81-82 : leaf, Mid keeps its default 0 */
      out_dim[emitted] = -1;
      out_mid[emitted] = 0.0f;
      out_id[emitted] = (int64_t)idn;
      ++emitted;
      continue; /* IndexBuilder.cs:94-97 */
    }
    int64_t pivot = (int64_t)(idn / (i128)count); /* IndexBuilder.cs:87, truncation toward zero */
    out_dim[emitted] = null_dim ? -3 : index;
    out_mid[emitted] = null_dim ? NAN : mid;
    out_id[emitted] = pivot;
    ++emitted;
    /* dbo.BuildIndex root (DDL.sql:104): `iif(S.Stdev = 0, iif(P.ID <= S.ID, 1, 2), iif(Value < Mean, 1, 2))` -- a
     * value equal to the mean goes high whatever its id (ids are > INT64_MIN) */
    if (sql && depth == 0 && !null_dim) pivot = INT64_MIN;

    /* IndexBuilder.cs:99,104 checked arithmetic */
    if (it.range_id > (INT64_MAX - 2) / 2) { rc = VIO_ERR_OVERFLOW; goto done; }

    /* IndexBuilder.cs:111-124 stable partition */
    int64_t nlo = 0, nhi = 0;
    int64_t* lo = perm + it.start;
    int64_t* hi = tmp + it.start;
    for (int64_t j = 0; j < count; ++j)
    {
      int64_t r = p[j];
      float value = rows[r * ld + index];
      int64_t id = ids[r];
      if (value > mid || (value == mid && id > pivot)) hi[nhi++] = r;
      else lo[nlo++] = r; /* in place: nlo <= j */
    }
    memcpy(perm + it.start + nlo, hi, (size_t)nhi * sizeof(int64_t));

    work_item wlo = {it.range_id * 2 + 1, it.start, nlo, !it.max};
    work_item whi = {it.range_id * 2 + 2, it.start + nlo, nhi, !it.max};
    if (!push(&st, wlo) || !push(&st, whi)) { rc = VIO_ERR_NOMEM; goto done; } /* IndexBuilder.cs:128-129 */
  }
  *out_count = emitted;

done:
  free(perm); free(tmp); free(mean); free(q); free(s1); free(s2); free(st.items);
  return rc;
}

/* ---- dbo.Search over IndexBuilder rows (DDL.sql:246-295) ------------------------
 * Table is given sorted by range_id (ascending) so that a row is found by binary search, mirroring the
 * clustered-index seek on TextIndex.RangeID.  A child row may be absent (empty range => no row).
 * Leaf <=> Dimension == -1, its TextID is Id; internal rows have TextID = null (DDL.sql:195-197).
 */
static int64_t find_row(const int64_t* range_id, int64_t nrows, int64_t key)
{
  int64_t lo = 0, hi = nrows - 1;
  while (lo <= hi)
  {
    int64_t m = lo + ((hi - lo) >> 1);
    if (range_id[m] == key) return m;
    if (range_id[m] < key) lo = m + 1; else hi = m - 1;
  }
  return -1;
}

/* One query. Writes up to cap ids (in DFS order, low branch first) and returns the total found in *out_n,
 * and the number of rows visited in *out_visits. */
int vio_search(int64_t nrows, const int64_t* range_id, const int32_t* dim, const float* mid,
               const int64_t* id, int32_t d, const float* query, float proximity, int64_t cap,
               int64_t* out_ids, int64_t* out_n, int64_t* out_visits)
{
  *out_n = 0;
  if (out_visits) *out_visits = 0;
  if (nrows == 0) return VIO_OK;
  work_stack st = {0, 0, 0};
  work_item root = {0, 0, 0, 0};
  if (!push(&st, root)) return VIO_ERR_NOMEM;
  int64_t found = 0, visits = 0;
  while (st.size > 0)
  {
    int64_t r = st.items[--st.size].range_id;
    int64_t row = find_row(range_id, nrows, r);
    if (row < 0) continue; /* join finds no I row */
    ++visits;
    int32_t k = dim[row];
    if (k < 0 && k != -3)
    {
      if (found < cap) out_ids[found] = id[row];
      ++found;
      continue;
    }
    if (k >= d) { free(st.items); return VIO_ERR_ARG; }
    if (r > (INT64_MAX - 2) / 2) continue;
    work_item c = {0, 0, 0, 0};
    if (k == -3)
    {
      /* a dbo.BuildIndex row with Dimension = null (mode 2): `N.Dimension is null or ...` follows both children
       * (DDL.sql:275,290) */
      c.range_id = 2 * r + 2; if (!push(&st, c)) { free(st.items); return VIO_ERR_NOMEM; }
      c.range_id = 2 * r + 1; if (!push(&st, c)) { free(st.items); return VIO_ERR_NOMEM; }
      continue;
    }
    float lo = query[k] - proximity; /* DDL.sql:249-250, SQL `real` arithmetic */
    float hi = query[k] + proximity;
    /* push high first so that low is visited first */
    if (mid[row] <= hi) { c.range_id = 2 * r + 2; if (!push(&st, c)) { free(st.items); return VIO_ERR_NOMEM; } } /* DDL.sql:280-293 */
    if (mid[row] >= lo) { c.range_id = 2 * r + 1; if (!push(&st, c)) { free(st.items); return VIO_ERR_NOMEM; } } /* DDL.sql:265-278 */
  }
  free(st.items);
  *out_n = found;
  if (out_visits) *out_visits = visits;
  return VIO_OK;
}

/* Batched form: CSR result. offsets[nq+1]; ids filled up to cap; *total = sum of counts. */
int vio_search_batch(int64_t nrows, const int64_t* range_id, const int32_t* dim, const float* mid,
                     const int64_t* id, int32_t d, int64_t nq, const float* queries, int64_t ldq,
                     float proximity, int64_t cap, int64_t* offsets, int64_t* out_ids, int64_t* total,
                     int64_t* total_visits)
{
  int64_t pos = 0, visits = 0;
  offsets[0] = 0;
  for (int64_t i = 0; i < nq; ++i)
  {
    int64_t n = 0, v = 0;
    int64_t room = cap > pos ? cap - pos : 0;
    int rc = vio_search(nrows, range_id, dim, mid, id, d, queries + i * ldq, proximity, room,
                        out_ids ? out_ids + (pos < cap ? pos : cap) : 0, &n, &v);
    if (rc != VIO_OK) return rc;
    pos += n;
    visits += v;
    offsets[i + 1] = pos;
  }
  *total = pos;
  if (total_visits) *total_visits = visits;
  return VIO_OK;
}

/* Verification contract of Find(vector, distance, predicate) (MemoryVectorIndex.cs:237-245): the index
 * returns candidates, the predicate decides.  Euclidean distance as in the reference's test helper
 * (MemoryVectorIndexTests.cs:209-217): float32 accumulation in index order, then sqrt. */
float vio_distance_l2(const float* a, const float* b, int32_t d)
{
  float s = 0.0f;
  for (int32_t i = 0; i < d; ++i)
  {
    float t = a[i] - b[i];
    s = s + t * t;
  }
  return sqrtf(s);
}

/* angular distance 1 - a.b / (|a| |b|), float32 accumulation in index order (the top-k layer's second metric) */
float vio_distance_angular(const float* a, const float* b, int32_t d)
{
  float dot = 0.0f, na = 0.0f, nb = 0.0f;
  for (int32_t i = 0; i < d; ++i)
  {
    dot = dot + a[i] * b[i];
    na = na + a[i] * a[i];
    nb = nb + b[i] * b[i];
  }
  return 1.0f - dot / (sqrtf(na) * sqrtf(nb));
}
