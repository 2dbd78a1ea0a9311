/*
 * vi_b200.h -- C ABI of the B200-native split-tree vector index (libvi_b200.so).
 *
 * This is the drop-in boundary for ONE hot path of nesterovsky-bros/vector-database: IndexBuilder.Build and the
 * RangeID-0 -> Low/High traversal search.  The reference is 100 % managed C# + T-SQL and has no FFI of its own;
 * each entry point below names the reference interface it replaces (paths relative to the reference root) and is
 * what a P/Invoke layer would bind.  Signatures use only plain pointers, sizes and int status
 * codes: no torch types, no exceptions, no callbacks on the hot path.
 *
 * Threading: one vi_ctx = one host thread ("instances of this class are not thread safe", MemoryRangeStore.cs:5).
 * There is NO CPU fallback: every entry point that computes needs a CUDA device and fails with VI_ERR_CUDA when
 * none is usable.
 */
#ifndef VI_B200_H
#define VI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VI_ABI_VERSION 1

/* Status codes.  The C# shim maps them back onto the reference's exception types. */
enum
{
  VI_OK = 0,
  VI_ERR_INVALID_ARG = 1,    /* ArgumentException("Invalid length of vector.") FileRangeStore.cs:59-64;
                                ArgumentException("Invalid vector size.") MemoryVectorIndex.cs:254 */
  VI_ERR_OVERFLOW = 2,       /* OverflowException from checked(rangeId*2+1|2) IndexBuilder.cs:99,104 (depth > 62) */
  VI_ERR_NOT_IMPLEMENTED = 3,/* NotImplementedException, root RangeStore.Add IndexBuilder.cs:206-209 */
  VI_ERR_STATE = 4,          /* call out of order (e.g. search before build) */
  VI_ERR_CAPACITY = 5,       /* caller buffer or reserved capacity too small */
  VI_ERR_OOM = 6,            /* device or host allocation failed */
  VI_ERR_CUDA = 7            /* CUDA runtime error, text in vi_last_error */
};

/* Statistics modes of vi_build. */
enum
{
  VI_MODE_EXACT = 0, /* literal float32 sequential Welford, IndexBuilder.cs:175-197: bit-identical range table */
  VI_MODE_FAST = 1,  /* qfx: order-independent exact integer sums of 26-bit fixed-point values,
                        HBM-bound, shardable; Mid within 1e-6 * max|x| of the exact mode's */
  VI_MODE_SQL = 2    /* the rules of the T-SQL builder dbo.BuildIndex (DDL.sql:44-202) over the qfx statistics: max stdev
                        at depth 0, MIN at depth 1, max below (DDL.sql:113,151,155); the root sends Value = Mean to the
                        high child (DDL.sql:104); a range whose chosen dimension has Stdev = 0 keeps splitting by ID but
                        its row has Dimension = null, Mid = null (DDL.sql:193-194) -- Dimension -3 / Mid NaN in
                        vi_ranges_copy, null in vi_textindex_copy -- and the search follows both its children
                        (DDL.sql:275,290).  Same kernels, same speed class and sharding as VI_MODE_FAST. */
};
#define VI_DIM_NULL (-3) /* vi_ranges_copy / vi_ranges_load: an internal row with Dimension = null (VI_MODE_SQL) */

typedef struct vi_ctx vi_ctx;

/* Per-build counters (all levels). */
typedef struct vi_build_info
{
  int64_t ranges;        /* rows in the range table */
  int32_t levels;        /* tree levels processed (max depth + 1) */
  int32_t mode;
  int64_t point_visits;  /* sum over levels of points in non-leaf ranges (A_l) */
  int64_t kernel_launches;
  double build_ms;       /* device time of the whole build, CUDA events */
  int32_t q_exponent;    /* fast mode: E with max|x| < 2^E */
  int32_t reserved;
  double subtree_ms;     /* fast mode: device time of the sub-tree kernel (ranges of <= 32 points, vi_subtree.cuh) */
  int64_t subtree_ranges;/* ranges handed to it */
  int32_t shared_retry;  /* multi-rank: 1 if the build was redone with fewer shared levels (a tightly clustered top range) */
  int32_t reserved2;
} vi_build_info;

/* Per-level record (profiling / roofline accounting, Program.cs has only a whole-build Stopwatch). */
typedef struct vi_level_info
{
  int32_t level;
  int32_t derived_points; /* of `points`: in big ranges whose integer sums were derived as parent - sibling (fast mode),
                             i.e. whose rows the statistics pass did not read */
  int64_t ranges;        /* non-leaf ranges processed at this level */
  int64_t points;        /* points in them (A_l) */
  int64_t rows_emitted;  /* table rows created by this level's partition pass */
  double stats_ms, partition_ms; /* level-synchronous passes only; sub-tree kernel time is vi_build_info.subtree_ms */
  int64_t in_subtrees;   /* of `points`, those processed inside the sub-tree kernel */
} vi_level_info;

/* ---- lifetime ------------------------------------------------------------------------------------------ */
int vi_abi_version(void);
/* device: CUDA ordinal.  Replaces nothing in the reference (it has no device); corresponds to constructing
 * the store factory, `new FileRangeStore(count, dimensions)` FileRangeStore.cs:18-27. */
int vi_create(int32_t device, vi_ctx** out);
void vi_destroy(vi_ctx* ctx); /* FileRangeStore.Dispose FileRangeStore.cs:32 */
const char* vi_last_error(const vi_ctx* ctx);

/* ---- ingest: the batched IRangeStore.Add (IRangeStore.cs:15, FileRangeStore.cs:57-75) ------------------------ */
/* Fixes the dimension count (FileRangeStore ctor arg `dimensions`) and reserves device memory for `capacity`
 * points (ctor arg `count`). Drops previously added points and any built table. */
int vi_points_reserve(vi_ctx* ctx, int64_t capacity, int32_t dims);
/* Appends n points from HOST memory: ids[n], rows row-major n x dims float32.  The library copies (as
 * FileRangeStore copies into its mapping, FileRangeStore.cs:140-151), so the caller's buffers are reusable on
 * return.  Grows the reservation if needed.  `dims` must equal the reserved dimension count, otherwise
 * VI_ERR_INVALID_ARG ("Invalid length of vector."). */
int vi_points_add(vi_ctx* ctx, const int64_t* ids, const float* rows, int64_t n, int32_t dims);
/* Same, from DEVICE memory on the ctx's device (device-to-device copy). */
int vi_points_add_device(vi_ctx* ctx, const int64_t* d_ids, const float* d_rows, int64_t n, int32_t dims);
/* The FileRangeStore record (FileRangeStore.cs:127-165): n records [int64 id][dims x float32], little endian, no
 * padding.  From a host buffer, or streamed from a file (n < 0: to the end of the file) through two pinned buffers:
 * the read of one batch overlaps the H2D copy and the de-interleave kernel of the one before.  read_ms / total_ms
 * (either may be NULL) report the time inside the file reads and the whole call. */
int vi_points_add_records(vi_ctx* ctx, const void* records, int64_t n, int32_t dims);
int vi_points_add_file(vi_ctx* ctx, const char* path, int64_t offset_bytes, int64_t n, int32_t dims, double* read_ms,
                       double* total_ms);
/* The reference's test program takes its real data set from an ANN-benchmarks HDF5 file: GetHdf5DatasetSize reads the
 * 2-D extent of `/train`, GetHdf5Dataset yields (row index, float[dimension]) in blocks of 100000 rows
 * (VectorIndex.MainTest/Program.cs:183-260, through HDF5-CSharp = libhdf5).  A minimal native reader of the container
 * (csrc/vi_hdf5.cu: superblock 0-3, symbol-table or compact-link groups, object headers 1 / 2, CONTIGUOUS unfiltered
 * little-endian data sets of rank 1 or 2; anything else is refused with a message):
 *   vi_hdf5_dataset_info  extent and element type of `dataset` ("/train", "test", "a/b"), and where its bytes start;
 *                         rank other than 1 or 2 -> VI_ERR_INVALID_ARG "Invalid rank." (Program.cs:203-206); ctx may be
 *                         NULL (no error text then);
 *   vi_hdf5_read_rows     rows [first_row, first_row + n) into a host buffer (the `/test` queries, `/neighbors`);
 *   vi_points_add_hdf5    appends float32 rows [first_row, first_row + n) (n < 0: to the end) with ids first_id,
 *                         first_id + 1, ... (Program.cs:252 yields the row index) through the pinned double-buffered
 *                         pipeline of vi_points_add_file: reading one batch overlaps the H2D copy of the one before. */
int vi_hdf5_dataset_info(vi_ctx* ctx, const char* path, const char* dataset, int64_t* rows, int64_t* cols,
                         int32_t* elem_class /* 0 integer, 1 IEEE float */, int32_t* elem_bytes, int64_t* data_offset);
const char* vi_hdf5_last_error(void); /* text of the calling thread's last failed vi_hdf5_* call made with ctx == NULL */
int vi_hdf5_read_rows(vi_ctx* ctx, const char* path, const char* dataset, int64_t first_row, int64_t n, void* out,
                      int64_t out_bytes);
int vi_points_add_hdf5(vi_ctx* ctx, const char* path, const char* dataset, int64_t first_row, int64_t n, int64_t first_id,
                       double* read_ms, double* total_ms);
int64_t vi_points_count(const vi_ctx* ctx);

/* ---- build: IndexBuilder.Build (IndexBuilder.cs:23-157) ------------------------------------------------------ */
/* Level-synchronous GPU build over every open range at once.  info may be NULL. */
int vi_build(vi_ctx* ctx, int32_t mode, vi_build_info* info);
/* vi_build followed by vi_ranges_copy, as IndexBuilder.Build's consumer sees them (rows only after the build), with the
 * device-to-host copy of the table overlapped with the build's last kernel: that kernel finishes the sub-trees of <= 32
 * points in up to eight slices whose row blocks are dense and contiguous, and each finished block is copied out (a second
 * stream) while the next slice runs.  Host buffers should be pinned for the overlap to be real; any of them may be NULL.
 * *rows = rows of the table; VI_ERR_CAPACITY if cap is smaller (the build itself is complete and vi_ranges_copy works). */
int vi_build_copy(vi_ctx* ctx, int32_t mode, vi_build_info* info, int64_t* range_id, int32_t* dimension, float* mid,
                  int64_t* id, int64_t cap, int64_t* rows);
/* Copies up to cap per-level records of the last build; returns the number of levels in *n. */
int vi_build_levels(const vi_ctx* ctx, vi_level_info* out, int32_t cap, int32_t* n);

/* ---- range table out: the (rangeId, RangeValue) stream of Build (IndexBuilder.cs:92, RangeValue.cs) ---------- */
int64_t vi_range_count(const vi_ctx* ctx);
/* One row per non-empty range, keyed by RangeID (row order is unspecified: the level-synchronous part comes
 * breadth-first, sub-trees finished by the sub-tree kernel come as dense depth-first blocks; the reference's
 * consumers key rows by rangeId, Program.cs:18-26).  Any output pointer may be NULL.
 * dimension == -1 marks a leaf (RangeValue.Dimension), id is the leaf's point id there and the tie-break pivot
 * elsewhere (RangeValue.Id). */
int vi_ranges_copy(const vi_ctx* ctx, int64_t* range_id, int32_t* dimension, float* mid, int64_t* id, int64_t cap);
/* The inverse of vi_ranges_copy: rows (RangeID, Dimension, Mid, Id) in any order -- the reference consumer's
 * Dictionary<long, RangeValue> (Program.cs:18-26) or its CSV "RangeID,Dimension,Mid,ID" (Program.cs:80,145-149) --
 * become the context's searchable table (children linked by 2r+1 / 2r+2, an absent child stays absent).  The
 * context then answers vi_search / vi_search_device; vi_search_verify needs the vectors and is refused.
 * VI_ERR_INVALID_ARG: duplicate or negative RangeID, no row for RangeID 0, Dimension outside [-1, dims). */
int vi_ranges_load(vi_ctx* ctx, const int64_t* range_id, const int32_t* dimension, const float* mid, const int64_t* id,
                   int64_t n, int32_t dims);
/* dbo.TextIndex form (DDL.sql:209-227): child RangeIDs or -1 for null, TextID = id for leaves and -1 (null)
 * for internal rows (DDL.sql:195-197), Dimension -1 / Mid NaN stand for null. */
int vi_textindex_copy(const vi_ctx* ctx, int64_t* range_id, int16_t* dimension, float* mid, int64_t* low_range_id,
                      int64_t* high_range_id, int64_t* text_id, int64_t cap);

/* ---- search: dbo.Search (DDL.sql:234-295) behind the Find(vector, distance, predicate) shape
 *      (MemoryVectorIndex.cs:242-245) ---------------------------------------------------------------------------- */
/* Batched traversal of nq HOST queries (row-major nq x dims) with one proximity.  CSR result: offsets[nq+1] and
 * ids (candidate TextIDs per query, DFS order, low branch first).  Two-call protocol: call with ids == NULL (or
 * cap too small) to get *total and offsets, then with cap >= *total.  Returns VI_ERR_CAPACITY when ids != NULL
 * and cap < *total (offsets and *total are still valid). */
int vi_search(vi_ctx* ctx, const float* queries, int64_t nq, int32_t dims, float proximity, int64_t* offsets,
              int64_t* ids, int64_t cap, int64_t* total);
/* The same search in two steps: begin stages the queries and counts (-> *total), fetch copies offsets[nq+1] and
 * ids[*total] (cap >= *total).  Batches whose queries visit many rows (proximity > 0) are walked ONCE, a warp per query:
 * begin keeps the candidates in a device pool and fetch only gathers them; point-lookup batches are walked one thread
 * per query, once to count and once to fill.  A build, reserve or any other search call in between invalidates the
 * pending search (fetch -> VI_ERR_STATE). */
int vi_search_begin(vi_ctx* ctx, const float* queries, int64_t nq, int32_t dims, float proximity, int64_t* total);
int vi_search_fetch(vi_ctx* ctx, int64_t* offsets, int64_t* ids, int64_t cap);
/* Device-resident form: d_queries, d_offsets[nq+1], d_ids[cap] are device pointers; *total is a host value.
 * visits (host, may be NULL) receives the number of table rows visited. */
int vi_search_device(vi_ctx* ctx, const float* d_queries, int64_t nq, int32_t dims, float proximity,
                     int64_t* d_offsets, int64_t* d_ids, int64_t cap, int64_t* total, int64_t* visits);
/* Candidate verification, the `predicate` half of the Find contract (MemoryVectorIndex.cs:237-241, 336-342):
 * keeps the candidates whose Euclidean distance to the query is <= distance (float32 sum in index order then
 * sqrt, as MemoryVectorIndexTests.cs:209-217).  Same CSR two-call protocol as vi_search. */
int vi_search_verify(vi_ctx* ctx, const float* queries, int64_t nq, int32_t dims, float proximity, float distance,
                     int64_t* offsets, int64_t* ids, int64_t cap, int64_t* total);
/* Top-k over the candidates of the traversal (the quality layer README.md:102-103 aims at): every candidate's
 * distance to its query -- float32 accumulation in index order like MemoryVectorIndexTests.cs:209-217 -- and the k
 * nearest per query, ties in traversal order.  metric 0: Euclidean, 1: angular (1 - cosine).  ids / dist are
 * [nq][k] (unused slots: -1 / +inf), count[q] = min(k, candidates of q); any of them may be NULL.
 * Recall against exact k-NN is governed by proximity: the traversal returns the points whose every coordinate is
 * within proximity of the query's, and more. */
int vi_search_topk(vi_ctx* ctx, const float* queries, int64_t nq, int32_t dims, float proximity, int32_t k, int32_t metric,
                   int64_t* ids, float* dist, int32_t* count, int64_t* candidates);

/* ---- multi-GPU (one process per GPU) ------------------------------------------------------------------------------ */
/* A multi-rank vi_build treats the points added to the `world` contexts as ONE data set in rank order (rank 0's points
 * first).  VI_MODE_EXACT (the chains of the literal recurrence run over the global order and cannot be split): every rank
 * receives all points (one all-gather), builds the levels above ceil(log2 world) redundantly -- same kernels, same data,
 * same bits, no communication -- and finishes only the sub-trees of the ranges it owns; the contexts end up in the same
 * shape as below.  VI_MODE_FAST / VI_MODE_SQL (integer sums, order-independent):  Top levels: every rank reduces its local
 * slice of every range and the sums meet in one all-reduce per level; then each range of level L = ceil(log2 world)+1
 * moves to one owner rank (a single all-to-all) and the owners finish their sub-trees without communication.  After
 * the build a context holds the rows of the shared top levels (replicated) plus the rows of the sub-trees it owns;
 * the union over ranks is the single-rank table.  A top range too tightly clustered for the integer statistics makes
 * the build start over with fewer shared levels (vi_build_info.shared_retry), down to handing everything to one rank.
 *
 * The library owns the collectives: NCCL (libnccl.so.2, bound at run time) on the context's own stream, so the shared
 * levels are enqueued without host synchronisation.  Rank 0 calls vi_comm_unique_id, the host passes the 128 bytes to
 * every rank by whatever means it has, every rank calls vi_comm_init (collective).  world == 1 leaves multi-rank mode. */
int vi_comm_unique_id(void* out, int32_t bytes /* >= 128 */);
int vi_comm_init(vi_ctx* ctx, const void* unique_id, int32_t bytes, int32_t rank, int32_t world);
/* calls / bytes sent so far by [all-reduce, all-to-all, all-gather] (either array may be NULL) */
int vi_comm_stats(const vi_ctx* ctx, int64_t* calls3, int64_t* bytes3);
/* Alternative for hosts that bring their own transport: the two collectives as host callbacks on DEVICE buffers
 * (they must have completed when they return; the library synchronises its stream before calling them):
 *   allreduce: sum of `count` uint64 words, in place, over all ranks;
 *   alltoallv: rank r's send_bytes[d] bytes (consecutive in d_send) go to rank d, which receives recv_bytes[r]
 *              bytes from rank r (consecutive in d_recv, ordered by source rank).
 * Both return 0 on success. */
typedef int (*vi_allreduce_u64_fn)(void* user, void* d_buf, int64_t count);
typedef int (*vi_alltoallv_fn)(void* user, const void* d_send, const int64_t* send_bytes, void* d_recv,
                               const int64_t* recv_bytes);
int vi_set_collective(vi_ctx* ctx, int32_t rank, int32_t world, vi_allreduce_u64_fn allreduce,
                      vi_alltoallv_fn alltoallv, void* user);
/* Rows [0, *shared_rows) of this context's table were numbered while the ranges were shared: the rows of levels < L
 * are identical on every rank; a level-L range's row is filled in by the rank that owns the range and is a placeholder
 * with Dimension == -2 on the others.  Rows from *shared_rows on belong to the sub-trees this rank owns. */
int vi_shared_rows(const vi_ctx* ctx, int64_t* shared_rows);
/* After a multi-rank build: gathers every rank's rows so that each context holds the WHOLE range table (row links
 * remapped) and can answer any query -- "search shards the query batch against a replicated range table".  One
 * all-gather of 32 bytes per row (plus a sum-all-reduce over the few shared rows).  Collective: every rank must call it.  The vectors stay with their owners, so
 * vi_search_verify is not available on a replicated table. */
int vi_table_replicate(vi_ctx* ctx);

/* ---- utilities -------------------------------------------------------------------------------------------------- */
/* Raw device pointers of the built table for zero-copy consumers (valid until the next build/reserve/destroy). */
int vi_table_device(const vi_ctx* ctx, const int64_t** range_id, const int32_t** dimension, const float** mid,
                    const int64_t** id, const int32_t** low_row, const int32_t** high_row);
/* CUDA stream (cudaStream_t) all work of this ctx is launched on; for external event timing. */
void* vi_stream(const vi_ctx* ctx);
/* Self-test of the exact mode's division-free recurrence step: compares it with IEEE division on `samples`
 * pseudo-random operand pairs on the device; *mismatches must come back 0. */
int vi_debug_divcheck(vi_ctx* ctx, uint64_t seed, int64_t samples, int64_t* mismatches);

#ifdef __cplusplus
}
#endif
#endif /* VI_B200_H */
