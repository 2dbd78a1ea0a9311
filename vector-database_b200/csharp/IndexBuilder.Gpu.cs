// IndexBuilder.Gpu.cs -- drop-in body for NesterovskyBros.VectorIndex.IndexBuilder.Build (IndexBuilder.cs:23-25):
// same signature, same (rangeId, RangeValue) stream, the work done by libvi_b200 on a B200.  SOURCE ONLY (no .NET
// SDK in the image).  `storeFactory` is accepted for signature compatibility: child ranges never leave the device
// (the reference's FileRangeStore arenas become the device permutation ping-pong), so it is not called.
// Pinned (`fixed`) calls live in plain helper methods: C# 12 and earlier refuse unsafe code inside iterators (CS1629).
using System;
using System.Collections.Generic;
using System.Runtime.InteropServices;
using System.Threading.Tasks;
using NesterovskyBros.VectorIndex.Native;

namespace NesterovskyBros.VectorIndex;

public partial class IndexBuilder
{
  public static int Device { get; set; } = 0;
  /// <summary>0 = exact (bit-identical to the managed builder), 1 = fast (order-independent integer sums),
  /// 2 = the T-SQL builder's rules (dbo.BuildIndex, DDL.sql:44-202).</summary>
  public static int Mode { get; set; } = NativeMethods.VI_MODE_EXACT;

  public static async IAsyncEnumerable<(long rangeId, RangeValue range)> Build(
    IAsyncEnumerable<(long id, Memory<float> vector)> points,
    Func<long, long, IRangeStore> storeFactory)
  {
    var index = await GpuVectorIndex.Create(points, Device, Mode);
    try
    {
      if (index == null) yield break;                              // IndexBuilder.cs:70-73: no points, no rows
      var (rid, dim, mid, oid) = index.Ranges();
      for (var i = 0L; i < rid.LongLength; ++i)
        yield return (rid[i], new RangeValue { Dimension = dim[i], Mid = mid[i], Id = oid[i] });
    }
    finally
    {
      index?.Dispose();
    }
  }
}

/// <summary>A built index that stays on the device: the (rangeId, RangeValue) rows and the search entry point with the
/// shape of MemoryVectorIndex&lt;R&gt;.Find (MemoryVectorIndex.cs:242-245) over the dbo.Search traversal
/// (DDL.sql:234-295).</summary>
public sealed class GpuVectorIndex : IDisposable
{
  private IntPtr ctx;
  private readonly int dims;
  private GpuVectorIndex(IntPtr ctx, int dims) { this.ctx = ctx; this.dims = dims; }

  public int Dimensions => dims;
  public long RangeCount => NativeMethods.vi_range_count(ctx);

  /// <summary>Drains `points` ONCE (the managed builder enumerates its input twice, IndexBuilder.cs:57,111) into
  /// 65 536-point batches -> vi_points_add, then vi_build.  Returns null for an empty input.</summary>
  public static async Task<GpuVectorIndex?> Create(
    IAsyncEnumerable<(long id, Memory<float> vector)> points, int device = 0, int mode = NativeMethods.VI_MODE_EXACT)
  {
    NativeMethods.Check(IntPtr.Zero, NativeMethods.vi_create(device, out var ctx));
    try
    {
      const int Batch = 65536;
      long[]? ids = null; float[]? rows = null; int dims = 0, fill = 0;
      await foreach (var (id, vector) in points)
      {
        if (rows == null)
        {
          dims = vector.Length;
          ids = new long[Batch]; rows = new float[(long)Batch * dims];
          NativeMethods.Check(ctx, NativeMethods.vi_points_reserve(ctx, Batch, dims));
        }
        if (vector.Length != dims) throw new ArgumentException("Invalid length of vector.", nameof(points));
        ids![fill] = id;
        vector.Span.CopyTo(rows.AsSpan(fill * dims, dims));
        if (++fill == Batch) { Add(ctx, ids, rows, fill, dims); fill = 0; }
      }
      if (rows == null) { NativeMethods.vi_destroy(ctx); return null; }
      if (fill > 0) Add(ctx, ids!, rows, fill, dims);
      NativeMethods.Check(ctx, NativeMethods.vi_build(ctx, mode, out _));
      return new GpuVectorIndex(ctx, dims);
    }
    catch
    {
      NativeMethods.vi_destroy(ctx);
      throw;
    }
  }

  /// <summary>The consumer's Dictionary&lt;long, RangeValue&gt; (Program.cs:18-26) or its CSV (Program.cs:145-149)
  /// back into a searchable index: vi_ranges_load.</summary>
  public static GpuVectorIndex Load(long[] rangeId, int[] dimension, float[] mid, long[] id, int dims, int device = 0)
  {
    NativeMethods.Check(IntPtr.Zero, NativeMethods.vi_create(device, out var ctx));
    try
    {
      LoadRows(ctx, rangeId, dimension, mid, id, dims);
      return new GpuVectorIndex(ctx, dims);
    }
    catch
    {
      NativeMethods.vi_destroy(ctx);
      throw;
    }
  }

  /// <summary>One rank of a multi-GPU build (one process per GPU): `uniqueId` comes from rank 0's
  /// vi_comm_unique_id; every rank adds its shard (rank order = data order) and calls this collectively.</summary>
  public static unsafe GpuVectorIndex CreateSharded(long[] ids, float[] rows, int dims, byte[] uniqueId, int rank, int world,
                                                    int device, bool replicateForSearch = true)
  {
    NativeMethods.Check(IntPtr.Zero, NativeMethods.vi_create(device, out var ctx));
    try
    {
      fixed (byte* u = uniqueId) NativeMethods.Check(ctx, NativeMethods.vi_comm_init(ctx, u, uniqueId.Length, rank, world));
      NativeMethods.Check(ctx, NativeMethods.vi_points_reserve(ctx, ids.LongLength, dims));
      Add(ctx, ids, rows, ids.Length, dims);
      NativeMethods.Check(ctx, NativeMethods.vi_build(ctx, NativeMethods.VI_MODE_FAST, out _));
      if (replicateForSearch) NativeMethods.Check(ctx, NativeMethods.vi_table_replicate(ctx));
      return new GpuVectorIndex(ctx, dims);
    }
    catch
    {
      NativeMethods.vi_destroy(ctx);
      throw;
    }
  }

  public (long[] rangeId, int[] dimension, float[] mid, long[] id) Ranges()
  {
    var k = NativeMethods.vi_range_count(ctx);
    var rid = new long[k]; var dim = new int[k]; var mid = new float[k]; var oid = new long[k];
    CopyRanges(ctx, rid, dim, mid, oid);
    return (rid, dim, mid, oid);
  }

  /// <summary>Candidates of dbo.Search for `vector` +- `distance` per coordinate; with no predicate the library's own
  /// Euclidean check (MemoryVectorIndexTests.cs:209-217) filters them, otherwise the caller's predicate decides, as in
  /// MemoryVectorIndex.Find (MemoryVectorIndex.cs:336-342).</summary>
  public IEnumerable<long> Find(ReadOnlyMemory<float> vector, float distance, Func<long, bool>? predicate = null)
  {
    if (vector.Length != dims) throw new ArgumentException("Invalid vector size.", nameof(vector));
    var ids = Search(ctx, vector.Span, dims, distance, verify: predicate == null);
    foreach (var id in ids)
      if (predicate == null || predicate(id)) yield return id;
  }

  public void Dispose()
  {
    if (ctx != IntPtr.Zero) NativeMethods.vi_destroy(ctx);
    ctx = IntPtr.Zero;
  }

  // ---- pinned calls (not iterators) ----------------------------------------------------------------------------------
  private static unsafe void Add(IntPtr ctx, long[] ids, float[] rows, int n, int dims)
  {
    fixed (long* pi = ids) fixed (float* pr = rows)
      NativeMethods.Check(ctx, NativeMethods.vi_points_add(ctx, pi, pr, n, dims));
  }

  private static unsafe void CopyRanges(IntPtr ctx, long[] rid, int[] dim, float[] mid, long[] oid)
  {
    fixed (long* pr = rid) fixed (int* pd = dim) fixed (float* pm = mid) fixed (long* po = oid)
      NativeMethods.Check(ctx, NativeMethods.vi_ranges_copy(ctx, pr, pd, pm, po, rid.LongLength));
  }

  private static unsafe void LoadRows(IntPtr ctx, long[] rid, int[] dim, float[] mid, long[] oid, int dims)
  {
    fixed (long* pr = rid) fixed (int* pd = dim) fixed (float* pm = mid) fixed (long* po = oid)
      NativeMethods.Check(ctx, NativeMethods.vi_ranges_load(ctx, pr, pd, pm, po, rid.LongLength, dims));
  }

  private static unsafe long[] Search(IntPtr ctx, ReadOnlySpan<float> q, int dims, float distance, bool verify)
  {
    var offsets = stackalloc long[2];
    fixed (float* pq = q)
    {
      long total;
      if (verify)
      {
        NativeMethods.Check(ctx, NativeMethods.vi_search_verify(ctx, pq, 1, dims, distance, distance, offsets, null, 0, out total));
        var ids = new long[Math.Max(total, 1)];
        fixed (long* pi = ids)
          NativeMethods.Check(ctx, NativeMethods.vi_search_verify(ctx, pq, 1, dims, distance, distance, offsets, pi, total, out total));
        return ids.AsSpan(0, (int)total).ToArray();
      }
      // one walk: begin counts and keeps the candidates on the device, fetch copies them
      NativeMethods.Check(ctx, NativeMethods.vi_search_begin(ctx, pq, 1, dims, distance, out total));
      var cand = new long[Math.Max(total, 1)];
      fixed (long* pi = cand)
        NativeMethods.Check(ctx, NativeMethods.vi_search_fetch(ctx, offsets, pi, cand.LongLength));
      return cand.AsSpan(0, (int)total).ToArray();
    }
  }
}
