// IndexBuilder.Gpu.cs -- drop-in body for NesterovskyBros.VectorIndex.IndexBuilder.Build (IndexBuilder.cs:23-25):
// same signature, same (rangeId, RangeValue) stream, the work done by libvi_b200 on a B200.  SOURCE ONLY (no .NET
// SDK in the image).  `storeFactory` is accepted for signature compatibility: child ranges never leave the device
// (the reference's FileRangeStore arenas become the device permutation ping-pong), so it is not called.
using System;
using System.Collections.Generic;
using System.Runtime.InteropServices;
using NesterovskyBros.VectorIndex.Native;

namespace NesterovskyBros.VectorIndex;

public partial class IndexBuilder
{
  public static int Device { get; set; } = 0;
  /// <summary>0 = exact (bit-identical to the managed builder), 1 = fast (order-independent integer sums).</summary>
  public static int Mode { get; set; } = NativeMethods.VI_MODE_EXACT;

  public static async IAsyncEnumerable<(long rangeId, RangeValue range)> Build(
    IAsyncEnumerable<(long id, Memory<float> vector)> points,
    Func<long, long, IRangeStore> storeFactory)
  {
    NativeMethods.Check(IntPtr.Zero, NativeMethods.vi_create(Device, out var ctx));
    try
    {
      const int Batch = 65536;
      long[]? ids = null; float[]? rows = null; int dims = 0, fill = 0;
      await foreach (var (id, vector) in points)                   // the input is enumerated ONCE (the managed
      {                                                            // builder enumerates it twice, :57 and :111)
        if (rows == null)
        {
          dims = vector.Length;
          ids = new long[Batch]; rows = new float[(long)Batch * dims];
          NativeMethods.Check(ctx, NativeMethods.vi_points_reserve(ctx, Batch, dims));
        }
        if (vector.Length != dims) throw new ArgumentException("Invalid length of vector.", nameof(points));
        ids![fill] = id;
        vector.Span.CopyTo(rows.AsSpan(fill * dims, dims));
        if (++fill == Batch) { Flush(ctx, ids, rows, fill, dims); fill = 0; }
      }
      if (rows == null) yield break;                               // IndexBuilder.cs:70-73
      if (fill > 0) Flush(ctx, ids!, rows, fill, dims);

      NativeMethods.Check(ctx, NativeMethods.vi_build(ctx, Mode, out _));
      var k = NativeMethods.vi_range_count(ctx);
      var rid = new long[k]; var dim = new int[k]; var mid = new float[k]; var oid = new long[k];
      unsafe
      {
        fixed (long* pr = rid) fixed (int* pd = dim) fixed (float* pm = mid) fixed (long* po = oid)
          NativeMethods.Check(ctx, NativeMethods.vi_ranges_copy(ctx, pr, pd, pm, po, k));
      }
      for (var i = 0L; i < k; ++i)
        yield return (rid[i], new RangeValue { Dimension = dim[i], Mid = mid[i], Id = oid[i] });
    }
    finally
    {
      NativeMethods.vi_destroy(ctx);
    }

    static unsafe void Flush(IntPtr ctx, long[] ids, float[] rows, int n, int dims)
    {
      fixed (long* pi = ids) fixed (float* pr = rows)
        NativeMethods.Check(ctx, NativeMethods.vi_points_add(ctx, pi, pr, n, dims));
    }
  }
}

/// <summary>Search entry point with the shape of MemoryVectorIndex&lt;R&gt;.Find (MemoryVectorIndex.cs:242-245) over
/// the dbo.Search traversal (DDL.sql:234-295).</summary>
public sealed class GpuVectorIndex : IDisposable
{
  private readonly IntPtr ctx;
  private readonly int dims;
  internal GpuVectorIndex(IntPtr ctx, int dims) { this.ctx = ctx; this.dims = dims; }

  public IEnumerable<long> Find(ReadOnlyMemory<float> vector, float distance, Func<long, bool>? predicate = null)
  {
    if (vector.Length != dims) throw new ArgumentException("Invalid vector size.", nameof(vector));
    long total; var offsets = new long[2];
    unsafe
    {
      fixed (float* q = vector.Span) fixed (long* po = offsets)
        NativeMethods.Check(ctx, predicate == null
          ? NativeMethods.vi_search_verify(ctx, q, 1, dims, distance, distance, po, null, 0, out total)
          : NativeMethods.vi_search(ctx, q, 1, dims, distance, po, null, 0, out total));
    }
    var ids = new long[Math.Max(total, 1)];
    unsafe
    {
      fixed (float* q = vector.Span) fixed (long* po = offsets) fixed (long* pi = ids)
        NativeMethods.Check(ctx, predicate == null
          ? NativeMethods.vi_search_verify(ctx, q, 1, dims, distance, distance, po, pi, total, out total)
          : NativeMethods.vi_search(ctx, q, 1, dims, distance, po, pi, total, out total));
    }
    for (var i = 0L; i < total; ++i)
      if (predicate == null || predicate(ids[i])) yield return ids[i];
  }

  public void Dispose() => NativeMethods.vi_destroy(ctx);
}
