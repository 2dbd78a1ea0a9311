// NativeMethods.cs -- P/Invoke layer over libvi_b200.so (include/vi_b200.h).  SOURCE ONLY: there is no .NET SDK in
// the build image or on the GPU box, so this file is not compiled or tested here (INTEGRATION.md).
using System;
using System.Runtime.InteropServices;

namespace NesterovskyBros.VectorIndex.Native;

internal static partial class NativeMethods
{
  private const string Lib = "vi_b200"; // libvi_b200.so on Linux

  public const int VI_OK = 0, VI_ERR_INVALID_ARG = 1, VI_ERR_OVERFLOW = 2, VI_ERR_NOT_IMPLEMENTED = 3,
    VI_ERR_STATE = 4, VI_ERR_CAPACITY = 5, VI_ERR_OOM = 6, VI_ERR_CUDA = 7;
  public const int VI_MODE_EXACT = 0, VI_MODE_FAST = 1;

  [StructLayout(LayoutKind.Sequential)]
  public struct BuildInfo
  {
    public long ranges; public int levels; public int mode; public long pointVisits; public long kernelLaunches;
    public double buildMs; public int qExponent; public int reserved; public double subtreeMs; public long subtreeRanges;
  }

  [LibraryImport(Lib)] public static partial int vi_abi_version();
  [LibraryImport(Lib)] public static partial int vi_create(int device, out IntPtr ctx);
  [LibraryImport(Lib)] public static partial void vi_destroy(IntPtr ctx);
  [LibraryImport(Lib)] public static partial IntPtr vi_last_error(IntPtr ctx);
  [LibraryImport(Lib)] public static partial int vi_points_reserve(IntPtr ctx, long capacity, int dims);
  [LibraryImport(Lib)] public static unsafe partial int vi_points_add(IntPtr ctx, long* ids, float* rows, long n, int dims);
  [LibraryImport(Lib)] public static partial long vi_points_count(IntPtr ctx);
  [LibraryImport(Lib)] public static partial int vi_build(IntPtr ctx, int mode, out BuildInfo info);
  [LibraryImport(Lib)] public static partial long vi_range_count(IntPtr ctx);
  [LibraryImport(Lib)] public static unsafe partial int vi_ranges_copy(IntPtr ctx, long* rangeId, int* dimension, float* mid, long* id, long cap);
  [LibraryImport(Lib)] public static unsafe partial int vi_ranges_load(IntPtr ctx, long* rangeId, int* dimension, float* mid, long* id, long n, int dims);
  [LibraryImport(Lib)] public static unsafe partial int vi_points_add_records(IntPtr ctx, void* records, long n, int dims);
  [LibraryImport(Lib, StringMarshalling = StringMarshalling.Utf8)] public static partial int vi_points_add_file(IntPtr ctx, string path, long offsetBytes, long n, int dims, out double readMs, out double totalMs);
  [LibraryImport(Lib)] public static unsafe partial int vi_textindex_copy(IntPtr ctx, long* rangeId, short* dimension, float* mid, long* low, long* high, long* textId, long cap);
  [LibraryImport(Lib)] public static unsafe partial int vi_search(IntPtr ctx, float* queries, long nq, int dims, float proximity, long* offsets, long* ids, long cap, out long total);
  [LibraryImport(Lib)] public static unsafe partial int vi_search_topk(IntPtr ctx, float* queries, long nq, int dims, float proximity, int k, int metric, long* ids, float* dist, int* count, out long candidates);
  [LibraryImport(Lib)] public static unsafe partial int vi_search_verify(IntPtr ctx, float* queries, long nq, int dims, float proximity, float distance, long* offsets, long* ids, long cap, out long total);

  /// <summary>Maps a status code back onto the exception the reference would have thrown.</summary>
  public static void Check(IntPtr ctx, int rc)
  {
    if (rc == VI_OK) return;
    var text = Marshal.PtrToStringUTF8(vi_last_error(ctx)) ?? "vi_b200 error";
    throw rc switch
    {
      VI_ERR_INVALID_ARG => new ArgumentException(text),          // FileRangeStore.cs:59-64, MemoryVectorIndex.cs:254
      VI_ERR_OVERFLOW => new OverflowException(text),             // IndexBuilder.cs:99,104 checked(rangeId*2+1)
      VI_ERR_NOT_IMPLEMENTED => new NotImplementedException(text),// IndexBuilder.cs:206-209
      VI_ERR_OOM => new OutOfMemoryException(text),
      _ => new InvalidOperationException(text),
    };
  }
}
