// NativeMethods.cs -- P/Invoke layer over libvi_b200.so (include/vi_b200.h): every entry point the header declares.
// SOURCE ONLY: there is no .NET SDK in the build image or on the GPU box, so this file is not compiled or tested here
// (INTEGRATION.md); the same ABI is exercised through ctypes and C++.
using System;
using System.Runtime.InteropServices;

namespace NesterovskyBros.VectorIndex.Native;

internal static partial class NativeMethods
{
  private const string Lib = "vi_b200"; // libvi_b200.so on Linux

  public const int VI_OK = 0, VI_ERR_INVALID_ARG = 1, VI_ERR_OVERFLOW = 2, VI_ERR_NOT_IMPLEMENTED = 3,
    VI_ERR_STATE = 4, VI_ERR_CAPACITY = 5, VI_ERR_OOM = 6, VI_ERR_CUDA = 7;
  public const int VI_MODE_EXACT = 0, VI_MODE_FAST = 1, VI_MODE_SQL = 2;
  public const int VI_DIM_NULL = -3;   // vi_ranges_copy: an internal row with Dimension = null (VI_MODE_SQL)

  [StructLayout(LayoutKind.Sequential)]
  public struct BuildInfo
  {
    public long ranges; public int levels; public int mode; public long pointVisits; public long kernelLaunches;
    public double buildMs; public int qExponent; public int reserved; public double subtreeMs; public long subtreeRanges;
    public int sharedRetry; public int reserved2;
  }

  [StructLayout(LayoutKind.Sequential)]
  public struct LevelInfo
  {
    public int level; public int derivedPoints; public long ranges; public long points; public long rowsEmitted;
    public double statsMs; public double partitionMs; public long inSubtrees;
  }

  // ---- lifetime
  [LibraryImport(Lib)] public static partial int vi_abi_version();
  [LibraryImport(Lib)] public static partial int vi_create(int device, out IntPtr ctx);
  [LibraryImport(Lib)] public static partial void vi_destroy(IntPtr ctx);
  [LibraryImport(Lib)] public static partial IntPtr vi_last_error(IntPtr ctx);
  // ---- ingest
  [LibraryImport(Lib)] public static partial int vi_points_reserve(IntPtr ctx, long capacity, int dims);
  [LibraryImport(Lib)] public static unsafe partial int vi_points_add(IntPtr ctx, long* ids, float* rows, long n, int dims);
  [LibraryImport(Lib)] public static partial int vi_points_add_device(IntPtr ctx, IntPtr dIds, IntPtr dRows, long n, int dims);
  [LibraryImport(Lib)] public static unsafe partial int vi_points_add_records(IntPtr ctx, void* records, long n, int dims);
  [LibraryImport(Lib, StringMarshalling = StringMarshalling.Utf8)] public static partial int vi_points_add_file(IntPtr ctx, string path, long offsetBytes, long n, int dims, out double readMs, out double totalMs);
  [LibraryImport(Lib, StringMarshalling = StringMarshalling.Utf8)] public static partial int vi_hdf5_dataset_info(IntPtr ctx, string path, string dataset, out long rows, out long cols, out int elemClass, out int elemBytes, out long dataOffset);
  [LibraryImport(Lib)] public static partial IntPtr vi_hdf5_last_error();
  [LibraryImport(Lib, StringMarshalling = StringMarshalling.Utf8)] public static unsafe partial int vi_hdf5_read_rows(IntPtr ctx, string path, string dataset, long firstRow, long n, void* dst, long dstBytes);
  [LibraryImport(Lib, StringMarshalling = StringMarshalling.Utf8)] public static partial int vi_points_add_hdf5(IntPtr ctx, string path, string dataset, long firstRow, long n, long firstId, out double readMs, out double totalMs);
  [LibraryImport(Lib)] public static partial long vi_points_count(IntPtr ctx);
  // ---- build, range table
  [LibraryImport(Lib)] public static partial int vi_build(IntPtr ctx, int mode, out BuildInfo info);
  [LibraryImport(Lib)] public static unsafe partial int vi_build_copy(IntPtr ctx, int mode, out BuildInfo info, long* rangeId, int* dimension, float* mid, long* id, long cap, out long rows);
  [LibraryImport(Lib)] public static unsafe partial int vi_build_levels(IntPtr ctx, LevelInfo* levels, int cap, out int n);
  [LibraryImport(Lib)] public static partial long vi_range_count(IntPtr ctx);
  [LibraryImport(Lib)] public static unsafe partial int vi_ranges_copy(IntPtr ctx, long* rangeId, int* dimension, float* mid, long* id, long cap);
  [LibraryImport(Lib)] public static unsafe partial int vi_ranges_load(IntPtr ctx, long* rangeId, int* dimension, float* mid, long* id, long n, int dims);
  [LibraryImport(Lib)] public static unsafe partial int vi_textindex_copy(IntPtr ctx, long* rangeId, short* dimension, float* mid, long* low, long* high, long* textId, long cap);
  // ---- search
  [LibraryImport(Lib)] public static unsafe partial int vi_search(IntPtr ctx, float* queries, long nq, int dims, float proximity, long* offsets, long* ids, long cap, out long total);
  [LibraryImport(Lib)] public static unsafe partial int vi_search_begin(IntPtr ctx, float* queries, long nq, int dims, float proximity, out long total);
  [LibraryImport(Lib)] public static unsafe partial int vi_search_fetch(IntPtr ctx, long* offsets, long* ids, long cap);
  [LibraryImport(Lib)] public static partial int vi_search_device(IntPtr ctx, IntPtr dQueries, long nq, int dims, float proximity, IntPtr dOffsets, IntPtr dIds, long cap, out long total, out long visits);
  [LibraryImport(Lib)] public static unsafe partial int vi_search_verify(IntPtr ctx, float* queries, long nq, int dims, float proximity, float distance, long* offsets, long* ids, long cap, out long total);
  [LibraryImport(Lib)] public static unsafe partial int vi_search_topk(IntPtr ctx, float* queries, long nq, int dims, float proximity, int k, int metric, long* ids, float* dist, int* count, out long candidates);
  // ---- multi-GPU: one process per GPU; the library owns NCCL (vi_comm_init), or the host brings two callbacks
  [LibraryImport(Lib)] public static unsafe partial int vi_comm_unique_id(byte* id128, int bytes);
  [LibraryImport(Lib)] public static unsafe partial int vi_comm_init(IntPtr ctx, byte* id128, int bytes, int rank, int world);
  [LibraryImport(Lib)] public static unsafe partial int vi_comm_stats(IntPtr ctx, long* calls3, long* bytes3);
  [UnmanagedFunctionPointer(CallingConvention.Cdecl)] public delegate int AllReduceU64(IntPtr user, IntPtr dBuf, long count);
  [UnmanagedFunctionPointer(CallingConvention.Cdecl)] public unsafe delegate int AllToAllV(IntPtr user, IntPtr dSend, long* sendBytes, IntPtr dRecv, long* recvBytes);
  [LibraryImport(Lib)] public static partial int vi_set_collective(IntPtr ctx, int rank, int world, IntPtr allReduce, IntPtr allToAllV, IntPtr user);
  [LibraryImport(Lib)] public static partial int vi_shared_rows(IntPtr ctx, out long sharedRows);
  [LibraryImport(Lib)] public static partial int vi_table_replicate(IntPtr ctx);
  // ---- utilities
  [LibraryImport(Lib)] public static partial int vi_table_device(IntPtr ctx, out IntPtr rangeId, out IntPtr dimension, out IntPtr mid, out IntPtr id, out IntPtr lowRow, out IntPtr highRow);
  [LibraryImport(Lib)] public static partial IntPtr vi_stream(IntPtr ctx);
  [LibraryImport(Lib)] public static partial int vi_debug_divcheck(IntPtr ctx, ulong seed, long samples, out long mismatches);

  /// <summary>Maps a status code back onto the exception the reference would have thrown.</summary>
  public static void Check(IntPtr ctx, int rc)
  {
    if (rc == VI_OK) return;
    var text = (ctx == IntPtr.Zero ? null : Marshal.PtrToStringUTF8(vi_last_error(ctx))) ?? "vi_b200 error " + rc;
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
