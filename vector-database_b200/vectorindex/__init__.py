"""Host-side mirror of NesterovskyBros.VectorIndex over the libvi_b200 C ABI (include/vi_b200.h).

Same names, argument meaning and error behaviour as the reference's C# surface for the split-tree path:

  IndexBuilder.Build(points, storeFactory)   VectorIndex/IndexBuilder.cs:23-25
  IRangeStore / MemoryRangeStore             VectorIndex/IRangeStore.cs:6-22, MemoryRangeStore.cs:7-32
  FileRangeStore(count, dimensions, buffer)  VectorIndex/FileRangeStore.cs:18, .NextStore :40-43
  RangeValue(Dimension, Mid, Id)             VectorIndex/RangeValue.cs:6-22
  VectorIndex.Find(vector, distance, predicate)  shape of MemoryVectorIndex.cs:242-245 over dbo.Search (DDL.sql:234-295)

All arithmetic happens in CUDA kernels behind the C ABI; this module only stages buffers.  There is no CPU
fallback: importing works anywhere, but the first call that needs the library raises if libvi_b200.so or a CUDA
device is missing.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import Callable, Iterable, Iterator, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libvi_b200.so")

VI_OK, VI_ERR_INVALID_ARG, VI_ERR_OVERFLOW, VI_ERR_NOT_IMPLEMENTED, VI_ERR_STATE, VI_ERR_CAPACITY, VI_ERR_OOM, \
    VI_ERR_CUDA = range(8)
MODE_EXACT = 0
MODE_FAST = 1
MODE_SQL = 2   # dbo.BuildIndex's rules (DDL.sql:44-202); Dimension DIM_NULL / Mid NaN = null
DIM_NULL = -3

_i64p = ctypes.POINTER(ctypes.c_int64)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i16p = ctypes.POINTER(ctypes.c_int16)
_f32p = ctypes.POINTER(ctypes.c_float)


class BuildInfo(ctypes.Structure):
    _fields_ = [("ranges", ctypes.c_int64), ("levels", ctypes.c_int32), ("mode", ctypes.c_int32),
                ("point_visits", ctypes.c_int64), ("kernel_launches", ctypes.c_int64), ("build_ms", ctypes.c_double),
                ("q_exponent", ctypes.c_int32), ("reserved", ctypes.c_int32), ("subtree_ms", ctypes.c_double),
                ("subtree_ranges", ctypes.c_int64), ("shared_retry", ctypes.c_int32), ("reserved2", ctypes.c_int32)]


class LevelInfo(ctypes.Structure):
    _fields_ = [("level", ctypes.c_int32), ("derived_points", ctypes.c_int32), ("ranges", ctypes.c_int64),
                ("points", ctypes.c_int64), ("rows_emitted", ctypes.c_int64), ("stats_ms", ctypes.c_double),
                ("partition_ms", ctypes.c_double), ("in_subtrees", ctypes.c_int64)]


ALLREDUCE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64)
ALLTOALLV_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, _i64p, ctypes.c_void_p, _i64p)

# every symbol include/vi_b200.h declares
EXPORTS = ["vi_abi_version", "vi_create", "vi_destroy", "vi_last_error", "vi_points_reserve", "vi_points_add",
           "vi_points_add_device", "vi_points_add_records", "vi_points_add_file", "vi_hdf5_dataset_info", "vi_hdf5_last_error", "vi_hdf5_read_rows",
           "vi_points_add_hdf5", "vi_points_count", "vi_build", "vi_build_copy",
           "vi_build_levels", "vi_range_count", "vi_ranges_copy", "vi_ranges_load", "vi_textindex_copy", "vi_search", "vi_search_begin",
           "vi_search_fetch", "vi_search_topk", "vi_search_device", "vi_search_verify",
           "vi_comm_unique_id", "vi_comm_init", "vi_comm_stats", "vi_set_collective", "vi_shared_rows", "vi_table_replicate",
           "vi_table_device", "vi_stream", "vi_debug_divcheck"]

_lib = None


def hdf5_dataset_info(path: str, dataset: str):
    """(rows, cols, numpy dtype, byte offset of the data) of a contiguous HDF5 data set -- Program.cs:183-222
    GetHdf5DatasetSize.  Host only (no device is touched).  ValueError("Invalid rank.") as the reference."""
    L = load_library()
    rows, cols, off = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
    cls, size = ctypes.c_int32(0), ctypes.c_int32(0)
    # no context (no device needed): the error text is the calling thread's vi_hdf5_last_error()
    rc = L.vi_hdf5_dataset_info(None, os.fsencode(path), dataset.encode(), ctypes.byref(rows), ctypes.byref(cols),
                                ctypes.byref(cls), ctypes.byref(size), ctypes.byref(off))
    if rc != 0:
        raise ValueError((L.vi_hdf5_last_error() or b"").decode() or f"vi_hdf5_dataset_info rc={rc}")
    kind = {0: "i", 1: "f"}[cls.value]
    return rows.value, cols.value, np.dtype(f"<{kind}{size.value}"), off.value


def hdf5_read(path: str, dataset: str, first_row: int = 0, n: Optional[int] = None) -> np.ndarray:
    """Rows [first_row, first_row + n) of a contiguous HDF5 data set into a numpy array (the `/test` queries)."""
    L = load_library()
    rows, cols, dt, _ = hdf5_dataset_info(path, dataset)
    if n is None:
        n = rows - first_row
    out = np.empty((n, cols), dt)
    rc = L.vi_hdf5_read_rows(None, os.fsencode(path), dataset.encode(), first_row, n, out.ctypes.data, out.nbytes)
    if rc != 0:
        raise ValueError((L.vi_hdf5_last_error() or b"").decode() or f"vi_hdf5_read_rows rc={rc}")
    return out


def load_library() -> ctypes.CDLL:
    """Loads libvi_b200.so (built in-tree by csrc/Makefile).  Raises if it is missing: no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           f"(make -C vector-database_b200/csrc). There is no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp = ctypes.c_void_p
    L.vi_abi_version.restype = ctypes.c_int
    L.vi_create.argtypes = [ctypes.c_int32, ctypes.POINTER(vp)]
    L.vi_destroy.argtypes = [vp]
    L.vi_destroy.restype = None
    L.vi_last_error.argtypes = [vp]
    L.vi_last_error.restype = ctypes.c_char_p
    L.vi_points_reserve.argtypes = [vp, ctypes.c_int64, ctypes.c_int32]
    L.vi_points_add.argtypes = [vp, _i64p, _f32p, ctypes.c_int64, ctypes.c_int32]
    L.vi_points_add_device.argtypes = [vp, vp, vp, ctypes.c_int64, ctypes.c_int32]
    L.vi_points_count.argtypes = [vp]
    L.vi_points_count.restype = ctypes.c_int64
    L.vi_build.argtypes = [vp, ctypes.c_int32, ctypes.POINTER(BuildInfo)]
    L.vi_build_copy.argtypes = [vp, ctypes.c_int32, ctypes.POINTER(BuildInfo), _i64p, _i32p, _f32p, _i64p, ctypes.c_int64, _i64p]
    L.vi_build_levels.argtypes = [vp, ctypes.POINTER(LevelInfo), ctypes.c_int32, _i32p]
    L.vi_range_count.argtypes = [vp]
    L.vi_range_count.restype = ctypes.c_int64
    L.vi_ranges_copy.argtypes = [vp, _i64p, _i32p, _f32p, _i64p, ctypes.c_int64]
    L.vi_textindex_copy.argtypes = [vp, _i64p, _i16p, _f32p, _i64p, _i64p, _i64p, ctypes.c_int64]
    L.vi_ranges_load.argtypes = [vp, _i64p, _i32p, _f32p, _i64p, ctypes.c_int64, ctypes.c_int32]
    L.vi_points_add_records.argtypes = [vp, vp, ctypes.c_int64, ctypes.c_int32]
    L.vi_points_add_file.argtypes = [vp, ctypes.c_char_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int32,
                                     ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    L.vi_hdf5_dataset_info.argtypes = [vp, ctypes.c_char_p, ctypes.c_char_p, _i64p, _i64p, _i32p, _i32p, _i64p]
    L.vi_hdf5_last_error.argtypes = []
    L.vi_hdf5_last_error.restype = ctypes.c_char_p
    L.vi_hdf5_read_rows.argtypes = [vp, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int64, ctypes.c_int64, vp, ctypes.c_int64]
    L.vi_points_add_hdf5.argtypes = [vp, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                     ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    L.vi_search.argtypes = [vp, _f32p, ctypes.c_int64, ctypes.c_int32, ctypes.c_float, _i64p, _i64p, ctypes.c_int64,
                            _i64p]
    L.vi_search_begin.argtypes = [vp, _f32p, ctypes.c_int64, ctypes.c_int32, ctypes.c_float, _i64p]
    L.vi_search_fetch.argtypes = [vp, _i64p, _i64p, ctypes.c_int64]
    L.vi_search_device.argtypes = [vp, vp, ctypes.c_int64, ctypes.c_int32, ctypes.c_float, vp, vp, ctypes.c_int64,
                                   _i64p, _i64p]
    L.vi_search_verify.argtypes = [vp, _f32p, ctypes.c_int64, ctypes.c_int32, ctypes.c_float, ctypes.c_float, _i64p,
                                   _i64p, ctypes.c_int64, _i64p]
    L.vi_search_topk.argtypes = [vp, _f32p, ctypes.c_int64, ctypes.c_int32, ctypes.c_float, ctypes.c_int32, ctypes.c_int32,
                                 _i64p, _f32p, _i32p, _i64p]
    L.vi_comm_unique_id.argtypes = [vp, ctypes.c_int32]
    L.vi_comm_init.argtypes = [vp, vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]
    L.vi_comm_stats.argtypes = [vp, _i64p, _i64p]
    L.vi_set_collective.argtypes = [vp, ctypes.c_int32, ctypes.c_int32, ALLREDUCE_FN, ALLTOALLV_FN, vp]
    L.vi_shared_rows.argtypes = [vp, _i64p]
    L.vi_table_replicate.argtypes = [vp]
    L.vi_table_device.argtypes = [vp] + [ctypes.POINTER(vp)] * 6
    L.vi_stream.argtypes = [vp]
    L.vi_debug_divcheck.argtypes = [vp, ctypes.c_uint64, ctypes.c_int64, _i64p]
    L.vi_stream.restype = vp
    for name in EXPORTS:
        if name not in ("vi_destroy", "vi_last_error", "vi_hdf5_last_error", "vi_points_count", "vi_range_count",
                        "vi_stream", "vi_abi_version"):
            getattr(L, name).restype = ctypes.c_int
    _lib = L
    return L


def comm_unique_id() -> bytes:
    """128 bytes identifying a new NCCL communicator (call on one rank, hand them to all)."""
    buf = (ctypes.c_ubyte * 128)()
    rc = load_library().vi_comm_unique_id(buf, 128)
    if rc != VI_OK:
        raise VectorIndexError(rc, "vi_comm_unique_id failed (libnccl.so.2 not loadable?)")
    return bytes(buf)


class VectorIndexError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"vi_b200 error {code}: {message}")
        self.code = code


def _raise(code: int, message: str):
    """Maps C-ABI status codes back onto the reference's exception types (SURVEY.md 8b 'Errors')."""
    if code == VI_ERR_INVALID_ARG:
        raise ValueError(message)  # ArgumentException
    if code == VI_ERR_OVERFLOW:
        raise OverflowError(message)  # OverflowException, IndexBuilder.cs:99,104
    if code == VI_ERR_NOT_IMPLEMENTED:
        raise NotImplementedError(message)  # IndexBuilder.cs:206-209
    if code == VI_ERR_OOM:
        raise MemoryError(message)
    raise VectorIndexError(code, message)


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t)


@dataclass(frozen=True)
class RangeValue:
    """VectorIndex/RangeValue.cs:6-22."""
    Dimension: int
    Mid: float
    Id: int


class Context:
    """Thin owner of one vi_ctx (one CUDA device, one host thread)."""

    def __init__(self, device: int = 0):
        self._L = load_library()
        h = ctypes.c_void_p()
        rc = self._L.vi_create(device, ctypes.byref(h))
        if rc != VI_OK:
            raise VectorIndexError(rc, "vi_create failed: no usable CUDA device (there is no CPU fallback)")
        self._h = h
        self.dims = 0
        self._cb = None

    def close(self):
        if getattr(self, "_h", None):
            self._L.vi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int):
        if rc != VI_OK:
            _raise(rc, self._L.vi_last_error(self._h).decode())

    # ---- ingest ----
    def reserve(self, capacity: int, dims: int):
        self._check(self._L.vi_points_reserve(self._h, capacity, dims))
        self.dims = dims

    def add(self, ids: np.ndarray, rows: np.ndarray):
        ids = np.ascontiguousarray(ids, np.int64)
        rows = np.ascontiguousarray(rows, np.float32)
        if rows.ndim != 2 or ids.shape[0] != rows.shape[0]:
            raise ValueError("Invalid length of vector.")
        self._check(self._L.vi_points_add(self._h, _p(ids, _i64p), _p(rows, _f32p), rows.shape[0], rows.shape[1]))

    def add_device(self, d_ids_ptr: int, d_rows_ptr: int, n: int, dims: int):
        self._check(self._L.vi_points_add_device(self._h, d_ids_ptr, d_rows_ptr, n, dims))

    def add_records(self, records, dims: int):
        """n records [int64 id][dims x float32] (the FileRangeStore layout, FileRangeStore.cs:127-165) in one buffer."""
        buf = np.ascontiguousarray(np.frombuffer(records, np.uint8) if not isinstance(records, np.ndarray) else records)
        rec = 8 + 4 * dims
        if buf.nbytes % rec:
            raise ValueError("Invalid length of vector.")
        self._check(self._L.vi_points_add_records(self._h, buf.ctypes.data, buf.nbytes // rec, dims))

    def add_file(self, path: str, dims: int, offset_bytes: int = 0, n: int = -1):
        """Streams records from a file (two pinned buffers). Returns (read_ms, total_ms)."""
        rd, tot = ctypes.c_double(0), ctypes.c_double(0)
        self._check(self._L.vi_points_add_file(self._h, os.fsencode(path), offset_bytes, n, dims, ctypes.byref(rd),
                                               ctypes.byref(tot)))
        return rd.value, tot.value

    def add_hdf5(self, path: str, dataset: str = "/train", first_row: int = 0, n: int = -1, first_id: Optional[int] = None):
        """Streams float32 rows of an HDF5 data set (Program.cs:224-260 GetHdf5Dataset); ids = row indexes unless
        first_id is given. Returns (read_ms, total_ms)."""
        rd, tot = ctypes.c_double(0), ctypes.c_double(0)
        self._check(self._L.vi_points_add_hdf5(self._h, os.fsencode(path), dataset.encode(), first_row, n,
                                               first_row if first_id is None else first_id, ctypes.byref(rd),
                                               ctypes.byref(tot)))
        return rd.value, tot.value

    @property
    def count(self) -> int:
        return int(self._L.vi_points_count(self._h))

    # ---- build ----
    def build(self, mode: int = MODE_EXACT) -> BuildInfo:
        info = BuildInfo()
        self._check(self._L.vi_build(self._h, mode, ctypes.byref(info)))
        return info

    def build_into(self, mode: int, rid: np.ndarray, dim: np.ndarray, mid: np.ndarray, oid: np.ndarray):
        """vi_build_copy: build and deliver the range table into caller-owned (ideally pinned) arrays, the D2H copy
        overlapped with the build's last kernel.  Returns (BuildInfo, rows)."""
        info = BuildInfo()
        k = ctypes.c_int64(0)
        self._check(self._L.vi_build_copy(self._h, mode, ctypes.byref(info), _p(rid, _i64p), _p(dim, _i32p), _p(mid, _f32p),
                                          _p(oid, _i64p), min(len(rid), len(dim), len(mid), len(oid)), ctypes.byref(k)))
        return info, k.value

    def levels(self):
        n = ctypes.c_int32(0)
        self._check(self._L.vi_build_levels(self._h, None, 0, ctypes.byref(n)))
        arr = (LevelInfo * max(n.value, 1))()
        self._check(self._L.vi_build_levels(self._h, arr, n.value, ctypes.byref(n)))
        return [arr[i] for i in range(n.value)]

    @property
    def range_count(self) -> int:
        return int(self._L.vi_range_count(self._h))

    def ranges(self):
        """(rangeId, Dimension, Mid, Id) columns, one row per range (unspecified order)."""
        k = self.range_count
        rid = np.empty(k, np.int64)
        dim = np.empty(k, np.int32)
        mid = np.empty(k, np.float32)
        oid = np.empty(k, np.int64)
        self._check(self._L.vi_ranges_copy(self._h, _p(rid, _i64p), _p(dim, _i32p), _p(mid, _f32p), _p(oid, _i64p), k))
        return rid, dim, mid, oid

    def ranges_into(self, rid: np.ndarray, dim: np.ndarray, mid: np.ndarray, oid: np.ndarray) -> int:
        """Same as ranges() into caller-owned (e.g. pinned) arrays; returns the row count."""
        k = self.range_count
        self._check(self._L.vi_ranges_copy(self._h, _p(rid, _i64p), _p(dim, _i32p), _p(mid, _f32p), _p(oid, _i64p),
                                           min(rid.shape[0], dim.shape[0], mid.shape[0], oid.shape[0])))
        return k

    def load_ranges(self, rid: np.ndarray, dim: np.ndarray, mid: np.ndarray, oid: np.ndarray, dims: int):
        """The inverse of ranges(): rows (RangeID, Dimension, Mid, Id) in any order become the searchable table."""
        rid = np.ascontiguousarray(rid, np.int64)
        dim = np.ascontiguousarray(dim, np.int32)
        mid = np.ascontiguousarray(mid, np.float32)
        oid = np.ascontiguousarray(oid, np.int64)
        if not (rid.shape == dim.shape == mid.shape == oid.shape) or rid.ndim != 1:
            raise ValueError("range table columns differ in length")
        self._check(self._L.vi_ranges_load(self._h, _p(rid, _i64p), _p(dim, _i32p), _p(mid, _f32p), _p(oid, _i64p),
                                           rid.shape[0], dims))
        self.dims = dims

    def textindex(self):
        """dbo.TextIndex columns (DDL.sql:209-227): RangeID, Dimension, Mid, LowRangeID, HighRangeID, TextID."""
        k = self.range_count
        rid = np.empty(k, np.int64)
        dim = np.empty(k, np.int16)
        mid = np.empty(k, np.float32)
        lo = np.empty(k, np.int64)
        hi = np.empty(k, np.int64)
        tid = np.empty(k, np.int64)
        self._check(self._L.vi_textindex_copy(self._h, _p(rid, _i64p), _p(dim, _i16p), _p(mid, _f32p), _p(lo, _i64p),
                                              _p(hi, _i64p), _p(tid, _i64p), k))
        return rid, dim, mid, lo, hi, tid

    # ---- search ----
    def search(self, queries: np.ndarray, proximity: float):
        """Batched dbo.Search. Returns CSR (offsets[nq+1], ids)."""
        queries = np.ascontiguousarray(queries, np.float32)
        if queries.ndim == 1:
            queries = queries[None, :]
        nq, d = queries.shape
        offsets = np.zeros(nq + 1, np.int64)
        total = ctypes.c_int64(0)
        self._check(self._L.vi_search_begin(self._h, _p(queries, _f32p), nq, d, ctypes.c_float(proximity), ctypes.byref(total)))
        ids = np.empty(max(total.value, 1), np.int64)
        self._check(self._L.vi_search_fetch(self._h, _p(offsets, _i64p), _p(ids, _i64p), ids.shape[0]))
        return offsets, ids[:total.value]

    def search_begin(self, queries: np.ndarray, proximity: float) -> int:
        """vi_search_begin: stages the queries, walks the table, returns the number of candidates."""
        queries = np.ascontiguousarray(queries, np.float32)
        nq, d = queries.shape
        total = ctypes.c_int64(0)
        self._check(self._L.vi_search_begin(self._h, _p(queries, _f32p), nq, d, ctypes.c_float(proximity), ctypes.byref(total)))
        return total.value

    def search_fetch(self, offsets: np.ndarray, ids: np.ndarray):
        """vi_search_fetch into caller-owned (e.g. pinned) arrays: offsets[nq+1], ids[>= total]."""
        self._check(self._L.vi_search_fetch(self._h, _p(offsets, _i64p), _p(ids, _i64p), ids.shape[0]))

    def search_two_call(self, queries: np.ndarray, proximity: float):
        """The size-then-fill protocol of vi_search itself (ids == NULL first, then cap >= total)."""
        queries = np.ascontiguousarray(queries, np.float32)
        if queries.ndim == 1:
            queries = queries[None, :]
        nq, d = queries.shape
        offsets = np.zeros(nq + 1, np.int64)
        total = ctypes.c_int64(0)
        self._check(self._L.vi_search(self._h, _p(queries, _f32p), nq, d, ctypes.c_float(proximity), _p(offsets, _i64p), None,
                                      0, ctypes.byref(total)))
        ids = np.empty(max(total.value, 1), np.int64)
        self._check(self._L.vi_search(self._h, _p(queries, _f32p), nq, d, ctypes.c_float(proximity), _p(offsets, _i64p),
                                      _p(ids, _i64p), total.value, ctypes.byref(total)))
        return offsets, ids[:total.value]

    def search_verify(self, queries: np.ndarray, proximity: float, distance: float):
        queries = np.ascontiguousarray(queries, np.float32)
        if queries.ndim == 1:
            queries = queries[None, :]
        nq, d = queries.shape
        offsets = np.zeros(nq + 1, np.int64)
        total = ctypes.c_int64(0)
        self._check(self._L.vi_search_verify(self._h, _p(queries, _f32p), nq, d, ctypes.c_float(proximity),
                                             ctypes.c_float(distance), _p(offsets, _i64p), None, 0, ctypes.byref(total)))
        ids = np.empty(max(total.value, 1), np.int64)
        self._check(self._L.vi_search_verify(self._h, _p(queries, _f32p), nq, d, ctypes.c_float(proximity),
                                             ctypes.c_float(distance), _p(offsets, _i64p), _p(ids, _i64p), total.value,
                                             ctypes.byref(total)))
        return offsets, ids[:total.value]

    def search_topk(self, queries: np.ndarray, proximity: float, k: int, metric: int = 0):
        """k nearest candidates per query: (ids[nq,k] (-1 = none), dist[nq,k] (+inf = none), count[nq], candidates)."""
        queries = np.ascontiguousarray(queries, np.float32)
        nq, d = queries.shape
        ids = np.empty((nq, k), np.int64)
        dist = np.empty((nq, k), np.float32)
        cnt = np.empty(nq, np.int32)
        cand = ctypes.c_int64(0)
        self._check(self._L.vi_search_topk(self._h, _p(queries, _f32p), nq, d, proximity, k, metric, _p(ids, _i64p),
                                           _p(dist, _f32p), _p(cnt, _i32p), ctypes.byref(cand)))
        return ids, dist, cnt, cand.value

    def search_device(self, d_queries_ptr: int, nq: int, dims: int, proximity: float, d_offsets_ptr: int,
                      d_ids_ptr: int, cap: int):
        total = ctypes.c_int64(0)
        visits = ctypes.c_int64(0)
        rc = self._L.vi_search_device(self._h, d_queries_ptr, nq, dims, ctypes.c_float(proximity), d_offsets_ptr,
                                      d_ids_ptr, cap, ctypes.byref(total), ctypes.byref(visits))
        self._check(rc)
        return total.value, visits.value

    def set_collective(self, rank: int, world: int, allreduce, alltoallv):
        """Registers the two collectives of a multi-rank build (see vectorindex.distributed).
        allreduce(ptr, count) and alltoallv(send_ptr, send_bytes[world], recv_ptr, recv_bytes[world]) get raw device
        pointers and return 0 on success."""
        def _ar(user, buf, count):
            try:
                return int(allreduce(buf, count) or 0)
            except Exception as e:  # an exception must not cross the C boundary
                print(f"allreduce callback failed: {e!r}")
                return 1

        def _a2a(user, send, sbytes, recv, rbytes):
            try:
                return int(alltoallv(send, [sbytes[i] for i in range(world)], recv, [rbytes[i] for i in range(world)]) or 0)
            except Exception as e:
                print(f"alltoallv callback failed: {e!r}")
                return 1

        self._cb = (ALLREDUCE_FN(_ar), ALLTOALLV_FN(_a2a))  # keep the thunks alive
        self._check(self._L.vi_set_collective(self._h, rank, world, self._cb[0], self._cb[1], None))

    def comm_init(self, unique_id: bytes, rank: int, world: int) -> None:
        """Library-owned NCCL communicator (collective call; unique_id: the 128 bytes of comm_unique_id() on rank 0)."""
        buf = (ctypes.c_ubyte * 128).from_buffer_copy(unique_id[:128] if world > 1 else bytes(128))
        self._check(self._L.vi_comm_init(self._h, buf, 128, rank, world))

    def comm_stats(self):
        calls = (ctypes.c_int64 * 3)()
        nbytes = (ctypes.c_int64 * 3)()
        self._check(self._L.vi_comm_stats(self._h, calls, nbytes))
        return {"allreduce": calls[0], "alltoallv": calls[1], "allgather": calls[2], "allreduce_bytes": nbytes[0],
                "alltoallv_bytes": nbytes[1], "allgather_bytes": nbytes[2]}

    def replicate(self) -> None:
        """multi-rank: every rank gets the whole table (collective call)"""
        self._check(self._L.vi_table_replicate(self._h))

    @property
    def shared_rows(self) -> int:
        k = ctypes.c_int64(0)
        self._check(self._L.vi_shared_rows(self._h, ctypes.byref(k)))
        return k.value

    def divcheck(self, seed: int, samples: int) -> int:
        bad = ctypes.c_int64(-1)
        self._check(self._L.vi_debug_divcheck(self._h, seed, samples, ctypes.byref(bad)))
        return bad.value

    @property
    def stream(self) -> int:
        return int(self._L.vi_stream(self._h) or 0)


# ---- IRangeStore family (host-side staging; children of the split tree live on the device) -------------------------
# ---- the formats on either side: CSV "RangeID,Dimension,Mid,ID" (Program.cs:80,145-149), FileRangeStore records ----
CSV_HEADER = "RangeID,Dimension,Mid,ID"


def _fmt_float32(x: np.float32) -> str:
    # shortest string that round-trips the float32 (what float.ToString() gives on .NET Core 3.0+)
    if np.isnan(x):
        return "NaN"
    if np.isinf(x):
        return "Infinity" if x > 0 else "-Infinity"
    return np.format_float_positional(x, unique=True, trim="-") if 1e-5 <= abs(float(x)) < 1e15 or x == 0 \
        else np.format_float_scientific(x, unique=True, trim="-", exp_digits=2).replace("e", "E")


def write_csv(path: str, rid, dim, mid, oid) -> None:
    """One line per range, as Program.cs:145-149 writes them (any row order)."""
    with open(path, "w", newline="") as f:
        f.write(CSV_HEADER + "\r\n")
        for r, d, m, i in zip(rid.tolist(), dim.tolist(), np.asarray(mid, np.float32), oid.tolist()):
            f.write(f"{r},{d},{_fmt_float32(m)},{i}\r\n")


def read_csv(path: str):
    """-> (rid int64, dim int32, mid float32, id int64); Mid is parsed to the nearest float32 (round trip exact)."""
    rid, dim, mid, oid = [], [], [], []
    with open(path, "r", newline="") as f:
        header = f.readline().strip()
        if header != CSV_HEADER:
            raise ValueError(f"not a range table CSV: header {header!r}")
        for line in f:
            line = line.strip()
            if not line:
                continue
            a, b, c, d = line.split(",")
            rid.append(int(a))
            dim.append(int(b))
            mid.append({"NaN": np.nan, "Infinity": np.inf, "-Infinity": -np.inf}.get(c, c))
            oid.append(int(d))
    # (each text is parsed straight to float32: going through float64 first could round twice)
    return (np.array(rid, np.int64), np.array(dim, np.int32), np.array([np.float32(x) for x in mid], np.float32),
            np.array(oid, np.int64))


def pack_records(ids: np.ndarray, rows: np.ndarray) -> np.ndarray:
    """[int64 id][dims x float32] records (FileRangeStore.cs:127-165) as one uint8 array."""
    ids = np.ascontiguousarray(ids, np.int64)
    rows = np.ascontiguousarray(rows, np.float32)
    n, d = rows.shape
    out = np.empty((n, 8 + 4 * d), np.uint8)
    out[:, :8] = ids.view(np.uint8).reshape(n, 8)
    out[:, 8:] = rows.view(np.uint8).reshape(n, 4 * d)
    return out.reshape(-1)


class IRangeStore:
    """VectorIndex/IRangeStore.cs:6-22."""

    def Add(self, id: int, vector) -> None:
        raise NotImplementedError

    def GetPoints(self) -> Iterator[Tuple[int, np.ndarray]]:
        raise NotImplementedError

    def Dispose(self) -> None:
        pass


class MemoryRangeStore(IRangeStore):
    """VectorIndex/MemoryRangeStore.cs:7-32: keeps references to the caller's vectors, insertion order."""

    def __init__(self):
        self._data = []

    def Add(self, id: int, vector) -> None:
        self._data.append((int(id), vector))

    def GetPoints(self):
        return iter(self._data)


class FileRangeStore:
    """VectorIndex/FileRangeStore.cs:10-182: a factory of stores with a fixed dimension count that COPIES what is
    added.  The reference backs it with a memory-mapped file of (8+4*D)*4*count bytes; the B200 equivalent of that
    scratch space is the device-resident permutation ping-pong, so this class only validates and stages."""

    def __init__(self, count: int, dimensions: int, buffer: int = 10000):
        if count < 0 or dimensions <= 0 or dimensions > 32767:
            raise ValueError("Invalid count or dimensions.")
        self.count = count
        self.dimensions = dimensions
        self.buffer = buffer

    def Dispose(self):
        pass

    def NextStore(self, rangeId: int, capacity: int) -> IRangeStore:
        return _FileStore(self, rangeId, capacity)


class _FileStore(IRangeStore):
    def __init__(self, container: FileRangeStore, rangeId: int, capacity: int):
        self._c = container
        self._ids = []
        self._rows = []

    def Add(self, id: int, vector) -> None:
        v = np.asarray(vector, np.float32)
        if v.shape[0] != self._c.dimensions:
            raise ValueError("Invalid length of vector.")  # FileRangeStore.cs:59-64
        self._ids.append(int(id))
        self._rows.append(v.copy())

    def GetPoints(self):
        return iter(zip(self._ids, self._rows))


class _RootStore(IRangeStore):
    """IndexBuilder.cs:200-212: wraps the input; Add is not implemented."""

    def __init__(self, points):
        self.points = points

    def Add(self, id, vector):
        raise NotImplementedError()

    def GetPoints(self):
        return iter(self.points)


def _drain(points, batch: int = 65536):
    """Yields (ids, rows) numpy batches from (ids, rows) arrays or an iterable of (id, vector)."""
    if isinstance(points, tuple) and len(points) == 2 and isinstance(points[1], np.ndarray) and points[1].ndim == 2:
        yield np.asarray(points[0], np.int64), points[1]
        return
    ids, rows = [], []
    for id_, v in points:
        ids.append(int(id_))
        rows.append(np.asarray(v, np.float32))
        if len(ids) >= batch:
            yield np.asarray(ids, np.int64), _stack(rows)
            ids, rows = [], []
    if ids:
        yield np.asarray(ids, np.int64), _stack(rows)


def _stack(rows):
    d = rows[0].shape[0]
    for r in rows:
        if r.shape[0] != d:
            raise ValueError("Invalid length of vector.")
    return np.stack(rows).astype(np.float32, copy=False)


class IndexBuilder:
    """VectorIndex/IndexBuilder.cs:12-198."""

    @staticmethod
    def Build(points, storeFactory: Optional[Callable[[int, int], IRangeStore]] = None, *, mode: int = MODE_EXACT,
              device: int = 0, context: Optional[Context] = None) -> Iterator[Tuple[int, RangeValue]]:
        """Yields (rangeId, RangeValue) like IndexBuilder.Build (IndexBuilder.cs:23-25, :92).

        `points`: iterable of (id, vector) -- enumerated once -- or a tuple (ids[n], rows[n, d]).
        `storeFactory(rangeId, capacity)` is accepted for signature compatibility; child ranges never leave the
        device, so it is not called.  Row order differs from the reference's depth-first order; consumers key by
        rangeId (Program.cs:18-26)."""
        ctx = context or Context(device)
        try:
            first = True
            for ids, rows in _drain(points):
                if first:
                    ctx.reserve(rows.shape[0], rows.shape[1])
                    first = False
                ctx.add(ids, rows)
            if first:
                return  # empty input: nothing is yielded (IndexBuilder.cs:70-73)
            ctx.build(mode)
            rid, dim, mid, oid = ctx.ranges()
            for i in range(rid.shape[0]):
                yield int(rid[i]), RangeValue(int(dim[i]), float(mid[i]), int(oid[i]))
        finally:
            if context is None:
                ctx.close()


class VectorIndex:
    """A built split-tree index resident on one GPU with the Find(vector, distance, predicate) entry shape of
    MemoryVectorIndex.cs:242-245; traversal semantics are dbo.Search's (DDL.sql:234-295)."""

    def __init__(self, ids: np.ndarray, rows: np.ndarray, *, mode: int = MODE_EXACT, device: int = 0):
        self.ctx = Context(device)
        rows = np.ascontiguousarray(rows, np.float32)
        self.ctx.reserve(rows.shape[0], rows.shape[1])
        self.ctx.add(ids, rows)
        self.info = self.ctx.build(mode)
        self._rows = rows
        self._ids = np.ascontiguousarray(ids, np.int64)
        self._by_id = None

    @property
    def Count(self) -> int:
        return self.ctx.count

    def Search(self, queries, proximity: float):
        return self.ctx.search(queries, proximity)

    def Find(self, vector, distance: float, predicate: Optional[Callable[[int, np.ndarray], bool]] = None):
        """Candidates of the box vector +- distance, each passed to predicate(id, vector) which verifies the match
        ("predicate should verify the match", MemoryVectorIndex.cs:237-241).  Without a predicate the GPU
        Euclidean verification kernel is used."""
        vector = np.asarray(vector, np.float32)
        if vector.shape[0] != self.ctx.dims:
            raise ValueError("Invalid vector size.")  # MemoryVectorIndex.cs:254
        if predicate is None:
            _, ids = self.ctx.search_verify(vector, distance, distance)
            yield from (int(i) for i in ids)
            return
        _, ids = self.ctx.search(vector, distance)
        if self._by_id is None:
            self._by_id = {int(i): k for k, i in enumerate(self._ids)}
        for i in ids:
            if predicate(int(i), self._rows[self._by_id[int(i)]]):
                yield int(i)

    def close(self):
        self.ctx.close()
