"""CPU ORACLE (test infrastructure only) for the split-tree index: ctypes binding of oracle/vi_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  The product path (vector-database_b200/) never does.  PARITY UNPINNED: see vi_oracle.c header.

Reference restated: VectorIndex/IndexBuilder.cs:23-198, VectorIndex/Stats.cs, VectorIndex/RangeValue.cs,
DDL.sql:246-295.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libvi_oracle.so")

MODE_LITERAL = 0
MODE_QFX = 1
MODE_SQL = 2  # dbo.BuildIndex's rules (DDL.sql:44-202) over the qfx statistics; Dimension -3 / Mid NaN = null

_i64p = ctypes.POINTER(ctypes.c_int64)
_i32p = ctypes.POINTER(ctypes.c_int32)
_f32p = ctypes.POINTER(ctypes.c_float)


def build_lib(force: bool = False) -> str:
    """Compile oracle/vi_oracle.c -> oracle/libvi_oracle.so (gcc, flags in oracle/Makefile)."""
    srcs = [os.path.join(_HERE, f) for f in ("vi_oracle.c", "vi_oracle_mt.c", "Makefile")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libvi_oracle.so"])
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build_lib()
        L = ctypes.CDLL(_LIB_PATH)
        L.vio_build.restype = ctypes.c_int
        L.vio_build.argtypes = [ctypes.c_int64, ctypes.c_int32, ctypes.c_int64, _i64p, _f32p, ctypes.c_int,
                                ctypes.c_int64, _i64p, _i32p, _f32p, _i64p, _i64p]
        L.vio_build_ex.restype = ctypes.c_int
        L.vio_build_ex.argtypes = L.vio_build.argtypes + [ctypes.c_int64, ctypes.c_int, ctypes.c_int32]
        L.vio_build_mt.restype = ctypes.c_int
        L.vio_build_mt.argtypes = [ctypes.c_int64, ctypes.c_int32, ctypes.c_int64, _i64p, _f32p, ctypes.c_int64, _i64p,
                                   _i32p, _f32p, _i64p, _i64p, ctypes.c_int]
        L.vio_search_batch.restype = ctypes.c_int
        L.vio_search_batch.argtypes = [ctypes.c_int64, _i64p, _i32p, _f32p, _i64p, ctypes.c_int32,
                                       ctypes.c_int64, _f32p, ctypes.c_int64, ctypes.c_float, ctypes.c_int64,
                                       _i64p, _i64p, _i64p, _i64p]
        L.vio_qfx_exponent.restype = ctypes.c_int
        L.vio_qfx_exponent.argtypes = [_f32p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int64]
        L.vio_distance_l2.restype = ctypes.c_float
        L.vio_distance_l2.argtypes = [_f32p, _f32p, ctypes.c_int32]
        L.vio_distance_angular.restype = ctypes.c_float
        L.vio_distance_angular.argtypes = [_f32p, _f32p, ctypes.c_int32]
        _lib = L
    return _lib


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t)


@dataclass
class RangeTable:
    """Rows (rangeId, RangeValue{Dimension, Mid, Id}) -- RangeValue.cs:6-22 -- sorted by rangeId."""
    range_id: np.ndarray  # int64
    dimension: np.ndarray  # int32, -1 = leaf
    mid: np.ndarray  # float32
    id: np.ndarray  # int64
    emission_order: np.ndarray | None = None  # rangeIds in the order the reference would yield them

    def __len__(self) -> int:
        return int(self.range_id.shape[0])

    def as_dict(self) -> dict:
        return {int(r): (int(d), float(m), int(i))
                for r, d, m, i in zip(self.range_id, self.dimension, self.mid, self.id)}


class OracleError(Exception):
    pass


def build(ids: np.ndarray, rows: np.ndarray, mode: int = MODE_LITERAL, root_rid: int = 0, root_depth: int = 0,
          qe: int | None = None) -> RangeTable:
    """IndexBuilder.Build restated (IndexBuilder.cs:23-157). rows: float32 [n, d], ids: int64 [n].
    root_rid / root_depth / qe: build the sub-tree of one range with a given fixed-point exponent (multi-rank tests)."""
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    assert rows.ndim == 2 and ids.shape[0] == rows.shape[0]
    n, d = rows.shape
    cap = 2 * n + n // 8 + 1024  # 2n-1 rows normally; a one-sided split adds a one-child row (same as the GPU table)
    rid = np.empty(cap, np.int64)
    dim = np.empty(cap, np.int32)
    mid = np.empty(cap, np.float32)
    oid = np.empty(cap, np.int64)
    cnt = ctypes.c_int64(0)
    rc = lib().vio_build_ex(n, d, d, _p(ids, _i64p), _p(rows, _f32p), mode, cap, _p(rid, _i64p), _p(dim, _i32p),
                            _p(mid, _f32p), _p(oid, _i64p), ctypes.byref(cnt), root_rid, 1 if root_depth % 2 == 0 else 0,
                            -2 ** 31 if qe is None else qe)
    if rc == -2:
        raise OverflowError("rangeId overflow (IndexBuilder.cs:99,104 checked arithmetic)")
    if rc == -4 and mode != MODE_LITERAL and np.isinf(rows).any():
        raise ValueError("the qfx modes refuse +-Inf in the data (no fixed-point image)")
    if rc != 0:
        raise OracleError(f"vio_build rc={rc}")
    k = cnt.value
    order = np.argsort(rid[:k], kind="stable")
    return RangeTable(rid[:k][order].copy(), dim[:k][order].copy(), mid[:k][order].copy(), oid[:k][order].copy(),
                      rid[:k].copy())


def build_mt(ids: np.ndarray, rows: np.ndarray, threads: int) -> RangeTable:
    """The literal build with `threads` host threads (vi_oracle_mt.c): same table as build(), used as the CPU baseline."""
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    n, d = rows.shape
    cap = 2 * n + n // 8 + 1024
    rid = np.empty(cap, np.int64)
    dim = np.empty(cap, np.int32)
    mid = np.empty(cap, np.float32)
    oid = np.empty(cap, np.int64)
    cnt = ctypes.c_int64(0)
    rc = lib().vio_build_mt(n, d, d, _p(ids, _i64p), _p(rows, _f32p), cap, _p(rid, _i64p), _p(dim, _i32p), _p(mid, _f32p),
                            _p(oid, _i64p), ctypes.byref(cnt), threads)
    if rc == -2:
        raise OverflowError("rangeId overflow (IndexBuilder.cs:99,104 checked arithmetic)")
    if rc != 0:
        raise OracleError(f"vio_build_mt rc={rc}")
    k = cnt.value
    order = np.argsort(rid[:k], kind="stable")
    return RangeTable(rid[:k][order].copy(), dim[:k][order].copy(), mid[:k][order].copy(), oid[:k][order].copy(), None)


def qfx_exponent(rows: np.ndarray) -> int:
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    n, d = rows.shape
    return int(lib().vio_qfx_exponent(_p(rows, _f32p), n, d, d))


def search(table: RangeTable, queries: np.ndarray, proximity: float):
    """dbo.Search restated (DDL.sql:246-295). Returns (offsets[nq+1], ids, total_visits); ids per query in
    DFS order (low branch first)."""
    queries = np.ascontiguousarray(queries, dtype=np.float32)
    if queries.ndim == 1:
        queries = queries[None, :]
    nq, d = queries.shape
    offsets = np.zeros(nq + 1, np.int64)
    total = ctypes.c_int64(0)
    visits = ctypes.c_int64(0)
    args = (len(table), _p(table.range_id, _i64p), _p(table.dimension, _i32p), _p(table.mid, _f32p),
            _p(table.id, _i64p), d, nq, _p(queries, _f32p), d, ctypes.c_float(proximity))
    rc = lib().vio_search_batch(*args, 0, _p(offsets, _i64p), None, ctypes.byref(total), ctypes.byref(visits))
    if rc != 0:
        raise OracleError(f"vio_search_batch rc={rc}")
    out = np.empty(max(total.value, 1), np.int64)
    rc = lib().vio_search_batch(*args, total.value, _p(offsets, _i64p), _p(out, _i64p), ctypes.byref(total),
                                ctypes.byref(visits))
    if rc != 0:
        raise OracleError(f"vio_search_batch rc={rc}")
    return offsets, out[:total.value], int(visits.value)


def distance_l2(a: np.ndarray, b: np.ndarray) -> float:
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return float(lib().vio_distance_l2(_p(a, _f32p), _p(b, _f32p), a.shape[0]))


def distance_angular(a: np.ndarray, b: np.ndarray) -> float:
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return float(lib().vio_distance_angular(_p(a, _f32p), _p(b, _f32p), a.shape[0]))
