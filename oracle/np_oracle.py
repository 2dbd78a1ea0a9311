"""Independent numpy restatement of the reference (small cases only) used to cross-check oracle/vi_oracle.c.

TEST INFRASTRUCTURE ONLY (same rules as oracle/__init__.py).  Written separately from the C oracle, directly from
VectorIndex/IndexBuilder.cs:23-198 and DDL.sql:246-295: numpy float32 arrays evaluate each `+ - * /` as one IEEE
binary32 operation per element, which is what the C# `float` code does; Python ints stand in for Int128.
"""
from __future__ import annotations

import numpy as np


def _trunc_div(a: int, b: int) -> int:
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b > 0) else -q


def _cmp_dotnet(a: np.float32, b: np.float32) -> int:
    # float.CompareTo: NaN lowest, -0 == +0
    if a < b:
        return -1
    if a > b:
        return 1
    if a == b:
        return 0
    if np.isnan(a):
        return 0 if np.isnan(b) else -1
    return 1


def build_literal(ids, rows):
    """Returns list of (rangeId, Dimension, Mid, Id) in the reference's emission order."""
    rows = np.asarray(rows, np.float32)
    ids = [int(i) for i in ids]
    out = []
    stack = [(0, list(range(len(ids))), True)]  # IndexBuilder.cs:33
    while stack:
        range_id, pts, mx = stack.pop()  # :37
        if not pts:
            continue  # :70-73
        mean = rows[pts[0]].copy()  # :159-173
        q = np.zeros_like(mean)
        idn = ids[pts[0]]
        count = 1
        with np.errstate(all="ignore"):
            for p in pts[1:]:  # :175-197
                v = rows[p]
                count += 1
                c = np.float32(count)
                pa = mean
                a = pa + (v - pa) / c
                q = q + (v - pa) * (v - a)
                mean = a
                idn += ids[p]
        if count == 1:
            out.append((range_id, -1, np.float32(0), idn))  # :81-82
            continue
        keys = q if mx else -q
        index = 0
        for i in range(1, len(keys)):  # MaxBy, strictly greater replaces (:77-79)
            if _cmp_dotnet(keys[i], keys[index]) > 0:
                index = i
        mid = mean[index]
        pivot = _trunc_div(idn, count)  # :87
        out.append((range_id, index, mid, pivot))
        if range_id > (2 ** 63 - 1 - 2) // 2:
            raise OverflowError("checked(rangeId * 2 + 2)")  # :99,104
        lo, hi = [], []
        for p in pts:  # :111-124
            value = rows[p, index]
            if value > mid or (value == mid and ids[p] > pivot):
                hi.append(p)
            else:
                lo.append(p)
        stack.append((2 * range_id + 1, lo, not mx))  # :128
        stack.append((2 * range_id + 2, hi, not mx))  # :129
    return out


def build_qfx(ids, rows, sql=False):
    """The fast-mode specification: exact integer sums of xi = rint(x * 2^(26-E)).

    sql=True: dbo.BuildIndex's rules (DDL.sql:44-202) over the same statistics -- `order by Stdev desc` at the root
    (:113), `iif(@level % 2 = 1, Stdev, -Stdev) desc` with @level = 0, 1, 3, 7, ... below it (:151,155): min at depth 1,
    max everywhere else; the root sends Value = Mean high unless Stdev = 0 (:104); Stdev = 0 nulls Dimension and Mid
    (:193-194), written -3 / NaN."""
    rows = np.asarray(rows, np.float32)
    ids = [int(i) for i in ids]
    finite = np.abs(rows[np.isfinite(rows)])
    amax = np.float32(finite.max()) if finite.size else np.float32(0)
    if np.isinf(rows).any():
        raise ValueError("the qfx modes refuse +-Inf (no fixed-point image); the literal mode takes it")
    if amax > 0:
        _, e = np.frexp(amax)
        e = int(e)
    else:
        e = 0
    e = min(max(e, -96), 128)
    k = np.float32(2.0) ** np.float32(26 - e)
    with np.errstate(all="ignore"):
        y = rows * k
    y = np.where(np.isnan(y), np.float32(0), y)
    xi = np.clip(np.rint(y.astype(np.float64)), -2 ** 31, 2 ** 31 - 1).astype(np.int64)
    out = []
    stack = [(0, list(range(len(ids))), True)]
    while stack:
        range_id, pts, mx = stack.pop()
        if not pts:
            continue
        n = len(pts)
        idn = sum(ids[p] for p in pts)
        if n == 1:
            out.append((range_id, -1, np.float32(0), idn))
            continue
        depth = (range_id + 1).bit_length() - 1
        if sql:
            mx = depth != 1
        stdev_zero = False
        sub = xi[pts]
        s1 = [int(v) for v in sub.sum(axis=0)]
        s2 = [sum(int(v) * int(v) for v in sub[:, j]) for j in range(sub.shape[1])]
        keys = [n * b - a * a for a, b in zip(s1, s2)]
        chosen = 0
        for i in range(1, len(keys)):
            if (keys[i] > keys[chosen]) if mx else (keys[i] < keys[chosen]):
                chosen = i
        if keys[chosen] < (n * n) << 10:
            # poorly resolved range: literal float32 Welford statistics (IndexBuilder.cs:159-197)
            mean = rows[pts[0]].copy()
            q = np.zeros_like(mean)
            with np.errstate(all="ignore"):
                for c, p in enumerate(pts[1:], start=2):
                    v = rows[p]
                    a = mean + (v - mean) / np.float32(c)
                    q = q + (v - mean) * (v - a)
                    mean = a
            fk = q if mx else -q
            index = 0
            for i in range(1, len(fk)):
                if _cmp_dotnet(fk[i], fk[index]) > 0:
                    index = i
            mid = mean[index]
            stdev_zero = sql and q[index] == 0
        else:
            index = 0
            for i in range(1, len(keys)):
                if (keys[i] > keys[index]) if mx else (keys[i] < keys[index]):
                    index = i
            mid = np.float32((np.float64(s1[index]) / np.float64(n)) * np.float64(2.0) ** (e - 26))
        pivot = _trunc_div(idn, n)
        out.append((range_id, -3, np.float32(np.nan), pivot) if stdev_zero else (range_id, index, mid, pivot))
        root_ties_high = sql and depth == 0 and not stdev_zero
        lo, hi = [], []
        for p in pts:
            value = rows[p, index]
            if value > mid or (value == mid and (root_ties_high or ids[p] > pivot)):
                hi.append(p)
            else:
                lo.append(p)
        stack.append((2 * range_id + 1, lo, not mx))
        stack.append((2 * range_id + 2, hi, not mx))
    return out


def search(table: dict, query, proximity):
    """dbo.Search over {rangeId: (Dimension, Mid, Id)} (DDL.sql:246-295). Returns ids in DFS order, low first."""
    query = np.asarray(query, np.float32)
    p = np.float32(proximity)
    out = []
    stack = [0]
    while stack:
        r = stack.pop()
        row = table.get(r)
        if row is None:
            continue
        dim, mid, rid = row
        if dim == -3:  # Dimension is null (DDL.sql:275,290): both children
            stack.append(2 * r + 2)
            stack.append(2 * r + 1)
            continue
        if dim < 0:
            out.append(rid)
            continue
        lo = np.float32(query[dim] - p)
        hi = np.float32(query[dim] + p)
        mid = np.float32(mid)
        if mid <= hi:
            stack.append(2 * r + 2)
        if mid >= lo:
            stack.append(2 * r + 1)
    return out
