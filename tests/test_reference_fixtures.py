"""The reference's own test fixtures, replayed against this index.

MempryVectorIndex.Tests/MemoryVectorIndexTests.cs holds the only executable expectations of the reference repo: for
fixed point sets, a fixed query point and a fixed distance, `index.Find(point, distance, predicate)` with the Euclidean
predicate must return exactly the records a plain scan finds (`:161-204`: no invalid match, no missed match, equal
counts).  They are written for MemoryVectorIndex<R> (the bisection trie), but the contract they state -- the index
yields a superset of the ball, the predicate verifies -- is the one `Find` keeps here over dbo.Search, so the same
fixtures pin both the CPU oracle (build + search + verify) and the CUDA path.  Distance: float32 `MathF.Sqrt(x*x + y*y)`
as `:206-214`.
"""
import numpy as np
import pytest

import oracle


def _grid(m, f):
    """records[i * m + j] = f(i, j), ids in insertion order (`records.Count`), MemoryVectorIndexTests.cs:14-27."""
    pts = np.array([f(i, j) for i in range(m) for j in range(m)], np.float32)
    return np.arange(m * m, dtype=np.int64), pts


FIXTURES = {
    # name: (points, query point, distance)   -- MemoryVectorIndexTests.cs line of the Test(...) call
    "Test_3_3": (lambda: _grid(3, lambda i, j: (i - 1, j - 1)), (0.5, 0.9), 0.6),                                   # :29
    "Test_10_10": (lambda: _grid(10, lambda i, j: ((i - 4.5) / 5, (j - 4.5) / 5)), (0.3, 0.3), 0.3),                # :49
    "Test_100_100": (lambda: _grid(100, lambda i, j: ((i - 49.5) / 50, (j - 49.5) / 50)), (0.3, 0.3), 0.1),         # :69
    "Test_1000_1000": (lambda: _grid(1000, lambda i, j: ((i - 499.5) / 500, (j - 499.5) / 500)), (0.3, 0.3), 0.05),  # :89
    "Test_100_100_NotNormalizedVectors": (lambda: _grid(100, lambda i, j: (i - 1, j - 1)), (0.3, 0.3), 0.3),         # :109
}


def _fixture(name):
    make, point, distance = FIXTURES[name]
    ids, rows = make()
    if name != "Test_3_3" and name != "Test_100_100_NotNormalizedVectors":
        # C#: `(i - 4.5f) / 5` is float32 arithmetic throughout
        m = int(round(np.sqrt(len(ids))))
        half, div = {10: (4.5, 5), 100: (49.5, 50), 1000: (499.5, 500)}[m]
        ax = ((np.arange(m, dtype=np.float32) - np.float32(half)) / np.float32(div)).astype(np.float32)
        rows = np.stack(np.meshgrid(ax, ax, indexing="ij"), -1).reshape(-1, 2).astype(np.float32)
    return ids, rows, np.array(point, np.float32), np.float32(distance)


def _plain_match(ids, rows, point, distance):
    # `records.Where(record => Distance(record.vector, point) <= distance)`, :163-165 with Distance of :206-214
    x = (rows[:, 0] - point[0]).astype(np.float32)
    y = (rows[:, 1] - point[1]).astype(np.float32)
    d = np.sqrt((x * x + y * y).astype(np.float32)).astype(np.float32)
    return set(ids[d <= distance].tolist())


@pytest.mark.parametrize("mode", [oracle.MODE_LITERAL, oracle.MODE_QFX, oracle.MODE_SQL])
@pytest.mark.parametrize("name", [n for n in FIXTURES if n != "Test_1000_1000"])
def test_reference_fixture_oracle(name, mode):
    ids, rows, point, distance = _fixture(name)
    plain = _plain_match(ids, rows, point, distance)
    table = oracle.build(ids, rows, mode)
    assert int((table.dimension == -1).sum()) == len(ids)            # `Assert.AreEqual(index.Count, records.Count)` :159
    offs, cand, _ = oracle.search(table, point[None, :], float(distance))
    by_id = {int(i): r for i, r in zip(ids, rows)}
    match = {int(c) for c in cand if oracle.distance_l2(by_id[int(c)], point) <= distance}
    assert match == plain                                            # :196-202
    assert len(plain) > 0 or name == "Test_100_100_NotNormalizedVectors"   # (that one matches nothing: 0 == 0)


@pytest.mark.gpu
@pytest.mark.parametrize("mode_name", ["exact", "fast", "sql"])
@pytest.mark.parametrize("name", list(FIXTURES))
def test_reference_fixture_gpu(name, mode_name):
    import vectorindex as vi
    mode = {"exact": vi.MODE_EXACT, "fast": vi.MODE_FAST, "sql": vi.MODE_SQL}[mode_name]
    ids, rows, point, distance = _fixture(name)
    plain = _plain_match(ids, rows, point, distance)
    index = vi.VectorIndex(ids, rows, mode=mode)
    try:
        assert index.Count == len(ids)
        calls = []

        def predicate(i, v):                                          # the test's predicate counts its calls, :170-176
            calls.append(i)
            return oracle.distance_l2(rows[i], point) <= distance

        match = list(index.Find(point, float(distance), predicate))
        assert set(match) == plain and len(match) == len(plain)       # no invalid match, no missed match, no duplicates
        assert len(calls) >= len(plain)
        # the library's own Euclidean verification kernel gives the same set
        assert set(index.Find(point, float(distance))) == plain
    finally:
        index.close()
