"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.
Bit-exact for both modes: exact mode vs the literal oracle, fast mode vs the qfx specification oracle."""
import numpy as np
import pytest

import oracle
import vectorindex as vi
from vectorindex import synthetic as ds

pytestmark = pytest.mark.gpu


def gpu_table(ids, rows, mode):
    with vi.Context(0) as ctx:
        ctx.reserve(len(ids), rows.shape[1])
        ctx.add(ids, rows)
        info = ctx.build(mode)
        rid, dim, mid, oid = ctx.ranges()
    assert len(np.unique(rid)) == len(rid), "a RangeID must appear once"
    order = np.argsort(rid, kind="stable")  # row order is unspecified (consumers key by rangeId, Program.cs:18-26)
    return rid[order], dim[order], mid[order], oid[order], info


def assert_same_table(ids, rows, mode):
    rid, dim, mid, oid, info = gpu_table(ids, rows, mode)
    ref = oracle.build(ids, rows, mode)
    assert len(rid) == len(ref)
    assert np.array_equal(rid, ref.range_id)
    assert np.array_equal(dim, ref.dimension)
    assert np.array_equal(mid.view(np.uint32), ref.mid.view(np.uint32)), "Mid must be bit-identical"
    assert np.array_equal(oid, ref.id)
    return info


MODES = [vi.MODE_EXACT, vi.MODE_FAST]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("n,d", [(1, 4), (2, 4), (3, 1), (5, 7), (33, 3), (600, 96), (1000, 33), (4097, 16)])
def test_small_shapes(mode, n, d):
    ids, rows = ds.uniform(n, d, seed=n + d)
    ids = (ids * 7 + 3)
    assert_same_table(ids, rows, mode)


@pytest.mark.parametrize("mode", MODES)
def test_empty(mode):
    with vi.Context(0) as ctx:
        ctx.reserve(0, 8)
        ctx.build(mode)
        assert ctx.range_count == 0


@pytest.mark.parametrize("mode", MODES)
def test_config1_100k_x_96_uniform(mode):
    # BASELINE.json configs[0]: MainTest-shaped build (Program.cs:163-181) at 100k x 96
    ids, rows = ds.uniform(100_000, 96, seed=1)
    info = assert_same_table(ids, rows, mode)
    assert info.ranges == 2 * 100_000 - 1


@pytest.mark.parametrize("mode", MODES)
def test_unit_gaussian_50k_x_96(mode):
    ids, rows = ds.unit_gaussian(50_000, 96, seed=2)
    assert_same_table(ids, rows, mode)


@pytest.mark.parametrize("mode", MODES)
def test_wide_rows_768(mode):
    # configs[4] shape (text-embedding width), reduced count
    ids, rows = ds.unit_gaussian(6000, 768, seed=3)
    assert_same_table(ids, rows, mode)


@pytest.mark.parametrize("mode", MODES)
def test_config5_width_200k_x_768(mode):
    # BASELINE configs[4] (1M x 768) at a size the oracle finishes in seconds: every range class at the full row width
    ids, rows = ds.unit_gaussian(200_000, 768, seed=31)
    info = assert_same_table(ids * 5 - 17, rows, mode)
    assert info.ranges >= 2 * 200_000 - 1


@pytest.mark.parametrize("mode", MODES)
def test_one_hot_crafted_set(mode):
    # Program.cs:54-66; literal mode must pick dimension 3 at the root, qfx dimension 0
    ids, rows = ds.one_hot(1536)
    rid, dim, mid, oid, _ = gpu_table(ids, rows, mode)
    assert dim[0] == (3 if mode == vi.MODE_EXACT else 0)
    assert_same_table(ids, rows, mode)


@pytest.mark.parametrize("mode", MODES)
def test_duplicates_constant_columns_negative_ids(mode):
    rng = np.random.default_rng(9)
    base = rng.random((700, 12), dtype=np.float32)
    rows = np.concatenate([base, base, base], 0)
    rows[:, 5] = 0.125
    ids = (np.arange(2100, dtype=np.int64)[::-1] - 1000) * 1_000_000_007
    assert_same_table(ids, rows, mode)


@pytest.mark.parametrize("mode", MODES)
def test_scaled_data(mode):
    ids, rows = ds.uniform(5000, 20, seed=11)
    assert_same_table(ids, (rows * np.float32(1e4)).astype(np.float32), mode)
    assert_same_table(ids, (rows * np.float32(1e-5)).astype(np.float32), mode)


@pytest.mark.parametrize("mode", MODES)
def test_skewed_tree(mode):
    # exponential spread: very unbalanced splits, deep tree, big and tiny ranges on the same level
    rng = np.random.default_rng(4)
    rows = np.exp(rng.standard_normal((20000, 6)) * 3).astype(np.float32)
    ids = np.arange(20000, dtype=np.int64)
    assert_same_table(ids, rows, mode)


@pytest.mark.parametrize("mode", MODES)
def test_inseparable_points_overflow(mode):
    # same vector, same id: never separate -> OverflowException at depth 62 (IndexBuilder.cs:99)
    rows = np.full((2, 3), 0.5, np.float32)
    ids = np.array([-4, -5], np.int64)
    with vi.Context(0) as ctx:
        ctx.reserve(2, 3)
        ctx.add(ids, rows)
        with pytest.raises(OverflowError):
            ctx.build(mode)


def test_wrong_vector_length_is_argument_error():
    with vi.Context(0) as ctx:
        ctx.reserve(4, 8)
        with pytest.raises(ValueError, match="Invalid length of vector"):
            ctx.add(np.arange(4), np.zeros((4, 7), np.float32))


def test_incremental_add_equals_single_add():
    ids, rows = ds.uniform(3000, 10, seed=5)
    with vi.Context(0) as ctx:
        ctx.reserve(10, 10)  # forces growth
        for s in range(0, 3000, 700):
            ctx.add(ids[s:s + 700], rows[s:s + 700])
        ctx.build(vi.MODE_EXACT)
        rid, dim, mid, oid = ctx.ranges()
    o = np.argsort(rid)
    rid, dim, mid, oid = rid[o], dim[o], mid[o], oid[o]
    ref = oracle.build(ids, rows, oracle.MODE_LITERAL)
    assert np.array_equal(rid, ref.range_id) and np.array_equal(dim, ref.dimension)
    assert np.array_equal(mid.view(np.uint32), ref.mid.view(np.uint32)) and np.array_equal(oid, ref.id)


def test_index_builder_mirror_yields_reference_rows():
    ids, rows = ds.uniform(500, 5, seed=8)
    pts = [(int(i), rows[k]) for k, i in enumerate(ids)]
    got = dict(vi.IndexBuilder.Build(pts, lambda rangeId, capacity: vi.MemoryRangeStore()))
    ref = oracle.build(ids, rows).as_dict()
    assert got.keys() == ref.keys()
    for k, v in got.items():
        assert (v.Dimension, v.Id) == (ref[k][0], ref[k][2])
        assert np.float32(v.Mid).tobytes() == np.float32(ref[k][1]).tobytes()


# ---- search ------------------------------------------------------------------------------------------------------
def _sorted_sets(offs, ids):
    return [np.sort(ids[offs[i]:offs[i + 1]]) for i in range(len(offs) - 1)]


@pytest.mark.parametrize("p", [0.0, 0.01, 0.05, 0.2])
def test_search_matches_oracle_traversal(p):
    ids, rows = ds.unit_gaussian(30_000, 96, seed=2)
    _, fresh = ds.unit_gaussian(300, 96, seed=77)
    queries = np.concatenate([rows[:300], fresh], 0)
    with vi.Context(0) as ctx:
        ctx.reserve(len(ids), 96)
        ctx.add(ids, rows)
        ctx.build(vi.MODE_EXACT)
        offs, out = ctx.search(queries, p)
    ref = oracle.build(ids, rows)
    roffs, rout, _ = oracle.search(ref, queries, p)
    assert np.array_equal(offs, roffs)
    assert np.array_equal(out, rout)  # same DFS order, not only the same sets
    if p == 0.0:
        for i in range(300):
            assert ids[i] in out[offs[i]:offs[i + 1]]


@pytest.mark.parametrize("mode", MODES)
def test_search_csr_parity_on_1m_point_index(mode):
    # a table with sub-tree blocks, one-child rows and 21+ levels: CSR arrays (not only sets) equal the oracle's
    n = 1_000_000
    ids, rows = ds.unit_gaussian(n, 96, seed=2)
    _, fresh = ds.unit_gaussian(500, 96, seed=78)
    queries = np.concatenate([rows[:250], rows[n - 250:], fresh], 0)
    ref = oracle.build(ids, rows, mode)
    with vi.Context(0) as ctx:
        ctx.reserve(n, 96)
        ctx.add(ids, rows)
        ctx.build(mode)
        for p in (0.0, 0.01, 0.03):
            offs, out = ctx.search(queries, p)
            roffs, rout, _ = oracle.search(ref, queries, p)
            assert np.array_equal(offs, roffs)
            assert np.array_equal(out, rout)
        # p = 0 finds every data point that is queried
        offs, out = ctx.search(queries[:500], 0.0)
        want = np.concatenate([ids[:250], ids[n - 250:]])
        for i in range(500):
            assert want[i] in out[offs[i]:offs[i + 1]]


def test_search_verify_equals_bruteforce():
    # MemoryVectorIndexTests.cs:161-204 property: Find + Euclidean predicate == brute force
    ids, rows = ds.grid2d(100)
    q = np.array([[0.1, -0.3], [0.9, 0.9], [-1.0, 0.0]], np.float32)
    dist = 0.07
    with vi.Context(0) as ctx:
        ctx.reserve(len(ids), 2)
        ctx.add(ids, rows)
        ctx.build(vi.MODE_EXACT)
        offs, out = ctx.search_verify(q, dist, dist)
    for i in range(q.shape[0]):
        want = sorted(int(k) for k in ids if oracle.distance_l2(rows[k], q[i]) <= np.float32(dist))
        assert sorted(out[offs[i]:offs[i + 1]].tolist()) == want
        assert len(want) > 0


def test_find_entry_point_shape():
    ids, rows = ds.grid2d(40)
    index = vi.VectorIndex(ids, rows)
    q = np.array([0.2, 0.2], np.float32)
    dist = 0.1
    calls = []

    def predicate(i, v):
        calls.append(i)
        return oracle.distance_l2(v, q) <= np.float32(dist)

    got = sorted(index.Find(q, dist, predicate))
    want = sorted(int(k) for k in ids if oracle.distance_l2(rows[k], q) <= np.float32(dist))
    assert got == want and len(calls) >= len(want)
    assert sorted(index.Find(q, dist)) == want
    with pytest.raises(ValueError, match="Invalid vector size"):
        list(index.Find(np.zeros(3, np.float32), dist))
    index.close()


def test_textindex_rows():
    ids, rows = ds.uniform(200, 4, seed=3)
    with vi.Context(0) as ctx:
        ctx.reserve(200, 4)
        ctx.add(ids, rows)
        ctx.build(vi.MODE_EXACT)
        rid, dim, mid, lo, hi, tid = ctx.textindex()
    have = set(rid.tolist())
    for k in range(len(rid)):
        if dim[k] < 0:
            assert tid[k] >= 0 and lo[k] == -1 and hi[k] == -1 and np.isnan(mid[k])
        else:
            assert tid[k] == -1
            assert lo[k] in (-1, 2 * rid[k] + 1) and hi[k] in (-1, 2 * rid[k] + 2)
            assert (lo[k] == -1) == (2 * rid[k] + 1 not in have)
            assert (hi[k] == -1) == (2 * rid[k] + 2 not in have)


def test_fast_division_is_correctly_rounded():
    # exact mode replaces (value - pa) / count by a reciprocal + two Markstein corrections (vi_stats_exact.cuh);
    # the device self-test compares it with IEEE division on 2e9 operand pairs
    with vi.Context(0) as ctx:
        assert ctx.divcheck(12345, 2_000_000_000) == 0
        assert ctx.divcheck(987654321, 500_000_000) == 0


@pytest.mark.parametrize("mode", MODES)
def test_mid_size_ranges_all_classes(mode):
    # 40k x 96: exercises chunked (>= 4096), warp-per-range and team-per-range classes of the fast mode
    ids, rows = ds.uniform(40_000, 96, seed=21)
    assert_same_table(ids, rows, mode)
    ids, rows = ds.unit_gaussian(9000, 128, seed=22)   # FULL <8,4> shape
    assert_same_table(ids, rows, mode)
    ids, rows = ds.unit_gaussian(5000, 100, seed=23)   # padded row width, guarded columns
    assert_same_table(ids, rows, mode)


@pytest.mark.parametrize("mode", MODES)
def test_cpp_host_mirror_main_test(mode):
    # cpp/main_test.cpp: Program.cs:54-66 crafted set through the C++ mirror of IndexBuilder.Build
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(vi.LIB_PATH), "main_test")
    assert os.path.exists(exe), "make -C vector-database_b200/csrc"
    out = subprocess.run([exe, str(mode), "256"], check=True, capture_output=True, text=True).stdout
    got = {}
    for line in out.strip().splitlines():
        r, d, bits, i = line.split(",")
        got[int(r)] = (int(d), int(bits), int(i))
    ids, rows = ds.one_hot(256)
    ref = oracle.build(ids, rows, mode)
    want = {int(r): (int(d), int(np.float32(m).view(np.uint32)), int(i))
            for r, d, m, i in zip(ref.range_id, ref.dimension, ref.mid, ref.id)}
    assert got == want


def test_cpp_host_mirror_hdf5_path(tmp_path):
    # cpp/main_test.cpp <mode> hdf5 <file>: Program.cs:86-150 -- `/train` of an HDF5 file through the native reader
    import os
    import subprocess
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from h5_writer import write_v0
    ids, rows = ds.unit_gaussian(5000, 32, seed=31)
    path = str(tmp_path / "train.hdf5")
    write_v0(path, {"train": rows, "test": rows[:4]})
    exe = os.path.join(os.path.dirname(vi.LIB_PATH), "main_test")
    for mode, omode in ((vi.MODE_FAST, oracle.MODE_QFX), (vi.MODE_SQL, oracle.MODE_SQL)):
        out = subprocess.run([exe, str(mode), "hdf5", path], check=True, capture_output=True, text=True).stdout
        got = {}
        for line in out.strip().splitlines():
            r, d, bits, i = line.split(",")
            got[int(r)] = (int(d), int(bits), int(i))
        ref = oracle.build(np.arange(5000, dtype=np.int64), rows, omode)   # ids = row indexes, Program.cs:252
        want = {int(r): (int(d), int(np.float32(m).view(np.uint32)), int(i))
                for r, d, m, i in zip(ref.range_id, ref.dimension, ref.mid, ref.id)}
        assert got == want
    bad = subprocess.run([exe, "1", "hdf5", path, "/nope"], capture_output=True, text=True)
    assert bad.returncode == 1 and "no object named" in bad.stderr


@pytest.mark.parametrize("mode", MODES)
def test_nan_rows_overflow_like_the_reference(mode):
    # NaN components never compare greater than Mid, such points go low forever and the reference
    # dies with OverflowException at depth 62 (IndexBuilder.cs:99).  Oracle and CUDA path must agree on that.
    ids, rows = ds.uniform(300, 5, seed=2)
    rows = rows.copy()
    rows[10] = np.nan
    rows[11] = np.nan
    with pytest.raises(OverflowError):
        oracle.build(ids, rows, mode)
    with vi.Context(0) as ctx:
        ctx.reserve(300, 5)
        ctx.add(ids, rows)
        with pytest.raises(OverflowError):
            ctx.build(mode)


@pytest.mark.parametrize("mode", MODES)
def test_single_nan_point_is_isolated(mode):
    # one NaN point ends up alone in a leaf: the build succeeds and matches the oracle
    ids, rows = ds.uniform(200, 4, seed=6)
    rows = rows.copy()
    rows[7, 2] = np.nan
    assert_same_table(ids, rows, mode)


def test_rebuild_and_mode_switch_on_one_context():
    # buffers are reused across builds and modes; results must not depend on what ran before
    ids, rows = ds.unit_gaussian(20_000, 96, seed=31)
    with vi.Context(0) as ctx:
        ctx.reserve(len(ids), 96)
        ctx.add(ids, rows)
        got = {}
        for mode in (vi.MODE_FAST, vi.MODE_EXACT, vi.MODE_FAST, vi.MODE_EXACT):
            ctx.build(mode)
            rid, dim, mid, oid = ctx.ranges()
            o = np.argsort(rid)
            cur = (rid[o], dim[o], mid[o].view(np.uint32), oid[o])
            if mode in got:
                assert all(np.array_equal(a, b) for a, b in zip(got[mode], cur))
            got[mode] = cur
        # smaller data set on the same context afterwards
        ctx.reserve(100, 96)
        ctx.add(ids[:100], rows[:100])
        ctx.build(vi.MODE_EXACT)
        rid, dim, mid, oid = ctx.ranges()
    ref = oracle.build(ids[:100], rows[:100])
    o = np.argsort(rid)
    assert np.array_equal(rid[o], ref.range_id) and np.array_equal(mid[o].view(np.uint32), ref.mid.view(np.uint32))


# ---- exact mode, top-level pipeline (vi_stats_exact_px.cuh): restart, safe-mode and tail paths ----------------------
def _exact_equals_oracle(ids, rows):
    assert_same_table(ids, rows, vi.MODE_EXACT)


@pytest.mark.parametrize("n", [513, 544, 545, 4096 + 17, 30_001])
def test_exact_pipeline_group_tails(n):
    # (n - 1) % 32 covers 0, 31, 1 and odd sizes: full groups go through the pipeline, the tail through the safe steps
    ids, rows = ds.unit_gaussian(n, 64, seed=n)
    _exact_equals_oracle(ids, rows)


def test_exact_pipeline_single_nan_restarts_then_safe_mode():
    # one NaN component poisons its chain from that point on: every later group fails the quotient check, the pipeline
    # restarts 16 times and then the variance warp finishes the range alone (same float32 recurrence, bit for bit)
    ids, rows = ds.unit_gaussian(20_000, 64, seed=77)
    rows = rows.copy()
    rows[1234, 5] = np.nan
    _exact_equals_oracle(ids, rows)


def test_exact_pipeline_tiny_zero_and_offset_operands():
    ids, rows = ds.unit_gaussian(24_000, 96, seed=78)
    rows = rows.copy()
    rows[:, 3] = 0.0                                   # d == 0 on every step
    rows[:, 4] = np.float32(1e-40)                     # denormal constant
    rows[::7, 5] = np.float32(3e-39)                   # denormal / zero mix: tiny non-zero differences (guard)
    rows[1::7, 5] = 0.0
    rows[:, 6] += np.float32(4096.0)                   # mean >> spread: quotients with heavy cancellation
    rows[::1000, 7] = np.float32(3e38)                 # huge values: d overflows to inf in places
    rows[:, 8] *= np.float32(1e-30)
    _exact_equals_oracle(ids, rows)


def test_exact_pipeline_count_with_all_ones_significand():
    # Count + 1 = 2^24 - 1 has an all-ones significand: the one float for which the speculative quotient's check is
    # not a proof, so its group takes the safe path (vi_stats_exact.cuh).  Needs > 2^24 points: 4 narrow dimensions.
    n = (1 << 24) + 100
    rng = np.random.default_rng(5)
    rows = rng.standard_normal((n, 4), dtype=np.float32)
    ids = np.arange(n, dtype=np.int64)
    _exact_equals_oracle(ids, rows)


# ---- formats on either side of the path: record ingest, range-table import ----------------
def _same_rows(got, want):
    rid, dim, mid, oid = got
    o = np.argsort(rid, kind="stable")
    assert np.array_equal(rid[o], want[0]) and np.array_equal(dim[o], want[1]) and np.array_equal(oid[o], want[3])
    assert np.array_equal(mid[o].view(np.uint32), want[2].view(np.uint32))


@pytest.mark.parametrize("mode", MODES)
def test_record_ingest_equals_array_ingest(mode, tmp_path):
    # the FileRangeStore record [int64 id][dims x float32] from a buffer and streamed from a file (several batches,
    # an offset, a padded row width) gives the same table as vi_points_add
    ids, rows = ds.unit_gaussian(70_000, 50, seed=41)
    ids = ids * 5 - 1000
    want = gpu_table(ids, rows, mode)
    rec = vi.pack_records(ids, rows)
    path = tmp_path / "points.bin"
    with open(path, "wb") as f:
        f.write(b"x" * 24)          # a header the caller skips
        f.write(rec.tobytes())
    with vi.Context(0) as ctx:
        ctx.reserve(10, 50)
        ctx.add_records(rec[: 1000 * (8 + 200)], 50)
        ctx.add_records(rec[1000 * (8 + 200):], 50)
        assert ctx.count == 70_000
        ctx.build(mode)
        _same_rows(ctx.ranges(), want)
    with vi.Context(0) as ctx:
        ctx.reserve(0, 50)
        read_ms, total_ms = ctx.add_file(str(path), 50, offset_bytes=24)
        assert ctx.count == 70_000 and total_ms > 0
        ctx.build(mode)
        _same_rows(ctx.ranges(), want)
        with pytest.raises(ValueError):
            ctx.add_file(str(path), 50, offset_bytes=24, n=70_001)   # shorter than asked
        with pytest.raises(ValueError):
            ctx.add_file(str(path), 49, offset_bytes=24)             # "Invalid length of vector."


def test_range_table_csv_round_trip_and_search(tmp_path):
    # build -> CSV "RangeID,Dimension,Mid,ID" (Program.cs:145-149) -> shuffled import into a fresh context -> the same
    # dbo.Search results, array for array
    ids, rows = ds.unit_gaussian(30_000, 32, seed=43)
    q = np.concatenate([rows[:200], ds.unit_gaussian(200, 32, seed=44)[1]])
    with vi.Context(0) as ctx:
        ctx.reserve(len(ids), 32)
        ctx.add(ids, rows)
        ctx.build(vi.MODE_EXACT)
        rid, dim, mid, oid = ctx.ranges()
        want0 = ctx.search(q, 0.0)
        want1 = ctx.search(q, 0.05)
        ti = ctx.textindex()
    path = str(tmp_path / "index.csv")
    vi.write_csv(path, rid, dim, mid, oid)
    r2, d2, m2, o2 = vi.read_csv(path)
    assert np.array_equal(r2, rid) and np.array_equal(d2, dim) and np.array_equal(o2, oid)
    assert np.array_equal(m2.view(np.uint32), mid.view(np.uint32))      # text round trip is exact
    perm = np.random.default_rng(1).permutation(len(r2))
    with vi.Context(0) as ctx:
        ctx.load_ranges(r2[perm], d2[perm], m2[perm], o2[perm], 32)
        assert ctx.range_count == len(rid)
        for want, p in ((want0, 0.0), (want1, 0.05)):
            offs, out = ctx.search(q, p)
            assert np.array_equal(offs, want[0]) and np.array_equal(out, want[1])
        # same dbo.TextIndex rows (keyed by RangeID)
        t2 = ctx.textindex()
        o1, o2_ = np.argsort(ti[0]), np.argsort(t2[0])
        for a, b in zip(ti, t2):
            assert np.array_equal(a[o1], b[o2_], equal_nan=True)
        with pytest.raises(vi.VectorIndexError):
            ctx.search_verify(q[:4], 0.05, 0.1)                           # no vectors behind an imported table
        # malformed tables
        with pytest.raises(ValueError):
            ctx.load_ranges(np.array([0, 1, 1]), np.array([0, -1, -1]), np.zeros(3, np.float32), np.array([1, 2, 3]), 32)
        with pytest.raises(ValueError):
            ctx.load_ranges(np.array([1, 2]), np.array([-1, -1]), np.zeros(2, np.float32), np.array([1, 2]), 32)
        with pytest.raises(ValueError):
            ctx.load_ranges(np.array([0]), np.array([32]), np.zeros(1, np.float32), np.array([1]), 32)
        ctx.load_ranges(np.zeros(0, np.int64), np.zeros(0, np.int32), np.zeros(0, np.float32), np.zeros(0, np.int64), 32)
        assert ctx.range_count == 0


# ---- top-k over the candidates ------------------------------------------------------------------
@pytest.mark.parametrize("metric", [0, 1])
def test_topk_equals_oracle_candidates_sorted_by_distance(metric):
    # oracle: dbo.Search candidates in traversal order, float32 distances accumulated in index order, stable sort
    ids, rows = ds.unit_gaussian(20_000, 16, seed=51)
    ids = ids * 3 + 1
    row_of = {int(i): r for r, i in enumerate(ids)}
    q = np.concatenate([rows[:30], ds.unit_gaussian(30, 16, seed=52)[1]])
    table = oracle.build(ids, rows, vi.MODE_EXACT)
    p, k = 0.25, 10
    offs, cand, _ = oracle.search(table, q, p)
    fn = oracle.distance_l2 if metric == 0 else oracle.distance_angular
    with vi.Context(0) as ctx:
        ctx.reserve(len(ids), 16)
        ctx.add(ids, rows)
        ctx.build(vi.MODE_EXACT)
        got_ids, got_dist, got_cnt, ncand = ctx.search_topk(q, p, k, metric)
        assert ncand == len(cand)
        few_ids, few_dist, few_cnt, _ = ctx.search_topk(q, 0.0, 4, metric)   # p = 0: at most one candidate per query
    assert (got_cnt > 3).any()
    for i in range(len(q)):
        c = cand[offs[i]:offs[i + 1]]
        d = np.array([fn(rows[row_of[int(x)]], q[i]) for x in c], np.float32)
        order = np.argsort(d, kind="stable")[:k]
        n = len(order)
        assert got_cnt[i] == n
        assert np.array_equal(got_ids[i, :n], c[order])
        assert np.array_equal(got_dist[i, :n].view(np.uint32), d[order].view(np.uint32))
        assert np.all(got_ids[i, n:] == -1) and np.all(np.isinf(got_dist[i, n:]))
    assert np.all(few_cnt[:30] == 1) and np.array_equal(few_ids[:30, 0], ids[:30]) and np.all(few_ids[:, 1:] == -1)


def test_topk_recall_reaches_one_when_the_box_covers_the_neighbours():
    # the traversal returns a superset of the L-infinity box of half-width p: with p >= the k-th neighbour's Euclidean
    # distance the top-k over candidates IS the exact k-NN (brute force, float64)
    ids, rows = ds.unit_gaussian(5000, 8, seed=53)
    q = ds.unit_gaussian(40, 8, seed=54)[1]
    k = 5
    d2 = ((rows[None, :, :].astype(np.float64) - q[:, None, :].astype(np.float64)) ** 2).sum(axis=2)
    exact = np.argsort(d2, axis=1, kind="stable")[:, :k]
    p = float(np.sqrt(np.take_along_axis(d2, exact, axis=1)[:, -1].max())) * 1.0001
    with vi.Context(0) as ctx:
        ctx.reserve(len(ids), 8)
        ctx.add(ids, rows)
        ctx.build(vi.MODE_FAST)
        got_ids, _, cnt, _ = ctx.search_topk(q, p, k, 0)
        with pytest.raises(ValueError):
            ctx.search_topk(q, p, 0, 0)
        with pytest.raises(ValueError):
            ctx.search_topk(q, p, 5, 7)
    assert np.all(cnt == k)
    for i in range(len(q)):
        assert set(got_ids[i].tolist()) == set(ids[exact[i]].tolist())


@pytest.mark.parametrize("mode", MODES)
def test_one_million_points_bit_exact(mode):
    # 1M x 96: 20+ levels -- the exact mode's pipeline kernel on six levels, the fast mode's sibling derivation on
    # ten, every range class and ~47 k sub-trees -- against the oracle row for row (the largest size the oracle
    # finishes in seconds; beyond it tests/test_gpu_fullsize.py checks properties)
    ids, rows = ds.unit_gaussian(1_000_000, 96, seed=61)
    info = assert_same_table(ids, rows, mode)
    assert info.levels >= 20


# ---- warp-per-query traversal: forced paths, stack spill, pool overflow ---------------------------------------------
_SEARCH_PATHS_CASE = []


def _search_paths_case():
    if not _SEARCH_PATHS_CASE:
        ids, rows = ds.unit_gaussian(60_000, 24, seed=12)
        _, fresh = ds.unit_gaussian(200, 24, seed=13)
        queries = np.concatenate([rows[:200], fresh], 0)
        ref = oracle.build(ids, rows, oracle.MODE_QFX)
        want = {p: oracle.search(ref, queries, p)[:2] for p in (0.0, 0.02, 0.1, 0.5)}
        _SEARCH_PATHS_CASE.append((ids, rows, queries, ref, want))
    return _SEARCH_PATHS_CASE[0]


@pytest.mark.parametrize("env", [
    {"VI_B200_SEARCH_PATH": "0"},                                        # one thread per query, count + fill
    {"VI_B200_SEARCH_PATH": "1"},                                        # warp per query, candidate pool + gather
    {"VI_B200_SEARCH_PATH": "1", "VI_B200_SEARCH_POOL": "0"},            # warp per query, count + fill
    {"VI_B200_SEARCH_PATH": "1", "VI_B200_SEARCH_STACK": "64"},          # shared stack spills to global memory
    {"VI_B200_SEARCH_PATH": "1", "VI_B200_SEARCH_POOL_SLOTS": "4096"},   # pool overflows: the rest is walked again
])
def test_search_paths_agree_with_the_oracle(env, monkeypatch):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    ids, rows, queries, ref, want = _search_paths_case()
    with vi.Context(0) as ctx:
        ctx.reserve(len(ids), 24)
        ctx.add(ids, rows)
        ctx.build(vi.MODE_FAST)
        for p in (0.0, 0.02, 0.1, 0.5):
            roffs, rout = want[p]
            offs, out = ctx.search(queries, p)
            assert np.array_equal(offs, roffs), p
            assert np.array_equal(out, rout), p
            offs, out = ctx.search_two_call(queries, p)
            assert np.array_equal(offs, roffs) and np.array_equal(out, rout), p
        # p = 0.5 on 24 dims: tens of thousands of candidates per query (deep pending lists)
        assert len(rout) > 200 * 10_000
        # verification and top-k walk with source rows (no pool)
        dist = 0.25
        voffs, vout = ctx.search_verify(queries[:50], 0.1, dist)
        roffs, rout, _ = oracle.search(ref, queries[:50], 0.1)
        for i in range(50):
            cand = rout[roffs[i]:roffs[i + 1]]
            want = [int(k) for k in cand if oracle.distance_l2(rows[k], queries[i]) <= np.float32(dist)]
            assert vout[voffs[i]:voffs[i + 1]].tolist() == want


@pytest.mark.parametrize("mode", [vi.MODE_EXACT, vi.MODE_FAST, vi.MODE_SQL])
@pytest.mark.parametrize("n", [5, 3000, 400_000])
def test_build_copy_equals_build_then_ranges_copy(mode, n):
    # vi_build_copy: at 400k points the sub-tree list is taken in four slices whose row blocks are copied out while the
    # next slice runs; the delivered table must be the one vi_ranges_copy returns, row for row
    import torch
    ids, rows = ds.unit_gaussian(n, 24, seed=n)
    ids = ids * 2 + 1
    cap = 2 * n + n // 8 + 1024
    pin = [torch.empty(cap, dtype=t, pin_memory=True).numpy() for t in (torch.int64, torch.int32, torch.float32, torch.int64)]
    for a in pin:
        a[:] = 0
    with vi.Context(0) as ctx:
        ctx.reserve(n, 24)
        ctx.add(ids, rows)
        info, k = ctx.build_into(mode, *pin)
        want = ctx.ranges()
        assert k == info.ranges == len(want[0])
        for got, w in zip(pin, want):
            assert np.array_equal(got[:k].view(np.uint8), w.view(np.uint8))
        # too small a destination: the build is complete, the copy is refused
        small = [np.zeros(10, a.dtype) for a in pin]
        if k > 10:
            with pytest.raises(vi.VectorIndexError):
                ctx.build_into(mode, *small)
            assert np.array_equal(ctx.ranges()[0], want[0])


def test_fast_modes_refuse_infinity_exact_mode_takes_it():
    # +-Inf has no fixed-point image: the qfx specification (oracle modes 1, 2) and the CUDA fast / SQL modes refuse it
    # alike; the literal mode runs the reference's arithmetic on it (NaN statistics and all)
    ids, rows = ds.unit_gaussian(3000, 8, seed=21)
    rows = rows.copy()
    rows[17, 3] = np.float32(np.inf)
    rows[900, 5] = np.float32(-np.inf)
    for omode in (oracle.MODE_QFX, oracle.MODE_SQL):
        with pytest.raises(ValueError):
            oracle.build(ids, rows, omode)
    with vi.Context(0) as ctx:
        ctx.reserve(len(ids), 8)
        ctx.add(ids, rows)
        for mode in (vi.MODE_FAST, vi.MODE_SQL):
            with pytest.raises(ValueError, match="Inf"):
                ctx.build(mode)
    assert_same_table(ids, rows, vi.MODE_EXACT)
    one_ids, one_rows = ids[:1], np.full((1, 8), np.inf, np.float32)   # a single point is a leaf whatever its value
    assert_same_table(one_ids, one_rows, vi.MODE_FAST)
