"""VI_MODE_SQL / oracle mode 2: the rules of the T-SQL builder dbo.BuildIndex (DDL.sql:44-202) -- SURVEY.md 8(f) rank 4.

Known answers worked out by hand from the SQL text: the C oracle and the independent numpy restatement must reproduce
them on the CPU, the CUDA path on the GPU; then bit-exact parity of the CUDA path against the oracle on seeded inputs."""
import numpy as np
import pytest

import oracle
from oracle import np_oracle

NAN = float("nan")


def _as_dict(rid, dim, mid, oid):
    return {int(r): (int(d), float(m), int(i)) for r, d, m, i in zip(rid, dim, mid, oid)}


def _same(got: dict, want: dict):
    assert sorted(got) == sorted(want)
    for r, (d, m, i) in want.items():
        gd, gm, gi = got[r]
        assert (gd, gi) == (d, i), (r, got[r], want[r])
        assert (np.isnan(gm) and np.isnan(m)) or np.float32(gm) == np.float32(m), (r, got[r], want[r])


# ---- hand-derived known answers (DDL.sql line numbers in the comments) ------------------------------------------------
def lattice16():
    """Full factorial of dim0 = +-0.4, dim1 = +-0.2, dim2 = +-0.1, dim3 = +-0.05; ids 0..15, dim0 outermost."""
    pts = [[a, b, c, e] for a in (-0.4, 0.4) for b in (-0.2, 0.2) for c in (-0.1, 0.1) for e in (-0.05, 0.05)]
    return np.arange(16, dtype=np.int64), np.array(pts, np.float32)


LATTICE16_WANT = {
    # depth 0 (:113 `order by Stdev desc`): dim 0, Mean 0, avg(ID) = 120 / 16 = 7
    0: (0, 0.0, 7),
    # depth 1 (:151 with @level = 0: `-Stdev desc` = MINIMUM): dim 0 is constant inside both children, Stdev = 0 wins;
    # such a row has Dimension = null, Mid = null (:193-194) and its points split on ID <= avg(ID) (:166)
    1: (-3, NAN, 3), 2: (-3, NAN, 11),
    # depth 2 (@level = 1: max): ids {0..3}, {4..7}, {8..11}, {12..15} have dims 0 and 1 constant: max is dim 2
    3: (2, 0.0, 1), 4: (2, 0.0, 5), 5: (2, 0.0, 9), 6: (2, 0.0, 13),
    # depth 3 (@level = 3: 3 % 2 = 1, max again -- IndexBuilder would take the minimum here): pairs differing in dim 3
    7: (3, 0.0, 0), 8: (3, 0.0, 2), 9: (3, 0.0, 4), 10: (3, 0.0, 6), 11: (3, 0.0, 8), 12: (3, 0.0, 10), 13: (3, 0.0, 12),
    14: (3, 0.0, 14),
    # depth 4: the points, in id order (every split sent the lower id low)
    **{15 + i: (-1, 0.0, i) for i in range(16)},
}


def root_ties():
    """1-d values -1, 0, 0, 1 with ids 10, 20, 30, 40."""
    return np.array([10, 20, 30, 40], np.int64), np.array([[-1.0], [0.0], [0.0], [1.0]], np.float32)


ROOT_TIES_WANT = {
    # root (:104 `iif(S.Stdev = 0, iif(P.ID <= S.ID, 1, 2), iif(Value < Mean, 1, 2))`): Mean 0, only -1 goes low; both
    # zeros go HIGH although id 20 <= avg(ID) = 25 (IndexBuilder would send id 20 low)
    0: (0, 0.0, 25),
    1: (-1, 0.0, 10),
    # depth 1: values 0, 0, 1 -> Mean 1/3; (:161-167) Value < Mean -> low: ids 20, 30; 1 -> high: id 40
    2: (0, float(np.float32(1.0 / 3.0)), 30),
    6: (-1, 0.0, 40),
    # depth 2: values 0, 0: Stdev = 0 -> null row, ID <= 25 low
    5: (-3, NAN, 25),
    11: (-1, 0.0, 20), 12: (-1, 0.0, 30),
}


@pytest.mark.parametrize("case,want", [(lattice16, LATTICE16_WANT), (root_ties, ROOT_TIES_WANT)])
def test_hand_derived_answers_oracle(case, want):
    ids, rows = case()
    t = oracle.build(ids, rows, oracle.MODE_SQL)
    _same(_as_dict(t.range_id, t.dimension, t.mid, t.id), want)
    _same({r: (d, float(m), i) for r, d, m, i in np_oracle.build_qfx(ids, rows, sql=True)}, want)


def test_lattice_differs_from_index_builder():
    # the same points through IndexBuilder's rules: min at depth 1 AND 3, constant dimension kept as Dimension/Mid
    ids, rows = lattice16()
    t = oracle.build(ids, rows, oracle.MODE_QFX)
    tab = _as_dict(t.range_id, t.dimension, t.mid, t.id)
    assert tab[1][0] == 0 and np.float32(tab[1][1]) == np.float32(-0.4)
    assert all(tab[r][0] != -3 for r in tab)


def test_search_follows_both_children_of_a_null_row():
    # dbo.Search (:275,290 `N.Dimension is null or ...`): the query (-0.4, -0.2, -0.1, -0.05) = point 0 with proximity 0
    # goes low at the root, BOTH ways at row 1 (null), low at rows 3 / 4 (dim 2) and 7 / 9 (dim 3): leaves 15 and 19,
    # i.e. ids 0 and 4 -- point 4 differs from the query only in dim 1, which no row on the way tests
    ids, rows = lattice16()
    t = oracle.build(ids, rows, oracle.MODE_SQL)
    offs, out, visits = oracle.search(t, rows[:1], 0.0)
    assert out.tolist() == [0, 4] and visits == 8
    tab = _as_dict(t.range_id, t.dimension, t.mid, t.id)
    assert np_oracle.search(tab, rows[0], 0.0) == [0, 4]


@pytest.mark.parametrize("seed", range(12))
def test_c_oracle_matches_numpy_restatement_sql(seed):
    rng = np.random.default_rng(seed)
    n, d = int(rng.integers(1, 300)), int(rng.integers(1, 9))
    rows = rng.uniform(-1, 1, (n, d)).astype(np.float32)
    kind = seed % 4
    if kind == 1:
        rows[:, 0] = 0.25                      # a constant column: Stdev = 0 wins the minimum at depth 1
    if kind == 2:
        rows = (np.round(rows * 4) / 4).astype(np.float32)   # many values equal to a Mean
    if kind == 3:
        rows[: n // 2] = rows[0]               # duplicates: null rows split by ID
    ids = (rng.permutation(n).astype(np.int64) * 3) - 50
    t = oracle.build(ids, rows, oracle.MODE_SQL)
    got = list(zip(t.range_id.tolist(), t.dimension.tolist(), t.mid.view(np.uint32).tolist(), t.id.tolist()))
    want = [(r, dm, int(np.float32(m).view(np.uint32)), i) for r, dm, m, i in sorted(np_oracle.build_qfx(ids, rows, sql=True))]
    assert got == want
    assert sorted(t.id[t.dimension == -1].tolist()) == sorted(ids.tolist())    # leaves are the points
    # search is a superset of the L-infinity box also through null rows
    q = rows[:4]
    offs, out, _ = oracle.search(t, q, 0.1)
    for k in range(len(q)):
        box = set(ids[(np.abs(rows - q[k]) <= np.float32(0.1)).all(axis=1)].tolist())
        assert box <= set(out[offs[k]:offs[k + 1]].tolist())


# ---- CUDA path -----------------------------------------------------------------------------------------------------------
def _gpu(ids, rows):
    import vectorindex as vi
    with vi.Context(0) as ctx:
        ctx.reserve(len(ids), rows.shape[1])
        ctx.add(ids, rows)
        ctx.build(vi.MODE_SQL)
        rid, dim, mid, oid = ctx.ranges()
        o = np.argsort(rid, kind="stable")
        return rid[o], dim[o], mid[o], oid[o]


@pytest.mark.gpu
@pytest.mark.parametrize("case,want", [(lattice16, LATTICE16_WANT), (root_ties, ROOT_TIES_WANT)])
def test_hand_derived_answers_gpu(case, want):
    ids, rows = case()
    _same(_as_dict(*_gpu(ids, rows)), want)


def _assert_gpu_equals_oracle(ids, rows):
    rid, dim, mid, oid = _gpu(ids, rows)
    ref = oracle.build(ids, rows, oracle.MODE_SQL)
    assert np.array_equal(rid, ref.range_id) and np.array_equal(dim, ref.dimension)
    assert np.array_equal(mid.view(np.uint32), ref.mid.view(np.uint32)), "Mid must be bit-identical (NaN payload included)"
    assert np.array_equal(oid, ref.id)
    return ref


@pytest.mark.gpu
@pytest.mark.parametrize("n,d", [(1, 4), (2, 4), (3, 1), (33, 3), (600, 96), (4097, 16), (50_000, 96), (6000, 768)])
def test_sql_mode_table_equals_oracle(n, d):
    from vectorindex import synthetic as ds
    ids, rows = ds.unit_gaussian(n, d, seed=n + d)
    _assert_gpu_equals_oracle(ids * 5 - 11, rows)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["constant_column", "grid", "duplicates", "all_identical"])
def test_sql_mode_null_rows_equal_oracle(kind):
    rng = np.random.default_rng(5)
    n, d = 20_000, 12
    rows = rng.uniform(-1, 1, (n, d)).astype(np.float32)
    if kind == "constant_column":
        rows[:, 3] = 0.125
    elif kind == "grid":
        rows = (np.round(rows * 8) / 8).astype(np.float32)
    elif kind == "duplicates":
        rows[: n // 2] = rows[rng.integers(0, 50, n // 2)]
    else:
        rows[:] = rows[0]
        n = 3000
        rows = rows[:n]
    # (non-negative ids: avg(ID) truncates toward zero, so two identical vectors with ids -3, -2 get the pivot -2 and
    # never separate -- dbo.BuildIndex loops and IndexBuilder overflows its rangeId on such input)
    ids = rng.permutation(n).astype(np.int64) * 7
    ref = _assert_gpu_equals_oracle(ids, rows)
    if kind in ("duplicates", "all_identical"):
        assert (ref.dimension == -3).any()


@pytest.mark.gpu
def test_sql_mode_search_and_textindex_equal_oracle():
    import vectorindex as vi
    rng = np.random.default_rng(9)
    n, d = 30_000, 8
    rows = rng.uniform(-1, 1, (n, d)).astype(np.float32)
    rows[: n // 3] = rows[rng.integers(n // 3, n // 3 + 400, n // 3)]            # copies: null rows near the leaves
    ids = rng.permutation(n).astype(np.int64)
    ref = oracle.build(ids, rows, oracle.MODE_SQL)
    assert (ref.dimension == -3).any()
    q = np.concatenate([rows[:200], rng.uniform(-1, 1, (200, d)).astype(np.float32)], 0)
    with vi.Context(0) as ctx:
        ctx.reserve(n, d)
        ctx.add(ids, rows)
        ctx.build(vi.MODE_SQL)
        for p in (0.0, 0.05, 0.3):
            offs, out = ctx.search(q, p)
            roffs, rout, _ = oracle.search(ref, q, p)
            assert np.array_equal(offs, roffs) and np.array_equal(out, rout), p
        rid, dim, mid, lo, hi, tid = ctx.textindex()
        table = ctx.ranges()
        # a table exported and imported again answers the same searches (null rows survive the round trip)
        with vi.Context(0) as imp:
            imp.load_ranges(*table, d)
            offs2, out2 = imp.search(q, 0.05)
            roffs, rout, _ = oracle.search(ref, q, 0.05)
            assert np.array_equal(offs2, roffs) and np.array_equal(out2, rout)
    o = np.argsort(rid)
    rid, dim, mid, lo, hi, tid = rid[o], dim[o], mid[o], lo[o], hi[o], tid[o]
    leaf = ref.dimension == -1
    null = ref.dimension < 0
    assert np.array_equal(rid, ref.range_id)
    assert np.array_equal(dim, np.where(null, -1, ref.dimension).astype(np.int16))            # :193
    assert np.isnan(mid[null]).all() and np.array_equal(mid[~null], ref.mid[~null])           # :194
    assert np.array_equal(lo, np.where(leaf, -1, rid * 2 + 1)) and np.array_equal(hi, np.where(leaf, -1, rid * 2 + 2))  # :195-196
    assert np.array_equal(tid, np.where(leaf, ref.id, -1))                                    # :197
