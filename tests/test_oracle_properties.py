"""CPU: size-independent properties of the reference algorithm, checked on the oracle for
many small random inputs in both modes.  The replay below re-derives every range's point set from the emitted rows
alone with the partition predicate of IndexBuilder.cs:115, i.e. it does not trust the oracle's own bookkeeping."""
import numpy as np
import pytest

import oracle

MODES = [oracle.MODE_LITERAL, oracle.MODE_QFX]


def _inputs(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 260))
    d = int(rng.integers(1, 9))
    kind = seed % 4
    if kind == 0:
        rows = rng.uniform(-1, 1, (n, d))
    elif kind == 1:
        rows = rng.standard_normal((n, d)) * 10 + 100            # offset data
    elif kind == 2:
        rows = rng.integers(0, 3, (n, d)).astype(np.float64)     # many duplicates / ties at Mid
    else:
        rows = rng.standard_normal((n, d))
        rows[:, 0] = 0.5                                          # a constant dimension
    ids = rng.permutation(n * 3)[:n].astype(np.int64)             # unique
    if kind < 2:
        ids -= n  # some negative ids -- not where whole vectors can coincide (the duplicate-heavy set; the constant
        #           dimension set when d == 1): two identical points whose ids do not exceed their mean id truncated
        #           TOWARD ZERO are inseparable and the reference itself dies with OverflowException
        #           (tests/test_oracle_kat.py::test_inseparable_points_overflow_at_depth_62)
    return ids, rows.astype(np.float32)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("seed", range(24))
def test_structure_and_partition_replay(seed, mode):
    ids, rows = _inputs(seed)
    n = len(ids)
    t = oracle.build(ids, rows, mode)
    table = {int(r): (int(a), np.float32(b), int(c)) for r, a, b, c in zip(t.range_id, t.dimension, t.mid, t.id)}
    assert len(table) == len(t), "a RangeID appears once"
    leaves = {r: v for r, v in table.items() if v[0] == -1}
    assert len(leaves) == n and sorted(v[2] for v in leaves.values()) == sorted(ids.tolist())
    assert all(v[1] == 0 for v in leaves.values())               # Mid stays default on leaves (IndexBuilder.cs:81-82)
    # replay from the root: the rows alone must reproduce every range's content
    stack = [(0, np.arange(n), True)]
    seen = 0
    while stack:
        r, idx, mx = stack.pop()
        assert r in table and len(idx) > 0
        dim, mid, pivot = table[r]
        seen += 1
        if len(idx) == 1:
            assert dim == -1 and pivot == ids[idx[0]]
            continue
        assert 0 <= dim < rows.shape[1]
        s = int(sum(int(ids[i]) for i in idx))
        assert pivot == (abs(s) // len(idx)) * (1 if s >= 0 else -1)   # Int128 division truncates toward zero
        v = rows[idx, dim]
        hi = (v > mid) | ((v == mid) & (ids[idx] > pivot))        # IndexBuilder.cs:115, order preserved
        for child, sel in ((2 * r + 1, ~hi), (2 * r + 2, hi)):
            if sel.any():
                stack.append((child, idx[sel], not mx))
            else:
                assert child not in table                          # an empty child emits no row (IndexBuilder.cs:70-73)
    assert seen == len(table)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("seed", range(8))
def test_search_contains_the_box_and_finds_dataset_points(seed, mode):
    ids, rows = _inputs(seed)
    t = oracle.build(ids, rows, mode)
    rng = np.random.default_rng(100 + seed)
    q = np.concatenate([rows[: min(10, len(rows))], rng.uniform(-1, 1, (6, rows.shape[1])).astype(np.float32)])
    for p in (0.0, 0.2):
        offs, out, _ = oracle.search(t, q, p)
        for i in range(len(q)):
            got = set(out[offs[i]:offs[i + 1]].tolist())
            lo, hi = (q[i] - np.float32(p)).astype(np.float32), (q[i] + np.float32(p)).astype(np.float32)
            box = np.all((rows >= lo) & (rows <= hi), axis=1)      # DDL.sql:249-250 bounds, float32
            assert set(ids[box].tolist()) <= got
    offs, out, _ = oracle.search(t, rows[: min(10, len(rows))], 0.0)
    for i in range(min(10, len(rows))):
        assert ids[i] in out[offs[i]:offs[i + 1]]
