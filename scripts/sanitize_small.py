"""Small builds / searches through every kernel family, for compute-sanitizer memcheck (scripts/sanitize.sh)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vector-database_b200"))
import numpy as np  # noqa: E402

import oracle  # noqa: E402
import vectorindex as vi  # noqa: E402
from vectorindex import synthetic as ds  # noqa: E402


def check(ids, rows, mode):
    ref = oracle.build(ids, rows, mode)
    with vi.Context(0) as ctx:
        ctx.reserve(len(ids), rows.shape[1])
        ctx.add(ids, rows)
        ctx.build(mode)
        rid, dim, mid, oid = ctx.ranges()
        q = rows[:64]
        offs, out = ctx.search(q, 0.05)
        ctx.search_topk(q, 0.05, 5, 0)
        ctx.search_verify(q, 0.05, 0.2)
    o = np.argsort(rid)
    assert np.array_equal(rid[o], ref.range_id) and np.array_equal(dim[o], ref.dimension)
    assert np.array_equal(mid[o].view(np.uint32), ref.mid.view(np.uint32)) and np.array_equal(oid[o], ref.id)
    return rid[o], dim[o], mid[o], oid[o]


ids, rows = ds.unit_gaussian(12_000, 96, seed=3)     # chunk kernel, sibling derivation, warp / team classes, sub-trees
t = check(ids, rows, vi.MODE_FAST)
check(ids, rows, vi.MODE_EXACT)                       # pipeline kernel (top levels), single-warp kernel, small kernel
ids2, rows2 = ds.uniform(3000, 50, seed=4)            # padded rows, non-vector cp.async path
check(ids2, rows2, vi.MODE_FAST)
check(ids2, rows2, vi.MODE_EXACT)
with vi.Context(0) as ctx:                            # table import + record ingest
    ctx.load_ranges(*t, 96)
    ctx.search(rows[:16], 0.02)
with vi.Context(0) as ctx:
    ctx.reserve(0, 50)
    ctx.add_records(vi.pack_records(ids2, rows2), 50)
    ctx.build(vi.MODE_FAST)

# ---- round 2 -----------------------------------------------------------------------------------------------------------
# every search path on one table: thread per query, warp count + fill, warp pool + gather, stack spill, pool overflow
ids3, rows3 = ds.unit_gaussian(40_000, 24, seed=6)
ref3 = oracle.build(ids3, rows3, oracle.MODE_QFX)
q3 = np.concatenate([rows3[:40], ds.unit_gaussian(40, 24, seed=7)[1]], 0)
want3 = {p: oracle.search(ref3, q3, p)[:2] for p in (0.0, 0.1, 0.4)}
for env in ({"VI_B200_SEARCH_PATH": "0"}, {"VI_B200_SEARCH_PATH": "1"}, {"VI_B200_SEARCH_PATH": "1", "VI_B200_SEARCH_POOL": "0"},
            {"VI_B200_SEARCH_PATH": "1", "VI_B200_SEARCH_STACK": "64"}, {"VI_B200_SEARCH_PATH": "1", "VI_B200_SEARCH_POOL_SLOTS": "4096"}):
    for k in ("VI_B200_SEARCH_PATH", "VI_B200_SEARCH_POOL", "VI_B200_SEARCH_STACK", "VI_B200_SEARCH_POOL_SLOTS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    with vi.Context(0) as ctx:
        ctx.reserve(len(ids3), 24)
        ctx.add(ids3, rows3)
        ctx.build(vi.MODE_FAST)
        for p, (roffs, rout) in want3.items():
            offs, out = ctx.search(q3, p)
            assert np.array_equal(offs, roffs) and np.array_equal(out, rout), (env, p)
        ctx.search_verify(q3[:8], 0.1, 0.3)
        ctx.search_topk(q3[:8], 0.1, 5, 1)
for k in ("VI_B200_SEARCH_PATH", "VI_B200_SEARCH_POOL", "VI_B200_SEARCH_STACK", "VI_B200_SEARCH_POOL_SLOTS"):
    os.environ.pop(k, None)
# SQL mode with null rows, build_copy with sliced sub-tree kernel, wide rows (column passes), sibling slots of warp ranges
rng = np.random.default_rng(8)
rows4 = rng.uniform(-1, 1, (30_000, 12)).astype(np.float32)
rows4[:10_000] = rows4[rng.integers(10_000, 10_300, 10_000)]
ids4 = rng.permutation(30_000).astype(np.int64)
t4 = check(ids4, rows4, vi.MODE_SQL)
assert (t4[1] == vi.DIM_NULL).any()
ids5, rows5 = ds.unit_gaussian(200_000, 16, seed=9)
with vi.Context(0) as ctx:
    ctx.reserve(len(ids5), 16)
    ctx.add(ids5, rows5)
    cap = 2 * len(ids5) + len(ids5) // 8 + 1024
    bufs = [np.zeros(cap, t) for t in (np.int64, np.int32, np.float32, np.int64)]
    info, k = ctx.build_into(vi.MODE_FAST, *bufs)
    want = ctx.ranges()
    assert all(np.array_equal(b[:k].view(np.uint8), w.view(np.uint8)) for b, w in zip(bufs, want))
ids6, rows6 = ds.unit_gaussian(3000, 768, seed=10)
check(ids6, rows6, vi.MODE_FAST)
check(ids6, rows6, vi.MODE_SQL)
print("sanitize_small: ok")
