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
print("sanitize_small: ok")
