import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vector-database_b200"))
import numpy as np, oracle, vectorindex as vi
from vectorindex import synthetic as ds
n, d = int(sys.argv[1]), int(sys.argv[2])
ids, rows = ds.uniform(n, d, seed=n + d)
ids = ids * 7 + 3
ref = oracle.build(ids, rows, vi.MODE_FAST)
with vi.Context(0) as ctx:
    ctx.reserve(n, d); ctx.add(ids, rows); ctx.build(vi.MODE_FAST)
    rid, dim, mid, oid = ctx.ranges()
o = np.argsort(rid); rid, dim, mid, oid = rid[o], dim[o], mid[o], oid[o]
print("rows", len(rid), len(ref))
want = {int(r): (int(a), float(b), int(c)) for r, a, b, c in zip(ref.range_id, ref.dimension, ref.mid, ref.id)}
bad = 0
for r, a, b, c in zip(rid, dim, mid, oid):
    w = want.get(int(r))
    if w != (int(a), float(b), int(c)):
        lvl = int(np.floor(np.log2(int(r) + 1)))
        print("mismatch rid", int(r), "level", lvl, "got", (int(a), float(b), int(c)), "want", w)
        bad += 1
        if bad > 8: break
print("bad", bad)
