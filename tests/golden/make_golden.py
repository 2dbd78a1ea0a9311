"""Writes the golden fixtures of this directory.

Two kinds, kept apart because they do not carry the same weight ("parity unpinned"):

* kat_hand_derived.json -- known answers worked out BY HAND from the reference's source (IndexBuilder.cs:23-198,
  DDL.sql:246-295; SURVEY.md 8c list).  This script only re-serialises the literals below; nothing is computed.
* frozen_*.npz -- outputs of the CPU oracle (oracle/vi_oracle.c, literal and fast-mode statements) on small seeded
  inputs, frozen so that neither the oracle nor the CUDA path can drift unnoticed.  They are NOT reference outputs: the
  reference (.NET 8 + SQL Server) cannot run here or on the GPU box.

    python tests/golden/make_golden.py        # from the repo root, after `make -C oracle`
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "vector-database_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

# rows: rangeId -> [Dimension, Mid, Id]
KATS = [
    {"name": "single point (IndexBuilder.cs:81-82)", "ids": [42], "rows": [[0.5, -0.25, 3.0]],
     "table": {"0": [-1, 0.0, 42]}},
    {"name": "two points, one dimension differs: even depth picks max variance, Mid = mean, Id = trunc(16/2)",
     "ids": [7, 9], "rows": [[1.0, 1.0, 1.0], [1.0, 0.0, 1.0]],
     "table": {"0": [1, 0.5, 8], "1": [-1, 0.0, 9], "2": [-1, 0.0, 7]}},
    {"name": "identical vectors: value == Mid and id > Id goes high (IndexBuilder.cs:115)", "ids": [3, 10],
     "rows": [[0.25, 0.25], [0.25, 0.25]], "table": {"0": [0, 0.25, 6], "1": [-1, 0.0, 3], "2": [-1, 0.0, 10]}},
    {"name": "Int128 division truncates toward zero: (-3 + -6) / 2 = -4 (IndexBuilder.cs:87)", "ids": [-3, -6],
     "rows": [[0.5], [0.5]], "table": {"0": [0, 0.5, -4], "1": [-1, 0.0, -6], "2": [-1, 0.0, -3]}},
    {"name": "odd depth picks MIN variance, lowest index on ties (IndexBuilder.cs:77-79): root splits dim 0 "
             "(values 0,0,4,4; Mid 2, Id trunc(10/4) = 2); both children hold identical vectors, every variance is 0, "
             "the first dimension wins, the id tie-break separates them",
     "ids": [1, 2, 3, 4], "rows": [[0.0, 1.0], [0.0, 1.0], [4.0, 1.0], [4.0, 1.0]],
     "table": {"0": [0, 2.0, 2], "1": [0, 0.0, 1], "2": [0, 4.0, 3], "3": [-1, 0.0, 1], "4": [-1, 0.0, 2],
               "5": [-1, 0.0, 3], "6": [-1, 0.0, 4]}},
    {"name": "odd depth picks the SMALLER non-zero variance: root (even) splits dim 0 (values 0,6,12,18: the float32 "
             "recurrence is exact here, mean 3 -> 6 -> 9, q 18 -> 72 -> 180; Id trunc(10/4) = 2); each child holds two "
             "points with q = 18 in dim 0 and (0.5)(0.25) = 0.125 in dim 1, so the min-variance level splits dim 1 at 0.25",
     "ids": [1, 2, 3, 4], "rows": [[0.0, 0.0], [6.0, 0.5], [12.0, 0.0], [18.0, 0.5]],
     "table": {"0": [0, 9.0, 2], "1": [1, 0.25, 1], "2": [1, 0.25, 3], "3": [-1, 0.0, 1], "4": [-1, 0.0, 2],
               "5": [-1, 0.0, 3], "6": [-1, 0.0, 4]}},
]

SEARCH_KAT = {
    "name": "dbo.Search over the two-point table above (DDL.sql:246-295): low iff Mid >= v - p, high iff Mid <= v + p",
    "ids": [7, 9], "rows": [[1.0, 1.0, 1.0], [1.0, 0.0, 1.0]],
    "cases": [{"query": [1.0, 0.0, 1.0], "p": 0.0, "ids": [9]},       # 0.5 >= 0 -> low only
              {"query": [1.0, 1.0, 1.0], "p": 0.0, "ids": [7]},       # 0.5 <= 1 -> high only
              {"query": [1.0, 0.5, 1.0], "p": 0.0, "ids": [9, 7]},    # Mid == v: both, low branch first
              {"query": [1.0, 0.9, 1.0], "p": 0.5, "ids": [9, 7]},    # 0.5 >= 0.4 and 0.5 <= 1.4
              {"query": [1.0, 0.9, 1.0], "p": 0.25, "ids": [7]}],     # 0.5 >= 0.65 fails
}


def main():
    import oracle
    from vectorindex import synthetic as ds
    with open(os.path.join(HERE, "kat_hand_derived.json"), "w") as f:
        json.dump({"tables": KATS, "search": SEARCH_KAT}, f, indent=1)

    def freeze(name, ids, rows, queries, prox):
        out = {"ids": ids, "rows": rows, "queries": queries, "proximity": np.float32(prox)}
        for mode, tag in ((oracle.MODE_LITERAL, "lit"), (oracle.MODE_QFX, "qfx")):
            t = oracle.build(ids, rows, mode)
            offs, cand, visits = oracle.search(t, queries, prox)
            out.update({f"{tag}_rid": t.range_id, f"{tag}_dim": t.dimension, f"{tag}_mid": t.mid, f"{tag}_id": t.id,
                        f"{tag}_offsets": offs, f"{tag}_candidates": cand, f"{tag}_visits": np.int64(visits)})
        np.savez_compressed(os.path.join(HERE, name), **out)

    ids, rows = ds.uniform(3000, 12, seed=7)
    freeze("frozen_uniform_3000x12.npz", ids * 2 + 5, rows, rows[::300].copy(), 0.1)
    ids, rows = ds.one_hot(96)                            # the crafted set of Program.cs:54-66 at d = 96
    freeze("frozen_one_hot_96.npz", ids, rows, rows[:8].copy(), 0.0)
    ids, rows = ds.unit_gaussian(2500, 96, seed=9)
    q = np.concatenate([rows[:5], ds.unit_gaussian(5, 96, seed=10)[1]])
    freeze("frozen_unit_gaussian_2500x96.npz", ids, rows, q, 0.08)


if __name__ == "__main__":
    main()
