#!/usr/bin/env python
"""Tuning sweep for the fast-mode build (run on a GPU box): the library re-reads its VI_B200_* knobs on every
vi_build, so one process with one resident data set can try many settings.  Prints per-level statistics times."""
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vector-database_b200"))

import torch  # noqa: E402

import bench  # noqa: E402
import vectorindex as vi  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 96
    dev = torch.device("cuda", 0)
    ids_d, rows_d = bench.gen_device(n, d, 2, dev)
    ctx = vi.Context(0)
    ctx.reserve(n, d)
    ctx.add_device(ids_d.data_ptr(), rows_d.data_ptr(), n, d)
    del rows_d
    grid = {
        "VI_B200_T_TEAM": os.environ.get("SWEEP_T_TEAM", "32,64,128").split(","),
        "VI_B200_T_BIG": os.environ.get("SWEEP_T_BIG", "1024,2048,4096").split(","),
        "VI_B200_BIG_UNROLL": os.environ.get("SWEEP_UNROLL", "2,4").split(","),
    }
    keys = list(grid)
    for combo in itertools.product(*[grid[k] for k in keys]):
        for k, v in zip(keys, combo):
            os.environ[k] = v
        ctx.build(vi.MODE_FAST)
        best = None
        for _ in range(3):
            info = ctx.build(vi.MODE_FAST)
            lv = ctx.levels()
            if best is None or info.build_ms < best[0]:
                best = (info.build_ms, lv)
        ms, lv = best
        stats = sum(l.stats_ms for l in lv)
        part = sum(l.partition_ms for l in lv)
        print(dict(zip(keys, combo)), f"build {ms:.2f} ms stats {stats:.2f} partition {part:.2f}", flush=True)
        print("   stats_ms/level:", " ".join(f"{l.stats_ms:.2f}" for l in lv), flush=True)


if __name__ == "__main__":
    main()
