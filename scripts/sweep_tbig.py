#!/usr/bin/env python
"""T_BIG sweep (GPU box): does the chunk path + sibling derivation beat the warp-per-range kernel for mid ranges?
Prints build ms, per-level statistics ms and the table checksum (must not change with the setting)."""
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
    import itertools
    for tb, wch in itertools.product(os.environ.get("SWEEP_T_BIG", "512,256,128,64").split(","),
                                     os.environ.get("SWEEP_WIDE_CH", "6").split(",")):
        os.environ["VI_B200_T_BIG"] = tb
        os.environ["VI_B200_WIDE_CH"] = wch
        ctx.build(vi.MODE_FAST)
        best = None
        for _ in range(3):
            info = ctx.build(vi.MODE_FAST)
            lv = ctx.levels()
            if best is None or info.build_ms < best[0]:
                best = (info.build_ms, lv, info)
        ms, lv, info = best
        cs = bench.table_checksum(*ctx.ranges())
        print(f"T_BIG={tb} WIDE_CH={wch} build {ms:.2f} ms stats {sum(l.stats_ms for l in lv):.2f} partition "
              f"{sum(l.partition_ms for l in lv):.2f} subtree {info.subtree_ms:.2f} launches {info.kernel_launches} "
              f"checksum {cs}", flush=True)
        print("   stats_ms/level:", " ".join(f"{l.stats_ms:.2f}" for l in lv[:20]), flush=True)
        print("   part_ms/level:", " ".join(f"{l.partition_ms:.2f}" for l in lv[:20]), flush=True)


if __name__ == "__main__":
    main()
