#!/usr/bin/env python
"""e2e (pinned host -> H2D -> vi_build_copy -> host) at 10M x 96 for several slice counts of the sub-tree kernel."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vector-database_b200"))
import torch
import bench
import vectorindex as vi
n, d = 10_000_000, 96
dev = torch.device("cuda", 0)
ids_d, rows_d = bench.gen_device(n, d, 2, dev)
rows_h = torch.empty((n, d), dtype=torch.float32, pin_memory=True); rows_h.copy_(rows_d)
ids_h = torch.empty((n,), dtype=torch.int64, pin_memory=True); ids_h.copy_(ids_d)
del rows_d, ids_d
cap = 2 * n + n // 8 + 1024
outs = [torch.empty(cap, dtype=t, pin_memory=True).numpy() for t in (torch.int64, torch.int32, torch.float32, torch.int64)]
ctx = vi.Context(0)
for sl in (1, 2, 3, 4, 6, 8):
    os.environ["VI_B200_COPY_SLICES"] = str(sl)
    ts = []
    for i in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ctx.reserve(n, d); ctx.add(ids_h.numpy(), rows_h.numpy()); t1 = time.perf_counter()
        info, k = ctx.build_into(vi.MODE_FAST, *outs); torch.cuda.synchronize(); t2 = time.perf_counter()
        if i >= 2: ts.append(((t2 - t0) * 1e3, (t2 - t1) * 1e3, info.build_ms, info.subtree_ms))
    print(f"slices={sl}: e2e {sum(t[0] for t in ts)/len(ts):.1f} ms, build+copy {sum(t[1] for t in ts)/len(ts):.1f}, "
          f"build alone {sum(t[2] for t in ts)/len(ts):.1f}, sub-tree kernel(s) {sum(t[3] for t in ts)/len(ts):.1f}", flush=True)
