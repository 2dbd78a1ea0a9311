#!/usr/bin/env python
"""Builds the 10M x 96 index once and runs the device-resident search at p = 0 and p = 0.01 (the workload of the
bench line's `search` object) -- the command the ncu captures of the traversal kernels are taken on."""
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
    nq = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    dev = torch.device("cuda", 0)
    ids_d, rows_d = bench.gen_device(n, d, 2, dev)
    ctx = vi.Context(0)
    ctx.reserve(n, d)
    ctx.add_device(ids_d.data_ptr(), rows_d.data_ptr(), n, d)
    ctx.build(vi.MODE_FAST)
    q_d = bench.gen_queries(rows_d, nq, seed=77)
    offs_d = torch.empty(nq + 1, dtype=torch.int64, device=dev)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    for p in (0.0, 0.01):
        total, visits = ctx.search_device(q_d.data_ptr(), nq, d, p, offs_d.data_ptr(), 0, 0)
        ids_out = torch.empty(max(total, 1), dtype=torch.int64, device=dev)
        ctx.search_device(q_d.data_ptr(), nq, d, p, offs_d.data_ptr(), ids_out.data_ptr(), total)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            ctx.search_device(q_d.data_ptr(), nq, d, p, offs_d.data_ptr(), ids_out.data_ptr(), total)
        b.record(stream)
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / max(reps, 1)
        print(f"p={p}: {ms:.2f} ms for {nq} queries = {nq / ms / 1e3:.1f} M queries/s, {total} candidates, {visits} visits "
              f"(path {os.environ.get('VI_B200_SEARCH_PATH', 'auto')})", flush=True)
        del ids_out


if __name__ == "__main__":
    main()
