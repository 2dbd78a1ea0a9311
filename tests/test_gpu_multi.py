"""GPU, one process per GPU over NCCL: the multi-rank builds (vi_build.cu build_sharded: fast / SQL mode with shared
levels + ownership exchange; build_sharded_exact: replicated points, common top, owned sub-trees) must give, as the union
of the per-rank tables, exactly the single-rank oracle table.  Needs >= 2 GPUs (skipped otherwise)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vector-database_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

pytestmark = pytest.mark.gpu


def _data(n, d, seed, kind):
    from vectorindex import synthetic as ds
    ids, rows = getattr(ds, "unit_gaussian" if kind in ("sql_dups", "exact") else "uniform" if kind == "clustered" else kind)(n, d, seed=seed)
    ids = ids * 3 + 7
    if kind == "clustered":
        # a tight cluster (|x| ~ 1e-7) plus outliers that set the quantisation scale: below the root the ranges are
        # poorly resolved for the integer statistics -> the build starts over with fewer shared levels
        rows = (rows * np.float32(1e-7)).astype(np.float32)
        rows[::7, 3] += np.float32(1.5)
    if kind == "sql_dups":
        # VI_MODE_SQL over copies of a few vectors: Stdev = 0 ranges (null Dimension rows) inside the owned sub-trees
        rows = rows.copy()
        rows[: n // 3] = rows[np.random.default_rng(seed).integers(n // 3, n // 3 + 300, n // 3)]
    return ids, rows


def _worker(rank, world, port, q, n, d, seed, kind, transport="nccl"):
    import torch.distributed as dist
    import vectorindex as vi
    from vectorindex import synthetic as ds
    from vectorindex.distributed import Collectives, init_nccl
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ids, rows = _data(n, d, seed, kind)
    mode = vi.MODE_SQL if kind == "sql_dups" else vi.MODE_EXACT if kind == "exact" else vi.MODE_FAST
    lo, hi = rank * n // world, (rank + 1) * n // world
    if kind == "uniform" and rank == world - 1:
        lo = hi = n  # an empty shard on the last rank ...
    if kind == "uniform" and rank == world - 2:
        hi = n       # ... its points go to the rank before it
    ctx = vi.Context(rank)
    ctx.reserve(max(hi - lo, 1), d)
    if hi > lo:
        ctx.add(ids[lo:hi], rows[lo:hi])
    if transport == "callbacks":
        coll = Collectives(torch.device("cuda", rank))  # the host's own transport (torch / NCCL) behind vi_set_collective
        coll.attach(ctx)
    else:
        init_nccl(ctx, torch.device("cuda", rank))      # library-owned NCCL communicator
    # a search before the table is replicated must be refused (this rank holds only its own sub-trees)
    info = ctx.build(mode)
    try:
        ctx.search(rows[:2], 0.0)
        refused = False
    except vi.VectorIndexError as e:
        refused = e.code == vi.VI_ERR_STATE
    assert refused
    rid, dim, mid, oid = ctx.ranges()
    shared = ctx.shared_rows
    calls = ctx.comm_stats()
    calls["retry"] = int(info.shared_retry)
    # replicate: every rank must now hold the whole table and answer any query like the oracle
    ctx.replicate()
    frid, fdim, fmid, foid = ctx.ranges()
    queries = np.concatenate([rows[: 40], rows[n - 40:]], 0)
    offs, out = ctx.search(queries, 0.02)
    q.put((rank, rid, dim, mid.view(np.uint32), oid, shared, calls, int(info.levels),
           (frid, fdim, fmid.view(np.uint32), foid, offs, out)))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("n,d,seed,kind,transport", [(200_000, 96, 3, "unit_gaussian", "nccl"), (5000, 16, 5, "uniform", "nccl"),
                                                     (30_000, 24, 7, "unit_gaussian", "callbacks"),
                                                     (20_000, 8, 9, "clustered", "nccl"),
                                                     (60_000, 12, 11, "sql_dups", "nccl"),
                                                     (150_000, 24, 13, "exact", "nccl"), (9000, 40, 15, "exact", "callbacks")])
def test_sharded_build_equals_oracle(world, n, d, seed, kind, transport):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import oracle
    import torch.multiprocessing as mp
    from vectorindex import synthetic as ds
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world * 7 + seed
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, n, d, seed, kind, transport)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        item = q.get(timeout=180)
        res[item[0]] = item[1:]
    for p in procs:
        p.join(60)
        if p.is_alive():
            p.kill()
    union = {}
    owned_rows = []
    for r in range(world):
        rid, dim, mid, oid, shared, calls, levels, full = res[r]
        if kind == "exact":
            assert calls["allgather"] >= 2 or transport == "callbacks"  # every rank receives every point: rows + ids
        else:
            assert calls["alltoallv"] >= 2  # rows + ids, once per attempt
        if kind == "clustered":
            assert calls["retry"] == 1
        for k in range(len(rid)):
            val = (int(dim[k]), int(mid[k]), int(oid[k]))
            if k < shared:
                if dim[k] == -2:
                    continue  # root row of a range owned by another rank (placeholder)
                assert union.setdefault(int(rid[k]), val) == val  # replicated top rows agree
            else:
                assert int(rid[k]) not in union
                union[int(rid[k])] = val
        owned_rows.append(len(rid) - shared)
    ids, rows = _data(n, d, seed, kind)
    ref = oracle.build(ids, rows, oracle.MODE_SQL if kind == "sql_dups" else oracle.MODE_LITERAL if kind == "exact" else oracle.MODE_QFX)
    if kind == "sql_dups":
        assert (ref.dimension == -3).any()
    want = {int(r): (int(dm), int(np.float32(m).view(np.uint32)), int(i))
            for r, dm, m, i in zip(ref.range_id, ref.dimension, ref.mid, ref.id)}
    assert union == want
    assert sum(owned_rows) + res[0][4] == len(want)
    # replicated tables: identical content on every rank, searches equal the oracle's
    queries = np.concatenate([rows[: 40], rows[n - 40:]], 0)
    roffs, rout, _ = oracle.search(ref, queries, 0.02)
    for r in range(world):
        frid, fdim, fmid, foid, offs, out = res[r][7]
        got = {int(a): (int(b), int(c), int(e)) for a, b, c, e in zip(frid, fdim, fmid, foid)}
        assert got == want
        assert np.array_equal(offs, roffs)
        for i in range(len(queries)):
            assert sorted(out[offs[i]:offs[i + 1]].tolist()) == sorted(rout[roffs[i]:roffs[i + 1]].tolist())
    if kind == "unit_gaussian":
        # ownership is balanced: no rank owns more than 1.5x its fair share of the rows
        assert max(owned_rows) <= 1.5 * len(want) / world
