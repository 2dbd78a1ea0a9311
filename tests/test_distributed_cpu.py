"""CPU, world_size 2, gloo: the N>1 host path.

1. vectorindex.distributed.Collectives (the callbacks libvi_b200 calls during a multi-rank build) on host pointers.
2. The multi-rank build PROTOCOL (DESIGN.md "Multi-GPU build"), restated in Python over the CPU oracle: shared top
   levels with all-reduced integer sums, LPT ownership, one all-to-all with pieces in source-rank order, owners finish
   their sub-trees alone.  The union of the per-rank tables must equal the single-rank oracle table bit for bit --
   the property the CUDA implementation (vi_build.cu build_sharded) is tested for on real GPUs
   (tests/test_gpu_multi.py).
"""
import math
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vector-database_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

QBITS = 26


def _init(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _worker_collectives(rank, world, port, q):
    _init(rank, world, port)
    from vectorindex.distributed import Collectives
    coll = Collectives(torch.device("cpu"))
    a = np.arange(5, dtype=np.uint64) + np.uint64(rank * 100)
    a[4] = np.uint64(2 ** 64 - 1) if rank == 0 else np.uint64(2)  # wraps like uint64
    coll.allreduce(a.ctypes.data, 5)
    # rank r sends (d + 1) * (r + 1) bytes to rank d
    sb = [(d + 1) * (rank + 1) for d in range(world)]
    rb = [(rank + 1) * (s + 1) for s in range(world)]
    send = np.concatenate([np.full(sb[d], 10 * rank + d, np.uint8) for d in range(world)])
    recv = np.zeros(sum(rb), np.uint8)
    coll.alltoallv(send.ctypes.data, sb, recv.ctypes.data, rb)
    q.put((rank, a.tolist(), recv.tolist(), rb))
    dist.destroy_process_group()


def test_collectives_on_gloo():
    world, port = 2, 29511
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_collectives, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
    for rank, a, recv, rb in res:
        assert a[:4] == [100, 102, 104, 106] and a[4] == 1  # (2^64 - 1 + 2) mod 2^64
        want = []
        for s in range(world):
            want += [10 * s + rank] * rb[s]
        assert recv == want


# ---- protocol emulation -------------------------------------------------------------------------------------------
def _gather_sum(obj):
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def emulate_rank(rank, world, ids, rows, coll):
    """One rank of the multi-rank fast-mode build, in Python over the oracle. Returns {rangeId: (dim, mid_bits, id)}
    for the shared rows (every rank) and the rows of the sub-trees this rank owns."""
    import oracle
    n_all = _gather_sum(len(ids))
    amax = max(_gather_sum(float(np.abs(rows).max()) if len(ids) else 0.0))
    e = int(np.frexp(np.float32(amax))[1]) if amax > 0 else 0
    k = np.float32(2.0) ** np.float32(QBITS - e)
    xi = np.rint((rows * k).astype(np.float64)).astype(np.int64)
    L = 1
    while (1 << (L - 1)) < world:
        L += 1
    table = {}
    segs = [dict(idx=np.arange(len(ids)), g=sum(n_all), rid=0)]
    if sum(n_all) == 1:
        segs = []
    for level in range(L):
        if not segs:
            break
        mx = level % 2 == 0
        nxt = []
        for sg in segs:
            sub = xi[sg["idx"]]
            s1 = [int(v) for v in sub.sum(axis=0)] if len(sub) else [0] * rows.shape[1]
            s2 = [sum(int(v) * int(v) for v in sub[:, j]) for j in range(rows.shape[1])]
            idn = sum(int(ids[i]) for i in sg["idx"])
            parts = _gather_sum((s1, s2, idn))
            S1 = [sum(p[0][j] for p in parts) for j in range(rows.shape[1])]
            S2 = [sum(p[1][j] for p in parts) for j in range(rows.shape[1])]
            IDN = sum(p[2] for p in parts)
            n = sg["g"]
            keys = [n * b - a * a for a, b in zip(S1, S2)]
            assert max(keys) >= (n * n) << 10, "poorly resolved range in the shared phase"
            dim = 0
            for j in range(1, len(keys)):
                if (keys[j] > keys[dim]) if mx else (keys[j] < keys[dim]):
                    dim = j
            mid = np.float32((np.float64(S1[dim]) / np.float64(n)) * np.float64(2.0) ** (e - QBITS))
            pivot = abs(IDN) // n * (1 if IDN >= 0 else -1)
            table[sg["rid"]] = (dim, int(mid.view(np.uint32)), pivot)
            v = rows[sg["idx"], dim]
            hi = (v > mid) | ((v == mid) & (ids[sg["idx"]] > pivot))
            for side, m in ((0, ~hi), (1, hi)):
                child = sg["idx"][m]
                g = sum(_gather_sum(len(child)))
                rid = 2 * sg["rid"] + 1 + side
                if g == 0:
                    continue
                if g == 1:
                    owner_id = sum(_gather_sum(int(ids[child[0]]) if len(child) else 0))
                    table[rid] = (-1, 0, owner_id)
                else:
                    nxt.append(dict(idx=child, g=g, rid=rid))
        segs = nxt
    depth = min(L, 10 ** 9) if segs else 0
    # ownership: largest range first to the least loaded rank
    order = sorted(range(len(segs)), key=lambda i: -segs[i]["g"])  # stable: ties keep the lower index first
    load = [0] * world
    owner = [0] * len(segs)
    for i in order:
        best = min(range(world), key=lambda g: (load[g], g))
        owner[i] = best
        load[best] += segs[i]["g"]
    counts = _gather_sum([len(s["idx"]) for s in segs])  # counts[g][i]
    d = rows.shape[1]
    send_rows, send_ids, sb = [], [], []
    for dst in range(world):
        nb = 0
        for i, sg in enumerate(segs):
            if owner[i] == dst:
                send_rows.append(rows[sg["idx"]])
                send_ids.append(ids[sg["idx"]])
                nb += len(sg["idx"])
        sb.append(nb)
    rb = [sum(counts[g][i] for i in range(len(segs)) if owner[i] == rank) for g in range(world)]
    srows = np.ascontiguousarray(np.concatenate(send_rows) if send_rows else np.zeros((0, d), np.float32))
    sids = np.ascontiguousarray(np.concatenate(send_ids) if send_ids else np.zeros(0, np.int64))
    rrows = np.zeros((sum(rb), d), np.float32)
    rids = np.zeros(sum(rb), np.int64)
    coll.alltoallv(srows.ctypes.data, [b * d * 4 for b in sb], rrows.ctypes.data, [b * d * 4 for b in rb])
    coll.alltoallv(sids.ctypes.data, [b * 8 for b in sb], rids.ctypes.data, [b * 8 for b in rb])
    # owned ranges: pieces in source-rank order
    src_off = np.cumsum([0] + rb[:-1]).tolist()
    for i, sg in enumerate(segs):
        if owner[i] != rank:
            continue
        pr, pi = [], []
        for g in range(world):
            ln = counts[g][i]
            pr.append(rrows[src_off[g]:src_off[g] + ln])
            pi.append(rids[src_off[g]:src_off[g] + ln])
            src_off[g] += ln
        t = oracle.build(np.concatenate(pi), np.concatenate(pr), oracle.MODE_QFX, root_rid=sg["rid"], root_depth=depth,
                         qe=e)
        for r, dm, m, i_ in zip(t.range_id, t.dimension, t.mid, t.id):
            table[int(r)] = (int(dm), int(np.float32(m).view(np.uint32)), int(i_))
    return table


def _worker_protocol(rank, world, port, q, n, d, seed):
    _init(rank, world, port)
    from vectorindex import synthetic as ds
    from vectorindex.distributed import Collectives
    ids, rows = ds.unit_gaussian(n, d, seed=seed)
    ids = ids * 3 + 7
    lo, hi = rank * n // world, (rank + 1) * n // world
    coll = Collectives(torch.device("cpu"))
    table = emulate_rank(rank, world, ids[lo:hi].copy(), rows[lo:hi].copy(), coll)
    q.put((rank, table))
    dist.destroy_process_group()


@pytest.mark.parametrize("n,d,seed", [(3000, 8, 4), (257, 5, 9)])
def test_sharded_protocol_equals_single_rank_oracle(n, d, seed):
    import oracle
    from vectorindex import synthetic as ds
    world, port = 2, 29512 + seed
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_protocol, args=(r, world, port, q, n, d, seed)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(60)
    union = {}
    for r in range(world):
        for key, val in res[r].items():
            assert union.setdefault(key, val) == val  # shared rows agree between ranks
    ids, rows = ds.unit_gaussian(n, d, seed=seed)
    ids = ids * 3 + 7
    ref = oracle.build(ids, rows, oracle.MODE_QFX)
    want = {int(r): (int(dm), int(np.float32(m).view(np.uint32)), int(i))
            for r, dm, m, i in zip(ref.range_id, ref.dimension, ref.mid, ref.id)}
    assert union == want
    # the split is real: both ranks own some sub-tree rows
    L = 2
    deep = [set(k for k in res[r] if math.floor(math.log2(k + 1)) > L) for r in range(world)]
    assert all(len(s) > 0 for s in deep) and not (deep[0] & deep[1])


# ---- exact-mode multi-rank protocol (vi_build.cu build_sharded_exact), restated over the literal oracle -------------------
@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("n,d,seed", [(4000, 6, 21), (300, 3, 22)])
def test_exact_mode_protocol_equals_single_rank_oracle(world, n, d, seed):
    """Every rank holds all points, replays levels 0 .. L-1 (L = ceil(log2 world)) of the literal build, owns some of
    the level-L ranges (largest first to the least loaded rank) and builds their sub-trees from the range's points in
    stable order; the common rows plus the owned sub-trees must be the single-rank table, rank by rank disjoint."""
    import oracle
    rng = np.random.default_rng(seed)
    rows = rng.uniform(-1, 1, (n, d)).astype(np.float32)
    ids = (rng.permutation(n).astype(np.int64) * 3) + 1
    ref = oracle.build(ids, rows, oracle.MODE_LITERAL)
    want = {int(r): (int(dm), int(np.float32(m).view(np.uint32)), int(i))
            for r, dm, m, i in zip(ref.range_id, ref.dimension, ref.mid, ref.id)}
    L = 1
    while (1 << L) < world:
        L += 1
    # the common top: partitions replayed from the rows of levels < L (what every rank computes for itself)
    ranges = {0: np.arange(n)}
    common = {}
    for level in range(L):
        nxt = {}
        for rid, idx in ranges.items():
            dim, midbits, pivot = want[rid]
            common[rid] = want[rid]
            if dim == -1:
                continue
            mid = np.array([midbits], np.uint32).view(np.float32)[0]
            v = rows[idx, dim]
            hi = (v > mid) | ((v == mid) & (ids[idx] > pivot))
            for side, m in ((1, ~hi), (2, hi)):
                if m.any():
                    nxt[2 * rid + side] = idx[m]           # stable: index order is kept
        ranges = nxt
    open_ranges = sorted(ranges)                            # position order = RangeID order inside a level
    sizes = [len(ranges[r]) for r in open_ranges]
    order = sorted(range(len(open_ranges)), key=lambda i: -sizes[i])
    load = [0] * world
    owner = {}
    for i in order:
        best = min(range(world), key=lambda g: (load[g], g))
        owner[open_ranges[i]] = best
        load[best] += sizes[i]
    union = dict(common)
    for rank in range(world):
        for rid in open_ranges:
            if owner[rid] != rank:
                continue
            idx = ranges[rid]
            sub = oracle.build(ids[idx], rows[idx], oracle.MODE_LITERAL, root_rid=rid, root_depth=L)
            for r, dm, m, i in zip(sub.range_id, sub.dimension, sub.mid, sub.id):
                assert int(r) not in union                  # no row is made twice
                union[int(r)] = (int(dm), int(np.float32(m).view(np.uint32)), int(i))
    assert union == want
