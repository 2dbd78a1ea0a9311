"""GPU, BASELINE.json full size (10M x 96 unit-Gaussian, configs[1]): the oracle would take ~40 s per mode here, so the
table is checked through size-independent properties of the reference algorithm:

  * one leaf per point, leaf ids are a permutation of the input ids, leaves have Mid == 0;
  * every non-leaf row has children 2r+1 / 2r+2 (both, except for a handful of one-sided splits) => 2N-1 (+ those)
    rows, every row except the root is some row's child, RangeIDs are unique;
  * depth parity: the split dimension of the root is the max-variance one (checked against a float64 recomputation);
  * the root's Mid is the mean of that dimension (exact mode: float32 Welford drift allowed; fast mode: 1e-6 * max|x|);
  * partition consistency on a sample: each sampled point is found by a p = 0 search (the walk from RangeID 0
    follows exactly the sides the build put it on), and a p > 0 search returns a superset of the L-infinity box;
  * exact and fast mode agree on the top of the tree (same split dimensions on the first 3 levels).
"""
import numpy as np
import pytest
import torch

import vectorindex as vi

pytestmark = pytest.mark.gpu

N, D = 10_000_000, 96


@pytest.fixture(scope="module")
def data():
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(2)
    rows = torch.empty((N, D), dtype=torch.float32, device=dev)
    for s in range(0, N, 1 << 20):
        e = min(N, s + (1 << 20))
        x = torch.randn((e - s, D), generator=g, device=dev, dtype=torch.float32)
        rows[s:e] = x / x.norm(dim=1, keepdim=True)
    ids = torch.arange(N, dtype=torch.int64, device=dev) * 3 + 11
    return ids, rows


def _build(ids, rows, mode):
    ctx = vi.Context(0)
    ctx.reserve(N, D)
    ctx.add_device(ids.data_ptr(), rows.data_ptr(), N, D)
    info = ctx.build(mode)
    return ctx, info


@pytest.mark.parametrize("mode", [vi.MODE_FAST, vi.MODE_EXACT])
def test_full_size_structure_and_search(data, mode):
    ids, rows = data
    ctx, info = _build(ids, rows, mode)
    rid, dim, mid, oid = ctx.ranges()
    # 2N-1 rows when every split has two non-empty sides; a one-sided split (two values closer than the arithmetic can
    # separate, the id tie-break sending both the same way) adds a one-child row -- a handful per million points
    extra = len(rid) - (2 * N - 1)
    assert info.ranges == len(rid) and 0 <= extra <= N // 10000
    assert len(np.unique(rid)) == len(rid)
    leaves = dim == -1
    assert int(leaves.sum()) == N
    assert np.all(mid[leaves] == 0)
    assert np.array_equal(np.sort(oid[leaves]), ids.cpu().numpy())  # leaf ids: a permutation of the input ids
    have = np.sort(rid)
    internal = rid[~leaves]
    present = []
    for child in (2 * internal + 1, 2 * internal + 2):
        pos = np.minimum(np.searchsorted(have, child), len(have) - 1)
        present.append(have[pos] == child)
    nchild = present[0].astype(np.int64) + present[1].astype(np.int64)
    assert np.all(nchild >= 1)                       # no internal row without children
    assert int((nchild == 1).sum()) == extra         # one-child rows account exactly for the extra rows
    assert int(nchild.sum()) == len(rid) - 1         # every row but the root is some row's child
    assert int(rid.max()).bit_length() <= 62
    # root: max-variance dimension and its mean
    root = int(np.nonzero(rid == 0)[0][0])
    var = rows.double().var(dim=0, unbiased=False).cpu().numpy()
    order = np.argsort(-var)
    assert int(dim[root]) in order[:2].tolist()  # the two leading variances may be closer than float32 noise
    mean = float(rows[:, int(dim[root])].double().mean())
    scale = float(rows.abs().max())
    tol = 1e-6 * scale if mode == vi.MODE_FAST else 2e-4 * scale  # float32 recurrence drift ~ sqrt(n) * ulp
    assert abs(float(mid[root]) - mean) <= tol
    # searches: sampled dataset points are found at p = 0; p > 0 returns a superset of the box
    pick = torch.randint(0, N, (2000,), device=rows.device, generator=torch.Generator(device=rows.device).manual_seed(5))
    q = rows[pick].cpu().numpy()
    offs, out = ctx.search(q, 0.0)
    want = ids[pick].cpu().numpy()
    for i in range(len(q)):
        assert want[i] in out[offs[i]:offs[i + 1]]
    p = 0.05
    offs, out = ctx.search(q[:20], p)
    for i in range(20):
        box = torch.nonzero(((rows - torch.from_numpy(q[i]).to(rows.device)).abs() <= p).all(dim=1)).flatten()
        got = set(out[offs[i]:offs[i + 1]].tolist())
        assert set(ids[box].cpu().numpy().tolist()) <= got
    ctx.close()


def test_full_size_modes_agree_on_top_levels(data):
    ids, rows = data
    tops = []
    for mode in (vi.MODE_FAST, vi.MODE_EXACT):
        ctx, _ = _build(ids, rows, mode)
        rid, dim, mid, oid = ctx.ranges()
        sel = rid < 7
        o = np.argsort(rid[sel])
        tops.append((dim[sel][o], mid[sel][o], oid[sel][o]))
        ctx.close()
    assert np.array_equal(tops[0][0], tops[1][0])            # same split dimensions on levels 0..2
    assert np.array_equal(tops[0][2][:1], tops[1][2][:1])    # same root pivot id
    assert np.all(np.abs(tops[0][1] - tops[1][1]) <= 2e-4)   # Mid differs by the float32 recurrence's drift only
