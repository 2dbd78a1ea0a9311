"""GPU, BASELINE.json full size (10M x 96 unit-Gaussian, configs[1]).

`test_full_size_bit_exact_vs_oracle`: the whole 10M x 96 table, both modes, bit for bit against oracle/vi_oracle.c
(about 45 s of CPU per mode), plus the exact-vs-fast divergence at full size (rows in both tables, same Dimension,
bit-identical rows, points whose root-to-leaf path is the same).  configs[4]'s width is covered at 200k x 768 in
tests/test_gpu_parity.py.  The remaining tests check size-independent properties of the reference algorithm
(SURVEY.md 8c (6), (7)):

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


def _sorted_table(ctx):
    rid, dim, mid, oid = ctx.ranges()
    o = np.argsort(rid, kind="stable")
    return rid[o], dim[o], mid[o], oid[o]


def test_full_size_bit_exact_vs_oracle(data):
    """BASELINE configs[1] itself: every row of the 10M x 96 table equals the oracle's, in both modes."""
    import oracle
    ids, rows = data
    ids_h = ids.cpu().numpy()
    rows_h = rows.cpu().numpy()
    tables = {}
    for mode, omode in ((vi.MODE_FAST, oracle.MODE_QFX), (vi.MODE_EXACT, oracle.MODE_LITERAL)):
        ctx, info = _build(ids, rows, mode)
        rid, dim, mid, oid = _sorted_table(ctx)
        ctx.close()
        ref = oracle.build(ids_h, rows_h, omode)
        assert len(rid) == len(ref) == info.ranges
        assert np.array_equal(rid, ref.range_id)
        assert np.array_equal(dim, ref.dimension)
        assert np.array_equal(mid.view(np.uint32), ref.mid.view(np.uint32)), "Mid must be bit-identical"
        assert np.array_equal(oid, ref.id)
        tables[mode] = (rid, dim, mid, oid)
        del ref
    # divergence of the fast mode from the literal one at full size (reported by bench.py as `divergence`):
    # the two tables describe different trees only where a point within float32 noise of a Mid changes side
    (fr, fd, fm, fo), (er, ed, em, eo) = tables[vi.MODE_FAST], tables[vi.MODE_EXACT]
    common, fi, ei = np.intersect1d(fr, er, assume_unique=True, return_indices=True)
    same_dim = fd[fi] == ed[ei]
    scale = float(rows.abs().max())
    # (measured at 10M x 96: 83 % of the RangeIDs exist in both tables, 64 % of those with the same Dimension -- once
    # one point changes sides near the top, every range below holds a slightly different set; bench.py reports the
    # figures as `divergence`.  Asserted here: the trees are recognisably the same tree, and identical at the top.)
    assert len(common) >= 0.7 * len(er)
    assert same_dim.mean() >= 0.5
    # levels 0..3 (ranges of >= 600k points): where both trees chose the same dimension, Mid differs by the literal
    # recurrence's own drift, O(sqrt(n) ulp) -- deeper down one point changing sides moves a small range's mean more
    # (the first range on which the two modes choose different dimensions is on level 4 at this size: below it the two
    # tables describe different point sets, so the comparison stops at level 3)
    top = (common < 15) & same_dim
    dmid = np.abs(fm[fi][top].astype(np.float64) - em[ei][top].astype(np.float64))
    assert float(dmid.max()) <= 1e-3 * scale
    # levels 0..2: identical split dimensions, and the same root pivot id
    assert np.array_equal(fd[fi][common < 7], ed[ei][common < 7])
    assert fo[0] == eo[0]
    # points whose root-to-leaf path is the same in both trees: same leaf RangeID for the same id
    fl, el = fd == -1, ed == -1
    fo_, eo_ = np.argsort(fo[fl]), np.argsort(eo[el])
    same_path = float((fr[fl][fo_] == er[el][eo_]).mean())
    print(f"10M divergence: rows in both {len(common)}/{len(er)}, same Dimension {int(same_dim.sum())}, "
          f"bit-identical {int((same_dim & (fm[fi].view(np.uint32) == em[ei].view(np.uint32)) & (fo[fi] == eo[ei])).sum())}, "
          f"same path {same_path:.4f}")
