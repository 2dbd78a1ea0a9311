"""Golden fixtures (tests/golden/, written by tests/golden/make_golden.py).

  kat_hand_derived.json   known answers derived by hand from IndexBuilder.cs / DDL.sql  -> oracle (CPU) and CUDA (GPU)
  frozen_*.npz            frozen oracle outputs on small seeded inputs                   -> oracle, numpy restatement, CUDA

The CPU tests pin the oracle; the GPU tests compare the CUDA path with the same files through the C ABI."""
import glob
import json
import os

import numpy as np
import pytest

import oracle
from oracle import np_oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KAT = json.load(open(os.path.join(GOLDEN, "kat_hand_derived.json")))
FROZEN = sorted(glob.glob(os.path.join(GOLDEN, "frozen_*.npz")))
MODES = [(oracle.MODE_LITERAL, "lit"), (oracle.MODE_QFX, "qfx")]


def _want(k):
    return {int(r): (int(v[0]), np.float32(v[1]).view(np.uint32).item(), int(v[2])) for r, v in k["table"].items()}


def _as_dict(rid, dim, mid, oid):
    return {int(r): (int(d), np.float32(m).view(np.uint32).item(), int(i)) for r, d, m, i in zip(rid, dim, mid, oid)}


def test_fixture_files_exist():
    assert len(KAT["tables"]) >= 5 and len(FROZEN) == 3


@pytest.mark.parametrize("mode,tag", MODES)
@pytest.mark.parametrize("k", KAT["tables"], ids=lambda k: k["name"][:40])
def test_oracle_reproduces_hand_derived_tables(k, mode, tag):
    t = oracle.build(np.array(k["ids"], np.int64), np.array(k["rows"], np.float32), mode)
    assert _as_dict(t.range_id, t.dimension, t.mid, t.id) == _want(k)


def test_oracle_reproduces_hand_derived_search():
    s = KAT["search"]
    t = oracle.build(np.array(s["ids"], np.int64), np.array(s["rows"], np.float32), oracle.MODE_LITERAL)
    for c in s["cases"]:
        _, out, _ = oracle.search(t, np.array([c["query"]], np.float32), c["p"])
        assert out.tolist() == c["ids"], c


@pytest.mark.parametrize("mode,tag", MODES)
@pytest.mark.parametrize("path", FROZEN, ids=os.path.basename)
def test_oracle_matches_frozen_outputs(path, mode, tag):
    z = np.load(path)
    t = oracle.build(z["ids"], z["rows"], mode)
    assert np.array_equal(t.range_id, z[f"{tag}_rid"]) and np.array_equal(t.dimension, z[f"{tag}_dim"])
    assert np.array_equal(t.mid.view(np.uint32), z[f"{tag}_mid"].view(np.uint32)) and np.array_equal(t.id, z[f"{tag}_id"])
    offs, cand, visits = oracle.search(t, z["queries"], float(z["proximity"]))
    assert np.array_equal(offs, z[f"{tag}_offsets"]) and np.array_equal(cand, z[f"{tag}_candidates"])
    assert visits == int(z[f"{tag}_visits"])


@pytest.mark.parametrize("path", [p for p in FROZEN if "2500x96" not in p], ids=os.path.basename)
def test_numpy_restatement_matches_frozen_outputs(path):
    # the independent numpy restatement (pure-Python loops: the two small sets only)
    z = np.load(path)
    for build, tag in ((np_oracle.build_literal, "lit"), (np_oracle.build_qfx, "qfx")):
        got = build(z["ids"], z["rows"])  # list of (rangeId, Dimension, Mid, Id)
        want = _as_dict(z[f"{tag}_rid"], z[f"{tag}_dim"], z[f"{tag}_mid"], z[f"{tag}_id"])
        assert {int(r): (int(d), np.float32(m).view(np.uint32).item(), int(i)) for r, d, m, i in got} == want


# ---- the CUDA path against the same files -----------------------------------------------------------------------------
def _gpu_table(ids, rows, mode):
    import vectorindex as vi
    with vi.Context(0) as ctx:
        ctx.reserve(len(ids), rows.shape[1])
        ctx.add(ids, rows)
        ctx.build(mode)
        return ctx.ranges()


@pytest.mark.gpu
@pytest.mark.parametrize("mode,tag", MODES)
@pytest.mark.parametrize("k", KAT["tables"], ids=lambda k: k["name"][:40])
def test_cuda_reproduces_hand_derived_tables(k, mode, tag):
    got = _gpu_table(np.array(k["ids"], np.int64), np.array(k["rows"], np.float32), mode)
    assert _as_dict(*got) == _want(k)


@pytest.mark.gpu
def test_cuda_reproduces_hand_derived_search():
    import vectorindex as vi
    s = KAT["search"]
    with vi.Context(0) as ctx:
        ctx.reserve(2, 3)
        ctx.add(np.array(s["ids"], np.int64), np.array(s["rows"], np.float32))
        ctx.build(vi.MODE_EXACT)
        for c in s["cases"]:
            _, out = ctx.search(np.array([c["query"]], np.float32), c["p"])
            assert out.tolist() == c["ids"], c


@pytest.mark.gpu
@pytest.mark.parametrize("mode,tag", MODES)
@pytest.mark.parametrize("path", FROZEN, ids=os.path.basename)
def test_cuda_matches_frozen_outputs(path, mode, tag):
    import vectorindex as vi
    z = np.load(path)
    with vi.Context(0) as ctx:
        ctx.reserve(len(z["ids"]), z["rows"].shape[1])
        ctx.add(z["ids"], z["rows"])
        ctx.build(mode)
        rid, dim, mid, oid = ctx.ranges()
        offs, cand = ctx.search(z["queries"], float(z["proximity"]))
    o = np.argsort(rid, kind="stable")
    assert np.array_equal(rid[o], z[f"{tag}_rid"]) and np.array_equal(dim[o], z[f"{tag}_dim"])
    assert np.array_equal(mid[o].view(np.uint32), z[f"{tag}_mid"].view(np.uint32)) and np.array_equal(oid[o], z[f"{tag}_id"])
    assert np.array_equal(offs, z[f"{tag}_offsets"]) and np.array_equal(cand, z[f"{tag}_candidates"])
