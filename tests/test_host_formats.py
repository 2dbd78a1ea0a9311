"""CPU: the text / binary formats on either side of the path (host helpers only; no GPU, no library calls).

CSV "RangeID,Dimension,Mid,ID" as VectorIndex.MainTest/Program.cs:80,145-149 writes it; records [int64 id][D x float32]
as FileRangeStore.cs:127-165 lays them out."""
import numpy as np
import pytest

import vectorindex as vi


def test_csv_round_trip_is_bit_exact(tmp_path):
    rng = np.random.default_rng(3)
    bits = rng.integers(0, 2 ** 32, 5000, dtype=np.uint64).astype(np.uint32)
    mid = bits.view(np.float32).copy()
    mid[:8] = [0.0, -0.0, 1e-45, -1e-45, 3.4028235e38, np.inf, -np.inf, 0.1]
    keep = ~np.isnan(mid)
    mid = mid[keep]
    n = len(mid)
    rid = rng.permutation(n).astype(np.int64) * 3
    dim = rng.integers(-1, 96, n).astype(np.int32)
    oid = rng.integers(-2 ** 62, 2 ** 62, n)
    path = str(tmp_path / "t.csv")
    vi.write_csv(path, rid, dim, mid, oid)
    with open(path, newline="") as f:
        assert f.readline() == "RangeID,Dimension,Mid,ID\r\n"
    r2, d2, m2, o2 = vi.read_csv(path)
    assert np.array_equal(r2, rid) and np.array_equal(d2, dim) and np.array_equal(o2, oid)
    assert np.array_equal(m2.view(np.uint32), mid.view(np.uint32))


def test_csv_nan_and_header(tmp_path):
    path = str(tmp_path / "t.csv")
    vi.write_csv(path, np.array([0]), np.array([2], np.int32), np.array([np.nan], np.float32), np.array([5]))
    assert "NaN" in open(path).read()
    assert np.isnan(vi.read_csv(path)[2][0])
    with open(path, "w") as f:
        f.write("a,b\n1,2\n")
    with pytest.raises(ValueError):
        vi.read_csv(path)


def test_float_text_is_shortest_round_trip():
    # what float.ToString() produces on .NET Core 3.0+: the shortest digits that parse back to the same float32
    assert vi._fmt_float32(np.float32(0.1)) == "0.1"
    assert vi._fmt_float32(np.float32(0.5)) == "0.5"
    assert vi._fmt_float32(np.float32(1.0)) == "1"
    assert vi._fmt_float32(np.float32(1.2345679e-07)) == "1.2345679E-07"
    assert vi._fmt_float32(np.float32(-3e20)) == "-3E+20"


def test_record_layout():
    ids = np.array([1, -2, 2 ** 40], np.int64)
    rows = np.arange(6, dtype=np.float32).reshape(3, 2)
    rec = vi.pack_records(ids, rows)
    assert rec.dtype == np.uint8 and rec.size == 3 * (8 + 2 * 4)
    r = rec.reshape(3, 16)
    assert np.array_equal(r[:, :8].copy().view(np.int64).ravel(), ids)          # little-endian long first
    assert np.array_equal(r[:, 8:].copy().view(np.float32).reshape(3, 2), rows)  # then the floats, no padding
