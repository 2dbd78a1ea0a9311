"""Native HDF5 reader (csrc/vi_hdf5.cu) -- SURVEY.md 8(f) rank 1: the `/train`, `/test` loader of
VectorIndex.MainTest/Program.cs:183-260.  Files are written by tests/h5_writer.py from the format specification (no
h5py / libhdf5 in this image).  CPU tests: container parsing (host only, no compute); GPU test: the rows streamed from
the file build the same index as the same rows added from memory."""
import numpy as np
import pytest

import vectorindex as vi
from h5_writer import write_v0, write_v2


def _data(seed=3, n=1000, d=24, nq=37):
    rng = np.random.default_rng(seed)
    train = rng.standard_normal((n, d), dtype=np.float32)
    test = rng.standard_normal((nq, d), dtype=np.float32)
    neighbors = rng.integers(0, n, (nq, 10)).astype(np.int32)
    distances = rng.random((nq, 10)).astype(np.float64)
    return train, test, neighbors, distances


def test_ann_benchmark_layout_v0(tmp_path):
    # the four data sets of an ann-benchmarks file, as h5py lays them out by default
    train, test, neighbors, distances = _data()
    path = str(tmp_path / "ann.hdf5")
    write_v0(path, {"train": train, "test": test, "neighbors": neighbors, "distances": distances})
    for name, arr in (("/train", train), ("test", test), ("/neighbors", neighbors), ("distances", distances)):
        rows, cols, dt, off = vi.hdf5_dataset_info(path, name)     # Program.cs:183-222 GetHdf5DatasetSize
        assert (rows, cols, dt) == (arr.shape[0], arr.shape[1], arr.dtype)
        assert np.array_equal(np.fromfile(path, dt, rows * cols, offset=off).reshape(rows, cols), arr)
        assert np.array_equal(vi.hdf5_read(path, name), arr)
    # Program.cs:233-238 reads blocks [index, index + step - 1]
    assert np.array_equal(vi.hdf5_read(path, "/train", 100, 250), train[100:350])
    assert vi.hdf5_read(path, "/train", 1000, 0).shape == (0, 24)
    with pytest.raises(ValueError, match="rows outside"):
        vi.hdf5_read(path, "/train", 900, 200)
    with pytest.raises(ValueError, match="no object named 'nope'"):
        vi.hdf5_dataset_info(path, "/nope")


def test_nested_groups_many_members_continuation_and_user_block(tmp_path):
    train, test, neighbors, _ = _data(seed=4, n=300, d=7)
    members = {f"m{i:02d}": np.full((2, 3), i, np.int32) for i in range(19)}   # three symbol table nodes
    members["train"] = (train, {"split_header": True})                         # layout message in a continuation block
    members["grp"] = {"inner": {"test": test}, "vec": neighbors[:, 0].copy()}
    path = str(tmp_path / "nested.h5")
    write_v0(path, members, userblock=512)                                     # addresses relative to the base address
    assert np.array_equal(vi.hdf5_read(path, "/train"), train)
    assert np.array_equal(vi.hdf5_read(path, "/grp/inner/test"), test)
    assert np.array_equal(vi.hdf5_read(path, "m17"), np.full((2, 3), 17, np.int32))
    rows, cols, dt, _ = vi.hdf5_dataset_info(path, "grp/vec")                  # rank 1: one column
    assert (rows, cols, dt) == (neighbors.shape[0], 1, np.dtype("<i4"))
    with pytest.raises(ValueError, match="no object named"):
        vi.hdf5_dataset_info(path, "/grp/missing/test")


def test_superblock_v2_object_headers_v2_and_link_messages(tmp_path):
    train, test, neighbors, _ = _data(seed=5, n=128, d=16)
    path = str(tmp_path / "latest.h5")
    write_v2(path, {"train": train, "test": test, "neighbors": neighbors})
    assert np.array_equal(vi.hdf5_read(path, "/train"), train)
    assert np.array_equal(vi.hdf5_read(path, "/test"), test)
    assert np.array_equal(vi.hdf5_read(path, "neighbors"), neighbors)


def test_refusals(tmp_path):
    train, _, _, _ = _data(seed=6, n=64, d=8)
    p = str(tmp_path / "bad.h5")
    write_v0(p, {"cube": np.zeros((2, 3, 4), np.float32), "chunked": (train, {"chunked": True}), "train": train})
    with pytest.raises(ValueError, match="Invalid rank"):                      # Program.cs:203-206
        vi.hdf5_dataset_info(p, "/cube")
    with pytest.raises(ValueError, match="chunked"):
        vi.hdf5_dataset_info(p, "/chunked")
    q = str(tmp_path / "bad2.h5")
    write_v2(q, {"chunked": (train, {"chunked": True})})
    with pytest.raises(ValueError, match="chunked"):
        vi.hdf5_dataset_info(q, "/chunked")
    not_h5 = tmp_path / "plain.bin"
    not_h5.write_bytes(b"\0" * 4096)
    with pytest.raises(ValueError, match="not an HDF5 file"):
        vi.hdf5_dataset_info(str(not_h5), "/train")
    with pytest.raises(ValueError, match="cannot open"):
        vi.hdf5_dataset_info(str(tmp_path / "absent.h5"), "/train")
    # a truncated file: the data set's storage lies outside what is left
    raw = open(p, "rb").read()
    cut = tmp_path / "cut.h5"
    cut.write_bytes(raw[:len(raw) // 2])
    with pytest.raises(ValueError):
        vi.hdf5_read(str(cut), "/train")


@pytest.mark.gpu
def test_hdf5_ingest_builds_the_same_index(tmp_path):
    rng = np.random.default_rng(8)
    n, d = 300_000, 96                                                        # several 64 MB batches
    train = rng.standard_normal((n, d), dtype=np.float32)
    train /= np.linalg.norm(train, axis=1, keepdims=True)
    path = str(tmp_path / "deep-image-like.hdf5")
    write_v0(path, {"train": train, "test": train[:10]})
    rows, cols, dt, _ = vi.hdf5_dataset_info(path, "/train")
    with vi.Context(0) as a, vi.Context(0) as b:
        a.reserve(rows, cols)
        a.add_hdf5(path, "/train")                                            # ids = row indexes (Program.cs:252)
        assert a.count == n
        b.reserve(n, d)
        b.add(np.arange(n, dtype=np.int64), train)
        for mode in (vi.MODE_FAST, vi.MODE_EXACT):
            a.build(mode)
            b.build(mode)
            ta, tb = a.ranges(), b.ranges()
            oa, ob = np.argsort(ta[0]), np.argsort(tb[0])
            for x, y in zip(ta, tb):
                assert np.array_equal(x[oa].view(np.uint8), y[ob].view(np.uint8))
        # a slice with its own ids, appended in two calls
        a.reserve(1000, d)
        a.add_hdf5(path, "/train", first_row=5000, n=600, first_id=70_000)
        a.add_hdf5(path, "/train", first_row=5600, n=400, first_id=70_600)
        b.reserve(1000, d)
        b.add(np.arange(70_000, 71_000, dtype=np.int64), train[5000:6000])
        a.build(vi.MODE_FAST)
        b.build(vi.MODE_FAST)
        ta, tb = a.ranges(), b.ranges()
        oa, ob = np.argsort(ta[0]), np.argsort(tb[0])
        for x, y in zip(ta, tb):
            assert np.array_equal(x[oa].view(np.uint8), y[ob].view(np.uint8))
        b.reserve(10, 95)
        with pytest.raises(ValueError, match="Invalid length of vector"):
            b.add_hdf5(path, "/train")


def test_corrupted_files_are_refused_or_read_never_crash(tmp_path):
    # random byte corruptions and truncations of both container generations: the parser must answer with an error or
    # with a data set that lies inside the file -- every read it then does is bounds-checked against the file size
    rng = np.random.default_rng(1)
    train = rng.standard_normal((200, 8), dtype=np.float32)
    a, b = str(tmp_path / "a.h5"), str(tmp_path / "b.h5")
    write_v0(a, {"train": train, "test": train[:5], "g": {"x": train[:3]}})
    write_v2(b, {"train": train, "test": train[:5]})
    answered = refused = 0
    for base in (a, b):
        raw = open(base, "rb").read()
        for trial in range(250):
            buf = bytearray(raw)
            for _ in range(int(rng.integers(1, 4))):
                buf[int(rng.integers(0, len(buf)))] = int(rng.integers(0, 256))
            if trial % 10 == 0:
                buf = buf[: int(rng.integers(8, len(buf)))]
            q = tmp_path / "c.h5"
            q.write_bytes(bytes(buf))
            for name in ("/train", "/test", "/g/x"):
                try:
                    rows, cols, dt, off = vi.hdf5_dataset_info(str(q), name)
                    assert rows >= 0 and cols >= 0 and off + rows * cols * dt.itemsize <= len(buf)
                    if rows * cols * dt.itemsize < (1 << 24):
                        vi.hdf5_read(str(q), name)
                    answered += 1
                except ValueError:
                    refused += 1
    assert answered > 0 and refused > 0
