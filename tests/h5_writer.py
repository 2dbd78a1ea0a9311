"""Test infrastructure: writes small HDF5 files byte by byte from the HDF5 File Format Specification, in the two shapes
the native reader (csrc/vi_hdf5.cu) accepts -- neither h5py nor libhdf5 exists in this image, so the fixtures cannot be
produced by the real library; this writer follows the layout libhdf5 itself emits:

  write_v0: superblock version 0 (what h5py writes by default, libver "earliest"): root group = object header v1 with a
            Symbol Table message -> B-tree v1 node ("TREE") -> symbol table node ("SNOD") + local heap ("HEAP") for the
            names; each data set = object header v1 with Dataspace (v1), Datatype (v1), Fill Value, Data Layout (v3,
            contiguous) messages, optionally split over a continuation block.  Nested groups are supported.
  write_v2: superblock version 2, object headers v2 ("OHDR", with checksums), new-style groups with compact Link
            messages; data layout v3 contiguous or (to test the refusal) chunked.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


def _dtype_msg(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind == "f":
        size = dt.itemsize
        exp_bits, man_bits = {4: (8, 23), 8: (11, 52)}[size]
        bias = (1 << (exp_bits - 1)) - 1
        head = bytes([0x11, 0x20, size * 8 - 1, 0x00]) + struct.pack("<I", size)
        props = struct.pack("<HHBBBBI", 0, size * 8, man_bits, exp_bits, 0, man_bits, bias)
        return head + props
    if dt.kind in "iu":
        head = bytes([0x10, 0x08 if dt.kind == "i" else 0x00, 0x00, 0x00]) + struct.pack("<I", dt.itemsize)
        return head + struct.pack("<HH", 0, dt.itemsize * 8)
    raise ValueError(dt)


def _dataspace_msg(shape) -> bytes:
    # version 1: version, rank, flags (bit 0: max dims present), reserved (5), dims, max dims
    return bytes([1, len(shape), 1, 0, 0, 0, 0, 0]) + b"".join(struct.pack("<Q", s) for s in shape) * 2


def _layout_msg(addr: int, size: int, chunked=False, rank=2) -> bytes:
    if chunked:
        # version 3, class 2: dimensionality (rank + 1), B-tree address, dimension sizes (4 bytes each), element size
        return bytes([3, 2, rank + 1]) + struct.pack("<Q", UNDEF) + b"".join(struct.pack("<I", 16) for _ in range(rank)) + struct.pack("<I", 4)
    return bytes([3, 1]) + struct.pack("<QQ", addr, size)


def _v1_msg(mtype: int, data: bytes) -> bytes:
    data = data + b"\0" * (-len(data) % 8)
    return struct.pack("<HHB3x", mtype, len(data), 0) + data


def _v1_header(msgs: list[bytes], nmsgs: int, total: int | None = None) -> bytes:
    body = b"".join(msgs)
    size = len(body) if total is None else total
    return struct.pack("<BBHII", 1, 0, nmsgs, 1, size) + b"\0" * 4 + body


class _File:
    def __init__(self):
        self.buf = bytearray()

    def tell(self) -> int:
        return len(self.buf)

    def align(self, a=8):
        self.buf += b"\0" * (-len(self.buf) % a)

    def put(self, b: bytes) -> int:
        self.align()
        at = len(self.buf)
        self.buf += b
        return at

    def patch(self, at: int, b: bytes):
        self.buf[at:at + len(b)] = b


def _write_dataset_v1(f: _File, arr: np.ndarray, split_header: bool, chunked: bool = False) -> int:
    arr = np.ascontiguousarray(arr)
    data_at = f.put(arr.tobytes()) if arr.size and not chunked else UNDEF
    m_space = _v1_msg(0x01, _dataspace_msg(arr.shape))
    m_type = _v1_msg(0x03, _dtype_msg(arr.dtype))
    m_fill = _v1_msg(0x05, bytes([2, 2, 2, 0]))  # fill value v2: allocate late, write if set, undefined
    m_layout = _v1_msg(0x08, _layout_msg(data_at, arr.nbytes, chunked, arr.ndim))
    m_filter = _v1_msg(0x0B, bytes([1, 1, 0, 0, 0, 0, 0, 0]) + struct.pack("<HHHH", 1, 0, 1, 1) + struct.pack("<I", 6) + b"\0" * 4) if chunked else b""
    if not split_header:
        msgs = [m_space, m_type, m_fill, m_layout] + ([m_filter] if chunked else [])
        return f.put(_v1_header(msgs, len(msgs)))
    # the layout message lives in a continuation block, as libhdf5 does when a header outgrows its first allocation
    cont = m_layout + _v1_msg(0x00, b"\0" * 8)
    cont_at = f.put(cont)
    m_cont = _v1_msg(0x10, struct.pack("<QQ", cont_at, len(cont)))
    msgs = [m_space, m_type, m_fill, m_cont]
    return f.put(_v1_header(msgs, 6))


def _write_group_v1(f: _File, members: dict) -> int:
    """members: name -> ndarray | dict (nested group). Returns the group's object header address."""
    entries = []
    for name in sorted(members):  # symbol table entries are sorted by name
        v = members[name]
        if isinstance(v, dict):
            entries.append((name, _write_group_v1(f, v)))
        else:
            arr, opts = (v if isinstance(v, tuple) else (v, {}))
            entries.append((name, _write_dataset_v1(f, arr, opts.get("split_header", False), opts.get("chunked", False))))
    # local heap: offset 0 holds the empty string, then the names, 8-byte aligned
    heap_data = bytearray(b"\0" * 8)
    name_off = {}
    for name, _ in entries:
        name_off[name] = len(heap_data)
        nb = name.encode() + b"\0"
        heap_data += nb + b"\0" * (-len(nb) % 8)
    heap_data += b"\0" * 16  # a free block
    seg_at = f.put(bytes(heap_data))
    heap_at = f.put(b"HEAP" + bytes([0, 0, 0, 0]) + struct.pack("<QQQ", len(heap_data), len(heap_data) - 16, seg_at))
    # symbol table node(s): at most 2K = 8 entries each
    snods = []
    for i in range(0, max(len(entries), 1), 8):
        part = entries[i:i + 8]
        body = b"SNOD" + bytes([1, 0]) + struct.pack("<H", len(part))
        for name, addr in part:
            body += struct.pack("<QQII16x", name_off[name], addr, 0, 0)
        body += b"\0" * (40 * (8 - len(part)))
        snods.append((f.put(body), name_off[part[-1][0]] if part else 0))
    # one B-tree leaf-level node over the symbol table nodes: key0 = 0 (the empty string), key i = last name of child i
    tree = b"TREE" + bytes([0, 0]) + struct.pack("<H", len(snods)) + struct.pack("<QQ", UNDEF, UNDEF) + struct.pack("<Q", 0)
    for at, last in snods:
        tree += struct.pack("<QQ", at, last)
    tree_at = f.put(tree)
    msgs = [_v1_msg(0x11, struct.pack("<QQ", tree_at, heap_at))]
    return f.put(_v1_header(msgs, 1))


def write_v0(path: str, members: dict, userblock: int = 0):
    f = _File()
    # superblock v0: 8 signature, versions, sizes, K values, flags, 4 addresses, root symbol table entry (40 bytes)
    sb_len = 8 + 8 + 4 + 4 + 4 * 8 + 40
    f.buf += b"\0" * sb_len
    root = _write_group_v1(f, members)
    eof = f.tell()
    sb = b"\x89HDF\r\n\x1a\n" + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", 4, 16, 0)
    sb += struct.pack("<QQQQ", userblock, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII16x", 0, root, 0, 0)
    assert len(sb) == sb_len
    f.patch(0, sb)
    with open(path, "wb") as out:
        # with a user block every address in the file is relative to the base address = the user block's size
        out.write(b"U" * userblock)
        out.write(bytes(f.buf))


# ---- version 2 object headers -------------------------------------------------------------------------------------------
def _lookup3_stub(b: bytes) -> int:
    # the reader does not verify checksums; any 32-bit value keeps the layout right
    return zlib.crc32(b) & 0xFFFFFFFF


def _v2_msg(mtype: int, data: bytes) -> bytes:
    return struct.pack("<BHB", mtype, len(data), 0) + data


def _v2_header(msgs: list[bytes]) -> bytes:
    body = b"".join(msgs)
    head = b"OHDR" + bytes([2, 0x01]) + struct.pack("<H", len(body))  # flags: chunk-0 size in 2 bytes
    blob = head + body
    return blob + struct.pack("<I", _lookup3_stub(blob))


def write_v2(path: str, datasets: dict):
    f = _File()
    f.buf += b"\0" * 48  # superblock v2: signature, version, sizes, flags, 4 addresses, checksum
    links = []
    for name in datasets:
        v = datasets[name]
        arr, opts = (v if isinstance(v, tuple) else (v, {}))
        arr = np.ascontiguousarray(arr)
        chunked = opts.get("chunked", False)
        data_at = f.put(arr.tobytes()) if arr.size and not chunked else UNDEF
        space = bytes([2, arr.ndim, 0, 1]) + b"".join(struct.pack("<Q", s) for s in arr.shape)  # dataspace v2, simple
        msgs = [_v2_msg(0x01, space), _v2_msg(0x03, _dtype_msg(arr.dtype)),
                _v2_msg(0x08, _layout_msg(data_at, arr.nbytes, chunked, arr.ndim))]
        links.append((name, f.put(_v2_header(msgs))))
    lmsgs = [_v2_msg(0x02, bytes([0, 0]) + struct.pack("<QQ", UNDEF, UNDEF))]  # link info: no fractal heap = compact
    for name, at in links:
        nb = name.encode()
        lmsgs.append(_v2_msg(0x06, bytes([1, 0x00, len(nb)]) + nb + struct.pack("<Q", at)))  # hard link, 1-byte name length
    root = f.put(_v2_header(lmsgs))
    eof = f.tell()
    sb = b"\x89HDF\r\n\x1a\n" + bytes([2, 8, 8, 0]) + struct.pack("<QQQQ", 0, UNDEF, eof, root)
    sb += struct.pack("<I", _lookup3_stub(sb))
    assert len(sb) == 48
    f.patch(0, sb)
    with open(path, "wb") as out:
        out.write(bytes(f.buf))
