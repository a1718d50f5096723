"""TensorFlow-free reader/writer for the TF-1 ``Saver`` V2 checkpoint bundle.

``common/deploy_network.py:48-49`` restores ``<prefix>.index`` +
``<prefix>.data-00000-of-00001`` (written by ``common/train_network.py:241,
337-339``; renamed by ``demo_pipeline.py:50-54``).  TensorFlow is not
installable here, so the on-disk format is implemented from its published
layout (tensorflow/core/util/tensor_bundle, tensorflow/core/lib/io/table):

``.index``  a LevelDB-style sorted table: data blocks of prefix-compressed
            entries (varint shared, varint non_shared, varint value_len, key
            suffix, value) followed by a fixed32 restart array + count; every
            block is trailed by a 1-byte compression tag and a masked crc32c;
            then a metaindex block, an index block (last-key -> BlockHandle)
            and a 48-byte footer ending in magic 0xdb4775248b80fb57.
            key ""   -> BundleHeaderProto{num_shards=1, endianness=2, version=3}
            key name -> BundleEntryProto{dtype=1, shape=2, shard_id=3, offset=4,
                                         size=5, crc32c=6 (fixed32, masked)}
``.data-*`` raw little-endian row-major tensor bytes at [offset, offset+size).

The ``.meta`` MetaGraphDef is not needed: the graph is ``build_FCN`` itself.
Only uncompressed blocks are supported (TF writes the index uncompressed).
"""
from __future__ import annotations

import os
import struct
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
_MASK_DELTA = 0xA282EAD8

# tensorflow DataType enum values -> numpy
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16,
           6: np.int8, 9: np.int64, 10: np.bool_, 17: np.uint16, 19: np.float16,
           22: np.uint32, 23: np.uint64}
_DTYPE_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}

# ----------------------------------------------------------------------------- crc32c

_CRC_TABLE: Optional[List[int]] = None
_native_crc = None


def _crc_table() -> List[int]:
    global _CRC_TABLE
    if _CRC_TABLE is None:
        tab = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            tab.append(c)
        _CRC_TABLE = tab
    return _CRC_TABLE


def set_native_crc32c(fn) -> None:
    """The C-ABI library registers its ``ukbb_crc32c`` here (see _lib.py)."""
    global _native_crc
    _native_crc = fn


def crc32c(data: bytes) -> int:
    """Castagnoli CRC (polynomial 0x1EDC6F41 reflected), as used by the bundle."""
    if _native_crc is not None and len(data) >= 256:
        return _native_crc(data)
    tab = _crc_table()
    c = 0xFFFFFFFF
    for b in data:
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(c: int) -> int:
    return (((c >> 15) | (c << 17)) + _MASK_DELTA) & 0xFFFFFFFF


def unmask_crc(m: int) -> int:
    r = (m - _MASK_DELTA) & 0xFFFFFFFF
    return ((r >> 17) | (r << 15)) & 0xFFFFFFFF


# ----------------------------------------------------------------------------- varint / proto

def _put_varint(v: int) -> bytes:
    out = bytearray()
    v &= (1 << 64) - 1
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _get_varint(buf: bytes, pos: int) -> Tuple[int, int]:
    shift = 0
    val = 0
    while True:
        if pos >= len(buf):
            raise ValueError("truncated varint")
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7
        if shift > 63:
            raise ValueError("varint too long")


def _proto_fields(buf: bytes) -> Iterator[Tuple[int, int, object]]:
    pos = 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v = buf[pos:pos + ln]
            if len(v) != ln:
                raise ValueError("truncated length-delimited field")
            pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield field, wt, v


def _encode_shape(shape: Tuple[int, ...]) -> bytes:
    out = b""
    for d in shape:
        dim = b"\x08" + _put_varint(int(d))                 # Dim.size = 1
        out += b"\x12" + _put_varint(len(dim)) + dim        # TensorShapeProto.dim = 2
    return out


def _decode_shape(buf: bytes) -> Tuple[int, ...]:
    dims = []
    for f, wt, v in _proto_fields(buf):
        if f == 2 and wt == 2:
            size = 0
            for f2, wt2, v2 in _proto_fields(v):
                if f2 == 1 and wt2 == 0:
                    size = v2 if v2 < (1 << 63) else v2 - (1 << 64)
            dims.append(size)
    return tuple(dims)


def _encode_entry(dtype_code: int, shape, shard_id: int, offset: int, size: int, crc_masked: int) -> bytes:
    out = b"\x08" + _put_varint(dtype_code)
    sh = _encode_shape(shape)
    out += b"\x12" + _put_varint(len(sh)) + sh
    if shard_id:
        out += b"\x18" + _put_varint(shard_id)
    if offset:
        out += b"\x20" + _put_varint(offset)
    out += b"\x28" + _put_varint(size)
    out += b"\x35" + struct.pack("<I", crc_masked)
    return out


def _decode_entry(buf: bytes) -> dict:
    e = {"dtype": 0, "shape": (), "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "slices": 0}
    for f, wt, v in _proto_fields(buf):
        if f == 1:
            e["dtype"] = v
        elif f == 2:
            e["shape"] = _decode_shape(v)
        elif f == 3:
            e["shard_id"] = v
        elif f == 4:
            e["offset"] = v
        elif f == 5:
            e["size"] = v
        elif f == 6:
            e["crc32c"] = v
        elif f == 7:
            e["slices"] += 1
    return e


# ----------------------------------------------------------------------------- table

def _build_block(items: List[Tuple[bytes, bytes]], restart_interval: int = 16) -> bytes:
    out = bytearray()
    restarts = []
    last = b""
    for i, (k, v) in enumerate(items):
        if i % restart_interval == 0:
            restarts.append(len(out))
            shared = 0
        else:
            shared = 0
            m = min(len(k), len(last))
            while shared < m and k[shared] == last[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v))
        out += k[shared:] + v
        last = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _block_with_trailer(block: bytes) -> bytes:
    crc = mask_crc(crc32c(block + b"\x00"))
    return block + b"\x00" + struct.pack("<I", crc)


def _parse_block(raw: bytes) -> List[Tuple[bytes, bytes]]:
    if len(raw) < 4:
        raise ValueError("table block too small")
    n_restarts = struct.unpack_from("<I", raw, len(raw) - 4)[0]
    end = len(raw) - 4 - 4 * n_restarts
    if end < 0:
        raise ValueError("corrupt restart array")
    pos = 0
    key = b""
    items = []
    while pos < end:
        shared, pos = _get_varint(raw, pos)
        non_shared, pos = _get_varint(raw, pos)
        vlen, pos = _get_varint(raw, pos)
        key = key[:shared] + raw[pos:pos + non_shared]
        pos += non_shared
        items.append((key, raw[pos:pos + vlen]))
        pos += vlen
    return items


def _read_block(buf: bytes, offset: int, size: int, verify: bool) -> List[Tuple[bytes, bytes]]:
    raw = buf[offset:offset + size]
    trailer = buf[offset + size:offset + size + 5]
    if len(raw) != size or len(trailer) != 5:
        raise ValueError("table block out of file bounds")
    if trailer[0] != 0:
        raise NotImplementedError("compressed table blocks (type %d) are not supported" % trailer[0])
    if verify:
        want = unmask_crc(struct.unpack("<I", trailer[1:])[0])
        if crc32c(raw + trailer[:1]) != want:
            raise ValueError("table block checksum mismatch at offset %d" % offset)
    return _parse_block(raw)


def read_index(path: str, verify: bool = True) -> Dict[str, dict]:
    with open(path, "rb") as f:
        buf = f.read()
    if len(buf) < 48:
        raise ValueError("%s: too small to be a checkpoint index" % path)
    footer = buf[-48:]
    if struct.unpack("<Q", footer[40:])[0] != TABLE_MAGIC:
        raise ValueError("%s: bad table magic (not a TF V2 checkpoint index)" % path)
    pos = 0
    _, pos = _get_varint(footer, pos)      # metaindex offset
    _, pos = _get_varint(footer, pos)      # metaindex size
    idx_off, pos = _get_varint(footer, pos)
    idx_size, pos = _get_varint(footer, pos)
    entries: Dict[str, dict] = {}
    header = None
    for _, handle in _read_block(buf, idx_off, idx_size, verify):
        off, p = _get_varint(handle, 0)
        size, p = _get_varint(handle, p)
        for k, v in _read_block(buf, off, size, verify):
            if k == b"":
                header = {f: val for f, _, val in _proto_fields(v)}
            else:
                entries[k.decode("utf-8")] = _decode_entry(v)
    if header is None:
        raise ValueError("%s: missing bundle header entry" % path)
    if header.get(2, 0) != 0:
        raise NotImplementedError("big-endian bundles are not supported")
    entries["__header__"] = {"num_shards": header.get(1, 0)}
    return entries


def read_bundle(prefix: str, verify: bool = True, names=None) -> Dict[str, np.ndarray]:
    """Load tensors of a V2 checkpoint ``prefix`` (``prefix.index`` + shards)."""
    index = read_index(prefix + ".index", verify)
    n_shards = index.pop("__header__")["num_shards"] or 1
    shards: Dict[int, np.memmap] = {}
    out: Dict[str, np.ndarray] = {}
    for name, e in index.items():
        if names is not None and name not in names:
            continue
        if e["slices"]:
            raise NotImplementedError("partitioned variable %r is not supported" % name)
        if e["dtype"] not in _DTYPES:
            continue                                     # e.g. string tensors: not ours
        sid = e["shard_id"]
        if sid not in shards:
            p = "%s.data-%05d-of-%05d" % (prefix, sid, n_shards)
            shards[sid] = np.memmap(p, dtype=np.uint8, mode="r") if os.path.getsize(p) else np.zeros(0, np.uint8)
        raw = bytes(shards[sid][e["offset"]:e["offset"] + e["size"]])
        dt = np.dtype(_DTYPES[e["dtype"]])
        n = int(np.prod(e["shape"])) if e["shape"] else 1
        if len(raw) != e["size"] or n * dt.itemsize != e["size"]:
            raise ValueError("tensor %r: size %d does not match shape %s" % (name, e["size"], e["shape"]))
        if verify and e["crc32c"] is not None and crc32c(raw) != unmask_crc(e["crc32c"]):
            raise ValueError("tensor %r: data checksum mismatch" % name)
        out[name] = np.frombuffer(raw, dtype=dt.newbyteorder("<")).astype(dt).reshape(e["shape"])
    return out


def write_bundle(prefix: str, tensors: Dict[str, np.ndarray], block_size: int = 4096) -> None:
    """Write ``tensors`` as a single-shard V2 bundle (the layout Saver produces with
    one shard); also drops an empty ``.meta`` so the trio of files of
    demo_pipeline.py:50-54 exists."""
    names = sorted(tensors.keys(), key=lambda s: s.encode("utf-8"))
    data = bytearray()
    items: List[Tuple[bytes, bytes]] = []
    header = b"\x08\x01" + b"\x1a\x02\x08\x01"          # num_shards=1, version{producer=1}
    items.append((b"", header))
    for name in names:
        a = np.asarray(tensors[name], order="C")
        if a.dtype not in _DTYPE_CODES:
            raise TypeError("unsupported dtype %s for %r" % (a.dtype, name))
        raw = a.astype(a.dtype.newbyteorder("<")).tobytes()
        items.append((name.encode("utf-8"),
                      _encode_entry(_DTYPE_CODES[a.dtype], a.shape, 0, len(data), len(raw), mask_crc(crc32c(raw)))))
        data += raw
    out = bytearray()
    index_items: List[Tuple[bytes, bytes]] = []
    cur: List[Tuple[bytes, bytes]] = []
    cur_bytes = 0

    def flush():
        nonlocal cur, cur_bytes
        if not cur:
            return
        blk = _build_block(cur)
        index_items.append((cur[-1][0], _put_varint(len(out)) + _put_varint(len(blk))))
        out.extend(_block_with_trailer(blk))
        cur, cur_bytes = [], 0

    for kv in items:
        cur.append(kv)
        cur_bytes += len(kv[0]) + len(kv[1]) + 3
        if cur_bytes >= block_size:
            flush()
    flush()
    meta = _build_block([])
    meta_handle = _put_varint(len(out)) + _put_varint(len(meta))
    out.extend(_block_with_trailer(meta))
    idx = _build_block(index_items, restart_interval=1)
    idx_handle = _put_varint(len(out)) + _put_varint(len(idx))
    out.extend(_block_with_trailer(idx))
    footer = meta_handle + idx_handle
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    out.extend(footer)
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))
    if not os.path.exists(prefix + ".meta"):
        with open(prefix + ".meta", "wb") as f:
            f.write(b"")
