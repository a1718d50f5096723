"""Minimal NIfTI-1 single-file (.nii / .nii.gz) reader and writer.

The reference does all image I/O through nibabel (``common/deploy_network.py:
80-82,134-151``: ``nib.load``, ``get_data``, ``Nifti1Image(pred, nim.affine)``,
``header['pixdim']`` copy, ``nib.save``).  nibabel is not available in the build
or GPU image, so this module provides exactly the subset the deploy path needs,
with nibabel's conventions: data in Fortran (X-fastest) order, best affine =
sform if sform_code > 0 else qform if qform_code > 0 else pixdim scaling, a new
image gets sform_code=2 / qform_code=0, scl_slope/inter applied on read when valid.
"""
from __future__ import annotations

import gzip
import zlib
import threading
import io
import os
from typing import Optional, Tuple

import numpy as np

_HDR = np.dtype([
    ("sizeof_hdr", "<i4"), ("data_type", "S10"), ("db_name", "S18"), ("extents", "<i4"),
    ("session_error", "<i2"), ("regular", "S1"), ("dim_info", "u1"),
    ("dim", "<i2", (8,)), ("intent_p1", "<f4"), ("intent_p2", "<f4"), ("intent_p3", "<f4"),
    ("intent_code", "<i2"), ("datatype", "<i2"), ("bitpix", "<i2"), ("slice_start", "<i2"),
    ("pixdim", "<f4", (8,)), ("vox_offset", "<f4"), ("scl_slope", "<f4"), ("scl_inter", "<f4"),
    ("slice_end", "<i2"), ("slice_code", "u1"), ("xyzt_units", "u1"),
    ("cal_max", "<f4"), ("cal_min", "<f4"), ("slice_duration", "<f4"), ("toffset", "<f4"),
    ("glmax", "<i4"), ("glmin", "<i4"), ("descrip", "S80"), ("aux_file", "S24"),
    ("qform_code", "<i2"), ("sform_code", "<i2"),
    ("quatern_b", "<f4"), ("quatern_c", "<f4"), ("quatern_d", "<f4"),
    ("qoffset_x", "<f4"), ("qoffset_y", "<f4"), ("qoffset_z", "<f4"),
    ("srow_x", "<f4", (4,)), ("srow_y", "<f4", (4,)), ("srow_z", "<f4", (4,)),
    ("intent_name", "S16"), ("magic", "S4"),
])
assert _HDR.itemsize == 348

_CODE2DT = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64,
            256: np.int8, 512: np.uint16, 768: np.uint32, 1024: np.int64, 1280: np.uint64}
_DT2CODE = {np.dtype(v): k for k, v in _CODE2DT.items()}


def _quat_to_affine(h) -> np.ndarray:
    b, c, d = float(h["quatern_b"]), float(h["quatern_c"]), float(h["quatern_d"])
    a2 = 1.0 - (b * b + c * c + d * d)
    a = np.sqrt(a2) if a2 > 1e-7 else 0.0
    if a2 <= 1e-7:
        nrm = 1.0 / np.sqrt(b * b + c * c + d * d)
        b, c, d = b * nrm, c * nrm, d * nrm
    R = np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                  [2 * (b * c + a * d), a * a + c * c - b * b - d * d, 2 * (c * d - a * b)],
                  [2 * (b * d - a * c), 2 * (c * d + a * b), a * a + d * d - b * b - c * c]])
    pd = np.asarray(h["pixdim"], dtype=np.float64)
    qfac = -1.0 if pd[0] < 0 else 1.0
    zooms = np.where(pd[1:4] > 0, pd[1:4], 1.0)
    aff = np.eye(4)
    aff[:3, :3] = R * (zooms * np.array([1.0, 1.0, qfac]))[None, :]
    aff[:3, 3] = [h["qoffset_x"], h["qoffset_y"], h["qoffset_z"]]
    return aff


def _affine_to_quat(aff: np.ndarray) -> Tuple[float, float, float, float, np.ndarray]:
    """nifti1 mat44_to_quatern (orthonormalisation skipped: rotation taken from the
    column-normalised matrix).  Returns (b, c, d, qfac, zooms)."""
    M = np.asarray(aff, dtype=np.float64)[:3, :3]
    zooms = np.sqrt((M * M).sum(axis=0))
    zooms = np.where(zooms > 0, zooms, 1.0)
    R = M / zooms[None, :]
    qfac = 1.0
    if np.linalg.det(R) < 0:
        R[:, 2] = -R[:, 2]
        qfac = -1.0
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    a = tr + 1.0
    if a > 0.5:
        a = 0.5 * np.sqrt(a)
        b = 0.25 * (R[2, 1] - R[1, 2]) / a
        c = 0.25 * (R[0, 2] - R[2, 0]) / a
        d = 0.25 * (R[1, 0] - R[0, 1]) / a
    else:
        xd, yd, zd = 1 + R[0, 0] - (R[1, 1] + R[2, 2]), 1 + R[1, 1] - (R[0, 0] + R[2, 2]), 1 + R[2, 2] - (R[0, 0] + R[1, 1])
        if xd > 1.0:
            b = 0.5 * np.sqrt(xd); c = 0.25 * (R[0, 1] + R[1, 0]) / b; d = 0.25 * (R[0, 2] + R[2, 0]) / b; a = 0.25 * (R[2, 1] - R[1, 2]) / b
        elif yd > 1.0:
            c = 0.5 * np.sqrt(yd); b = 0.25 * (R[0, 1] + R[1, 0]) / c; d = 0.25 * (R[1, 2] + R[2, 1]) / c; a = 0.25 * (R[0, 2] - R[2, 0]) / c
        else:
            d = 0.5 * np.sqrt(zd); b = 0.25 * (R[0, 2] + R[2, 0]) / d; c = 0.25 * (R[1, 2] + R[2, 1]) / d; a = 0.25 * (R[1, 0] - R[0, 1]) / d
        if a < 0:
            b, c, d = -b, -c, -d
    return float(b), float(c), float(d), qfac, zooms


class Nifti1Image:
    """The slice of nibabel's ``Nifti1Image`` used by deploy_network.py."""

    def __init__(self, data: np.ndarray, affine: Optional[np.ndarray], header: Optional[np.ndarray] = None):
        self._data = data
        if header is None:
            header = np.zeros((), dtype=_HDR)
            header["sizeof_hdr"] = 348
            header["regular"] = b"r"
            header["pixdim"] = 1.0
            header["vox_offset"] = 352.0
            header["scl_slope"] = np.nan
            header["scl_inter"] = np.nan
            header["magic"] = b"n+1"
            if affine is None:
                affine = np.eye(4)
            affine = np.asarray(affine, dtype=np.float64)
            b, c, d, qfac, zooms = _affine_to_quat(affine)
            header["pixdim"][0] = qfac
            nd = min(data.ndim, 3)
            header["pixdim"][1:1 + nd] = zooms[:nd]
            header["quatern_b"], header["quatern_c"], header["quatern_d"] = b, c, d
            header["qoffset_x"], header["qoffset_y"], header["qoffset_z"] = affine[:3, 3]
            header["qform_code"] = 0
            header["sform_code"] = 2
            header["srow_x"], header["srow_y"], header["srow_z"] = affine[0], affine[1], affine[2]
        else:
            header = header.copy()
            if affine is not None:
                affine = np.asarray(affine, dtype=np.float64)
        self.header = header
        self._affine = affine if affine is not None else self._best_affine()
        self._sync_shape()

    def _sync_shape(self):
        d = self._data
        if d.ndim > 7:
            raise ValueError("NIfTI-1 supports at most 7 dimensions")
        dim = np.ones(8, dtype=np.int16)
        dim[0] = d.ndim
        dim[1:1 + d.ndim] = d.shape
        self.header["dim"] = dim
        dt = np.dtype(d.dtype).newbyteorder("=")
        if dt not in _DT2CODE:
            raise TypeError("dtype %s cannot be stored in NIfTI-1 by this writer" % d.dtype)
        self.header["datatype"] = _DT2CODE[dt]
        self.header["bitpix"] = dt.itemsize * 8

    def _best_affine(self) -> np.ndarray:
        h = self.header
        if int(h["sform_code"]) > 0:
            aff = np.eye(4)
            aff[0], aff[1], aff[2] = h["srow_x"], h["srow_y"], h["srow_z"]
            return aff
        if int(h["qform_code"]) > 0:
            return _quat_to_affine(h)
        pd = np.asarray(h["pixdim"], dtype=np.float64)
        aff = np.diag([pd[1] or 1.0, pd[2] or 1.0, pd[3] or 1.0, 1.0])
        return aff

    @property
    def affine(self) -> np.ndarray:
        return self._affine

    @property
    def shape(self):
        return self._data.shape

    def get_data(self) -> np.ndarray:
        """Fortran-ordered array; scl_slope/scl_inter applied when valid (nibabel
        semantics: slope 0 or NaN = no scaling)."""
        h = self.header
        slope, inter = float(h["scl_slope"]), float(h["scl_inter"])
        if np.isfinite(slope) and slope != 0.0 and not (slope == 1.0 and (inter == 0.0 or not np.isfinite(inter))):
            inter = inter if np.isfinite(inter) else 0.0
            return self._data * slope + inter
        return self._data

    get_fdata = get_data


def _parse_header(raw348: bytes, path: str):
    hdr = np.frombuffer(raw348, dtype=_HDR)[0].copy()
    if int(hdr["sizeof_hdr"]) != 348:
        raise ValueError("%s: not a little-endian NIfTI-1 file (sizeof_hdr=%d)" % (path, int(hdr["sizeof_hdr"])))
    if bytes(hdr["magic"])[:3] != b"n+1":
        raise ValueError("%s: only single-file NIfTI-1 (magic n+1) is supported" % path)
    nd = int(hdr["dim"][0])
    if not 1 <= nd <= 7:
        raise ValueError("%s: bad dim[0]=%d" % (path, nd))
    shape = tuple(int(x) for x in hdr["dim"][1:1 + nd])
    code = int(hdr["datatype"])
    if code not in _CODE2DT:
        raise TypeError("%s: unsupported NIfTI datatype code %d" % (path, code))
    dt = np.dtype(_CODE2DT[code]).newbyteorder("<")
    off = int(hdr["vox_offset"]) or 352
    return hdr, shape, dt, off


def _member_index(raw: bytes):
    """Members of a .gz file written by `gzip_parallel`: every member header carries an extra subfield ('U', 'K') with the
    compressed size of the member and the size of its payload (the BGZF idea), so the members can be located without
    inflating and inflated independently.  Returns [(deflate offset, deflate length, payload offset, payload length)] or None
    when the file was written by something else (then it is one sequential stream)."""
    pos, out_pos, idx = 0, 0, []
    n = len(raw)
    while pos < n:
        if n - pos < 28 or raw[pos:pos + 4] != b"\x1f\x8b\x08\x04":
            return None
        xlen = int.from_bytes(raw[pos + 10:pos + 12], "little")
        if xlen != 12 or raw[pos + 12:pos + 14] != b"UK" or raw[pos + 14:pos + 16] != b"\x08\x00":
            return None
        msize = int.from_bytes(raw[pos + 16:pos + 20], "little")
        psize = int.from_bytes(raw[pos + 20:pos + 24], "little")
        if msize < 32 or pos + msize > n:
            return None
        idx.append((pos + 24, msize - 32, out_pos, psize))
        pos += msize
        out_pos += psize
    return idx


def _inflate_member(args):
    raw, off, ln, dst, out_off, psize = args
    chunk = zlib.decompress(raw[off:off + ln], -15)
    if len(chunk) != psize:
        raise ValueError("corrupt gzip member: %d bytes inflated, index says %d" % (len(chunk), psize))
    if (zlib.crc32(chunk) & 0xffffffff) != int.from_bytes(raw[off + ln:off + ln + 4], "little"):
        raise ValueError("corrupt gzip member: CRC mismatch")
    dst[out_off:out_off + psize] = np.frombuffer(chunk, dtype=np.uint8)


def read_decompressed(path: str, alloc=None, threads: Optional[int] = None):
    """The decompressed bytes of a .nii / .nii.gz file as a uint8 array.  `alloc(nbytes)` may supply the destination (e.g. a view
    of pinned host memory).  Files written by `save` (indexed multi-member gzip) are inflated member-parallel on a thread pool
    (zlib releases the GIL) straight into the destination; any other .gz is one sequential stream and is inflated in pieces
    into the destination, so callers get their parallelism from decoding several files at once."""
    with open(path, "rb") as f:
        raw = f.read()
    alloc = alloc or (lambda nbytes: np.empty(nbytes, dtype=np.uint8))
    if raw[:2] != b"\x1f\x8b":
        dst = alloc(len(raw))
        dst[:len(raw)] = np.frombuffer(raw, dtype=np.uint8)
        return dst[:len(raw)]
    idx = _member_index(raw)
    if idx is not None:
        total = idx[-1][2] + idx[-1][3]
        dst = alloc(total)
        jobs = [(raw, off, ln, dst, out_off, psize) for off, ln, out_off, psize in idx]
        if len(jobs) == 1:
            _inflate_member(jobs[0])
        else:
            list(_pool().map(_inflate_member, jobs))
        return dst[:total]
    # foreign writer (nibabel, dcm2niix, ...): sequential stream(s), inflated in 8 MB pieces
    pieces, produced = [], 0
    d, data = zlib.decompressobj(31), memoryview(raw)
    while True:
        piece = d.decompress(data, 8 << 20)
        if piece:
            pieces.append(piece)
            produced += len(piece)
        if d.unconsumed_tail:
            data = d.unconsumed_tail
            continue
        if d.eof and d.unused_data:                                     # next member of a plain multi-member stream
            data, d = d.unused_data, zlib.decompressobj(31)
            continue
        if not d.eof and not piece:
            raise ValueError("%s: truncated gzip stream" % path)
        if d.eof:
            break
        data = b""
    dst = alloc(produced)
    o = 0
    for piece in pieces:
        dst[o:o + len(piece)] = np.frombuffer(piece, dtype=np.uint8)
        o += len(piece)
    return dst[:produced]


def load(path: str, alloc=None) -> Nifti1Image:
    """Read a NIfTI-1 file.  With `alloc`, the decompressed file lands in caller-provided memory and the image data is a VIEW of
    it whenever the on-disk dtype is native little-endian (no copy: UK Biobank volumes are float32, data/biobank_utils.py:314)."""
    raw = read_decompressed(path, alloc)
    if len(raw) < 352:
        raise ValueError("%s: too small for a NIfTI-1 file" % path)
    hdr, shape, dt, off = _parse_header(raw[:348].tobytes(), path)
    n = int(np.prod(shape))
    if len(raw) < off + n * dt.itemsize:
        raise ValueError("%s: truncated voxel data" % path)
    data = np.frombuffer(raw, dtype=dt, count=n, offset=off)
    if dt.newbyteorder("=") != dt or alloc is None:
        data = data.astype(dt.newbyteorder("="))
    return Nifti1Image(data.reshape(shape, order="F"), None, hdr)


_PAR_CHUNK = 8 << 20          # bytes of payload per gzip member when compressing in parallel
_PAR_MIN = 32 << 20           # smaller payloads are written as one member
_POOL = None
_POOL_LOCK = threading.Lock()


def _pool():
    """One process-wide thread pool for gzip members (deflate and inflate release the GIL); its workers never submit work
    themselves, so readers / writers of several files may share it."""
    global _POOL
    with _POOL_LOCK:
        if _POOL is None:
            from concurrent.futures import ThreadPoolExecutor
            _POOL = ThreadPoolExecutor(max_workers=max(2, os.cpu_count() or 2), thread_name_prefix="ukbb-gz")
        return _POOL


def _gzip_member(args) -> bytes:
    """One gzip member (RFC 1952) with an extra subfield 'UK' = (member size, payload size), see `_member_index`.
    args = (prefix bytes, source (bytes-like, or an array slice converted to `dtype` here, in the worker), dtype, level)."""
    prefix, src, dtype, level = args[:4]
    strategy = args[4] if len(args) > 4 else zlib.Z_DEFAULT_STRATEGY
    if dtype is not None:
        src = np.ascontiguousarray(src, dtype=dtype)
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    body = (c.compress(prefix) if prefix else b"") + c.compress(src) + c.flush()
    view = memoryview(src).cast("B")
    plen = len(prefix) + len(view)
    crc = zlib.crc32(view, zlib.crc32(prefix)) & 0xffffffff
    msize = 24 + len(body) + 8
    head = (b"\x1f\x8b\x08\x04" + b"\x00\x00\x00\x00" + (b"\x04" if level == 1 else b"\x00") + b"\xff" + b"\x0c\x00" + b"UK" + b"\x08\x00" +
            msize.to_bytes(4, "little") + plen.to_bytes(4, "little"))
    return head + body + crc.to_bytes(4, "little") + (plen & 0xffffffff).to_bytes(4, "little")


def gzip_parallel(payload, compresslevel: int = 1, threads: Optional[int] = None) -> bytes:
    """gzip `payload` as a sequence of independent members compressed by a thread pool (zlib releases the GIL).
    A multi-member stream is a valid .gz file (RFC 1952 section 2.2): `gzip.decompress`, `gzip.open`, zlib's gzread -- hence
    nibabel -- return the concatenation.  The reference writes a float64 label volume per sequence (160 MB for one SA
    subject, deploy_network.py:134-137), whose single-threaded deflate dominates the wall clock of the drop-in CLI."""
    view = memoryview(payload).cast("B")
    if len(view) < _PAR_MIN:
        return _gzip_member((b"", view, None, compresslevel))
    chunks = [(b"", view[o:o + _PAR_CHUNK], None, compresslevel) for o in range(0, len(view), _PAR_CHUNK)]
    return b"".join(_pool().map(_gzip_member, chunks))


def save(img: Nifti1Image, path: str, compresslevel: int = 1, dtype=None, label_data: bool = False) -> None:
    """Write a single-file NIfTI-1 image.  `dtype` stores the data converted to another type (e.g. uint8 label maps as the
    reference's float64 volumes, deploy_network.py:92,134-137): the conversion happens chunk by chunk inside the gzip workers,
    so the 8x larger array never exists in memory.  `label_data` selects zlib's run-length strategy, which deflates label
    volumes (long runs of identical values) about twice as fast as the default strategy and slightly smaller."""
    strategy = zlib.Z_RLE if label_data else zlib.Z_DEFAULT_STRATEGY
    img._sync_shape()
    h = img.header.copy()
    data = np.asarray(img._data)
    out_dt = np.dtype(dtype if dtype is not None else data.dtype).newbyteorder("=")
    if out_dt not in _DT2CODE:
        raise TypeError("dtype %s cannot be stored in NIfTI-1 by this writer" % out_dt)
    h["datatype"] = _DT2CODE[out_dt]
    h["bitpix"] = out_dt.itemsize * 8
    h["vox_offset"] = 352.0
    h["magic"] = b"n+1"
    head = h.tobytes() + b"\x00\x00\x00\x00"
    flat = data.reshape(-1, order="F")                 # a view for Fortran-ordered data (NIfTI order), else one copy
    le = out_dt.newbyteorder("<")
    tmp = path + ".tmp%d_%d" % (os.getpid(), threading.get_ident())
    with open(tmp, "wb") as f:
        if not path.endswith(".gz"):
            f.write(head)
            f.write(np.ascontiguousarray(flat, dtype=le).tobytes())
        elif flat.size * out_dt.itemsize < _PAR_MIN:
            f.write(_gzip_member((head, flat, le, compresslevel, strategy)))
        else:
            step = _PAR_CHUNK // out_dt.itemsize
            jobs = [(head if o == 0 else b"", flat[o:o + step], le, compresslevel, strategy) for o in range(0, flat.size, step)]
            for member in _pool().map(_gzip_member, jobs):
                f.write(member)
    os.replace(tmp, path)
