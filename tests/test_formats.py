"""Checkpoint bundle, NIfTI and weight-contract tests (CPU)."""
import os

import numpy as np
import pytest

from ukbb_cardiac_b200 import nifti, synth, tf_bundle
from ukbb_cardiac_b200 import weights as W


def test_crc32c_known_answers():
    assert tf_bundle.crc32c(b"123456789") == 0xE3069283
    assert tf_bundle.crc32c(b"") == 0
    assert tf_bundle.crc32c(bytes(32)) == 0x8A9136AA                     # rfc3720 B.4
    assert tf_bundle.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
    assert tf_bundle.unmask_crc(tf_bundle.mask_crc(0xDEADBEEF)) == 0xDEADBEEF


def test_native_crc32c_matches_python():
    from ukbb_cardiac_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for n in (0, 1, 7, 8, 9, 255, 256, 1000, 4097):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        tf_bundle.set_native_crc32c(None)
        py = tf_bundle.crc32c(data)
        assert int(lib.ukbb_crc32c(data, len(data))) == py
    _lib._lib = None
    _lib.load()


def test_bundle_roundtrip_and_contract(tmp_path):
    t = synth.with_optimizer_slots(synth.make_weights(0, 4))
    t["global_step"] = np.asarray(50000, dtype=np.int64)
    prefix = str(tmp_path / "FCN_sa")
    tf_bundle.write_bundle(prefix, t)
    for ext in (".index", ".data-00000-of-00001", ".meta"):
        assert os.path.exists(prefix + ext)
    back = tf_bundle.read_bundle(prefix)
    assert set(back) == set(t)
    for k in t:
        np.testing.assert_array_equal(back[k], t[k])
        assert back[k].dtype == t[k].dtype and back[k].shape == np.asarray(t[k]).shape
    assert W.validate(back) == 4                      # Adam slots / global_step are ignored
    assert W.n_parameters(4) == 1989012


def test_bundle_detects_corruption(tmp_path):
    t = synth.make_weights(1, 2)
    prefix = str(tmp_path / "m")
    tf_bundle.write_bundle(prefix, t)
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[1234] ^= 0x40
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    with pytest.raises(ValueError, match="checksum"):
        tf_bundle.read_bundle(prefix)
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[-1] ^= 0xFF
    open(prefix + ".index", "wb").write(bytes(idx))
    with pytest.raises(ValueError, match="magic"):
        tf_bundle.read_bundle(prefix)


def test_validate_fails_loudly():
    t = synth.make_weights(0, 3)
    assert W.validate(t) == 3
    bad = dict(t); del bad["batch_normalization_7/gamma"]
    with pytest.raises(KeyError):
        W.validate(bad)
    bad = dict(t); bad["conv2d_5/kernel"] = np.zeros((3, 3, 64, 32), np.float32)
    with pytest.raises(ValueError):
        W.validate(bad)
    with pytest.raises(KeyError):
        W.validate({"foo": np.zeros(1, np.float32)})
    # n_class is inferred from the last kernel (seg4 model: 6 classes)
    assert W.validate(synth.make_weights(0, 6)) == 6


def test_layer_table_names():
    assert W.conv_name(0) == "conv2d" and W.conv_name(20) == "conv2d_20"
    assert W.bn_name(0) == "batch_normalization" and W.bn_name(19) == "batch_normalization_19"
    shapes = W.expected_shapes(4)
    assert shapes["conv2d/kernel"] == (3, 3, 1, 16)
    assert shapes["conv2d_13/kernel"] == (1, 1, 16, 32)
    assert shapes["conv2d_18/kernel"] == (1, 1, 160, 64)
    assert shapes["conv2d_20/bias"] == (4,)
    assert "batch_normalization_20/gamma" not in shapes


@pytest.mark.parametrize("dtype", [np.float32, np.int16, np.float64, np.uint8])
def test_nifti_roundtrip(tmp_path, dtype):
    rng = np.random.default_rng(2)
    data = np.asfortranarray((rng.random((7, 5, 3, 2)) * 100).astype(dtype))
    aff = np.array([[-1.8, 0, 0, 90.0], [0, 1.8, 0.1, -120.0], [0, -0.1, 10.0, 33.0], [0, 0, 0, 1]])
    img = nifti.Nifti1Image(data, aff)
    img.header["pixdim"][4] = 0.031
    p = str(tmp_path / "sa.nii.gz")
    nifti.save(img, p)
    back = nifti.load(p)
    np.testing.assert_array_equal(back.get_data(), data)
    assert back.get_data().dtype == dtype and back.get_data().flags.f_contiguous
    np.testing.assert_allclose(back.affine, aff, rtol=1e-6)
    assert back.header["pixdim"][4] == np.float32(0.031)
    np.testing.assert_allclose(back.header["pixdim"][1:4], [1.8, np.sqrt(1.8 ** 2 + 0.1 ** 2), np.sqrt(0.1 ** 2 + 100)], rtol=1e-6)


def test_nifti_label_volume_like_reference(tmp_path):
    """deploy_network.py:134-137: Nifti1Image(pred float64, nim.affine) + pixdim copy."""
    src = nifti.Nifti1Image(np.zeros((4, 4, 2, 3), np.float32, order="F"), np.diag([1.8, 1.8, 10.0, 1.0]))
    src.header["pixdim"][4] = 0.05
    pred = np.asfortranarray(np.random.default_rng(0).integers(0, 4, (4, 4, 2, 3)).astype(np.float64))
    out = nifti.Nifti1Image(pred, src.affine)
    out.header["pixdim"] = src.header["pixdim"]
    p = str(tmp_path / "seg_sa.nii.gz")
    nifti.save(out, p)
    back = nifti.load(p)
    assert back.get_data().dtype == np.float64
    np.testing.assert_array_equal(back.get_data(), pred)
    np.testing.assert_array_equal(back.header["pixdim"], src.header["pixdim"])
    np.testing.assert_allclose(back.affine, src.affine)
    # plain .nii and scl_slope handling
    src.header["scl_slope"], src.header["scl_inter"] = 2.0, 1.0
    nifti.save(src, str(tmp_path / "x.nii"))
    np.testing.assert_array_equal(nifti.load(str(tmp_path / "x.nii")).get_data(), np.ones((4, 4, 2, 3)))


def test_nifti_rejects_garbage(tmp_path):
    p = str(tmp_path / "bad.nii")
    open(p, "wb").write(b"\x00" * 400)
    with pytest.raises(ValueError):
        nifti.load(p)


def test_synth_stack_properties():
    v = synth.make_stack(3, (32, 48, 2, 3))
    assert v.shape == (32, 48, 2, 3) and v.dtype == np.float32 and v.flags.f_contiguous
    assert np.all(v == np.rint(v)) and v.min() >= 0 and v.max() <= 4095
    np.testing.assert_array_equal(v, synth.make_stack(3, (32, 48, 2, 3)))
    assert not np.array_equal(v, synth.make_stack(4, (32, 48, 2, 3)))


def test_parallel_gzip_is_a_plain_gz_stream(tmp_path):
    """Large volumes are deflated as a multi-member gzip stream by a thread pool; any gzip reader returns the same bytes."""
    import gzip
    import time
    rng = np.random.default_rng(0)
    vol = np.zeros((96, 104, 10, 50), dtype=np.float64, order="F")            # 40 MB of label-like data (mostly zeros)
    vol[30:60, 40:70] = rng.integers(0, 4, size=(30, 30, 10, 50))
    img = nifti.Nifti1Image(vol, np.diag([1.8, 1.8, 10.0, 1.0]))
    p = str(tmp_path / "seg.nii.gz")
    t0 = time.perf_counter()
    nifti.save(img, p)
    dt = time.perf_counter() - t0
    raw = open(p, "rb").read()
    assert raw[:2] == b"\x1f\x8b" and raw.count(b"\x1f\x8b\x08") >= 2     # several members
    plain = gzip.decompress(raw)                                               # stdlib reader: concatenation of the members
    assert len(plain) == 352 + vol.nbytes
    with gzip.open(p, "rb") as g:
        assert g.read() == plain
    back = nifti.load(p)
    np.testing.assert_array_equal(back.get_data(), vol)
    small = nifti.gzip_parallel(b"abc" * 100)                                  # small payload: one member
    assert gzip.decompress(small) == b"abc" * 100 and len(nifti._member_index(small)) == 1
    print("saved %d MB in %.2f s" % (vol.nbytes >> 20, dt))


def test_member_index_parallel_inflate_and_foreign_files(tmp_path):
    """SURVEY 8(f) rank 2, input side: files written by nifti.save carry a per-member index (gzip extra subfield) and are inflated
    member-parallel straight into caller-provided memory; files from any other writer (single- or multi-member) are read as a
    sequential stream; both give the same array, and a foreign reader (stdlib gzip) reads our files."""
    import gzip
    vol = synth.make_stack(2, (96, 104, 10, 12))                               # 4.8 MB of float32 -> written in 8 MB members? no: force it
    img = nifti.Nifti1Image(vol, np.diag([1.8, 1.8, 10.0, 1.0]))
    old_chunk, old_min = nifti._PAR_CHUNK, nifti._PAR_MIN
    nifti._PAR_CHUNK, nifti._PAR_MIN = 1 << 20, 1 << 20
    try:
        p = str(tmp_path / "sa.nii.gz")
        nifti.save(img, p)
    finally:
        nifti._PAR_CHUNK, nifti._PAR_MIN = old_chunk, old_min
    raw = open(p, "rb").read()
    idx = nifti._member_index(raw)
    assert idx is not None and len(idx) == -(-(352 + vol.nbytes) // (1 << 20))
    assert sum(m[3] for m in idx) == 352 + vol.nbytes
    assert gzip.decompress(raw)[352:] == vol.tobytes(order="F")
    arena = np.zeros(352 + vol.nbytes + 64, dtype=np.uint8)
    back = nifti.load(p, alloc=lambda n: arena)
    np.testing.assert_array_equal(back.get_data(), vol)
    assert np.shares_memory(back.get_data(), arena)                            # decoded in place: no copy between inflate and upload
    np.testing.assert_array_equal(nifti.load(p).get_data(), vol)
    plain = gzip.decompress(raw)
    for name, blob in (("single", gzip.compress(plain, 1)), ("multi", gzip.compress(plain[:5000], 6) + gzip.compress(plain[5000:], 1)),
                       ("raw", plain)):
        q = str(tmp_path / (name + (".nii" if name == "raw" else ".nii.gz")))
        open(q, "wb").write(blob)
        assert name == "raw" or nifti._member_index(blob) is None
        np.testing.assert_array_equal(nifti.load(q).get_data(), vol)
        np.testing.assert_array_equal(nifti.load(q, alloc=lambda n: arena).get_data(), vol)
    with pytest.raises(ValueError):
        open(p, "wb").write(raw[:len(raw) // 2 + 7])
        nifti.load(p)
    bad = bytearray(raw)
    bad[idx[1][0] + 5] ^= 0xff                                                  # corrupt the deflate stream of member 1
    open(p, "wb").write(bytes(bad))
    with pytest.raises(Exception):
        nifti.load(p)


def test_x2_plane_format_model():
    """The FP16 + E4M3 activation planes of the fp16x2 mode as the GPU tests build and decode them (tests/gpu_util.py, a restatement of
    tc_common.cuh split_pack4): hi = rn_fp16(x), lo plane per 16 channels = [e4m3((x - hi) 2^11) x 16 | e4m3(hi) x 16].  Decoding
    hi + lo8 2^-11 reproduces x to 2^-15 relative + 2^-21 absolute (a 4-bit correction of a 2^-11 residual), exact zeros stay zero, and the byte layout
    is the one the kernels read (the lo plane is as large as an FP16 plane)."""
    import torch
    from gpu_util import e4m3, x2_decode, x2_planes
    rng = np.random.default_rng(0)
    x = np.abs(rng.normal(0.0, 1.0, size=(2, 3, 5, 32))).astype(np.float32)
    x[0, 0, 0, :4] = 0.0
    x[1, 2, 4, 5] = 300.0                                         # inside the E4M3 range (448)
    hi, lo8, hi8, plane = x2_planes(x)
    assert plane.dtype == np.float16 and plane.shape == x.shape
    raw = np.ascontiguousarray(plane).view(np.uint8).reshape(x.shape[:-1] + (2, 32))
    assert (raw[..., :16] == e4m3((x - hi) * 2048.0)[0].reshape(x.shape[:-1] + (2, 16))).all()
    assert (raw[..., 16:] == e4m3(hi)[0].reshape(x.shape[:-1] + (2, 16))).all()
    dec = x2_decode(torch.stack([torch.from_numpy(hi).to(torch.float16), torch.from_numpy(plane)]))
    assert (dec[0, 0, 0, :4] == 0).all()
    assert np.abs(dec - x).max() <= 2.0 ** -15 * np.abs(x).max()
    # per element: 2^-4 of a residual <= 2^-11 |x|, or half an E4M3 subnormal step (2^-10) of the scaled residual for small values
    assert (np.abs(dec - x) <= 2.0 ** -15 * np.abs(x) + 2.0 ** -21).all()
    # E4M3 codes: round to nearest even, saturating at 448
    codes, vals = e4m3(np.array([0.0, 1.0, 1.0625, 1.1875, 448.0, 1000.0, -3.0], np.float32))
    assert vals.tolist() == [0.0, 1.0, 1.0, 1.25, 448.0, 448.0, -3.0]
