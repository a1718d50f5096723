"""GPU parity tests, FP32 exactness mode: CUDA path (through the C ABI) vs the CPU oracle.

Tolerances (BASELINE.json north_star): logits within 1e-4 relative error
(max|d| / max|ref|), label maps identical to the oracle's argmax-of-float32-softmax; a
mismatching pixel is accepted only if the float64 oracle shows it is a near-tie
(top-2 gap < 1e-5 * max|logit|), and at most 1e-5 of the pixels may be such near-ties.
"""
import os

import numpy as np
import pytest
import torch

from oracle import deploy_oracle as do
from oracle import fcn_oracle as fo
from ukbb_cardiac_b200 import synth
from ukbb_cardiac_b200.fcn import FCNEngine, pad16

from gpu_util import adjudicate_labels, from_device_labels, from_device_logits, to_device_layout

pytestmark = pytest.mark.gpu

LOGIT_RTOL = 1e-4


def _check_forward(eng, weights, img_nxyc, x_pre=0, y_pre=0, x=None, y=None):
    n, X2, Y2, _ = img_nxyc.shape
    x = X2 - x_pre if x is None else x
    y = Y2 - y_pre if y is None else y
    labels, logits, prob = eng.forward(to_device_layout(img_nxyc), x_pre, y_pre, x, y, want_logits=True, want_prob=True)
    torch.cuda.synchronize()
    lg = from_device_logits(logits)
    ref32 = fo.build_fcn(img_nxyc, weights, torch.float32)
    ref64 = fo.build_fcn(img_nxyc, weights, torch.float64)
    scale = np.abs(ref64).max()
    err = np.abs(lg - ref64).max() / scale
    err32 = np.abs(ref32 - ref64).max() / scale
    assert err <= LOGIT_RTOL, "logits rel err %g (oracle f32 itself: %g)" % (err, err32)
    # prob = softmax(logits)
    p_ref, pred_ref = fo.softmax_argmax(lg)
    np.testing.assert_allclose(from_device_logits(prob), p_ref, rtol=2e-6, atol=1e-7)
    lab = from_device_labels(labels)
    crop = (slice(None), slice(x_pre, x_pre + x), slice(y_pre, y_pre + y))
    assert lab.shape == (n, x, y)
    # (1) the device argmax agrees with argmax-of-softmax on its own logits
    np.testing.assert_array_equal(lab, pred_ref[crop])
    # (2) and with the oracle's labels, up to float64-adjudicated near-ties
    _, pred32 = fo.softmax_argmax(ref32)
    nbad, _ = adjudicate_labels(lab, ref64[crop], 1e-5 * scale)
    assert nbad <= max(1, int(1e-5 * lab.size)), "%d near-tie pixels" % nbad
    nbad32 = int((lab != pred32[crop]).sum())
    assert nbad32 <= max(2, int(2e-5 * lab.size))
    return err


@pytest.mark.parametrize("n_class", [4, 2, 3, 6])
def test_forward_small(n_class):
    w = synth.make_weights(0, n_class)
    rng = np.random.default_rng(n_class)
    with FCNEngine(w, mode="fp32") as eng:
        assert eng.n_class == n_class
        for shape in [(2, 32, 48, 1), (1, 16, 16, 1), (3, 64, 32, 1)]:
            _check_forward(eng, w, rng.random(shape).astype(np.float32))


def test_forward_sa_size_and_crop():
    w = synth.make_weights(0, 4)
    vol = synth.make_stack(0)
    img = do.rescale_intensity(vol.copy(order="F"), (1, 99))
    fr = np.transpose(img[:, :, 3:5, 7], (2, 0, 1)).astype(np.float32)[..., None]       # (2,192,208,1)
    with FCNEngine(w, mode="fp32") as eng:
        _check_forward(eng, w, fr)
        # crop branch: pretend the real image was 171 x 210 inside the 192 x 208... use LA-like padding
        la = np.zeros((1, 176, 224, 1), np.float32)
        la[0, 2:173, 7:217, 0] = np.random.default_rng(0).random((171, 210))
        _check_forward(eng, w, la, x_pre=2, y_pre=7, x=171, y=210)


def test_class_counts_and_determinism():
    w = synth.make_weights(0, 4)
    img = np.random.default_rng(9).random((5, 48, 64, 1)).astype(np.float32)
    with FCNEngine(w, mode="fp32") as eng:
        dev = to_device_layout(img)
        l1, _, _ = eng.forward(dev, 1, 2, 40, 60)
        cnt = eng.class_counts(5).cpu().numpy()
        l2, _, _ = eng.forward(dev, 1, 2, 40, 60)
        torch.cuda.synchronize()
        assert torch.equal(l1, l2)
        lab = l1.cpu().numpy()
        assert lab.shape == (5, 60, 40)
        for i in range(5):
            np.testing.assert_array_equal(cnt[i], np.bincount(lab[i].ravel(), minlength=4))
        # batch invariance: slice i alone gives the same labels
        l3, _, _ = eng.forward(dev[2:3].contiguous(), 1, 2, 40, 60)
        assert torch.equal(l3[0], l1[2])


def test_preprocess_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "rescale_reference.npz"))
    names = sorted({k.split("/")[0] for k in g.files})
    w = synth.make_weights(0, 2)
    with FCNEngine(w, mode="fp32") as eng:
        for nm in names:
            arr = g[nm + "/input"]
            shp = arr.shape
            x, y = shp[0], shp[1]
            n_slices = int(np.prod(shp[2:])) if arr.ndim > 2 else 1
            vol = torch.from_numpy(np.ascontiguousarray(arr.reshape(-1, order="F"))).cuda()
            out, vlvh, (x_pre, y_pre) = eng.preprocess(vol, n_slices, x, y, clip_in_place=True)
            torch.cuda.synchronize()
            np.testing.assert_array_equal(vlvh.cpu().numpy(), g[nm + "/vl_vh"], err_msg=nm)      # bit-exact float64
            np.testing.assert_array_equal(vol.cpu().numpy(), g[nm + "/clipped"].reshape(-1, order="F"), err_msg=nm)
            x2, _ = pad16(x); y2, _ = pad16(y)
            o = out.cpu().numpy()
            assert o.shape == (n_slices, y2, x2)
            ref = g[nm + "/rescaled_f32"].reshape(x, y, n_slices, order="F")               # (X, Y, N)
            exp = np.zeros((n_slices, y2, x2), np.float32)
            exp[:, y_pre:y_pre + y, x_pre:x_pre + x] = np.transpose(ref, (2, 1, 0))
            np.testing.assert_array_equal(o, exp, err_msg=nm)                                # bit-exact float32


@pytest.mark.parametrize("seq,shape,n_class", [("sa", (40, 52, 3, 4), 4), ("la_2ch", (50, 43, 1, 5), 2)])
def test_segment_volume_matches_deploy_oracle(seq, shape, n_class):
    w = synth.make_weights(0, n_class)
    vol = synth.make_stack(5, shape)
    pred_ref, clipped = do.deploy_sequence(vol.copy(order="F"), do.make_runner(w))
    with FCNEngine(w, mode="fp32") as eng:
        lab, (vl, vh), counts = eng.segment_volume(vol)
    assert lab.shape == shape and lab.dtype == np.uint8
    assert vl == do.percentile_linear(vol, 1) and vh == do.percentile_linear(vol, 99)
    mism = int((lab != pred_ref).sum())
    assert mism <= max(2, int(2e-5 * lab.size)), "%d / %d voxels differ" % (mism, lab.size)
    # counts[t, z, k] = sum(pred == k) -> ES frame rule of deploy_network.py:125-131
    X, Y, Z, T = shape
    for t in range(T):
        for z in range(Z):
            np.testing.assert_array_equal(counts[t, z], np.bincount(lab[:, :, z, t].ravel(), minlength=n_class))
    if mism == 0:
        cnt1 = counts[:, :, 1].sum(axis=1)
        es = int(np.argmin(cnt1)) if seq == "sa" else int(np.argmax(cnt1))
        assert es == do.es_frame(pred_ref, seq)


def test_session_run_contract():
    w = synth.make_weights(0, 3)
    img = np.random.default_rng(2).random((2, 32, 48, 1)).astype(np.float32)
    with FCNEngine(w, mode="fp32") as eng:
        prob, pred = eng.run(["prob:0", "pred:0"], feed_dict={"image:0": img, "training:0": False})
        assert prob.shape == (2, 32, 48, 3) and prob.dtype == np.float32
        assert pred.shape == (2, 32, 48) and pred.dtype == np.int32
        p_ref, pred_ref = fo.session_run(img, w)
        np.testing.assert_allclose(prob, p_ref, rtol=1e-4, atol=1e-6)
        assert (pred != pred_ref).sum() <= 2
        with pytest.raises(ValueError):
            eng.run(["pred:0"], {"image:0": img[:, :30]})
        with pytest.raises(KeyError):
            eng.run(["nope:0"], {"image:0": img})
        with pytest.raises(NotImplementedError):
            eng.run(["pred:0"], {"image:0": img, "training:0": True})


def test_full_size_properties():
    """Size-independent properties on one full SA subject (192x208x10x50, 500 slices):
    determinism, batch invariance against an independent call on a subset, counts ==
    histogram of the label volume."""
    w = synth.make_weights(0, 4)
    vol = synth.make_stack(1)
    with FCNEngine(w, mode="fp32") as eng:
        lab, (vl, vh), counts = eng.segment_volume(vol)
        lab2, _, _ = eng.segment_volume(vol)
        np.testing.assert_array_equal(lab, lab2)
        assert counts.sum() == vol.size
        np.testing.assert_array_equal(counts.sum(axis=(0, 1)), np.bincount(lab.ravel(), minlength=4))
        assert (np.bincount(lab.ravel(), minlength=4) > 0.01 * lab.size).all()       # every class populated
        # subset: frames 10..11 pushed through preprocess+forward by hand with the same (vl, vh)
        img = do.rescale_intensity(vol.copy(order="F"), (1, 99))
        fr = np.transpose(img[:, :, :, 10], (2, 0, 1)).astype(np.float32)[..., None]
        l, _, _ = eng.forward(to_device_layout(fr))
        np.testing.assert_array_equal(from_device_labels(l), np.transpose(lab[:, :, :, 10], (2, 0, 1)))
    assert vl == do.percentile_linear(vol, 1) and vh == do.percentile_linear(vol, 99)


def test_error_paths():
    w = synth.make_weights(0, 4)
    with FCNEngine(w, mode="fp32") as eng:
        from ukbb_cardiac_b200._lib import UkbbError
        with pytest.raises(UkbbError):
            eng.forward(torch.zeros((1, 30, 32), device="cuda"))            # not a multiple of 16
        with pytest.raises(UkbbError):
            eng.forward(torch.zeros((1, 32, 32), device="cuda"), 4, 4, 32, 32)  # crop outside
    bad = dict(w); bad["conv2d_3/kernel"] = np.zeros((3, 3, 32, 16), np.float32)
    with pytest.raises(ValueError):
        FCNEngine(bad, mode="fp32")


def _rescale_oracle(arr):
    """reference arithmetic (image_utils.py:70-77 under numpy 2) via the oracle restatement; returns (vl, vh, clipped, out)."""
    a = arr.copy(order="F")
    out = do.rescale_intensity(a, (1, 99))
    return do.percentile_linear(arr, 1), do.percentile_linear(arr, 99), a, np.asarray(out, dtype=np.float32)


@pytest.mark.parametrize("shape", [(37, 29, 3, 5), (40, 28, 3, 5), (64, 48, 2, 3)])
@pytest.mark.parametrize("case", ["int12", "int16_high", "constant", "negative", "fraction", "minus_zero", "int_but_one_65536"])
def test_preprocess_integer_fast_path_and_fallback(monkeypatch, case, shape):
    """Integer-valued volumes take the one-pass counting select + lookup-table rescale; anything else (a negative voxel, a
    fractional one, -0.0, a level above 65535) takes the three-pass radix select on the same call.  Both are bit-exact against
    the reference arithmetic, and the fast path equals the generic path (ukbb_fcn_debug_flags bit 0) bit for bit."""
    # (37, 29): odd sizes, pad branches, n % 4 tail, scalar rescale; (40, 28) and (64, 48): rows of whole float4s (x_pre = 4 / 0),
    # the table-lookup rescale kernel
    rng = np.random.default_rng(11)
    a = np.floor(rng.gamma(2.0, 300.0, size=shape)).astype(np.float32)
    if case == "int12":
        a = np.minimum(a, 4095.0)
    elif case == "int16_high":
        a = np.minimum(a * 20.0, 65535.0)                     # most levels above the shared-memory histogram
        a.flat[7] = 65535.0
    elif case == "constant":
        a[:] = 123.0
    elif case == "negative":
        a.flat[100] = -3.0
    elif case == "fraction":
        a.flat[1234] += 0.5
    elif case == "minus_zero":
        a.flat[5] = -0.0
    elif case == "int_but_one_65536":
        a.flat[9] = 65536.0
    a = np.asfortranarray(a)
    x, y = shape[0], shape[1]
    n_slices = shape[2] * shape[3]
    w = synth.make_weights(0, 2)
    results = []
    for no_int in (False, True):
        with FCNEngine(w, mode="fp32") as eng:
            eng.debug_flags(1 if no_int else 0)
            vol = torch.from_numpy(np.ascontiguousarray(a.reshape(-1, order="F"))).cuda()
            out, vlvh, (x_pre, y_pre) = eng.preprocess(vol, n_slices, x, y, clip_in_place=True)
            torch.cuda.synchronize()
            results.append((out.cpu().numpy(), vlvh.cpu().numpy(), vol.cpu().numpy()))
    for k in range(3):
        np.testing.assert_array_equal(results[0][k], results[1][k])
    if case == "constant":
        return                                                # vh == vl: 0 / 0 in the reference too, nothing more to compare
    vl, vh, clipped, ref = _rescale_oracle(a)
    o, vv, vol_after = results[0]
    np.testing.assert_array_equal(vv, np.array([vl, vh]))
    np.testing.assert_array_equal(vol_after, clipped.reshape(-1, order="F"))
    x2, _ = pad16(x); y2, _ = pad16(y)
    exp = np.zeros((n_slices, y2, x2), np.float32)
    exp[:, y_pre:y_pre + y, x_pre:x_pre + x] = np.transpose(ref.reshape(x, y, n_slices, order="F"), (2, 1, 0))
    np.testing.assert_array_equal(o, exp)
