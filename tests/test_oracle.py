"""CPU tests of the oracle: known-answer tests authored from the reference's
definitions (the reference ships no tests) and the golden
vectors produced by the reference's own image_utils (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import deploy_oracle as do
from oracle import fcn_oracle as fo
from ukbb_cardiac_b200 import synth


# ---------------------------------------------------------------- network.py:117-135
def test_linear_kernels():
    np.testing.assert_array_equal(fo.linear_1d(3), np.array([.5, 1, .5], np.float32))
    np.testing.assert_array_equal(fo.linear_1d(7), np.array([.25, .5, .75, 1, .75, .5, .25], np.float32))
    with pytest.raises(NotImplementedError):
        fo.linear_1d(4)
    w = fo.linear_2d(3)
    np.testing.assert_array_equal(w, np.outer([.5, 1, .5], [.5, 1, .5]).astype(np.float32))


# ---------------------------------------------------------------- network.py:138-167
@pytest.mark.parametrize("f", [2, 4, 8, 16])
def test_upsample_matches_closed_form(f):
    rng = np.random.default_rng(f)
    x = rng.random((1, 3, 5, 4))            # N C H W
    up = fo.transpose_upsample2d(torch.from_numpy(x), f).numpy()
    assert up.shape == (1, 3, 5 * f, 4 * f)
    ref = fo.upsample_closed_form_1d(x, f)                       # along W
    ref = np.swapaxes(fo.upsample_closed_form_1d(np.swapaxes(ref, 2, 3), f), 2, 3)   # along H
    np.testing.assert_allclose(up, ref, rtol=0, atol=1e-12)


def test_upsample_impulse_and_borders():
    # f=2: up[0] = .5*x0, up[1] = x0, up[2] = .5*(x0+x1)  
    x = torch.tensor([[[[1.0, 3.0]]]], dtype=torch.float64)     # 1x1x1x2
    up = fo.transpose_upsample2d(x, 2).numpy()[0, 0]
    # rows: H=1 -> 2 rows with vertical weights [.5, 1]
    np.testing.assert_allclose(up[1], [0.5, 1.0, 2.0, 3.0])
    np.testing.assert_allclose(up[0], [0.25, 0.5, 1.0, 1.5])
    ones = torch.ones(1, 1, 4, 4, dtype=torch.float64)
    for f in (2, 4, 8, 16):
        u = fo.transpose_upsample2d(ones, f).numpy()[0, 0]
        c = u[f:-f, f:-f]
        np.testing.assert_allclose(c, 1.0)                     # partition of unity in the interior
        assert u[0, 0] < 1.0                                     # tapered border


# ---------------------------------------------------------------- TF SAME semantics
def test_same_pad_arithmetic():
    assert fo.same_pad(192, 3, 1) == (192, 1, 1)
    assert fo.same_pad(192, 3, 2) == (96, 0, 1)      # even input, stride 2: pad AFTER only
    assert fo.same_pad(13, 3, 2) == (7, 1, 1)
    assert fo.same_pad(26, 3, 2) == (13, 0, 1)
    assert fo.same_pad(5, 1, 1) == (5, 0, 0)


def test_stride2_same_asymmetry():
    w = np.ones((3, 3, 1, 1), np.float32)
    x = np.zeros((1, 6, 6, 1), np.float32)
    x[0, 0, 0, 0] = 1
    y = fo.conv2d_same_numpy(x, w, 2)[0, :, :, 0]
    exp = np.zeros((3, 3)); exp[0, 0] = 1                       # out[i] = sum_k in[2i+k]
    np.testing.assert_array_equal(y, exp)
    x[:] = 0; x[0, 5, 5, 0] = 1
    y = fo.conv2d_same_numpy(x, w, 2)[0, :, :, 0]
    exp = np.zeros((3, 3)); exp[2, 2] = 1                       # (5,5) = 2*2+1 only reaches out (2,2)
    np.testing.assert_array_equal(y, exp)
    x[:] = 0; x[0, 2, 2, 0] = 1                                 # in[2] feeds out[0] (k=2) and out[1] (k=0)
    y = fo.conv2d_same_numpy(x, w, 2)[0, :, :, 0]
    exp = np.zeros((3, 3)); exp[:2, :2] = 1
    np.testing.assert_array_equal(y, exp)


@pytest.mark.parametrize("stride,k,cin,cout,h,w", [(1, 3, 3, 5, 7, 9), (2, 3, 4, 6, 8, 10), (1, 1, 8, 4, 5, 6), (2, 3, 2, 3, 13, 9)])
def test_torch_conv_matches_numpy_taploop(stride, k, cin, cout, h, w):
    rng = np.random.default_rng(5)
    x = rng.normal(size=(2, h, w, cin))
    wt = rng.normal(size=(k, k, cin, cout))
    a = fo.conv2d_same(torch.from_numpy(np.transpose(x, (0, 3, 1, 2))).double(), wt, stride)
    a = a.permute(0, 2, 3, 1).numpy()
    b = fo.conv2d_same_numpy(x, wt, stride)
    np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-12)


def test_kernel_axes_are_hw():
    # an asymmetric kernel: W[kh=0, kw=2] = 1 -> out[h, w] = in[h-1, w+1]
    wt = np.zeros((3, 3, 1, 1)); wt[0, 2, 0, 0] = 1
    x = np.arange(20, dtype=np.float64).reshape(1, 4, 5, 1)
    y = fo.conv2d_same(torch.from_numpy(np.transpose(x, (0, 3, 1, 2))), wt, 1).numpy()[0, 0]
    assert y[1, 0] == x[0, 0, 1, 0] and y[3, 3] == x[0, 2, 4, 0] and y[0, 0] == 0 and y[1, 4] == 0


def test_bn_eps_and_formula():
    x = torch.tensor([[[[2.0]], [[-1.0]]]], dtype=torch.float64)       # N=1,C=2
    y = fo.bn_relu(x, np.array([2.0, 1.0]), np.array([0.5, 0.1]), np.array([1.0, 0.0]), np.array([0.999, 3.999])).numpy()
    np.testing.assert_allclose(y[0, :, 0, 0], [(2 - 1) * 2 / np.sqrt(1.0) + 0.5, 0.0], rtol=1e-12)   # 0.999+1e-3 = 1
    assert fo.BN_EPS == 1e-3


def test_softmax_tie_lowest_index():
    lg = np.array([[1.0, 3.0, 3.0, 0.0], [2.0, 2.0, 2.0, 2.0], [0.0, 1.0, 2.0, 5.0]], np.float32)
    prob, pred = fo.softmax_argmax(lg)
    np.testing.assert_array_equal(pred, [1, 0, 3])
    assert pred.dtype == np.int32 and prob.dtype == np.float32
    np.testing.assert_allclose(prob.sum(-1), 1.0, rtol=1e-6)


def test_layer_table_and_flops():
    tab = fo.layer_table(4)
    assert len(tab) == 21
    assert [t[2] for t in tab[:13]] == [16, 16, 32, 32, 64, 64, 64, 128, 128, 128, 256, 256, 256]
    assert [t[4] for t in tab[:13]] == [1, 1, 2, 1, 2, 1, 1, 2, 1, 1, 2, 1, 1]
    assert tab[18][1:3] == (160, 64) and tab[20][1:3] == (64, 4)
    # SURVEY 8d / BASELINE.md section 2
    assert abs(fo.flops_per_slice(192, 208, 4) / 1e9 - 3.1374) < 1e-4
    assert abs(fo.flops_per_slice(224, 176, 2) / 1e9 - 3.0871) < 1e-4
    assert abs(fo.flops_per_slice(224, 176, 3) / 1e9 - 3.0921) < 1e-4
    n_par = sum(int(np.prod(v.shape)) for v in synth.make_weights(0, 4).values())
    assert n_par == 1989012


def test_oracle_selfpin(golden_dir):
    g = np.load(os.path.join(golden_dir, "oracle_selfpin.npz"))
    for nc in (2, 3, 4, 6):
        lg = fo.build_fcn(g["image"], synth.make_weights(0, nc), torch.float64)
        np.testing.assert_allclose(lg, g["logits%d" % nc], rtol=1e-9, atol=1e-9)
    lg32 = fo.build_fcn(g["image"], synth.make_weights(0, 4), torch.float32)
    assert np.abs(lg32 - g["logits4"]).max() / np.abs(g["logits4"]).max() < 1e-5


# ---------------------------------------------------------------- deploy_network.py / image_utils.py
def test_pad_arithmetic():
    assert do.pad16(192) == (192, 0, 0)
    assert do.pad16(210) == (224, 7, 7)
    assert do.pad16(171) == (176, 2, 3)


def test_percentile_closed_form():
    a = np.arange(1000, dtype=np.float32)
    rng = np.random.default_rng(0)
    rng.shuffle(a)
    assert do.percentile_linear(a, 1) == pytest.approx(9.99, abs=1e-9)
    assert do.percentile_linear(a, 99) == pytest.approx(989.01, abs=1e-9)
    for q in (0, 1, 37.5, 50, 99, 100):
        # array-valued q, as in the reference call np.percentile(image, (1, 99)): float64 lerp
        assert do.percentile_linear(a, q) == float(np.percentile(a, (q, q))[0])


def test_rescale_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "rescale_reference.npz"))
    names = sorted({k.split("/")[0] for k in g.files})
    assert len(names) >= 6
    for nm in names:
        a = np.asfortranarray(g[nm + "/input"].copy())
        res = do.rescale_intensity(a, (1, 99))
        vl, vh = g[nm + "/vl_vh"]
        assert do.percentile_linear(g[nm + "/input"], 1) == vl, nm
        assert do.percentile_linear(g[nm + "/input"], 99) == vh, nm
        np.testing.assert_array_equal(a, g[nm + "/clipped"], err_msg=nm)          # in-place side effect
        np.testing.assert_array_equal(res, g[nm + "/rescaled_f64"], err_msg=nm)
        assert res.dtype == np.float64


def test_dice_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "dice_reference.npz"))
    for k in range(4):
        assert fo.categorical_dice(g["a"], g["b"], k) == pytest.approx(float(g["dice"][k]), rel=1e-6)


def test_es_rule():
    pred = np.zeros((4, 4, 1, 3))
    pred[:2, :, 0, 0] = 1      # 8 voxels
    pred[:1, :, 0, 1] = 1      # 4 voxels
    pred[:3, :, 0, 2] = 1      # 12 voxels
    assert do.es_frame(pred, "sa") == 1
    assert do.es_frame(pred, "la_4ch", seg4=True) == 1
    assert do.es_frame(pred, "la_2ch") == 2
    assert do.es_frame(pred, "la_4ch") == 2


def test_deploy_sequence_shapes_and_crop():
    calls = []

    def run(fr):
        calls.append(fr.shape)
        assert fr.dtype == np.float32
        # label = 1 where the padded input is exactly 0 (padding or p1-clipped), else 2
        return np.where(fr[..., 0] == 0, 1, 2).astype(np.int32)

    img = np.asfortranarray(np.random.default_rng(1).integers(0, 1000, size=(21, 18, 2, 3)).astype(np.float32))
    pred, clipped = do.deploy_sequence(img, run)
    assert calls == [(2, 32, 32, 1)] * 3
    assert pred.shape == (21, 18, 2, 3) and pred.dtype == np.float64
    assert clipped is img                                       # the reference clips its input in place
    assert set(np.unique(pred)) <= {1.0, 2.0}
