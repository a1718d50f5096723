"""Aortic UNet + BiConvLSTM path (BASELINE config C5).

CPU: known-answer tests pinning the TensorFlow semantics the restatement relies on (oracle/ao_oracle.py; TensorFlow is absent, so
this path is PARITY UNPINNED like the FCN network), and the equivalence of the restatement's feature-reuse shortcut with the literal
reference loop.  GPU: the device path (ukbb_ao_segment through the C ABI) against the oracle.
"""
import numpy as np
import pytest
import torch

from oracle import ao_oracle as ao
from oracle import fcn_oracle as fo
from ukbb_cardiac_b200 import aorta, synth


def test_conv2d_transpose_same_known_answers():
    """tf.layers.conv2d_transpose(k = 3, strides = 2, 'same'): out = 2 in; big[y] = sum_{i, k: 2 i + k = y} small[i] w[k] (pad_before 0)."""
    x = torch.zeros((1, 1, 3, 3), dtype=torch.float64)
    x[0, 0, 1, 1] = 1.0
    w = np.arange(1, 10, dtype=np.float64).reshape(3, 3, 1, 1)         # [kh, kw, out, in]
    y = ao.conv2d_transpose_same(x, w, 2)[0, 0].numpy()
    assert y.shape == (6, 6)
    expect = np.zeros((6, 6))
    expect[2:5, 2:5] = w[:, :, 0, 0]                                    # impulse at i = 1 lands at y = 2 .. 4
    np.testing.assert_array_equal(y, expect)
    x[:] = 0.0
    x[0, 0, 2, 2] = 1.0                                                 # last input pixel: taps k = 2 fall outside the 6-wide output (cropped)
    y = ao.conv2d_transpose_same(x, w, 2)[0, 0].numpy()
    expect = np.zeros((6, 6))
    expect[4:6, 4:6] = w[:2, :2, 0, 0]
    np.testing.assert_array_equal(y, expect)
    # channel layout [kh, kw, out, in]
    x2 = torch.ones((1, 2, 1, 1), dtype=torch.float64)
    w2 = np.zeros((3, 3, 3, 2))
    w2[0, 0, 1, 0] = 5.0
    w2[0, 0, 2, 1] = 7.0
    y2 = ao.conv2d_transpose_same(x2, w2, 2)[0].numpy()
    assert y2[1, 0, 0] == 5.0 and y2[2, 0, 0] == 7.0 and y2[0].sum() == 0.0


def test_conv_lstm_step_known_answers():
    """Conv2DLSTMCell: gate order (i, j, f, o), forget_bias = 1, c' = sig(f + 1) c + sig(i) tanh(j), h' = tanh(c') sig(o)."""
    nh = 2
    x = torch.zeros((1, 1, 1, 1), dtype=torch.float64)
    h = torch.zeros((1, nh, 1, 1), dtype=torch.float64)
    c = torch.full((1, nh, 1, 1), 0.5, dtype=torch.float64)
    k = np.zeros((3, 3, 1 + nh, 4 * nh))
    b = np.array([0.3, -0.2, 1.5, 0.7, -0.4, 0.1, 0.9, -1.1])          # i0 i1 j0 j1 f0 f1 o0 o1
    h2, c2 = ao.conv_lstm_step(x, h, c, k, b)
    sg = lambda v: 1.0 / (1.0 + np.exp(-v))
    for u in range(nh):
        i, j, f, o = b[u], b[nh + u], b[2 * nh + u], b[3 * nh + u]
        ce = sg(f + 1.0) * 0.5 + sg(i) * np.tanh(j)
        assert abs(float(c2[0, u, 0, 0]) - ce) < 1e-12
        assert abs(float(h2[0, u, 0, 0]) - np.tanh(ce) * sg(o)) < 1e-12
    # the kernel's input-channel axis is concat([x, h]): a weight on channel 1 sees h[0]
    k[1, 1, 1, 0] = 2.0
    h[0, 0, 0, 0] = 0.25
    _, c3 = ao.conv_lstm_step(x, h, c, k, b)
    assert abs(float(c3[0, 0, 0, 0]) - (sg(b[4] + 1.0) * 0.5 + sg(b[0] + 0.5) * np.tanh(b[2]))) < 1e-12


def test_window_weights_and_normalise():
    w = ao.window_weights(5, 0.1)
    assert len(w) == 9 and w[4] == 1.0 and np.allclose(w, w[::-1])
    assert abs(w[0] - (1 - 4 / 5.0) ** 0.1) < 1e-15
    assert np.array_equal(ao.window_weights(5, 0), np.ones(9))
    img = np.arange(100, dtype=np.float32).reshape(5, 5, 1, 4)
    z = ao.normalise_intensity(img, 10.0)
    roi = img[img >= np.percentile(img, 10.0)]
    np.testing.assert_allclose(z, (img - roi.mean()) / (roi.std() + 1e-6), rtol=1e-6)
    np.testing.assert_array_equal(z, aorta.normalise_intensity(img, 10.0))


def test_variable_table_matches_oracle_names():
    assert aorta.variable_table() == ao.weight_names()
    w = synth.make_ao_weights(0)
    assert aorta.validate(w) == (16, 16, 3)
    bad = dict(w)
    bad["UNet/conv2_up/conv2d_transpose/kernel"] = np.zeros((3, 3, 128, 64), np.float32)      # conv layout instead of [.., out, in]
    with pytest.raises(ValueError, match="conv2d_transpose"):
        aorta.validate(bad)
    del bad["LSTM/backward/conv_lstm_cell/biases"]
    with pytest.raises(ValueError):
        aorta.validate(bad)


def test_feature_reuse_equals_the_literal_window_loop():
    """deploy_network_ao.py:146-177 feeds every 9-frame window through the WHOLE graph; the restatement evaluates the UNet once per
    frame.  Same probabilities (the UNet acts per frame), checked on a tiny cine against the literal loop."""
    w = synth.make_ao_weights(1)
    img = synth.make_ao_stack(0, (24, 20, 1, 10))
    old = ao.IMAGE_SIZE
    ao.IMAGE_SIZE = 32
    try:
        pred, prob = ao.deploy_sequence(img, w)
        X, Y, Z, T = img.shape
        image = np.pad(ao.normalise_intensity(img, 10.0), ((4, 4), (6, 6), (0, 0), (0, 0)), 'constant')
        prob2 = np.zeros((X, Y, Z, T, 3), dtype=np.float32)
        weight = np.zeros((1, 1, 1, T, 1))
        ww = np.reshape(ao.window_weights(5, 0.1), (1, 1, 1, 9, 1))
        for t in range(T):
            idx = [(i + T) % T for i in range(t - 4, t + 5)]
            image_idx = np.expand_dims(np.transpose(image[:, :, :, idx], axes=(2, 3, 0, 1)).astype(np.float32), -1)
            prob_idx = np.transpose(ao.model_prob(image_idx, w), axes=(2, 3, 0, 1, 4))
            prob2[:, :, :, idx] += prob_idx[4:4 + X, 6:6 + Y] * ww
            weight[:, :, :, idx] += ww
        prob2 /= weight
    finally:
        ao.IMAGE_SIZE = old
    np.testing.assert_allclose(prob, prob2, atol=2e-6)
    assert (pred == np.argmax(prob2, -1)).mean() > 0.999


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("shape,size", [((40, 36, 1, 12), 48), ((30, 44, 2, 9), 64), ((64, 64, 1, 10), 64)])
def test_ao_engine_matches_oracle(shape, size):
    w = synth.make_ao_weights(0)
    img = synth.make_ao_stack(1, shape)
    old = ao.IMAGE_SIZE
    ao.IMAGE_SIZE = size
    try:
        pred_ref, prob_ref = ao.deploy_sequence(img, w)
        _, prob64 = ao.deploy_sequence(img, w, dtype=torch.float64)
    finally:
        ao.IMAGE_SIZE = old
    with aorta.AortaEngine(w) as eng:
        pred, prob = eng.segment_sequence(img, want_prob=True, image_size=size)
    assert pred.shape == shape and pred.dtype == np.int32
    err = np.abs(prob - prob64).max()
    assert err < 1e-4, err                                                # float32 device vs float64 oracle
    bad = pred != np.argmax(prob64, -1)
    if bad.any():                                                         # only near-ties may differ
        srt = np.sort(prob64[bad], -1)
        assert (srt[:, -1] - srt[:, -2]).max() < 2e-4
    assert (pred == pred_ref).mean() >= 0.9995
    assert len(np.unique(pred_ref)) == 3                                  # the fixture exercises every class


@pytest.mark.gpu
def test_ao_rejects_short_sequences_and_bad_sizes():
    from ukbb_cardiac_b200._lib import UkbbError
    w = synth.make_ao_weights(0)
    with aorta.AortaEngine(w) as eng:
        with pytest.raises(UkbbError, match="fewer than the time window"):
            eng.segment_sequence(synth.make_ao_stack(0, (32, 32, 1, 5)), image_size=32)
        with pytest.raises(ValueError, match="does not fit"):
            eng.segment_sequence(synth.make_ao_stack(0, (40, 32, 1, 9)), image_size=32)
        with pytest.raises(UkbbError, match="multiple of 16"):
            eng.segment_sequence(synth.make_ao_stack(0, (20, 20, 1, 9)), image_size=24)
