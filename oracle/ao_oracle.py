"""CPU restatement of the aortic UNet + bidirectional ConvLSTM deploy path (TEST INFRASTRUCTURE ONLY).

Follows
  * ``common/network_ao.py:18-64``     UNet: conv2d_bn_relu blocks (stride 2 at the first conv of levels 1..), learned
    ``conv2d_transpose_bn_relu`` (3x3, stride 2, SAME; ``common/network.py:28-34``) + skip concat ``[conv_l, up]`` + n_block convs
  * ``common/network_ao.py:255-319``   BiConv_LSTM: forward cell over t = 0..n-1, backward cell over t = n-1..0, per-step
    concat ``[fw_t, bw_t]`` -> 1x1 conv (bias) -> logits
  * ``common/network_ao.py:322-399``   UNet_LSTM_Model: features = net['conv0_up'], prob = softmax(outputs)
  * ``common/deploy_network_ao.py:83-183`` z-score normalisation, fixed pad to 256 x 256, circular window of 2 R - 1 frames around
    every frame, weighted overlap-add of the window probabilities, argmax
  * ``common/image_utils.py:60-67``    normalise_intensity
and the TensorFlow-1.x semantics those call sites rely on (TensorFlow is absent; version unpinned upstream):
  * ``tf.layers.conv2d_transpose(k=3, strides=2, padding='same')``: gradient of a SAME stride-2 conv, i.e.
    big[y] = sum_{i, k : 2 i + k = y} small[i] * w[k] (pad_before = 0), kernel layout [kh, kw, out, in];
  * ``tf.contrib.rnn.Conv2DLSTMCell(kernel_shape=[3, 3])``: ONE SAME conv of concat([x, h]) with kernel [3, 3, Cx + Ch, 4 Ch] plus
    bias, split along channels into (i, j, f, o); c' = sigmoid(f + 1) * c + sigmoid(i) * tanh(j); h' = tanh(c') * sigmoid(o);
    zero initial state.
PARITY UNPINNED: no TensorFlow, no reference tests or golden vectors for this path; pinned by hand-authored known-answer tests only.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F

from . import fcn_oracle as fo

N_LEVEL = 5
N_BLOCK = [2, 2, 2, 2, 2]          # train_network_ao.py:284
N_HIDDEN = 16                      # --num_hidden
N_CLASS = 3                        # deploy_network_ao.py:98
IMAGE_SIZE = 256                   # deploy_network_ao.py:104


def n_filter(f0: int = 16) -> List[int]:
    return [f0 * 2 ** i for i in range(N_LEVEL)]      # train_network_ao.py:275-277


def normalise_intensity(image: np.ndarray, thres_roi: float = 10.0) -> np.ndarray:
    """image_utils.py:60-67."""
    val_l = np.percentile(image, thres_roi)
    roi = (image >= val_l)
    mu, sigma = np.mean(image[roi]), np.std(image[roi])
    eps = 1e-6
    return (image - mu) / (sigma + eps)


def conv2d_transpose_same(x: torch.Tensor, w_tf: np.ndarray, stride: int = 2) -> torch.Tensor:
    """x NCHW; w_tf [kh, kw, out, in] (tf.layers.conv2d_transpose variable layout)."""
    w = fo._t(np.transpose(w_tf, (3, 2, 0, 1)), x.dtype)              # torch: [in, out, kh, kw]
    full = F.conv_transpose2d(x, w, stride=stride)
    k = w_tf.shape[0]
    pb = max(k - stride, 0) // 2                                      # pad_before of the forward SAME conv (0 for k = 3, s = 2)
    return full[:, :, pb:pb + x.shape[2] * stride, pb:pb + x.shape[3] * stride]


def _bn_relu(x, w, scope):
    return fo.bn_relu(x, w[scope + "/gamma"], w[scope + "/beta"], w[scope + "/moving_mean"], w[scope + "/moving_variance"])


def _suffix(i: int) -> str:
    return "" if i == 0 else "_%d" % i


def unet_features(image_nxyc: np.ndarray, w: Dict[str, np.ndarray], dtype=torch.float32) -> torch.Tensor:
    """network_ao.py:18-64 up to net['conv0_up'] (the conv_out logits are not used by the UNet-LSTM model).  Returns NCHW."""
    x = fo._t(np.transpose(image_nxyc, (0, 3, 1, 2)), dtype)
    net = {}
    for l in range(N_LEVEL):
        sc = "UNet/conv%d" % l
        for i in range(N_BLOCK[l]):
            x = fo.conv2d_same(x, w["%s/conv2d%s/kernel" % (sc, _suffix(i))], 2 if (l > 0 and i == 0) else 1)
            x = _bn_relu(x, w, "%s/batch_normalization%s" % (sc, _suffix(i)))
        net[l] = x
    up = net[N_LEVEL - 1]
    for l in range(N_LEVEL - 2, -1, -1):
        sc = "UNet/conv%d_up" % l
        x = conv2d_transpose_same(up, w[sc + "/conv2d_transpose/kernel"], 2)
        x = _bn_relu(x, w, sc + "/batch_normalization")
        x = torch.cat([net[l], x], dim=1)
        for i in range(N_BLOCK[l]):
            x = fo.conv2d_same(x, w["%s/conv2d%s/kernel" % (sc, _suffix(i))], 1)
            x = _bn_relu(x, w, "%s/batch_normalization%s" % (sc, _suffix(i + 1)))
        up = x
    return up


def conv_lstm_step(x, h, c, kernel, bias):
    """tf.contrib.rnn.Conv2DLSTMCell.call: x, h, c NCHW; kernel [3, 3, Cx + Ch, 4 Ch]; gates (i, j, f, o), forget_bias = 1."""
    g = fo.conv2d_same(torch.cat([x, h], dim=1), kernel, 1) + fo._t(bias, x.dtype).view(1, -1, 1, 1)
    i, j, f, o = torch.chunk(g, 4, dim=1)
    c2 = torch.sigmoid(f + 1.0) * c + torch.sigmoid(i) * torch.tanh(j)
    h2 = torch.tanh(c2) * torch.sigmoid(o)
    return h2, c2


def biconv_lstm_logits(features: torch.Tensor, w: Dict[str, np.ndarray]) -> torch.Tensor:
    """features [N, T, C, H, W] -> logits [N, T, n_class, H, W] (network_ao.py:255-319)."""
    n, t_n = features.shape[:2]
    nh = w["LSTM/forward/conv_lstm_cell/biases"].shape[0] // 4
    out_fw, out_bw = [None] * t_n, [None] * t_n
    for name, order, store in (("forward", range(t_n), out_fw), ("backward", range(t_n - 1, -1, -1), out_bw)):
        h = torch.zeros((n, nh) + tuple(features.shape[3:]), dtype=features.dtype)
        c = torch.zeros_like(h)
        for t in order:
            h, c = conv_lstm_step(features[:, t], h, c, w["LSTM/%s/conv_lstm_cell/kernel" % name], w["LSTM/%s/conv_lstm_cell/biases" % name])
            store[t] = h
    outs = []
    for t in range(t_n):
        y = fo.conv2d_same(torch.cat([out_fw[t], out_bw[t]], dim=1), w["LSTM/output/conv2d/kernel"], 1)
        outs.append(y + fo._t(w["LSTM/output/conv2d/bias"], features.dtype).view(1, -1, 1, 1))
    return torch.stack(outs, dim=1)


def model_prob(image_ntxyc: np.ndarray, w: Dict[str, np.ndarray], dtype=torch.float32) -> np.ndarray:
    """sess.run('prob:0', {'image:0': image NTXYC}) of the UNet-LSTM model -> prob NTXYC."""
    n, t_n = image_ntxyc.shape[:2]
    feat = unet_features(image_ntxyc.reshape((n * t_n,) + image_ntxyc.shape[2:]), w, dtype)
    feat = feat.reshape((n, t_n) + tuple(feat.shape[1:]))
    logits = biconv_lstm_logits(feat, w)                              # [N, T, C, X, Y]
    prob = torch.softmax(logits, dim=2)
    return prob.permute(0, 1, 3, 4, 2).contiguous().numpy()


def window_weights(weight_R: int, weight_r: float) -> np.ndarray:
    """deploy_network_ao.py:134-143."""
    time_window = weight_R * 2 - 1
    rad = int((time_window - 1) / 2)
    w = []
    for t in range(time_window):
        d = abs(t - rad)
        w += [pow(1 - float(d) / weight_R, weight_r) if d <= weight_R else 0]
    return np.array(w)


def deploy_sequence(image_xyzt: np.ndarray, w: Dict[str, np.ndarray], weight_R: int = 5, weight_r: float = 0.1, time_step: int = 1,
                    z_score: bool = True, dtype=torch.float32, features=None):
    """deploy_network_ao.py:83-183 for the UNet-LSTM model.  Returns (pred int32 (X, Y, Z, T), prob float32 (X, Y, Z, T, C))."""
    X, Y, Z, T = image_xyzt.shape
    image = normalise_intensity(image_xyzt, 10.0) if z_score else image_xyzt
    n_class = w["LSTM/output/conv2d/kernel"].shape[-1]
    prob = np.zeros((X, Y, Z, T, n_class), dtype=np.float32)
    X2 = Y2 = IMAGE_SIZE
    x_pre, y_pre = int((X2 - X) / 2), int((Y2 - Y) / 2)
    image = np.pad(image, ((x_pre, X2 - X - x_pre), (y_pre, Y2 - Y - y_pre), (0, 0), (0, 0)), 'constant')
    time_window = weight_R * 2 - 1
    rad = int((time_window - 1) / 2)
    weight = np.zeros((1, 1, 1, T, 1))
    ww = np.reshape(window_weights(weight_R, weight_r), (1, 1, 1, time_window, 1))
    # the UNet features of a frame do not depend on the window it appears in: evaluate them once per frame
    fr = np.transpose(image, (2, 3, 0, 1)).astype(np.float32)[..., None]          # [Z, T, X2, Y2, 1]
    feat_all = unet_features(fr.reshape((Z * T, X2, Y2, 1)), w, dtype)
    feat_all = feat_all.reshape((Z, T) + tuple(feat_all.shape[1:]))
    for t in range(0, T, time_step):
        idx = [(i + T) if i < 0 else (i - T) if i >= T else i for i in range(t - rad, t + rad + 1)]
        logits = biconv_lstm_logits(feat_all[:, idx], w)
        prob_idx = torch.softmax(logits, dim=2).permute(3, 4, 0, 1, 2).contiguous().numpy().astype(np.float32)    # XYNTC
        prob[:, :, :, idx] += prob_idx[x_pre:x_pre + X, y_pre:y_pre + Y] * ww
        weight[:, :, :, idx] += ww
    prob /= weight
    pred = np.argmax(prob, axis=-1).astype(np.int32)
    return pred, prob


def weight_names(f0: int = 16, n_hidden: int = N_HIDDEN, n_class: int = N_CLASS):
    """(name, shape) of every variable the deploy path reads, TF-1 tf.layers naming inside the variable scopes of network_ao.py
    (inferred, like the FCN names: no checkpoint and no TensorFlow here to verify them)."""
    nf = n_filter(f0)
    out = []
    cin = 1
    for l in range(N_LEVEL):
        for i in range(N_BLOCK[l]):
            out.append(("UNet/conv%d/conv2d%s/kernel" % (l, _suffix(i)), (3, 3, cin, nf[l])))
            out += [("UNet/conv%d/batch_normalization%s/%s" % (l, _suffix(i), v), (nf[l],)) for v in ("gamma", "beta", "moving_mean", "moving_variance")]
            cin = nf[l]
    for l in range(N_LEVEL - 2, -1, -1):
        sc = "UNet/conv%d_up" % l
        out.append((sc + "/conv2d_transpose/kernel", (3, 3, nf[l], nf[l + 1])))
        out += [("%s/batch_normalization/%s" % (sc, v), (nf[l],)) for v in ("gamma", "beta", "moving_mean", "moving_variance")]
        cin = 2 * nf[l]
        for i in range(N_BLOCK[l]):
            out.append(("%s/conv2d%s/kernel" % (sc, _suffix(i)), (3, 3, cin, nf[l])))
            out += [("%s/batch_normalization%s/%s" % (sc, _suffix(i + 1), v), (nf[l],)) for v in ("gamma", "beta", "moving_mean", "moving_variance")]
            cin = nf[l]
    for d in ("forward", "backward"):
        out.append(("LSTM/%s/conv_lstm_cell/kernel" % d, (3, 3, nf[0] + n_hidden, 4 * n_hidden)))
        out.append(("LSTM/%s/conv_lstm_cell/biases" % d, (4 * n_hidden,)))
    out.append(("LSTM/output/conv2d/kernel", (1, 1, 2 * n_hidden, n_class)))
    out.append(("LSTM/output/conv2d/bias", (n_class,)))
    return out
