"""CPU restatement of ``build_FCN`` + softmax/argmax (TEST INFRASTRUCTURE ONLY).

Follows the reference graph definition
  * ``common/network.py:19-25``   conv2d_bn_relu (tf.layers.conv2d SAME, no bias -> BN -> ReLU)
  * ``common/network.py:117-135`` linear_1d / linear_2d
  * ``common/network.py:138-167`` transpose_upsample2d (diagonal constant filter, conv2d_transpose SAME)
  * ``common/network.py:170-230`` build_FCN topology
  * ``common/train_network.py:174-199`` hyper-parameters, prob = softmax, pred = int32(argmax)
and the TensorFlow-1.x op semantics those call sites rely on (TF is a
third-party dependency that is absent here; version unpinned upstream):
  * SAME padding: out = ceil(in/s); pad_total = max((out-1)*s + k - in, 0);
    pad_before = pad_total // 2 (the extra pixel goes AFTER).
  * FusedBatchNorm inference: y = (x - mean) * (rsqrt(var + eps) * gamma) + beta,
    eps = 1e-3 (tf.layers.batch_normalization default).
  * conv2d_transpose SAME: gradient of a SAME stride-f conv; full transposed
    output cropped by pad_before = (k - f) // 2 at the start.
  * argmax returns the first maximal index.

PARITY UNPINNED (see ``oracle/__init__.py``).

Two evaluations are provided from the same code: float64 (ground truth, used to
adjudicate label near-ties) and float32 (stand-in for "TF on CPU", also the timed
CPU baseline).  Convolutions use torch's CPU ``conv2d`` with EXPLICIT padding;
``conv2d_same_numpy`` is an independent tap-loop used by the tests to pin the
torch path on small cases.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3  # tf.layers.batch_normalization default epsilon

# build_FCN hyper-parameters, common/train_network.py:174-195
N_LEVEL = 5
N_FILTER = [16, 32, 64, 128, 256]
N_BLOCK = [2, 2, 3, 3, 3]
SAME_DIM = 32
FC = 64


def conv_name(i: int) -> str:
    """tf.layers default variable scope of the i-th conv2d created in the graph."""
    return "conv2d" if i == 0 else "conv2d_%d" % i


def bn_name(i: int) -> str:
    return "batch_normalization" if i == 0 else "batch_normalization_%d" % i


def layer_table(n_class: int) -> List[Tuple[str, int, int, int, int]]:
    """(role, cin, cout, ksize, stride) in graph-creation order
    (network.py:179-189 encoder, :201-204 same_dim, :227-229 head)."""
    tab = []
    cin = 1
    for l in range(N_LEVEL):
        for b in range(N_BLOCK[l]):
            stride = 2 if (l > 0 and b == 0) else 1
            tab.append(("enc%d_%d" % (l, b), cin, N_FILTER[l], 3, stride))
            cin = N_FILTER[l]
    for l in range(N_LEVEL):
        tab.append(("same%d" % l, N_FILTER[l], SAME_DIM, 1, 1))
    tab.append(("fc0", SAME_DIM * N_LEVEL, FC, 1, 1))
    tab.append(("fc1", FC, FC, 1, 1))
    tab.append(("logits", FC, n_class, 1, 1))
    return tab


def same_pad(in_size: int, k: int, s: int) -> Tuple[int, int, int]:
    """TF SAME: returns (out, pad_before, pad_after)."""
    out = -(-in_size // s)
    total = max((out - 1) * s + k - in_size, 0)
    before = total // 2
    return out, before, total - before


def linear_1d(sz: int) -> np.ndarray:
    """network.py:117-124"""
    if sz % 2 == 0:
        raise NotImplementedError("`Linear kernel` requires odd filter size.")
    c = (sz + 1) // 2
    h = np.array(list(range(1, c + 1)) + list(range(c - 1, 0, -1)), dtype=np.float32)
    h /= float(c)
    return h


def linear_2d(sz: int) -> np.ndarray:
    """network.py:127-135"""
    h = linear_1d(sz)
    return (h[:, None] * h[None, :]).astype(np.float32)


def _t(x: np.ndarray, dtype) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(x)).to(dtype)


def conv2d_same(x: torch.Tensor, w_hwio: np.ndarray, stride: int) -> torch.Tensor:
    """x: NCHW torch tensor. w: HWIO numpy. TF-SAME conv, no bias."""
    kh, kw = w_hwio.shape[:2]
    _, pt, pb = same_pad(x.shape[2], kh, stride)
    _, pl, pr = same_pad(x.shape[3], kw, stride)
    xp = F.pad(x, (pl, pr, pt, pb))
    w = _t(np.transpose(w_hwio, (3, 2, 0, 1)), x.dtype)
    return F.conv2d(xp, w, stride=stride)


def conv2d_same_numpy(x_nhwc: np.ndarray, w_hwio: np.ndarray, stride: int) -> np.ndarray:
    """Independent tap-loop SAME conv (float64 accumulate) for pinning conv2d_same."""
    n, h, w, cin = x_nhwc.shape
    kh, kw, _, cout = w_hwio.shape
    oh, pt, pb = same_pad(h, kh, stride)
    ow, pl, pr = same_pad(w, kw, stride)
    xp = np.pad(x_nhwc.astype(np.float64), ((0, 0), (pt, pb), (pl, pr), (0, 0)))
    out = np.zeros((n, oh, ow, cout), dtype=np.float64)
    for a in range(kh):
        for b in range(kw):
            patch = xp[:, a:a + (oh - 1) * stride + 1:stride, b:b + (ow - 1) * stride + 1:stride, :]
            out += patch @ w_hwio[a, b].astype(np.float64)
    return out


def bn_relu(x: torch.Tensor, gamma, beta, mean, var, eps: float = BN_EPS) -> torch.Tensor:
    """FusedBatchNorm(is_training=False) + ReLU, Eigen order of operations."""
    dt = x.dtype
    g, b, m, v = (_t(a, dt).view(1, -1, 1, 1) for a in (gamma, beta, mean, var))
    scaling = torch.rsqrt(v + torch.tensor(eps, dtype=dt)) * g
    return torch.relu((x - m) * scaling + b)


def transpose_upsample2d(x: torch.Tensor, factor: int) -> torch.Tensor:
    """network.py:138-167 evaluated per channel (the constant filter is diagonal
    in channels, so groups=C is the same arithmetic without the zeros)."""
    sz = 2 * factor - 1
    c = x.shape[1]
    w = _t(linear_2d(sz), x.dtype).view(1, 1, sz, sz).repeat(c, 1, 1, 1)
    full = F.conv_transpose2d(x, w, stride=factor, groups=c)
    # gradient-of-SAME-conv crop: forward conv out*f -> out has pad_total = k - f = f - 1
    pb = (factor - 1) // 2
    h, wd = x.shape[2] * factor, x.shape[3] * factor
    return full[:, :, pb:pb + h, pb:pb + wd]


def upsample_closed_form_1d(x: np.ndarray, factor: int) -> np.ndarray:
    """SURVEY R5 closed form along the last axis: up[y] = sum_i x[i]*max(0, 1-|y-(i*f+pb)|/f),
    pb = (f-1)//2 ... written from the transposed-conv definition for the KATs."""
    n = x.shape[-1]
    pb = (factor - 1) // 2
    out = np.zeros(x.shape[:-1] + (n * factor,), dtype=np.float64)
    for y in range(n * factor):
        for i in range(n):
            j = y + pb - i * factor  # tap index into the 2f-1 filter
            if 0 <= j < 2 * factor - 1:
                out[..., y] += x[..., i] * (1.0 - abs(j - (factor - 1)) / factor)
    return out


def build_fcn(image_nhwc: np.ndarray, weights: Dict[str, np.ndarray], dtype=torch.float32,
              return_features: bool = False):
    """image: (N, H, W, 1) with H, W multiples of 16.  Returns logits (N, H, W, n_class)
    as a numpy array of `dtype`.  `weights` uses the TF variable names."""
    n_class = weights[conv_name(20) + "/kernel"].shape[-1]
    tab = layer_table(n_class)
    x = _t(np.transpose(image_nhwc, (0, 3, 1, 2)), dtype)
    feats = {}
    li = 0
    level_out = []
    for l in range(N_LEVEL):
        for b in range(N_BLOCK[l]):
            _, cin, cout, k, s = tab[li]
            x = conv2d_same(x, weights[conv_name(li) + "/kernel"], s)
            bn = bn_name(li)
            x = bn_relu(x, weights[bn + "/gamma"], weights[bn + "/beta"],
                        weights[bn + "/moving_mean"], weights[bn + "/moving_variance"])
            if return_features:
                feats[tab[li][0]] = x
            li += 1
        level_out.append(x)
    ups = []
    for l in range(N_LEVEL):
        y = conv2d_same(level_out[l], weights[conv_name(li) + "/kernel"], 1)
        bn = bn_name(li)
        y = bn_relu(y, weights[bn + "/gamma"], weights[bn + "/beta"],
                    weights[bn + "/moving_mean"], weights[bn + "/moving_variance"])
        if return_features:
            feats[tab[li][0]] = y
        li += 1
        ups.append(y if l == 0 else transpose_upsample2d(y, 2 ** l))
    x = torch.cat(ups, dim=1)
    for _ in range(2):
        x = conv2d_same(x, weights[conv_name(li) + "/kernel"], 1)
        bn = bn_name(li)
        x = bn_relu(x, weights[bn + "/gamma"], weights[bn + "/beta"],
                    weights[bn + "/moving_mean"], weights[bn + "/moving_variance"])
        if return_features:
            feats[tab[li][0]] = x
        li += 1
    logits = conv2d_same(x, weights[conv_name(li) + "/kernel"], 1)
    logits = logits + _t(weights[conv_name(li) + "/bias"], dtype).view(1, -1, 1, 1)
    out = logits.permute(0, 2, 3, 1).contiguous().numpy()
    if return_features:
        return out, {k: v.permute(0, 2, 3, 1).contiguous().numpy() for k, v in feats.items()}
    return out


def softmax_argmax(logits: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """train_network.py:198-199: prob = softmax(logits) in the logits dtype,
    pred = int32(argmax(prob, -1)) -- first maximal index wins."""
    m = logits.max(axis=-1, keepdims=True)
    e = np.exp(logits - m)
    prob = e / e.sum(axis=-1, keepdims=True, dtype=logits.dtype)
    pred = np.argmax(prob, axis=-1).astype(np.int32)
    return prob.astype(logits.dtype), pred


def session_run(image_nxyc: np.ndarray, weights: Dict[str, np.ndarray], dtype=torch.float32):
    """Equivalent of sess.run(['prob:0','pred:0'], {'image:0': image, 'training:0': False})
    (deploy_network.py:110-111)."""
    logits = build_fcn(image_nxyc, weights, dtype)
    return softmax_argmax(logits)


def categorical_dice(pred: np.ndarray, truth: np.ndarray, k: int) -> float:
    """image_utils.py:171-175 np_categorical_dice."""
    a = (pred == k).astype(np.float32)
    b = (truth == k).astype(np.float32)
    return float(2 * np.sum(a * b) / (np.sum(a) + np.sum(b)))


def flops_per_slice(h: int, w: int, n_class: int) -> float:
    """Algorithmic FLOPs: 2*Hout*Wout*k^2*Cin*Cout per conv + 4-tap bilinear."""
    tab = layer_table(n_class)
    total = 0.0
    hh, ww = h, w
    li = 0
    sizes = []
    for l in range(N_LEVEL):
        for b in range(N_BLOCK[l]):
            _, cin, cout, k, s = tab[li]
            hh, ww = -(-hh // s), -(-ww // s)
            total += 2.0 * hh * ww * k * k * cin * cout
            li += 1
        sizes.append((hh, ww))
    for l in range(N_LEVEL):
        _, cin, cout, k, s = tab[li]
        total += 2.0 * sizes[l][0] * sizes[l][1] * cin * cout
        li += 1
    for _ in range(3):
        _, cin, cout, k, s = tab[li]
        total += 2.0 * h * w * cin * cout
        li += 1
    total += (N_LEVEL - 1) * 2.0 * 4 * h * w * SAME_DIM
    return total
