"""CPU restatement of the deploy_network.py host loop (TEST INFRASTRUCTURE ONLY).

Follows ``common/deploy_network.py:80-131`` (process_seq) and ``:166-201``
(ED/ES frames) and ``common/image_utils.py:70-77`` (rescale_intensity) as they
execute under numpy 2.x in the build container:

  * ``np.percentile(image, (1, 99))`` over ALL voxels of the array, method
    'linear'; for float32 input numpy forms ``diff = b - a`` in float32 and the
    lerp ``a + diff*t`` (``b - diff*(1-t)`` when t >= 0.5) in float64 -> float64.
  * clip IN PLACE (the float64 thresholds are rounded to the array dtype on
    assignment), then ``(float32(x) - vl) / (vh - vl)`` in float64, rounded to
    float32 when the frame is fed to the network (deploy_network.py:106).
  * zero-pad X, Y to multiples of 16, pre = floor, extra pixel after (:97-100).
  * per frame t: (X,Y,Z)->(Z,X,Y,1) float32, run, transpose back, crop (:103-116).
  * ES = argmin_t / argmax_t of the class-1 voxel count (:125-131).

The preprocessing half is PINNED by ``tests/golden/rescale_*.npz`` (outputs of
the reference's own function, see tests/golden/make_golden.py).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Tuple

import numpy as np

from . import fcn_oracle


def percentile_linear(image: np.ndarray, q: float) -> float:
    """numpy 'linear' percentile restated from first principles (float32 input)."""
    n = image.size
    vi = (n - 1) * (np.float64(q) / 100.0)
    lo = int(math.floor(vi))
    hi = min(lo + 1, n - 1)
    # order statistics lo and hi (np.partition places the k-th smallest at index k)
    a = np.partition(image.reshape(-1), sorted({lo, hi}))
    t = vi - lo
    diff = a[hi] - a[lo]  # in the array dtype, like numpy's _lerp
    if t >= 0.5:
        return float(np.float64(a[hi]) - np.float64(diff) * (1.0 - t))
    return float(np.float64(a[lo]) + np.float64(diff) * t)


def rescale_intensity(image: np.ndarray, thres=(1.0, 99.0)) -> np.ndarray:
    """image_utils.py:70-77, including the in-place clipping of `image`."""
    val_l = np.float64(percentile_linear(image, thres[0]))
    val_h = np.float64(percentile_linear(image, thres[1]))
    image[image < val_l] = val_l
    image[image > val_h] = val_h
    return (image.astype(np.float32) - val_l) / (val_h - val_l)


def pad16(x: int) -> Tuple[int, int, int]:
    """deploy_network.py:97-99 -> (X2, x_pre, x_post)."""
    x2 = int(math.ceil(x / 16.0)) * 16
    pre = int((x2 - x) / 2)
    return x2, pre, (x2 - x) - pre


def deploy_sequence(image_xyzt: np.ndarray, run: Callable[[np.ndarray], np.ndarray]
                    ) -> Tuple[np.ndarray, np.ndarray]:
    """deploy_network.py:83-116.  `run(image_fr NXYC float32) -> pred (N,X,Y) int32`.
    Returns (pred float64 (X,Y,Z,T), clipped input image)."""
    image = image_xyzt
    X, Y, Z, T = image.shape
    orig = image
    image = rescale_intensity(image, (1, 99))
    pred = np.zeros(image.shape)
    X2, x_pre, x_post = pad16(X)
    Y2, y_pre, y_post = pad16(Y)
    image = np.pad(image, ((x_pre, x_post), (y_pre, y_post), (0, 0), (0, 0)), "constant")
    for t in range(T):
        fr = np.transpose(image[:, :, :, t], axes=(2, 0, 1)).astype(np.float32)
        fr = np.expand_dims(fr, axis=-1)
        p = run(fr)
        p = np.transpose(p, axes=(1, 2, 0))
        pred[:, :, :, t] = p[x_pre:x_pre + X, y_pre:y_pre + Y]
    return pred, orig


def deploy_volume(image_xyz: np.ndarray, run) -> np.ndarray:
    """deploy_network.py:166-201 (one ED or ES volume; 2-D input gets a Z axis)."""
    image = image_xyz
    X, Y = image.shape[:2]
    if image.ndim == 2:
        image = np.expand_dims(image, axis=2)
    image = rescale_intensity(image, (1, 99))
    X2, x_pre, x_post = pad16(X)
    Y2, y_pre, y_post = pad16(Y)
    image = np.pad(image, ((x_pre, x_post), (y_pre, y_post), (0, 0)), "constant")
    image = np.expand_dims(np.transpose(image, axes=(2, 0, 1)).astype(np.float32), axis=-1)
    p = run(image)
    p = np.transpose(p, axes=(1, 2, 0))
    return p[x_pre:x_pre + X, y_pre:y_pre + Y]


def es_frame(pred: np.ndarray, seq_name: str, seg4: bool = False) -> int:
    """deploy_network.py:125-131."""
    cnt = np.sum(pred == 1, axis=(0, 1, 2))
    if seq_name == "sa" or (seq_name == "la_4ch" and seg4):
        return int(np.argmin(cnt))
    return int(np.argmax(cnt))


def make_runner(weights: Dict[str, np.ndarray], dtype=None):
    import torch
    dt = torch.float32 if dtype is None else dtype

    def run(fr):
        _, pred = fcn_oracle.session_run(fr, weights, dt)
        return pred
    return run
