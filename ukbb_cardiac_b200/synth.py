"""Deterministic synthetic inputs: cine stacks in NIfTI
memory order and random-init ``build_FCN`` weight sets under the checkpoint
names of ``weights.py``.  There is no network in the build or GPU environment, so
neither UK Biobank images nor the trained models of ``demo_pipeline.py:28-54``
are reachable; every parity and throughput figure is on these inputs.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

from . import weights as W

SA_SHAPE = (192, 208, 10, 50)      # BASELINE.json configs[0]
LA_SHAPE = (210, 171, 1, 50)       # exercises the 7/7 and 2/3 pad branches


def make_stack(seed: int, shape: Tuple[int, int, int, int] = SA_SHAPE) -> np.ndarray:
    """MR-like, integer-valued float32 (X,Y,Z,T) array, Fortran (NIfTI) memory order:
    smooth background field + a bright blood-pool disc inside a darker ring whose radius
    beats sinusoidally with t + Rayleigh noise, rounded and clipped to [0, 4095]."""
    X, Y, Z, T = shape
    rng = np.random.default_rng(1000 + seed)
    xs = np.arange(X, dtype=np.float32)[:, None]
    ys = np.arange(Y, dtype=np.float32)[None, :]
    field = np.zeros((X, Y), dtype=np.float32)
    for _ in range(8):
        cx, cy = rng.uniform(0, X), rng.uniform(0, Y)
        sg = rng.uniform(0.15, 0.5) * max(X, Y)
        amp = rng.uniform(40, 300)
        field += amp * np.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / (2 * sg * sg)).astype(np.float32)
    cx0, cy0 = X * rng.uniform(0.4, 0.6), Y * rng.uniform(0.4, 0.6)
    r0 = 0.12 * min(X, Y) * rng.uniform(0.8, 1.2)
    dist = np.sqrt((xs - cx0) ** 2 + (ys - cy0) ** 2).astype(np.float32)
    out = np.empty(shape, dtype=np.float32, order="F")
    phase = rng.uniform(0, 2 * np.pi)
    for t in range(T):
        beat = 1.0 - 0.3 * (0.5 - 0.5 * np.cos(2 * np.pi * t / max(T, 1) + phase))
        for z in range(Z):
            taper = 1.0 - 0.5 * abs(z - (Z - 1) / 2.0) / max(Z, 1)
            r = r0 * beat * taper
            pool = 1400.0 / (1.0 + np.exp((dist - r) / 1.5))
            wall = -500.0 * np.exp(-((dist - 1.35 * r) ** 2) / (2 * (0.18 * r + 1.0) ** 2))
            noise = rng.rayleigh(30.0, size=(X, Y)).astype(np.float32)
            out[:, :, z, t] = np.clip(np.rint(field + pool + wall + noise), 0, 4095)
    return out


# Final-layer bias per (seed, n_class), calibrated once with the float64 oracle on
# make_stack(0) so that every class covers a few per cent of the pixels (a Dice test
# on an empty class is 0/0).  See tests/golden/make_golden.py::calibrate_bias.
_CALIBRATED_BIAS: Dict[Tuple[int, int], Tuple[float, ...]] = {
    (0, 4): (7.5353, 3.3673, -1.6585, -2.7473),
    (0, 2): (-2.4877, -4.0748),
    (0, 3): (-0.9091, -6.6617, 2.1724),
    (0, 6): (1.4588, -3.2817, -3.3534, -2.3251, -8.2807, -0.3654),
}


def make_weights(seed: int, n_class: int, calibrated: bool = True) -> Dict[str, np.ndarray]:
    """Random-init weight set under TF names: conv kernels Glorot-uniform (the
    tf.layers default) x sqrt(2); BN gamma~U(.8,1.2), beta~U(-.1,.1), mean~N(0,.05),
    var~U(.8,1.2); final bias ~U(-.5,.5) unless a calibrated one is tabulated."""
    rng = np.random.default_rng(seed)
    t: Dict[str, np.ndarray] = {}
    for i, sp in enumerate(W.layer_table(n_class)):
        fan_in, fan_out = sp.ksize * sp.ksize * sp.cin, sp.ksize * sp.ksize * sp.cout
        lim = np.sqrt(6.0 / (fan_in + fan_out)) * np.sqrt(2.0)
        if i == W.N_CONV - 1:
            lim *= 8.0          # O(1) logit contrast, so FP32 near-ties are rare on the fixtures
        t[W.conv_name(i) + "/kernel"] = rng.uniform(
            -lim, lim, size=(sp.ksize, sp.ksize, sp.cin, sp.cout)).astype(np.float32)
        if i < W.N_BN:
            b = W.bn_name(i)
            t[b + "/gamma"] = rng.uniform(0.8, 1.2, sp.cout).astype(np.float32)
            t[b + "/beta"] = rng.uniform(-0.1, 0.1, sp.cout).astype(np.float32)
            t[b + "/moving_mean"] = rng.normal(0.0, 0.05, sp.cout).astype(np.float32)
            t[b + "/moving_variance"] = rng.uniform(0.8, 1.2, sp.cout).astype(np.float32)
    bias = rng.uniform(-0.5, 0.5, n_class).astype(np.float32)
    if calibrated and (seed, n_class) in _CALIBRATED_BIAS:
        bias = np.asarray(_CALIBRATED_BIAS[(seed, n_class)], dtype=np.float32)
    t[W.conv_name(W.N_CONV - 1) + "/bias"] = bias
    return t


def with_optimizer_slots(t: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """Add the Adam slots a real training checkpoint carries (train_network.py:241
    saves all globals) so loaders are tested against them."""
    out = dict(t)
    for k, v in t.items():
        if k.endswith("/kernel") or k.endswith("/bias") or k.endswith("/gamma") or k.endswith("/beta"):
            out[k + "/Adam"] = np.zeros_like(v)
            out[k + "/Adam_1"] = np.zeros_like(v)
    out["beta1_power"] = np.asarray(0.9, dtype=np.float32)
    out["beta2_power"] = np.asarray(0.999, dtype=np.float32)
    return out


AO_SHAPE = (240, 196, 1, 100)     # BASELINE config C5: synthetic aortic cine


def make_ao_weights(seed: int = 0, f0: int = 16, n_hidden: int = 16, n_class: int = 3) -> Dict[str, np.ndarray]:
    """Random-init weights of the UNet-LSTM aortic model under the checkpoint names of ``aorta.variable_table``: He-scaled kernels,
    BN statistics around the identity, small LSTM biases, an output bias that gives every class some area."""
    from . import aorta
    rng = np.random.default_rng(7000 + seed)
    w = {}
    for name, shape in aorta.variable_table(f0, n_hidden, n_class):
        if name.endswith("kernel"):
            fan_in = int(np.prod(shape[:2])) * (shape[3] if "conv2d_transpose" in name else shape[2])
            w[name] = (rng.normal(size=shape) * np.sqrt(2.0 / fan_in)).astype(np.float32)
        elif name.endswith("gamma") or name.endswith("moving_variance"):
            w[name] = rng.uniform(0.8, 1.2, size=shape).astype(np.float32)
        elif name.endswith("biases"):
            w[name] = rng.uniform(-0.2, 0.2, size=shape).astype(np.float32)
        elif name.endswith("bias"):
            w[name] = rng.uniform(-0.3, 0.3, size=shape).astype(np.float32)
        else:
            w[name] = rng.uniform(-0.1, 0.1, size=shape).astype(np.float32) if name.endswith("beta") else rng.normal(0, 0.05, size=shape).astype(np.float32)
    w["LSTM/output/conv2d/kernel"] *= 6.0          # spread the class scores so that every class covers some area
    return w


def make_ao_stack(seed: int, shape: Tuple[int, int, int, int] = AO_SHAPE) -> np.ndarray:
    """MR-like aortic cine: the same generator as the SA stacks (pulsating bright discs on a smooth field)."""
    return make_stack(5000 + seed, shape)
