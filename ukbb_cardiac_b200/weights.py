"""The checkpoint contract of the FCN deploy path.

``common/deploy_network.py:48-49`` restores every variable BY NAME from the
TF-1 checkpoint written by ``common/train_network.py:241,337-339``.  The
variable names are the ``tf.layers`` defaults in graph-creation order
(``common/network.py:179-189`` encoder, ``:201-204`` same_dim, ``:227-229``
head): ``conv2d[_i]/kernel`` (HWIO), ``batch_normalization[_i]/{gamma,beta,
moving_mean,moving_variance}``, ``conv2d_20/bias``.  Names cannot be verified
against a real checkpoint here (none is reachable), so loading validates every
tensor by shape and fails loudly.
"""
from __future__ import annotations

from typing import Dict, List, NamedTuple

import numpy as np

N_LEVEL = 5
N_FILTER = (16, 32, 64, 128, 256)   # train_network.py:176-181 (num_filter=16, x2 per level)
N_BLOCK = (2, 2, 3, 3, 3)           # train_network.py:192
SAME_DIM = 32                       # train_network.py:195
FC = 64
N_CONV = 21
N_BN = 20
BN_EPS = 1e-3                       # tf.layers.batch_normalization default

# train_network.py:156-171; the seg4 la_4ch model has 6 (cardiac_utils.py:147)
N_CLASS = {"sa": 4, "la_2ch": 2, "la_4ch": 3}


class ConvSpec(NamedTuple):
    role: str
    cin: int
    cout: int
    ksize: int
    stride: int
    level: int      # resolution level the OUTPUT lives at (0 = full res)


def conv_name(i: int) -> str:
    return "conv2d" if i == 0 else "conv2d_%d" % i


def bn_name(i: int) -> str:
    return "batch_normalization" if i == 0 else "batch_normalization_%d" % i


def layer_table(n_class: int) -> List[ConvSpec]:
    tab: List[ConvSpec] = []
    cin = 1
    for l in range(N_LEVEL):
        for b in range(N_BLOCK[l]):
            tab.append(ConvSpec("enc%d_%d" % (l, b), cin, N_FILTER[l], 3, 2 if (l > 0 and b == 0) else 1, l))
            cin = N_FILTER[l]
    for l in range(N_LEVEL):
        tab.append(ConvSpec("same%d" % l, N_FILTER[l], SAME_DIM, 1, 1, l))
    tab.append(ConvSpec("fc0", SAME_DIM * N_LEVEL, FC, 1, 1, 0))
    tab.append(ConvSpec("fc1", FC, FC, 1, 1, 0))
    tab.append(ConvSpec("logits", FC, n_class, 1, 1, 0))
    return tab


def expected_shapes(n_class: int) -> Dict[str, tuple]:
    shapes: Dict[str, tuple] = {}
    for i, sp in enumerate(layer_table(n_class)):
        shapes[conv_name(i) + "/kernel"] = (sp.ksize, sp.ksize, sp.cin, sp.cout)
        if i < N_BN:
            for v in ("gamma", "beta", "moving_mean", "moving_variance"):
                shapes[bn_name(i) + "/" + v] = (sp.cout,)
    shapes[conv_name(N_CONV - 1) + "/bias"] = (n_class,)
    return shapes


def infer_n_class(tensors: Dict[str, np.ndarray]) -> int:
    key = conv_name(N_CONV - 1) + "/kernel"
    if key not in tensors:
        raise KeyError("checkpoint has no %r: not a build_FCN checkpoint "
                       "(names follow tf.layers defaults, see weights.py)" % key)
    return int(tensors[key].shape[-1])


def validate(tensors: Dict[str, np.ndarray]) -> int:
    """Check every variable the deploy graph needs; ignore optimizer slots
    (``*/Adam``, ``*/Adam_1``, ``beta1_power``, ``beta2_power``) and anything else.
    Returns n_class."""
    n_class = infer_n_class(tensors)
    for name, shape in expected_shapes(n_class).items():
        if name not in tensors:
            raise KeyError("checkpoint is missing variable %r" % name)
        if tuple(tensors[name].shape) != shape:
            raise ValueError("variable %r has shape %s, build_FCN expects %s"
                             % (name, tuple(tensors[name].shape), shape))
        if tensors[name].dtype != np.float32:
            raise TypeError("variable %r has dtype %s, expected float32" % (name, tensors[name].dtype))
    return n_class


def n_parameters(n_class: int) -> int:
    return int(sum(int(np.prod(s)) for s in expected_shapes(n_class).values()))
