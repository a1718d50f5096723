"""Host side of the aortic cine segmentation path: UNet + bidirectional ConvLSTM,
``common/deploy_network_ao.py`` with ``--model UNet-LSTM`` (reference lines cited inline).

The reference restores the graph with ``tf.train.import_meta_graph`` / ``saver.restore`` (``deploy_network_ao.py:59-60``) and runs
``sess.run('prob:0', {'image:0': window})`` once per time frame on the 9-frame window around it (``:146-171``).  ``AortaEngine`` reads
the same checkpoint variables (TF-free bundle reader), and segments a whole cine with ONE device call (``ukbb_ao_segment``).
PyTorch only owns device memory and streams; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from . import _lib, tf_bundle

N_LEVEL = 5                 # --num_level (train_network_ao.py:41)
N_BLOCK = 2                 # n_block = [2, 2, 2, 2, 2] (train_network_ao.py:284)
N_CLASS = 3                 # deploy_network_ao.py:98
IMAGE_SIZE = 256            # deploy_network_ao.py:104
BN_EPS = 1e-3
_BN = ("gamma", "beta", "moving_mean", "moving_variance")


def _suffix(i: int) -> str:
    return "" if i == 0 else "_%d" % i


def variable_table(f0: int = 16, n_hidden: int = 16, n_class: int = N_CLASS) -> List[Tuple[str, Tuple[int, ...]]]:
    """(name, shape) of every checkpoint variable the UNet-LSTM deploy path reads: tf.layers default names inside the variable scopes
    of network_ao.py:23-63 ('UNet/conv{l}', 'UNet/conv{l}_up') and :273-310 ('LSTM/forward', 'LSTM/backward', 'LSTM/output').
    Inferred like the FCN names (no checkpoint, no TensorFlow here): the loader validates by shape and fails loudly."""
    nf = [f0 * 2 ** i for i in range(N_LEVEL)]
    out = []
    cin = 1
    for l in range(N_LEVEL):
        for i in range(N_BLOCK):
            out.append(("UNet/conv%d/conv2d%s/kernel" % (l, _suffix(i)), (3, 3, cin, nf[l])))
            out += [("UNet/conv%d/batch_normalization%s/%s" % (l, _suffix(i), v), (nf[l],)) for v in _BN]
            cin = nf[l]
    for l in range(N_LEVEL - 2, -1, -1):
        sc = "UNet/conv%d_up" % l
        out.append((sc + "/conv2d_transpose/kernel", (3, 3, nf[l], nf[l + 1])))
        out += [("%s/batch_normalization/%s" % (sc, v), (nf[l],)) for v in _BN]
        cin = 2 * nf[l]
        for i in range(N_BLOCK):
            out.append(("%s/conv2d%s/kernel" % (sc, _suffix(i)), (3, 3, cin, nf[l])))
            out += [("%s/batch_normalization%s/%s" % (sc, _suffix(i + 1), v), (nf[l],)) for v in _BN]
            cin = nf[l]
    for d in ("forward", "backward"):
        out.append(("LSTM/%s/conv_lstm_cell/kernel" % d, (3, 3, nf[0] + n_hidden, 4 * n_hidden)))
        out.append(("LSTM/%s/conv_lstm_cell/biases" % d, (4 * n_hidden,)))
    out.append(("LSTM/output/conv2d/kernel", (1, 1, 2 * n_hidden, n_class)))
    out.append(("LSTM/output/conv2d/bias", (n_class,)))
    return out


def validate(tensors: Dict[str, np.ndarray]) -> Tuple[int, int, int]:
    """Returns (f0, n_hidden, n_class) inferred from the shapes; raises ValueError naming the first missing / mis-shaped variable."""
    k0 = tensors.get("UNet/conv0/conv2d/kernel")
    ko = tensors.get("LSTM/output/conv2d/kernel")
    if k0 is None or ko is None:
        raise ValueError("checkpoint has no 'UNet/conv0/conv2d/kernel' / 'LSTM/output/conv2d/kernel': not a UNet-LSTM aortic model")
    f0, n_class, n_hidden = int(k0.shape[-1]), int(ko.shape[-1]), int(ko.shape[-2]) // 2
    for name, shape in variable_table(f0, n_hidden, n_class):
        if name not in tensors:
            raise ValueError("checkpoint variable %r is missing" % name)
        if tuple(tensors[name].shape) != shape:
            raise ValueError("checkpoint variable %r has shape %s, the UNet-LSTM graph expects %s" % (name, tuple(tensors[name].shape), shape))
    return f0, n_hidden, n_class


def normalise_intensity(image: np.ndarray, thres_roi: float = 10.0) -> np.ndarray:
    """image_utils.py:60-67, the reference's own arithmetic (called at deploy_network_ao.py:92): percentile, region mean / std,
    z-score.  4.7 M voxels per aortic cine: host numpy, a few ms next to the device call."""
    val_l = np.percentile(image, thres_roi)
    roi = (image >= val_l)
    mu, sigma = np.mean(image[roi]), np.std(image[roi])
    eps = 1e-6
    return (image - mu) / (sigma + eps)


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class AoWeights(C.Structure):
    _fields_ = [("n_level", C.c_int), ("down", C.POINTER(_lib.ConvWeights)), ("up_transpose", C.POINTER(_lib.ConvWeights)),
                ("up", C.POINTER(_lib.ConvWeights)), ("lstm_kernel", C.POINTER(C.c_float) * 2), ("lstm_bias", C.POINTER(C.c_float) * 2),
                ("out_kernel", C.POINTER(C.c_float)), ("out_bias", C.POINTER(C.c_float)), ("n_hidden", C.c_int), ("n_class", C.c_int),
                ("bn_eps", C.c_float)]


class AortaEngine:
    """UNet + BiConvLSTM inference engine bound to one CUDA device."""

    def __init__(self, tensors: Dict[str, np.ndarray], device: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("AortaEngine needs a CUDA device: libukbb_fcn has no CPU fallback")
        self.lib = _lib.load()
        self.f0, self.n_hidden, self.n_class = validate(tensors)
        self.device = torch.device("cuda", device)
        keep = []

        def conv(scope, kname, bn, ks, cin, cout, stride):
            cw = _lib.ConvWeights()
            k = np.ascontiguousarray(tensors["%s/%s/kernel" % (scope, kname)], dtype=np.float32)
            arrs = [np.ascontiguousarray(tensors["%s/%s/%s" % (scope, bn, v)], dtype=np.float32) for v in _BN]
            keep.extend([k] + arrs)
            cw.kernel = _fptr(k)
            cw.ksize, cw.cin, cw.cout, cw.stride = ks, cin, cout, stride
            cw.gamma, cw.beta, cw.moving_mean, cw.moving_variance = (_fptr(a) for a in arrs)
            return cw

        nf = [self.f0 * 2 ** i for i in range(N_LEVEL)]
        down = (_lib.ConvWeights * (2 * N_LEVEL))()
        cin = 1
        for l in range(N_LEVEL):
            for i in range(N_BLOCK):
                down[2 * l + i] = conv("UNet/conv%d" % l, "conv2d" + _suffix(i), "batch_normalization" + _suffix(i), 3, cin, nf[l],
                                       2 if (l > 0 and i == 0) else 1)
                cin = nf[l]
        upt = (_lib.ConvWeights * (N_LEVEL - 1))()
        up = (_lib.ConvWeights * (2 * (N_LEVEL - 1)))()
        for j, l in enumerate(range(N_LEVEL - 2, -1, -1)):
            sc = "UNet/conv%d_up" % l
            upt[j] = conv(sc, "conv2d_transpose", "batch_normalization", 3, nf[l + 1], nf[l], 2)
            cin = 2 * nf[l]
            for i in range(N_BLOCK):
                up[2 * j + i] = conv(sc, "conv2d" + _suffix(i), "batch_normalization" + _suffix(i + 1), 3, cin, nf[l], 1)
                cin = nf[l]
        w = AoWeights()
        w.n_level, w.down, w.up_transpose, w.up = N_LEVEL, down, upt, up
        for d, name in enumerate(("forward", "backward")):
            k = np.ascontiguousarray(tensors["LSTM/%s/conv_lstm_cell/kernel" % name], dtype=np.float32)
            b = np.ascontiguousarray(tensors["LSTM/%s/conv_lstm_cell/biases" % name], dtype=np.float32)
            keep.extend([k, b])
            w.lstm_kernel[d], w.lstm_bias[d] = _fptr(k), _fptr(b)
        ko = np.ascontiguousarray(tensors["LSTM/output/conv2d/kernel"], dtype=np.float32).reshape(2 * self.n_hidden, self.n_class)
        bo = np.ascontiguousarray(tensors["LSTM/output/conv2d/bias"], dtype=np.float32)
        keep.extend([ko, bo])
        w.out_kernel, w.out_bias = _fptr(ko), _fptr(bo)
        w.n_hidden, w.n_class, w.bn_eps = self.n_hidden, self.n_class, BN_EPS
        handle = C.c_void_p()
        _lib.check(self.lib.ukbb_ao_create(C.byref(w), device, C.byref(handle)))
        self._h = handle

    @classmethod
    def from_checkpoint(cls, model_path: str, device: int = 0) -> "AortaEngine":
        """``saver.restore(sess, model_path)``: reads ``model_path.index`` / ``.data-*``."""
        _lib.load()
        return cls(tf_bundle.read_bundle(model_path), device=device)

    def close(self) -> None:
        if getattr(self, "_h", None):
            self.lib.ukbb_ao_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def launch_count(self) -> int:
        return int(self.lib.ukbb_ao_launch_count(self._h))

    def segment_frames(self, image_tyx: torch.Tensor, x_pre: int, y_pre: int, x: int, y: int, weight_R: int = 5, weight_r: float = 0.1,
                       want_prob: bool = False):
        """image_tyx: cuda float32 [T, Y2, X2] (normalised, zero-padded).  Returns (labels uint8 [T, Y, X], prob [T, Y, X, C] or None)."""
        assert image_tyx.is_cuda and image_tyx.dtype == torch.float32 and image_tyx.dim() == 3 and image_tyx.is_contiguous()
        t, y2, x2 = image_tyx.shape
        labels = torch.empty((t, y, x), dtype=torch.uint8, device=self.device)
        prob = torch.empty((t, y, x, self.n_class), dtype=torch.float32, device=self.device) if want_prob else None
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ukbb_ao_segment(self._h, image_tyx.data_ptr(), t, x2, y2, x_pre, y_pre, x, y, weight_R, float(weight_r),
                                                labels.data_ptr(), prob.data_ptr() if want_prob else None,
                                                torch.cuda.current_stream(self.device).cuda_stream))
        return labels, prob

    def segment_sequence(self, image: np.ndarray, weight_R: int = 5, weight_r: float = 0.1, z_score: bool = True, want_prob: bool = False,
                         image_size: int = IMAGE_SIZE):
        """deploy_network_ao.py:92-186 for one cine ``image`` (X, Y, Z, T): z-score, pad to image_size x image_size, the window loop,
        argmax, crop.  Returns (pred int32 (X, Y, Z, T), prob float32 (X, Y, Z, T, C) or None)."""
        X, Y, Z, T = image.shape
        img = normalise_intensity(image, 10.0) if z_score else image
        X2 = Y2 = image_size
        if X > X2 or Y > Y2:
            raise ValueError("image %dx%d does not fit the fixed %dx%d network input (deploy_network_ao.py:104)" % (X, Y, X2, Y2))
        x_pre, y_pre = int((X2 - X) / 2), int((Y2 - Y) / 2)
        img = np.pad(img, ((x_pre, X2 - X - x_pre), (y_pre, Y2 - Y - y_pre), (0, 0), (0, 0)), 'constant')
        pred = np.zeros((X, Y, Z, T), dtype=np.int32)
        prob = np.zeros((X, Y, Z, T, self.n_class), dtype=np.float32) if want_prob else None
        for z in range(Z):
            # (X2, Y2, T) -> device [T][Y2][X2]: the NIfTI memory order of the slice's frames
            dev = torch.from_numpy(np.ascontiguousarray(np.transpose(img[:, :, z, :], (2, 1, 0)), dtype=np.float32)).to(self.device)
            lab, pr = self.segment_frames(dev, x_pre, y_pre, X, Y, weight_R, weight_r, want_prob)
            pred[:, :, z, :] = np.transpose(lab.cpu().numpy(), (2, 1, 0))
            if want_prob:
                prob[:, :, z] = np.transpose(pr.cpu().numpy(), (2, 1, 0, 3))
        return pred, prob
