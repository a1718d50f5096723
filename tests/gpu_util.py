"""Helpers shared by the GPU parity tests."""
import numpy as np
import torch

from oracle import fcn_oracle as fo


def to_device_layout(image_nxyc: np.ndarray) -> torch.Tensor:
    """TF [N, X, Y, 1] -> device [N, Y, X] float32 cuda."""
    return torch.from_numpy(np.ascontiguousarray(np.transpose(image_nxyc[..., 0], (0, 2, 1)))).cuda()


def from_device_logits(logits: torch.Tensor) -> np.ndarray:
    """device [N, Y, X, C] -> TF [N, X, Y, C]."""
    return logits.permute(0, 2, 1, 3).contiguous().cpu().numpy()


def from_device_labels(labels: torch.Tensor) -> np.ndarray:
    return labels.permute(0, 2, 1).contiguous().cpu().numpy()


def adjudicate_labels(labels: np.ndarray, logits64: np.ndarray, gap_tol: float):
    """Compare labels with argmax of the float64 logits; every mismatch must be a near-tie
    (float64 top-2 gap below gap_tol).  Returns (n_mismatch, worst_gap)."""
    ref = np.argmax(logits64, axis=-1)
    bad = labels != ref
    if not bad.any():
        return 0, 0.0
    srt = np.sort(logits64[bad], axis=-1)
    gaps = srt[:, -1] - srt[:, -2]
    chosen = np.take_along_axis(logits64[bad], labels[bad][:, None].astype(np.int64), axis=-1)[:, 0]
    # the chosen class must itself be within gap_tol of the maximum
    worst = float(np.max(srt[:, -1] - chosen))
    assert worst <= gap_tol, "label mismatch that is not a near-tie: float64 gap %g > %g" % (worst, gap_tol)
    return int(bad.sum()), float(gaps.max())
