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


def round16(a: np.ndarray, dt) -> np.ndarray:
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dt).to(torch.float32).numpy()


def e4m3(a: np.ndarray):
    """float32 -> (uint8 codes, decoded float32) of FP8 E4M3, round to nearest even, saturating at +-448."""
    q = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).clamp(-448.0, 448.0).to(torch.float8_e4m3fn)
    return q.view(torch.uint8).numpy(), q.to(torch.float32).numpy()


def x2_planes(x32: np.ndarray):
    """Activation tensor [..., C] -> the two planes of the x2 scheme (tc_common.cuh): hi = rn_fp16(x); lo plane per 16 channels =
    [e4m3((x - hi) 2^11) x 16 | e4m3(hi) x 16].  Returns (hi, decoded lo8, decoded hi8, lo plane as float16 bit patterns [..., C])."""
    hi = round16(x32, torch.float16)
    lo_c, lo_f = e4m3((np.asarray(x32, dtype=np.float32) - hi) * 2048.0)
    hi_c, hi_f = e4m3(hi)
    g = x32.shape[:-1] + (x32.shape[-1] // 16, 16)
    plane = np.concatenate([lo_c.reshape(g), hi_c.reshape(g)], axis=-1)              # [..., C / 16, 32] bytes
    return hi, lo_f, hi_f, np.ascontiguousarray(plane).view(np.float16).reshape(x32.shape)


def x2_decode(o: torch.Tensor) -> np.ndarray:
    """[2][...][C] float16 tensor written by an x2 kernel -> float64 values hi + lo8 2^-11."""
    hi = o[0].float().cpu().numpy().astype(np.float64)
    b = o[1].contiguous().view(torch.uint8).cpu()
    b = b.reshape(b.shape[:-1] + (b.shape[-1] // 32, 32))[..., :16].contiguous()
    lo = b.view(torch.float8_e4m3fn).to(torch.float32).numpy().reshape(hi.shape).astype(np.float64)
    return hi + lo / 2048.0
