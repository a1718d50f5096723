"""Ventricular volumes from the segmentation: the per-frame class voxel counts that
`short_axis/eval_ventricular_volume.py:34-81` takes from `seg_sa.nii.gz` are already produced on the device by the
classifier stage of the forward (`ukbb_fcn_class_counts`, per-slice counts [T, Z, C]), so the clinical measures of one
subject follow from 2 x 3 integers and the NIfTI header without touching the label volume again.

Labels of the SA network (`cardiac_utils.py`): 1 = LV cavity, 2 = LV myocardium, 3 = RV cavity."""
from __future__ import annotations

import numpy as np

COLUMNS = ["LVEDV (mL)", "LVESV (mL)", "LVSV (mL)", "LVEF (%)", "LVCO (L/min)", "LVM (g)",
           "RVEDV (mL)", "RVESV (mL)", "RVSV (mL)", "RVEF (%)"]            # eval_ventricular_volume.py:77-79
DENSITY = 1.05                                                               # g/mL, eval_ventricular_volume.py:43


def frame_counts_from_labels(seg: np.ndarray, n_class: int = 4) -> np.ndarray:
    """[T, C] voxel counts of a label volume (X, Y, Z, T) -- what `np.sum(seg == k, axis=(0, 1, 2))` gives the reference."""
    seg = np.asarray(seg)
    if seg.ndim == 3:
        seg = seg[..., None]
    t = seg.shape[3]
    out = np.zeros((t, n_class), dtype=np.int64)
    for k in range(n_class):
        out[:, k] = (seg == k).sum(axis=(0, 1, 2))
    return out


def ventricular_volumes(frame_counts: np.ndarray, pixdim: np.ndarray, n_frames: int) -> dict:
    """eval_ventricular_volume.py:40-71.  frame_counts: [T, C] class voxel counts per time frame (or [T, Z, C] per-slice
    counts as the device emits them); pixdim: the 8-element NIfTI `pixdim` of sa.nii.gz; n_frames: `dim[4]`."""
    fc = np.asarray(frame_counts, dtype=np.int64)
    if fc.ndim == 3:
        fc = fc.sum(axis=1)
    pixdim = np.asarray(pixdim)
    volume_per_pix = pixdim[1] * pixdim[2] * pixdim[3] * 1e-3               # :41-42 (mL per voxel), in the header's float32
    duration_per_cycle = n_frames * pixdim[4]                                # :46
    heart_rate = 60.0 / duration_per_cycle                                   # :47
    vol_t = fc[:, 1] * volume_per_pix                                        # :54
    frame = {"ED": 0, "ES": int(np.argmin(vol_t))}                           # :53-55
    val = {}
    for name, fr in frame.items():                                           # :58-62
        val["LV%sV" % name] = fc[fr, 1] * volume_per_pix
        val["LV%sM" % name] = fc[fr, 2] * volume_per_pix * DENSITY
        val["RV%sV" % name] = fc[fr, 3] * volume_per_pix
    val["LVSV"] = val["LVEDV"] - val["LVESV"]                                # :64-66
    val["LVCO"] = val["LVSV"] * heart_rate * 1e-3
    val["LVEF"] = val["LVSV"] / val["LVEDV"] * 100
    val["RVSV"] = val["RVEDV"] - val["RVESV"]                                # :68-70
    val["RVCO"] = val["RVSV"] * heart_rate * 1e-3
    val["RVEF"] = val["RVSV"] / val["RVEDV"] * 100
    val["ES_frame"] = frame["ES"]
    return val


def table_row(val: dict) -> list:
    """The CSV line of eval_ventricular_volume.py:72-74."""
    return [val["LVEDV"], val["LVESV"], val["LVSV"], val["LVEF"], val["LVCO"], val["LVEDM"],
            val["RVEDV"], val["RVESV"], val["RVSV"], val["RVEF"]]
