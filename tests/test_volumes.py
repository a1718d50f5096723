"""ventricular volumes (short_axis/eval_ventricular_volume.py:34-81) from class counts."""
import os
import subprocess
import sys

import numpy as np
import pytest

from ukbb_cardiac_b200 import nifti, synth, volumes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_measures(seg, pixdim, dim4):
    """The arithmetic of eval_ventricular_volume.py:40-74 written out as the reference does it, on the label volume."""
    volume_per_pix = pixdim[1] * pixdim[2] * pixdim[3] * 1e-3
    density = 1.05
    heart_rate = 60.0 / (dim4 * pixdim[4])
    vol_t = np.sum(seg == 1, axis=(0, 1, 2)) * volume_per_pix
    frame = {"ED": 0, "ES": np.argmin(vol_t)}
    val = {}
    for fr_name, fr in frame.items():
        val["LV{0}V".format(fr_name)] = np.sum(seg[:, :, :, fr] == 1) * volume_per_pix
        val["LV{0}M".format(fr_name)] = np.sum(seg[:, :, :, fr] == 2) * volume_per_pix * density
        val["RV{0}V".format(fr_name)] = np.sum(seg[:, :, :, fr] == 3) * volume_per_pix
    val["LVSV"] = val["LVEDV"] - val["LVESV"]
    val["LVCO"] = val["LVSV"] * heart_rate * 1e-3
    val["LVEF"] = val["LVSV"] / val["LVEDV"] * 100
    val["RVSV"] = val["RVEDV"] - val["RVESV"]
    val["RVCO"] = val["RVSV"] * heart_rate * 1e-3
    val["RVEF"] = val["RVSV"] / val["RVEDV"] * 100
    return [val["LVEDV"], val["LVESV"], val["LVSV"], val["LVEF"], val["LVCO"], val["LVEDM"],
            val["RVEDV"], val["RVESV"], val["RVSV"], val["RVEF"]]


def beating_labels(shape=(24, 20, 5, 8), seed=3):
    """A label volume whose LV cavity shrinks and re-expands over the cycle (ES in the middle)."""
    x, y, z, t = shape
    rng = np.random.default_rng(seed)
    seg = np.zeros(shape, dtype=np.uint8)
    yy, xx = np.meshgrid(np.arange(y), np.arange(x))
    for fr in range(t):
        r_lv = 4.0 - 2.0 * np.sin(np.pi * fr / (t - 1))
        for k in range(z):
            d = np.hypot(xx - 9, yy - 10)
            seg[..., k, fr][d < r_lv + 2.5] = 2
            seg[..., k, fr][d < r_lv] = 1
            seg[..., k, fr][np.hypot(xx - 17, yy - 10) < 3.0 + 0.3 * rng.random()] = 3
    return seg


def test_measures_match_reference_arithmetic():
    seg = beating_labels()
    pixdim = np.array([1.0, 1.8, 1.8, 10.0, 0.031, 0, 0, 0], dtype=np.float32)
    want = reference_measures(seg, pixdim, seg.shape[3])
    fc = volumes.frame_counts_from_labels(seg)
    val = volumes.ventricular_volumes(fc, pixdim, seg.shape[3])
    got = volumes.table_row(val)
    assert got == want                                        # same operations on the same integers: bit-identical doubles
    assert val["ES_frame"] == int(np.argmin(fc[:, 1])) and 0 < val["ES_frame"] < seg.shape[3] - 1
    # per-slice counts [T, Z, C], the layout the device emits, reduce to the same numbers
    per_slice = np.stack([[[(seg[:, :, k, fr] == c).sum() for c in range(4)] for k in range(seg.shape[2])] for fr in range(seg.shape[3])])
    assert volumes.table_row(volumes.ventricular_volumes(per_slice, pixdim, seg.shape[3])) == want


def test_cli_writes_reference_csv(tmp_path):
    import pandas as pd
    seg = beating_labels()
    for name in ("subjA", "subjB"):
        d = tmp_path / name
        d.mkdir()
        img = nifti.Nifti1Image(np.zeros(seg.shape, np.float32, order="F"), np.diag([1.8, 1.8, 10.0, 1.0]))
        img.header["pixdim"][4] = 0.031
        nifti.save(img, str(d / "sa.nii.gz"))
        lab = nifti.Nifti1Image(seg.astype(np.float64), img.affine)          # the reference stores labels as float64 (deploy_network.py:134-137)
        lab.header["pixdim"] = img.header["pixdim"]
        nifti.save(lab, str(d / "seg_sa.nii.gz"))
    (tmp_path / "empty").mkdir()                                              # no files: skipped like the reference does
    out = tmp_path / "table.csv"
    subprocess.run([sys.executable, os.path.join(ROOT, "short_axis", "eval_ventricular_volume.py"), "--data_dir", str(tmp_path),
                    "--output_csv", str(out)], check=True, capture_output=True)
    df = pd.read_csv(out, index_col=0)
    assert list(df.columns) == volumes.COLUMNS and list(df.index) == ["subjA", "subjB"]
    pixdim = nifti.load(str(tmp_path / "subjA" / "sa.nii.gz")).header["pixdim"]
    want = reference_measures(seg, pixdim, seg.shape[3])
    np.testing.assert_allclose(df.loc["subjA"].values, want, rtol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_device_counts_give_the_same_volumes(mode):
    """The class counts the classifier stage emits on the device equal the counts of the label volume it wrote, so the
    clinical measures need no second pass over seg_sa."""
    from ukbb_cardiac_b200.fcn import FCNEngine
    w = synth.make_weights(0, 4)
    vol = synth.make_stack(2, (64, 48, 4, 6))
    with FCNEngine(w, mode=mode) as eng:
        lab, _, counts = eng.segment_volume(vol)
    assert counts.shape == (6, 4, 4)
    fc = volumes.frame_counts_from_labels(lab)
    assert (counts.sum(axis=1) == fc).all()
    pixdim = np.array([1.0, 1.8, 1.8, 10.0, 0.031, 0, 0, 0], dtype=np.float32)
    a = volumes.ventricular_volumes(counts, pixdim, 6)
    b = volumes.ventricular_volumes(fc, pixdim, 6)
    # a random-init network may leave a class empty in a frame: 0 / 0 = NaN in both (assert_array_equal treats NaNs as equal)
    np.testing.assert_array_equal(np.array(volumes.table_row(a)), np.array(volumes.table_row(b)))
    assert a["ES_frame"] == b["ES_frame"]
