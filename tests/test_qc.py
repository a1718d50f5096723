"""Quality-control gates.

CPU: the oracle restatement (oracle/qc_oracle.py) against tests/golden/qc_reference.npz, which holds outputs of the reference's
own get_largest_cc / remove_small_cc / sa_pass_quality_control / la_pass_quality_control / atrium_pass_quality_control
(tests/golden/make_golden_qc.py).  GPU: ukbb_cc_stats against scipy labelling on random masks, and the product gates
(ukbb_cardiac_b200/qc.py, device statistics) against the same golden verdicts.
"""
import os

import numpy as np
import pytest
from scipy import ndimage

from oracle import qc_oracle as qo

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "qc_reference.npz"))
SA = sorted(k[len("sa_verdict_"):] for k in GOLD.files if k.startswith("sa_verdict_"))
LA = sorted(k[len("la_verdict_"):] for k in GOLD.files if k.startswith("la_verdict_"))
AT = sorted(k[len("at_verdict_"):] for k in GOLD.files if k.startswith("at_verdict_"))
AO = sorted(k[len("ao_verdict_"):] for k in GOLD.files if k.startswith("ao_verdict_"))


def test_golden_covers_both_verdicts():
    for names, pre in ((SA, "sa"), (LA, "la"), (AT, "at"), (AO, "ao")):
        verdicts = {bool(GOLD["%s_verdict_%s" % (pre, n)]) for n in names}
        assert verdicts == {True, False}, pre


@pytest.mark.parametrize("i", range(6))
def test_oracle_cc_helpers_match_reference(i):
    m = GOLD["cc_in_%d" % i]
    np.testing.assert_array_equal(qo.get_largest_cc(m).astype(np.uint8), GOLD["cc_largest_%d" % i])
    np.testing.assert_array_equal(qo.remove_small_cc(m).astype(np.uint8), GOLD["cc_clean_%d" % i])


def test_oracle_gates_match_reference(capsys):
    for n in SA:
        assert qo.sa_pass_quality_control(GOLD["sa_" + n]) == bool(GOLD["sa_verdict_" + n]), n
    for n in LA:
        assert qo.la_pass_quality_control(GOLD["la_" + n]) == bool(GOLD["la_verdict_" + n]), n
    for n in AT:
        assert qo.atrium_pass_quality_control(GOLD["at_" + n], {'LA': 1, 'RA': 2}) == bool(GOLD["at_verdict_" + n]), n
    for n in AO:
        assert qo.aorta_pass_quality_control(GOLD["ao_img_" + n], GOLD["ao_seg_" + n]) == bool(GOLD["ao_verdict_" + n]), n


# ------------------------------------------------------------------------------------------ GPU
def _scipy_stats(mask, conn, thres):
    structure = ndimage.generate_binary_structure(2, conn)
    cc, n = ndimage.label(mask, structure=structure)
    areas = np.bincount(cc.ravel(), minlength=n + 1)[1:]
    return [int(mask.sum()), n, int((areas > thres).sum()), int(areas.max()) if n else 0, int(areas[areas >= thres].sum())]


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(5, 40, 36), (3, 208, 192), (2, 176, 224), (4, 17, 255)])
@pytest.mark.parametrize("conn", [1, 2])
def test_cc_stats_matches_scipy(shape, conn):
    import torch
    from ukbb_cardiac_b200 import qc
    rng = np.random.default_rng(shape[1] + conn)
    n, y, x = shape
    lab = np.zeros(shape, np.uint8)
    for i in range(n):
        f = ndimage.gaussian_filter(rng.normal(size=(y, x)), 1.0 + i)
        lab[i] = np.digitize(f, np.quantile(f, [0.5, 0.7, 0.85, 0.95]))         # classes 0..4, blobs of all sizes incl. single pixels
    lab[0, ::2, ::2] = 3                                                         # checkerboard: thousands of one-pixel components (conn 1) ...
    lab[1 % n, :, :] = 2                                                         # ... and one component covering a whole slice
    spiral = np.zeros((y, x), np.uint8)                                          # a long thin spiral: worst case for label propagation
    for k in range(0, min(x, y) // 2 - 1, 2):
        spiral[k, k:x - k] = 1; spiral[k:y - k, x - k - 1] = 1; spiral[y - k - 1, k:x - k] = 1; spiral[k + 2:y - k, k] = 1
        spiral[k + 1, k] = 0
    lab[n - 1] = spiral * 4
    classes = [1, 2, 3, 4]
    st = qc.cc_stats(torch.from_numpy(lab).cuda(), classes, connectivity=conn, thres=10)
    for i in range(n):
        for j, c in enumerate(classes):
            ref = _scipy_stats(lab[i] == c, conn, 10)
            got = [int(v) for v in st[i, j, [0, 1, 2, 3, 5]]]
            assert got == ref, (i, c, got, ref)
            if ref[3]:
                root = int(st[i, j, 4])
                assert lab[i].ravel()[root] == c                                 # the reported pixel belongs to a largest component
                cc, _ = ndimage.label(lab[i] == c, structure=ndimage.generate_binary_structure(2, conn))
                assert (cc == cc.ravel()[root]).sum() == ref[3]


@pytest.mark.gpu
def test_product_gates_match_reference(capsys):
    from ukbb_cardiac_b200 import qc
    for n in SA:
        assert qc.sa_pass_quality_control(GOLD["sa_" + n]) == bool(GOLD["sa_verdict_" + n]), n
    for n in LA:
        assert qc.la_pass_quality_control(GOLD["la_" + n]) == bool(GOLD["la_verdict_" + n]), n
    for n in AT:
        assert qc.atrium_pass_quality_control(GOLD["at_" + n], {'LA': 1, 'RA': 2}) == bool(GOLD["at_verdict_" + n]), n
    for n in AO:
        assert qc.aorta_pass_quality_control(GOLD["ao_img_" + n], GOLD["ao_seg_" + n]) == bool(GOLD["ao_verdict_" + n]), n
    out = capsys.readouterr().out
    assert "There is missing segmentation between the slices." in out and "abrupt change of area at time frame" in out


@pytest.mark.gpu
def test_gates_on_engine_output(tmp_path):
    """The gates run on what the engine produces: a label NIfTI written by the deploy path and the in-memory label volume give the
    same verdict as the oracle on the same labels."""
    from ukbb_cardiac_b200 import nifti, qc, synth
    from ukbb_cardiac_b200.fcn import FCNEngine
    w = synth.make_weights(0, 4)
    vol = synth.make_stack(2, (96, 112, 8, 3))
    with FCNEngine(w) as eng:
        lab, _, _ = eng.segment_volume(vol)
    ed = np.asfortranarray(lab[:, :, :, 0])
    p = str(tmp_path / "seg_sa_ED.nii.gz")
    nifti.save(nifti.Nifti1Image(ed, np.eye(4)), p, dtype=np.float64, label_data=True)
    v_file, v_mem = qc.sa_pass_quality_control(p), qc.sa_pass_quality_control(ed)
    assert v_file == v_mem == qo.sa_pass_quality_control(ed)
