"""End-to-end drop-in test on the GPU: checkpoint files in the reference's TF layout ->
common/deploy_network.py CLI -> label NIfTI files, compared with the CPU restatement of the
reference loop (oracle/deploy_oracle.py) on the same synthetic subjects."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import deploy_oracle as do
from ukbb_cardiac_b200 import nifti, synth, tf_bundle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seq,shape,n_class,mode", [("sa", (40, 52, 3, 4), 4, "fp32"), ("la_2ch", (50, 43, 1, 5), 2, "fp32"),
                                                    ("la_4ch", (50, 43, 1, 5), 3, "fp16"), ("sa", (40, 52, 3, 4), 4, None),
                                                    ("la_4ch", (50, 43, 1, 5), 3, None)])
def test_cli_process_seq(tmp_path, seq, shape, n_class, mode):
    data = tmp_path / "data"
    w = synth.make_weights(0, n_class)
    tf_bundle.write_bundle(str(tmp_path / "model" / ("FCN_" + seq)), synth.with_optimizer_slots(w))
    vols = {}
    NS = 4                                          # enough subjects for the decode / device / encode pipeline to overlap
    for i in range(NS):
        d = data / ("100000%d" % i)
        os.makedirs(d)
        vols[i] = synth.make_stack(10 + i, shape)
        img = nifti.Nifti1Image(vols[i], np.diag([1.8, 1.8, 10.0, 1.0]))
        if i == 1:                                  # a file from a foreign writer: plain single-member gzip
            import gzip
            nifti.save(img, str(d / (seq + ".nii")))
            open(str(d / (seq + ".nii.gz")), "wb").write(gzip.compress(open(str(d / (seq + ".nii")), "rb").read(), 6))
            os.remove(str(d / (seq + ".nii")))
        else:
            nifti.save(img, str(d / (seq + ".nii.gz")))
    cmd = [sys.executable, os.path.join(ROOT, "common", "deploy_network.py"), "--seq_name", seq, "--data_dir", str(data),
           "--model_path", str(tmp_path / "model" / ("FCN_" + seq))] + (["--mode", mode] if mode else [])      # None = the default mode (fp16x2)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Start deployment on the data set ..." in r.stdout and "for processing %d subjects" % NS in r.stdout
    # the per-subject blocks come out in subject order with the reference's lines
    names = [ln for ln in r.stdout.splitlines() if ln.startswith("100000")]
    assert names == ["100000%d" % i for i in range(NS)]
    for i in range(NS):
        d = data / ("100000%d" % i)
        pred_ref, clipped = do.deploy_sequence(vols[i].copy(order="F"), do.make_runner(w))
        seg = nifti.load(str(d / ("seg_%s.nii.gz" % seq))).get_data()
        assert seg.dtype == np.float64 and seg.shape == shape
        agree = (seg == pred_ref).mean()
        assert agree >= {"fp32": 1 - 2e-5, "fp16": 0.996, None: 0.999}[mode], agree
        es = do.es_frame(seg, seq)
        assert "ED frame = 0, ES frame = %d" % es in r.stdout
        np.testing.assert_array_equal(nifti.load(str(d / ("seg_%s_ES.nii.gz" % seq))).get_data(), seg[:, :, :, es])
        np.testing.assert_array_equal(nifti.load(str(d / ("%s_ED.nii.gz" % seq))).get_data(), clipped[:, :, :, 0])
    # second run: everything is skipped (resume rule)
    r2 = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r2.returncode == 0 and "for processing 0 subjects" in r2.stdout


def test_cli_aortic(tmp_path):
    """common/deploy_network_ao.py drop-in: checkpoint bundle with the UNet-LSTM variable names -> CLI -> seg_ao.nii.gz (int32), compared
    with the CPU restatement of deploy_network_ao.py:83-193 (at the reference's fixed 256 x 256 network input)."""
    from oracle import ao_oracle as ao
    w = synth.make_ao_weights(0)
    tf_bundle.write_bundle(str(tmp_path / "model" / "UNet-LSTM_ao.ckpt-20000"), w)
    data = tmp_path / "data"
    vols = {}
    for i in range(2):
        d = data / ("200000%d" % i)
        os.makedirs(d)
        vols[i] = synth.make_ao_stack(i, (60, 52, 1, 10))
        img = nifti.Nifti1Image(vols[i], np.diag([1.6, 1.6, 6.0, 1.0]))
        img.header["pixdim"][4] = 0.01
        nifti.save(img, str(d / "ao.nii.gz"))
    os.makedirs(data / "2000009")                                       # a subject without ao.nii.gz
    cmd = [sys.executable, os.path.join(ROOT, "common", "deploy_network_ao.py"), "--data_dir", str(data),
           "--model_path", str(tmp_path / "model" / "UNet-LSTM_ao.ckpt-20000")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Start evaluating on the test set ..." in r.stdout and "for processing 2 subjects" in r.stdout
    assert "does not contain an image with file name ao.nii.gz" in r.stdout
    for i in range(2):
        seg_img = nifti.load(str(data / ("200000%d" % i) / "seg_ao.nii.gz"))
        seg = seg_img.get_data()
        assert seg.dtype == np.int32 and seg.shape == vols[i].shape
        assert abs(float(seg_img.header["pixdim"][4]) - 0.01) < 1e-7
        pred_ref, _ = ao.deploy_sequence(vols[i], w)
        assert (seg == pred_ref).mean() >= 0.9995
    r2 = subprocess.run(cmd + ["--noprocess_seq"], capture_output=True, text=True, timeout=300)
    assert r2.returncode == 0 and "UNet-LSTM does not support frame-wise segmentation" in r2.stdout


def test_cli_native_int16_volume(tmp_path):
    """A short-axis file stored as int16 (round-1 advisor finding): the reference computes the percentiles and clips IN PLACE in the
    file's dtype, so the thresholds are truncated on assignment and the saved <seq>_ED/ES frames are int16; the drop-in reproduces
    both (host arithmetic of image_utils.py:70-77 for such files, network on the device)."""
    seq, shape = "sa", (40, 52, 3, 4)
    w = synth.make_weights(0, 4)
    tf_bundle.write_bundle(str(tmp_path / "model" / "FCN_sa"), w)
    d = tmp_path / "data" / "1000000"
    os.makedirs(d)
    vol = synth.make_stack(21, shape).astype(np.int16)
    nifti.save(nifti.Nifti1Image(np.asfortranarray(vol), np.diag([1.8, 1.8, 10.0, 1.0])), str(d / "sa.nii.gz"))
    cmd = [sys.executable, os.path.join(ROOT, "common", "deploy_network.py"), "--seq_name", seq, "--data_dir", str(tmp_path / "data"),
           "--model_path", str(tmp_path / "model" / "FCN_sa"), "--mode", "fp32"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    ref_in = np.array(vol, order="F")                                  # int16, clipped in place by the restated reference loop
    pred_ref, clipped = do.deploy_sequence(ref_in, do.make_runner(w))
    assert clipped.dtype == np.int16
    seg = nifti.load(str(d / "seg_sa.nii.gz")).get_data()
    assert (seg == pred_ref).mean() >= 1 - 2e-5
    ed = nifti.load(str(d / "sa_ED.nii.gz")).get_data()
    assert ed.dtype == np.int16
    np.testing.assert_array_equal(ed, clipped[:, :, :, 0])
