"""Host-side logic of the deploy CLI (CPU: the device engine is replaced by a stub that
implements the same ``segment_volume`` contract with the oracle's preprocessing)."""
import io
import os

import numpy as np
import pytest

from oracle import deploy_oracle as do
from ukbb_cardiac_b200 import deploy, nifti, synth


class StubEngine:
    """labels = 1 inside a disc whose radius shrinks then grows with t (so ES is well defined)."""
    n_class = 4

    def __init__(self):
        self.calls = 0

    def segment_volume(self, image):
        self.calls += 1
        img = image if image.ndim == 4 else image.reshape(image.shape[0], image.shape[1], -1, 1)
        X, Y, Z, T = img.shape
        vl, vh = do.percentile_linear(image, 1), do.percentile_linear(image, 99)
        xs, ys = np.meshgrid(np.arange(X), np.arange(Y), indexing="ij")
        lab = np.zeros((X, Y, Z, T), np.uint8, order="F")
        for t in range(T):
            r = 3 + abs(t - 2) * 2
            lab[:, :, :, t] = (((xs - X // 2) ** 2 + (ys - Y // 2) ** 2) < r * r)[:, :, None]
        counts = np.zeros((T, Z, 4), np.int64)
        for t in range(T):
            for z in range(Z):
                counts[t, z] = np.bincount(lab[:, :, z, t].ravel(), minlength=4)
        return lab.reshape(image.shape, order="F"), (vl, vh), counts


def make_subject(root, name, shape=(24, 20, 2, 5), seq="sa", seed=0):
    d = os.path.join(root, name)
    os.makedirs(d, exist_ok=True)
    vol = synth.make_stack(seed, shape)
    img = nifti.Nifti1Image(vol, np.diag([1.8, 1.8, 10.0, 1.0]))
    img.header["pixdim"][4] = 0.03
    nifti.save(img, os.path.join(d, seq + ".nii.gz"))
    return vol


def test_flag_parsing_tf_style():
    f = deploy.parse_flags(["--seq_name", "la_4ch", "--data_dir=/x", "--model_path", "m/FCN", "--noprocess_seq", "--seg4",
                            "--save_seg=false", "--unknown_flag", "7"])
    assert (f.seq_name, f.data_dir, f.model_path) == ("la_4ch", "/x", "m/FCN")
    assert f.process_seq is False and f.seg4 is True and f.save_seg is False
    d = deploy.parse_flags([])
    assert (d.seq_name, d.data_dir, d.model_path, d.process_seq, d.save_seg, d.seg4) == ("sa", "ukbb_cardiac_demo", "", True, True, False)
    assert deploy.parse_flags(["--process_seq=True"]).process_seq is True
    with pytest.raises(SystemExit):
        deploy.parse_flags(["--seq_name", "ao"])
    assert deploy.seg_prefix(f) == "seg4" and deploy.seg_prefix(d) == "seg"


def test_es_rule_from_counts():
    counts = np.zeros((3, 2, 4), np.int64)
    counts[:, 0, 1] = [8, 4, 12]
    assert deploy.es_frame_from_counts(counts, "sa", False) == 1
    assert deploy.es_frame_from_counts(counts, "la_4ch", True) == 1
    assert deploy.es_frame_from_counts(counts, "la_2ch", False) == 2
    assert deploy.es_frame_from_counts(counts, "la_4ch", False) == 2


def test_sharding_is_deterministic_and_complete():
    items = ["s%03d" % i for i in range(11)]
    parts = [deploy.shard(items, r, 4) for r in range(4)]
    assert sorted(sum(parts, [])) == items
    assert parts[1] == ["s001", "s005", "s009"]
    assert deploy.shard(items, 0, 1) == items


def test_deploy_sequence_files_prints_and_resume(tmp_path):
    root = str(tmp_path)
    vol = make_subject(root, "1000001")
    make_subject(root, "1000002", seed=1)
    os.makedirs(os.path.join(root, "1000003"))                     # no image -> "Skip."
    flags = deploy.parse_flags(["--seq_name", "sa", "--data_dir", root, "--model_path", "unused"])
    eng = StubEngine()
    buf = io.StringIO()
    assert deploy.deploy(flags, engine=eng, out=buf) == 0
    text = buf.getvalue()
    assert text.startswith("Start deployment on the data set ...\n1000001\n  Reading ")
    assert "  Segmenting full sequence ...\n  Segmentation time = " in text
    assert "  ED frame = 0, ES frame = 2\n  Saving segmentation ...\n" in text
    assert "does not contain an image with file name sa.nii.gz. Skip." in text
    assert "Average segmentation time = " in text and "for processing 2 subjects" in text
    d = os.path.join(root, "1000001")
    seg = nifti.load(os.path.join(d, "seg_sa.nii.gz"))
    assert seg.get_data().dtype == np.float64 and seg.shape == vol.shape          # deploy_network.py:92
    assert seg.header["pixdim"][4] == np.float32(0.03)
    np.testing.assert_allclose(seg.affine, np.diag([1.8, 1.8, 10.0, 1.0]))
    ref_lab, (vl, vh), _ = StubEngine().segment_volume(vol)
    np.testing.assert_array_equal(seg.get_data(), ref_lab.astype(np.float64))
    for fr, k in (("ED", 0), ("ES", 2)):
        s = nifti.load(os.path.join(d, "seg_sa_%s.nii.gz" % fr)).get_data()
        np.testing.assert_array_equal(s, ref_lab[:, :, :, k].astype(np.float64))
        im = nifti.load(os.path.join(d, "sa_%s.nii.gz" % fr)).get_data()
        clipped = vol.copy(order="F")
        do.rescale_intensity(clipped, (1, 99))                                     # clips in place
        np.testing.assert_array_equal(im, clipped[:, :, :, k])                     # saved images are CLIPPED
    # resume: nothing is recomputed when seg_sa.nii.gz exists (deploy_network.py:62-67)
    buf2 = io.StringIO()
    deploy.deploy(flags, engine=eng, out=buf2)
    assert eng.calls == 2 and "for processing 0 subjects" in buf2.getvalue()
    # --nosave_seg writes nothing
    os.remove(os.path.join(d, "seg_sa.nii.gz")); os.remove(os.path.join(d, "seg_sa_ED.nii.gz"))
    deploy.deploy(deploy.parse_flags(["--data_dir", root, "--nosave_seg"]), engine=eng, out=io.StringIO())
    assert eng.calls == 3 and not os.path.exists(os.path.join(d, "seg_sa.nii.gz"))


def test_deploy_ed_es_branch_and_seg4_prefix(tmp_path):
    root = str(tmp_path)
    d = os.path.join(root, "2000001")
    os.makedirs(d)
    for fr, seed in (("ED", 0), ("ES", 1)):
        vol = synth.make_stack(seed, (24, 20, 3, 1))[:, :, :, 0]
        nifti.save(nifti.Nifti1Image(np.asfortranarray(vol), np.eye(4)), os.path.join(d, "la_4ch_%s.nii.gz" % fr))
    os.makedirs(os.path.join(root, "2000002"))
    flags = deploy.parse_flags(["--seq_name", "la_4ch", "--data_dir", root, "--noprocess_seq", "--seg4"])
    buf = io.StringIO()
    eng = StubEngine()
    deploy.deploy(flags, engine=eng, out=buf)
    text = buf.getvalue()
    assert "  Segmenting ED frame ...\n" in text and "  Segmenting ES frame ...\n" in text
    assert "does not contain an image with file name la_4ch_ED.nii.gz or la_4ch_ES.nii.gz. Skip." in text
    assert "Average segmentation time = " in text and "s per frame" in text
    for fr in ("ED", "ES"):
        s = nifti.load(os.path.join(d, "seg4_la_4ch_%s.nii.gz" % fr))
        assert s.get_data().dtype == np.int32 and s.shape == (24, 20, 3)
    assert eng.calls == 2


def test_two_rank_sharding_with_gloo(tmp_path):
    """World-size-2 run of the host-side sharding (gloo, CPU): the two ranks split the sorted
    subject list without overlap and the host-side gather of per-rank results is complete."""
    import torch.multiprocessing as mp
    root = str(tmp_path / "data")
    for i in range(5):
        make_subject(root, "30000%02d" % i, shape=(16, 16, 1, 3), seed=i)
    mp.spawn(_gloo_worker, args=(2, root, str(tmp_path / "rdzv")), nprocs=2, join=True)
    for i in range(5):
        assert os.path.exists(os.path.join(root, "30000%02d" % i, "seg_sa.nii.gz"))


def _gloo_worker(rank, world, root, rdzv):
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="file://" + rdzv, rank=rank, world_size=world)
    flags = deploy.parse_flags(["--data_dir", root, "--shard_index", str(rank), "--num_shards", str(world)])
    eng = StubEngine()
    deploy.deploy(flags, engine=eng, out=io.StringIO())
    mine = deploy.shard(sorted(os.listdir(root)), rank, world)
    assert eng.calls == len(mine)
    # host-side gather of what each rank processed (the only "collective" of the path)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    assert sorted(sum(gathered, [])) == sorted(os.listdir(root))
    assert not set(gathered[0]) & set(gathered[1])
    dist.barrier()
    dist.destroy_process_group()


def test_worker_devices_respect_callers_mask():
    """--gpus N maps worker r through the caller's CUDA_VISIBLE_DEVICES instead of overwriting it (round-1 advisor finding)."""
    assert deploy.worker_devices(3, env={}) == ["0", "1", "2"]
    assert deploy.worker_devices(2, env={"CUDA_VISIBLE_DEVICES": "4,5,6,7"}) == ["4", "5"]
    assert deploy.worker_devices(1, env={"CUDA_VISIBLE_DEVICES": " GPU-abc "}) == ["GPU-abc"]
    with pytest.raises(SystemExit):
        deploy.worker_devices(3, env={"CUDA_VISIBLE_DEVICES": "0,1"})


def test_run_workers_drains_pipes_concurrently(monkeypatch, capsys):
    """Every worker prints far more than a pipe holds (64 KB) before ANY of them may exit: with one reader per worker the run
    finishes; draining the pipes one after the other in rank order would deadlock here (the round-1 behaviour serialised the GPUs)."""
    import subprocess as sp
    import sys
    real_popen = sp.Popen
    script = ("import sys, os, time\n"
              "r = int(sys.argv[sys.argv.index('--shard_index') + 1])\n"
              "d = os.environ['UKBB_TEST_RENDEZVOUS']\n"
              "sys.stdout.write(('worker %d line\\n' % r) * 20000); sys.stdout.flush()\n"       # ~280 KB each
              "open(os.path.join(d, 'done%d' % r), 'w').close()\n"
              "t0 = time.time()\n"
              "while not all(os.path.exists(os.path.join(d, 'done%d' % k)) for k in range(3)):\n"
              "    assert time.time() - t0 < 60\n"
              "    time.sleep(0.01)\n")

    def fake_popen(cmd, **kw):
        assert kw["env"]["CUDA_VISIBLE_DEVICES"] == cmd[cmd.index("--shard_index") + 1]
        return real_popen([sys.executable, "-c", script] + cmd[3:], **kw)

    import tempfile
    with tempfile.TemporaryDirectory() as d:
        monkeypatch.setenv("UKBB_TEST_RENDEZVOUS", d)
        monkeypatch.delenv("CUDA_VISIBLE_DEVICES", raising=False)
        monkeypatch.setattr(deploy.subprocess, "Popen", fake_popen)
        flags = deploy.parse_flags(["--gpus", "3"])
        assert deploy.run_workers(flags, ["--gpus", "3"]) == 0
    out = capsys.readouterr().out.splitlines()
    assert len(out) == 60000
    assert out[0] == "worker 0 line" and out[20000] == "worker 1 line" and out[-1] == "worker 2 line"     # replayed in rank order


def test_default_mode_is_the_compliant_one():
    assert deploy.parse_flags([]).mode == "fp16x2"
    assert deploy.parse_flags(["--mode", "bf16"]).mode == "bf16"
    with pytest.raises(SystemExit):
        deploy.parse_flags(["--mode", "int8"])


def test_split_blocks_partition():
    """SplitEngine's block partition: contiguous, covering, sizes within one of each other, empty blocks allowed."""
    from ukbb_cardiac_b200.fcn import split_blocks
    for n in (0, 1, 2, 21, 500, 501):
        for g in (1, 2, 3, 8):
            bl = split_blocks(n, g)
            assert len(bl) == g and bl[0][0] == 0 and bl[-1][1] == n
            assert all(bl[i][1] == bl[i + 1][0] for i in range(g - 1))
            sizes = [b - a for a, b in bl]
            assert min(sizes) >= 0 and max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
