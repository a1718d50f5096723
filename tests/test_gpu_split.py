"""SURVEY 8(e), single subject over G GPUs: the slices of a sequence are independent once the percentile thresholds of the whole
sequence are known, so SplitEngine gives contiguous (z, t) blocks to G engines.  Labels, thresholds and class counts must be
IDENTICAL to one FCNEngine.segment_volume call.  [0, 0] / [0, 0, 0] put every engine on GPU 0 (runs on a one-GPU box: exercises the
block split, ukbb_fcn_rescale with given thresholds and the reassembly); [0, 1] needs two GPUs."""
import numpy as np
import pytest
import torch

from ukbb_cardiac_b200 import synth
from ukbb_cardiac_b200.fcn import FCNEngine, SplitEngine

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["fp32", "fp16x2"])
@pytest.mark.parametrize("devices", [[0, 0], [0, 0, 0], [0, 1]])
def test_split_sequence_is_identical_to_one_engine(devices, mode):
    if max(devices) >= torch.cuda.device_count():
        pytest.skip("needs %d GPUs" % (max(devices) + 1))
    w = synth.make_weights(0, 4)
    vol = synth.make_stack(2, (96, 112, 3, 7))               # 21 slices: blocks of 11 + 10 / 7 + 7 + 7
    with FCNEngine(w, mode=mode) as one:
        lab1, (vl1, vh1), c1 = one.segment_volume(vol)
    with SplitEngine(w, devices, mode=mode) as many:
        lab2, (vl2, vh2), c2 = many.segment_volume(vol)
    assert (vl1, vh1) == (vl2, vh2)
    assert lab2.shape == lab1.shape and lab2.flags.f_contiguous
    assert (lab1 == lab2).all()
    assert (c1 == c2).all()


def test_split_more_engines_than_slices():
    w = synth.make_weights(0, 2)
    vol = synth.make_stack(4, (50, 43, 1, 2))                # 2 slices over 3 engines: one block is empty
    with FCNEngine(w) as one:
        lab1, thr1, c1 = one.segment_volume(vol)
    with SplitEngine(w, [0, 0, 0]) as many:
        lab2, thr2, c2 = many.segment_volume(vol)
    assert thr1 == thr2 and (lab1 == lab2).all() and (c1 == c2).all()
