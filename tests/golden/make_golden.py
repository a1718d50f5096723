"""Generates the committed golden fixtures.  Run ONCE in the build container
(`python tests/golden/make_golden.py`): it imports the reference's own
``common/image_utils.py`` from /root/reference (with ``tensorflow`` and ``nibabel``
stubbed out -- neither is installed and neither is touched by the functions used)
and records the outputs of the REAL ``rescale_intensity`` / ``np_categorical_dice``
on small seeded inputs.  /root/reference does not exist on the GPU box, so tests only
read the .npz files written here.

Also records logits of the float64 oracle on tiny inputs so that later edits of the
oracle cannot silently change its semantics (self-pinning, not reference-pinning:
TensorFlow is unavailable, see oracle/__init__.py).
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def import_reference_image_utils():
    sys.modules.setdefault("tensorflow", types.ModuleType("tensorflow"))
    sys.modules.setdefault("nibabel", types.ModuleType("nibabel"))
    spec = importlib.util.spec_from_file_location("ref_image_utils", "/root/reference/common/image_utils.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def rescale_cases():
    rng = np.random.default_rng(7)
    cases = {}
    # MR-like integer-valued 4-D stack (ties at the percentile ranks are likely)
    cases["int4d"] = np.rint(rng.gamma(2.0, 150.0, size=(24, 20, 3, 5))).astype(np.float32)
    # continuous values: the lerp between neighbours is not representable in float32
    cases["cont4d"] = (rng.normal(300.0, 120.0, size=(21, 19, 2, 4))).astype(np.float32)
    # negative values, 3-D (ED/ES volume branch, deploy_network.py:179)
    cases["neg3d"] = (rng.normal(0.0, 50.0, size=(18, 17, 4))).astype(np.float32)
    # tiny: percentile ranks fall between the first/last two samples
    cases["tiny"] = np.array([[[5.0, 1.0, 9.0]], [[3.0, 7.0, 2.5]]], dtype=np.float32).reshape(2, 3, 1)
    # heavy ties: most of the mass on two values
    t = np.full((16, 16, 2, 2), 100.0, dtype=np.float32)
    t[:3] = 7.0
    t[-1, -1] = 4000.0
    cases["ties"] = t
    # a sequence whose voxel count makes gamma >= 0.5 for the low percentile
    cases["odd"] = np.rint(rng.uniform(0, 3000, size=(13, 11, 3, 7))).astype(np.float32)
    return cases


def main():
    ref = import_reference_image_utils()
    out = {}
    for name, arr in rescale_cases().items():
        a = np.asfortranarray(arr.copy())
        vl, vh = np.percentile(a, (1, 99))
        res = ref.rescale_intensity(a, (1, 99))          # clips `a` in place
        out[name + "/input"] = arr
        out[name + "/vl_vh"] = np.array([vl, vh], dtype=np.float64)
        out[name + "/clipped"] = np.asarray(a)
        out[name + "/rescaled_f64"] = np.asarray(res, dtype=np.float64)
        out[name + "/rescaled_f32"] = np.asarray(res).astype(np.float32)   # deploy_network.py:106
        assert res.dtype == np.float64, res.dtype
    np.savez_compressed(os.path.join(HERE, "rescale_reference.npz"), **out)

    # np_categorical_dice of the reference on a fixed pair of label maps
    rng = np.random.default_rng(11)
    a = rng.integers(0, 4, size=(32, 32, 3))
    b = a.copy()
    flip = rng.random(a.shape) < 0.1
    b[flip] = rng.integers(0, 4, size=int(flip.sum()))
    dice = np.array([ref.np_categorical_dice(a, b, k) for k in range(4)], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "dice_reference.npz"), a=a.astype(np.uint8), b=b.astype(np.uint8), dice=dice)

    # oracle self-pin: float64 logits on a tiny input for each n_class
    import torch
    from oracle import fcn_oracle
    from ukbb_cardiac_b200 import synth
    pin = {}
    img = np.random.default_rng(3).random((2, 32, 48, 1)).astype(np.float32)
    pin["image"] = img
    for nc in (2, 3, 4, 6):
        w = synth.make_weights(0, nc)
        pin["logits%d" % nc] = fcn_oracle.build_fcn(img, w, torch.float64)
    np.savez_compressed(os.path.join(HERE, "oracle_selfpin.npz"), **pin)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
