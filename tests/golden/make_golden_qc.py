"""Generates tests/golden/qc_reference.npz.  Run ONCE in the build container (`python tests/golden/make_golden_qc.py`): it imports
the reference's own ``common/image_utils.py`` and ``common/cardiac_utils.py`` from /root/reference -- with the modules that are
not installed here (tensorflow, nibabel, vtk, matplotlib, scikit-image) stubbed out; none of them is touched by get_largest_cc,
remove_small_cc, sa_pass_quality_control or la_pass_quality_control except ``nib.load``, which the stub serves from memory -- and
records the outputs of the REAL functions on seeded synthetic label maps.  atrium_pass_quality_control calls
``skimage.measure.label(..., connectivity=2)``; its stub here is scipy.ndimage.label with the 3-D face + edge structure, so those
verdicts pin the reference's control flow and thresholds but not scikit-image's labelling itself (stated in oracle/qc_oracle.py)."""
import importlib.util
import os
import sys
import types
import warnings

import numpy as np
from scipy import ndimage

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")
VOLUMES = {}


def import_reference():
    for name in ("tensorflow", "vtk", "vtk.util", "vtk.util.numpy_support", "matplotlib", "matplotlib.pyplot", "skimage", "skimage.measure"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["vtk.util"].numpy_support = sys.modules["vtk.util.numpy_support"]
    sys.modules["vtk"].util = sys.modules["vtk.util"]
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sk = sys.modules["skimage.measure"]
    sys.modules["skimage"].measure = sk

    def sk_label(mask, connectivity=None, return_num=False):
        cc, n = ndimage.label(mask, structure=ndimage.generate_binary_structure(mask.ndim, connectivity or mask.ndim))
        return (cc, n) if return_num else cc
    sk.label = sk_label
    nib = types.ModuleType("nibabel")

    class _Img:
        def __init__(self, a):
            self.dataobj = a
    nib.load = lambda name: _Img(VOLUMES[name])
    sys.modules["nibabel"] = nib

    def load(modname, path):
        spec = importlib.util.spec_from_file_location(modname, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
        return mod
    for pkg in ("ukbb_cardiac", "ukbb_cardiac.common"):
        sys.modules.setdefault(pkg, types.ModuleType(pkg))
    iu = load("ukbb_cardiac.common.image_utils", "/root/reference/common/image_utils.py")
    cu = load("ukbb_cardiac.common.cardiac_utils", "/root/reference/common/cardiac_utils.py")
    return iu, cu


def disc(X, Y, cx, cy, r):
    xs, ys = np.meshgrid(np.arange(X), np.arange(Y), indexing="ij")
    return (xs - cx) ** 2 + (ys - cy) ** 2 < r * r


def sa_volume(rng, X=56, Y=64, Z=10, missing=(), tiny_rv=False, specks=True, first=1, last=8):
    seg = np.zeros((X, Y, Z), dtype=np.int16)
    for z in range(Z):
        if z < first or z > last or z in missing:
            continue
        r = 7 + 3 * np.sin(np.pi * (z - first) / max(last - first, 1))
        cx, cy = X / 2 + rng.uniform(-2, 2), Y / 2 + rng.uniform(-2, 2)
        seg[:, :, z][disc(X, Y, cx, cy, r + 4)] = 2
        seg[:, :, z][disc(X, Y, cx, cy, r)] = 1
        rv = disc(X, Y, cx - 4, cy + r + 9, 1.6 if tiny_rv else 6) & (seg[:, :, z] == 0)
        seg[:, :, z][rv] = 3
        if specks:                                  # isolated small components of every class
            for l in (1, 2, 3):
                for _ in range(3):
                    x, y = rng.integers(2, X - 3), rng.integers(2, 12)
                    seg[x:x + rng.integers(1, 3), y:y + rng.integers(1, 3), z] = l
    return seg


def la_slice(rng, X=64, Y=56, drop=None, frag_myo=False):
    seg = np.zeros((X, Y, 1), dtype=np.int16)
    s = seg[:, :, 0]
    s[disc(X, Y, 24, 24, 12)] = 2
    s[disc(X, Y, 24, 24, 8)] = 1
    s[disc(X, Y, 24, 44, 6)] = 3
    s[disc(X, Y, 46, 22, 7)] = 4
    s[disc(X, Y, 48, 42, 6)] = 5
    if frag_myo:                                    # myocardium only as specks below the threshold
        s[s == 2] = 0
        for k in range(6):
            s[4 + 4 * k:6 + 4 * k, 2:4] = 2
    if drop:
        s[s == drop] = 0
    for _ in range(4):
        x, y = rng.integers(2, X - 3), rng.integers(2, Y - 3)
        if s[x, y] == 0:
            s[x, y] = rng.integers(1, 6)
    return seg


def atrium_seq(rng, X=48, Y=40, T=12, vanish=None, fragment=None, jump=None, diag=False, amp=2.0):
    lab = np.zeros((X, Y, 1, T), dtype=np.int16)
    for t in range(T):
        r = 7 + amp * np.sin(2 * np.pi * t / T)
        if jump is not None and t == jump:
            r *= 1.7
        lab[:, :, 0, t][disc(X, Y, 16, 14, r)] = 1
        lab[:, :, 0, t][disc(X, Y, 32, 26, r * 0.8)] = 2
        if fragment is not None and t == fragment:
            lab[40:45, 2:6, 0, t] = 1                # a second component of 20 pixels
        if diag and t == 3:                          # touches the main blob only through a corner: one component for connectivity 2
            xs = np.argwhere(lab[:, :, 0, t] == 1)
            x, y = xs[np.argmax(xs[:, 0] + xs[:, 1])]
            lab[x + 1:x + 5, y + 1:y + 5, 0, t] = 1
    if vanish is not None:
        lab[:, :, 0, vanish][lab[:, :, 0, vanish] == 2] = 0
    return lab


def main():
    iu, cu = import_reference()
    rng = np.random.default_rng(11)
    out = {}
    # helpers on random blobs (ties of the largest area included)
    for i in range(6):
        m = (ndimage.gaussian_filter(rng.normal(size=(40, 36)), 2.0) > 0.05).astype(np.uint8)
        if i == 5:
            m[:] = 0
            m[2:6, 2:6] = 1; m[20:24, 20:24] = 1; m[30, 30] = 1         # two largest components of equal area
        out["cc_in_%d" % i] = m
        out["cc_largest_%d" % i] = iu.get_largest_cc(m).astype(np.uint8)
        out["cc_clean_%d" % i] = iu.remove_small_cc(m).astype(np.uint8)
    sa_cases = {"good": {}, "missing_slice": {"missing": (4,)}, "few_slices": {"first": 3, "last": 6}, "tiny_rv": {"tiny_rv": True},
                "no_specks": {"specks": False}, "short_stack": {"Z": 7, "first": 0, "last": 6}}
    for name, kw in sa_cases.items():
        seg = sa_volume(rng, **kw)
        if name == "few_slices":
            seg[seg == 3] = 3
        VOLUMES["sa_" + name] = seg
        out["sa_" + name] = seg
        out["sa_verdict_" + name] = np.array(bool(cu.sa_pass_quality_control("sa_" + name)))
    empty = sa_volume(rng); empty[empty == 3] = 0
    VOLUMES["sa_no_rv"] = empty; out["sa_no_rv"] = empty
    out["sa_verdict_no_rv"] = np.array(bool(cu.sa_pass_quality_control("sa_no_rv")))
    for name, kw in {"good": {}, "no_la": {"drop": 4}, "frag_myo": {"frag_myo": True}, "no_lv": {"drop": 1}}.items():
        seg = la_slice(rng, **kw)
        VOLUMES["la_" + name] = seg
        out["la_" + name] = seg
        out["la_verdict_" + name] = np.array(bool(cu.la_pass_quality_control("la_" + name)))
    for name, kw in {"good": {}, "vanish": {"vanish": 5}, "fragment": {"fragment": 7}, "jump": {"jump": 4}, "diag": {"diag": True}}.items():
        lab = atrium_seq(rng, **kw)
        out["at_" + name] = lab
        out["at_verdict_" + name] = np.array(bool(cu.atrium_pass_quality_control(lab, {'LA': 1, 'RA': 2})))
    for name, kw in {"good": {}, "vanish": {"vanish": 5}, "fragment": {"fragment": 7}, "jump": {"jump": 4}, "noisy": {}, "breathing": {}}.items():
        lab = atrium_seq(rng, amp=0.8, **kw)
        if name == "breathing":                      # slow growth: adjacent ratios fine, max / min >= 2
            lab[:] = 0
            for t in range(lab.shape[3]):
                r = 5.0 + 0.32 * (t if t <= 6 else 12 - t)
                lab[:, :, 0, t][disc(48, 40, 16, 14, r)] = 1
                lab[:, :, 0, t][disc(48, 40, 32, 26, 6)] = 2
        img = (200 + 30 * rng.standard_normal(lab.shape)).astype(np.float32)
        img[lab == 1] += 300
        img[lab == 2] += 250
        if name == "noisy":
            xs = np.argwhere(lab[:, :, 0, 6] == 2)[0]
            img[xs[0], xs[1], 0, 6] = 2500.0
        out["ao_seg_" + name] = lab
        out["ao_img_" + name] = img
        out["ao_verdict_" + name] = np.array(bool(cu.aorta_pass_quality_control(img, lab)))
    np.savez_compressed(os.path.join(HERE, "qc_reference.npz"), **out)
    for k in sorted(out):
        if "verdict" in k:
            print(k, bool(out[k]))


if __name__ == "__main__":
    main()
