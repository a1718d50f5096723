"""conv0_0 variants (UKBB_NO_FIRST: own FP32 kernel, UKBB_FIRST_FP32: fused, FP32 CUDA cores, default: fused, tensor pipe hi/lo split):
logit error against the float64 oracle and pairwise differences."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import fcn_oracle as fo
from ukbb_cardiac_b200 import synth
from ukbb_cardiac_b200.fcn import FCNEngine
from gpu_util import to_device_layout, from_device_logits, from_device_labels
w = synth.make_weights(0, 4)
img = np.random.default_rng(7).random((3, 64, 96, 1)).astype(np.float32)
dev = to_device_layout(img)
ref = fo.build_fcn(img, w, torch.float64)
pred = np.argmax(ref, -1)
for mode in ("bf16", "fp16"):
    out = {}
    for name, env in (("separate", {"UKBB_NO_FIRST": "1"}), ("fused_fp32", {"UKBB_FIRST_FP32": "1"}), ("fused_tc", {})):
        for k in ("UKBB_NO_FIRST", "UKBB_FIRST_FP32"):
            os.environ.pop(k, None)
        os.environ.update(env)
        with FCNEngine(w, mode=mode) as eng:
            l, g, _ = eng.forward(dev, want_logits=True)
            torch.cuda.synchronize()
            out[name] = (from_device_logits(g), from_device_labels(l))
    for name, (g, l) in out.items():
        e = np.abs(g - ref)
        print("%s %-10s vs float64: max rel %.4f  rms rel %.5f  label agreement %.5f" % (mode, name, e.max() / np.abs(ref).max(),
              np.sqrt((e ** 2).mean()) / np.sqrt((ref ** 2).mean()), (l == pred).mean()))
    a, b = out["fused_tc"][0], out["fused_fp32"][0]
    d = np.abs(a - b)
    print("%s fused_tc vs fused_fp32: max rel %.4f  rms rel %.5f  labels equal %.5f; separate vs fused_fp32 max %.2g" % (
        mode, d.max() / np.abs(b).max(), np.sqrt((d ** 2).mean()) / np.sqrt((b ** 2).mean()), (out["fused_tc"][1] == out["fused_fp32"][1]).mean(),
        np.abs(out["separate"][0] - b).max()))
