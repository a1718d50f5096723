"""Which kernel is off?  Intermediate tensors of one forward over SA frames (20 slices, 192x208) against the float64 oracle features."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import deploy_oracle as do, fcn_oracle as fo
from ukbb_cardiac_b200 import synth
from ukbb_cardiac_b200.fcn import FCNEngine
from gpu_util import to_device_layout
w = synth.make_weights(0, 4)
vol = synth.make_stack(0)
img = do.rescale_intensity(vol.copy(order="F"), (1, 99))
fr = np.concatenate([np.transpose(img[:, :, :, t], (2, 0, 1)) for t in (0, 20)]).astype(np.float32)[..., None]
if len(sys.argv) > 2 and sys.argv[2] == "rand":
    fr = np.random.default_rng(0).random(fr.shape).astype(np.float32)
_, feats = fo.build_fcn(fr, w, torch.float64, return_features=True)
tensors = {(1, 0): "enc0_1", (0, 1): "enc1_0", (1, 1): "enc1_1", (1, 2): "enc2_1", (0, 2): "enc2_2", (1, 3): "enc3_1", (0, 3): "enc3_2", (1, 4): "enc4_1", (0, 4): "enc4_2"}
for mode in sys.argv[1].split(","):
    with FCNEngine(w, mode=mode) as eng:
        eng.forward(to_device_layout(fr)); torch.cuda.synchronize()
        for (which, level), name in tensors.items():
            ref = np.transpose(feats[name], (0, 2, 1, 3))
            got = eng.debug_read(which, level, ref.shape).cpu().numpy()
            err = np.abs(got - ref)
            bad = err > 1e-3 * np.abs(ref).max()
            print("%s %-7s rel err %.3g  bad %.4g%%  first bad %s" % (mode, name, err.max() / np.abs(ref).max(), 100 * bad.mean(),
                  tuple(int(v) for v in np.argwhere(bad)[0]) if bad.any() else None), flush=True)
            if name == "enc0_1" and bad.any():
                idx = np.argwhere(bad)
                print("   bad slices", np.unique(idx[:, 0])[:20], "rows", np.unique(idx[:, 1])[:40], "cols", np.unique(idx[:, 2])[:60], "ch", np.unique(idx[:, 3]))
