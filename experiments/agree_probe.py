"""Label agreement / Dice of the tensor-core modes vs the float32 restatement on the random-init SA fixture,
for the head variants (env UKBB_HEAD_V2 = head_mma, default = head_tc) -- 4 frames x 10 slices."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import deploy_oracle as do, fcn_oracle as fo
from ukbb_cardiac_b200 import synth
from ukbb_cardiac_b200.fcn import FCNEngine
from gpu_util import from_device_labels, to_device_layout
w = synth.make_weights(0, 4)
vol = synth.make_stack(0)
img = do.rescale_intensity(vol.copy(order="F"), (1, 99))
fr = np.concatenate([np.transpose(img[:, :, :, t], (2, 0, 1)) for t in (0, 10, 20, 30)]).astype(np.float32)[..., None]
_, pred = fo.session_run(fr, w)
for env in ({}, {"UKBB_HEAD_V2": "1"}, {"UKBB_NO_GROUP": "1"}, {"UKBB_HEAD_V2": "1", "UKBB_NO_GROUP": "1"}):
    for k in ("UKBB_HEAD_V2", "UKBB_NO_GROUP"): os.environ.pop(k, None)
    os.environ.update(env)
    for mode in ("bf16", "fp16"):
        with FCNEngine(w, mode=mode) as eng:
            labels, _, _ = eng.forward(to_device_layout(fr)); torch.cuda.synchronize()
        lab = from_device_labels(labels)
        print(env, mode, "agreement %.5f" % (lab == pred).mean(), "dice", ["%.4f" % fo.categorical_dice(lab, pred, k) for k in range(4)], flush=True)
