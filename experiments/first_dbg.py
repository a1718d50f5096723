import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from ukbb_cardiac_b200 import synth
from ukbb_cardiac_b200.fcn import FCNEngine
from gpu_util import to_device_layout
w = synth.make_weights(0, 4)
img = np.random.default_rng(1).random((3, 64, 96, 1)).astype(np.float32)
dev = to_device_layout(img)
with FCNEngine(w, mode="bf16") as eng:
    l1, g1, _ = eng.forward(dev, want_logits=True)
    torch.cuda.synchronize()
    print("ok", eng.launch_count)
