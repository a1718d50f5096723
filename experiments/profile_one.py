"""One warm-up subject + one profiled subject (device-resident SA volume) for ncu launch lists."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ukbb_cardiac_b200 import synth
from ukbb_cardiac_b200.fcn import FCNEngine
mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
X, Y, Z, T = synth.SA_SHAPE
eng = FCNEngine(synth.make_weights(0, 4), mode=mode)
vol = torch.from_numpy(synth.make_stack(0).reshape(-1, order="F").copy()).cuda()
for i in range(reps):
    l0 = eng.launch_count
    padded, vlvh, (xp, yp) = eng.preprocess(vol, Z * T, X, Y)
    labels, _, _ = eng.forward(padded, xp, yp, X, Y)
    torch.cuda.synchronize()
    print("rep", i, "launches", eng.launch_count - l0)
