"""Diagnostics for the tensor-core path: per-layer errors + dumps for offline analysis."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_bf16 import bf16_round, layer_reference, LAYER_LEVEL
from ukbb_cardiac_b200 import synth, weights as W
from ukbb_cardiac_b200.fcn import FCNEngine

w = synth.make_weights(0, 4)
MODE = sys.argv[1] if len(sys.argv) > 1 else "bf16"
from test_gpu_bf16 import TDT
eng = FCNEngine(w, mode=MODE)
dump = {}
for li in range(1, 20):
    sp = W.layer_table(4)[li]
    lvl = LAYER_LEVEL[li]
    lvl_in = lvl - 1 if sp.stride == 2 else lvl
    n, H, Wd = 3, 32 >> lvl_in, 48 >> lvl_in
    x = bf16_round(np.random.default_rng(li).normal(size=(n, H, Wd, sp.cin)), TDT[MODE])
    try:
        out = eng.debug_conv(li, torch.from_numpy(x).to(TDT[MODE]).cuda(), lvl).float().cpu().numpy()
    except Exception as e:
        print("layer %d %s: EXCEPTION %s" % (li, sp.role, e)); break
    ref = layer_reference(w, li, x, TDT[MODE])
    err = np.abs(out - ref); tol = 2.0 ** -7 * np.abs(ref) + 2e-3
    print("layer %2d %-7s cin %3d cout %3d s%d k%d in %dx%d: max err %.4g, frac bad %.4g, |ref| max %.3g, out nz %.3f"
          % (li, sp.role, sp.cin, sp.cout, sp.stride, sp.ksize, H, Wd, err.max(), (err > tol).mean(), np.abs(ref).max(), (out != 0).mean()))
    if (err > tol).any() and len(dump) < 6:
        dump["in%d" % li] = x; dump["out%d" % li] = out; dump["ref%d" % li] = ref
if dump:
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "diag_bf16_dump.npz"), **dump)
eng.close()
