"""Per-layer precision budget (CPU emulation, round 2): which layers of build_FCN need split
(hi+lo) operands for the tensor-core path to reach >= 99.9 % label agreement / Dice >= 0.999?
Rounding points mirror the kernels: every stored activation is rounded to `act` precision, every
weight to `w` precision, accumulation in fp32."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fcn_oracle as fo, deploy_oracle as do
from ukbb_cardiac_b200 import synth

def q(x, bits):
    """round to `bits` significant bits (None = keep fp32); 11 = fp16-like, 8 = bf16-like, 22 = fp16 hi+lo"""
    if bits is None:
        return x
    if bits == 11:
        return x.to(torch.float16).to(torch.float32)
    if bits == 8:
        return x.to(torch.bfloat16).to(torch.float32)
    if bits == 22:
        hi = x.to(torch.float16).to(torch.float32)
        lo = (x - hi).to(torch.float16).to(torch.float32)
        return hi + lo
    if bits == 16:
        hi = x.to(torch.bfloat16).to(torch.float32)
        lo = (x - hi).to(torch.bfloat16).to(torch.float32)
        return hi + lo
    m, e = torch.frexp(x)
    s = float(1 << bits)
    return torch.ldexp(torch.round(m * s) / s, e)

GROUPS = ["l0", "l1", "l2", "l3", "l4", "side", "fc0", "fc1"]

def group_of(li):
    if li < 2: return "l0"
    if li < 4: return "l1"
    if li < 7: return "l2"
    if li < 10: return "l3"
    if li < 13: return "l4"
    if li < 18: return "side"
    if li == 18: return "fc0"
    return "fc1"

def forward(img, w, prec):
    """prec: dict group -> (act_bits, w_bits); act_bits applies to the group's OUTPUT storage"""
    n_class = w["conv2d_20/kernel"].shape[-1]
    tab = fo.layer_table(n_class)
    x = torch.from_numpy(np.transpose(img, (0, 3, 1, 2))).float()
    def conv(x, li, s):
        ab, wb = prec[group_of(li)]
        bn = fo.bn_name(li)
        g, b, m, v = (w[bn + "/" + k] for k in ("gamma", "beta", "moving_mean", "moving_variance"))
        k = torch.from_numpy(w[fo.conv_name(li) + "/kernel"]).float()
        if li >= 13 or li == 0:   # head / side / conv0_0 kernels fold the BN scale into the weights before rounding
            sc = torch.from_numpy(g / np.sqrt(v + fo.BN_EPS)).float()
            k = q(k * sc.view(1, 1, 1, -1), wb)
            y = fo.conv2d_same(x, k.numpy(), s) + torch.from_numpy(b - m * sc.numpy()).float().view(1, -1, 1, 1)
            return q(torch.relu(y), ab)
        k = q(k, wb)
        y = fo.conv2d_same(x, k.numpy(), s)
        return q(fo.bn_relu(y, g, b, m, v), ab)
    li = 0; lv = []
    for l in range(5):
        for b in range(fo.N_BLOCK[l]):
            x = conv(x, li, tab[li][4]); li += 1
        lv.append(x)
    ups = []
    for l in range(5):
        ups.append((conv(lv[l], li, 1), l)); li += 1
    # fc0 commuted through the upsample: t_l = W_l s_l (stored at act precision of "fc0"), then upsample
    ab, wb = prec["fc0"]
    bn = fo.bn_name(li)
    g, b, m, v = (w[bn + "/" + k] for k in ("gamma", "beta", "moving_mean", "moving_variance"))
    sc = torch.from_numpy(g / np.sqrt(v + fo.BN_EPS)).float()
    k = torch.from_numpy(w[fo.conv_name(li) + "/kernel"]).float() * sc.view(1, 1, 1, -1)
    acc = 0
    for y, l in ups:
        kl = q(k[:, :, 32 * l:32 * l + 32, :], wb)
        t = fo.conv2d_same(y, kl.numpy(), 1)
        if l > 0:
            t = fo.transpose_upsample2d(q(t, ab), 2 ** l)
        acc = acc + t
    x = q(torch.relu(acc + torch.from_numpy(b - m * sc.numpy()).float().view(1, -1, 1, 1)), ab)
    li += 1
    x = conv(x, li, 1); li += 1
    kk = w[fo.conv_name(li) + "/kernel"]
    lg = fo.conv2d_same(x, kk, 1) + torch.from_numpy(w[fo.conv_name(li) + "/bias"]).view(1, -1, 1, 1)
    return lg.permute(0, 2, 3, 1).numpy()

def report(name, lg, ref, pred):
    p = lg.argmax(-1)
    d = [fo.categorical_dice(p, pred, k) for k in range(ref.shape[-1])]
    print("%-44s agree %.5f  flips %6d  rms err %.3e  min dice %.5f" % (
        name, (p == pred).mean(), (p != pred).sum(), np.sqrt(((lg - ref) ** 2).mean()) / ref.std(), min(d)), flush=True)

if __name__ == "__main__":
    torch.set_num_threads(8)
    w = synth.make_weights(0, 4)
    vol = synth.make_stack(0)
    img = do.rescale_intensity(vol.copy(order="F"), (1, 99))
    fr = np.concatenate([np.transpose(img[:, :, 2:8:2, t], (2, 0, 1)) for t in (0, 12, 25, 37)]).astype(np.float32)[..., None]
    exact = {g: (None, None) for g in GROUPS}
    ref = forward(fr, w, exact)
    ref64 = fo.build_fcn(fr, w, torch.float64)
    pred = ref64.argmax(-1)
    report("fp32 everywhere (vs f64)", ref, ref64, pred)
    lo = (11, 11)
    report("fp16/fp16 everywhere", forward(fr, w, {g: lo for g in GROUPS}), ref64, pred)
    report("split22/22 everywhere", forward(fr, w, {g: (22, 22) for g in GROUPS}), ref64, pred)
    report("bf16 split16/16 everywhere", forward(fr, w, {g: (16, 16) for g in GROUPS}), ref64, pred)
    for g in GROUPS:
        p = dict(exact); p[g] = lo
        report("only %s fp16/fp16" % g, forward(fr, w, p), ref64, pred)
    for g in GROUPS:
        p = {k: (22, 22) for k in GROUPS}; p[g] = lo
        report("all split, %s fp16/fp16" % g, forward(fr, w, p), ref64, pred)
    p = {k: (22, 22) for k in GROUPS}; p["l3"] = lo; p["l4"] = lo
    report("all split, l3+l4 fp16/fp16", forward(fr, w, p), ref64, pred)
    p = {k: (22, 22) for k in GROUPS}; p["l2"] = lo; p["l3"] = lo; p["l4"] = lo
    report("all split, l2+l3+l4 fp16/fp16", forward(fr, w, p), ref64, pred)
    report("act 22 / w 11 everywhere", forward(fr, w, {g: (22, 11) for g in GROUPS}), ref64, pred)
    report("act 11 / w 22 everywhere", forward(fr, w, {g: (11, 22) for g in GROUPS}), ref64, pred)
    report("13/13 everywhere", forward(fr, w, {g: (13, 13) for g in GROUPS}), ref64, pred)
    report("12/12 everywhere", forward(fr, w, {g: (12, 12) for g in GROUPS}), ref64, pred)
