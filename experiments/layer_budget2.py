"""Finer precision budget (round 2): per-tensor activation bits and per-layer weight bits.
Which stored tensors may stay single 16-bit (hi only) while the rest are hi+lo split?"""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fcn_oracle as fo, deploy_oracle as do
from ukbb_cardiac_b200 import synth
from experiments.layer_budget import q, report

def forward(img, w, ab, wb, tb):
    """ab[li]: bits of the stored output of conv li (0..19); wb[li]: bits of the weights of conv li; tb: bits of t_l"""
    n_class = w["conv2d_20/kernel"].shape[-1]
    tab = fo.layer_table(n_class)
    x = torch.from_numpy(np.transpose(img, (0, 3, 1, 2))).float()
    def fold(li):
        bn = fo.bn_name(li)
        g, b, m, v = (w[bn + "/" + k] for k in ("gamma", "beta", "moving_mean", "moving_variance"))
        sc = torch.from_numpy(g / np.sqrt(v + fo.BN_EPS)).float()
        sh = torch.from_numpy(b - m * sc.numpy()).float().view(1, -1, 1, 1)
        return sc, sh, (g, b, m, v)
    def conv(x, li, s):
        sc, sh, bnp = fold(li)
        k = torch.from_numpy(w[fo.conv_name(li) + "/kernel"]).float()
        if li >= 13 or li == 0:
            k = q(k * sc.view(1, 1, 1, -1), wb[li])
            return q(torch.relu(fo.conv2d_same(x, k.numpy(), s) + sh), ab[li])
        k = q(k, wb[li])
        return q(fo.bn_relu(fo.conv2d_same(x, k.numpy(), s), *bnp), ab[li])
    li = 0; lv = []
    for l in range(5):
        for b in range(fo.N_BLOCK[l]):
            x = conv(x, li, tab[li][4]); li += 1
        lv.append(x)
    ups = []
    for l in range(5):
        ups.append((conv(lv[l], li, 1), l)); li += 1
    sc, sh, _ = fold(li)
    k = torch.from_numpy(w[fo.conv_name(li) + "/kernel"]).float() * sc.view(1, 1, 1, -1)
    acc = 0
    for y, l in ups:
        kl = q(k[:, :, 32 * l:32 * l + 32, :], wb[li])
        t = fo.conv2d_same(y, kl.numpy(), 1)
        if l > 0:
            t = fo.transpose_upsample2d(q(t, tb[l]), 2 ** l)
        acc = acc + t
    x = q(torch.relu(acc + sh), ab[li]); li += 1
    x = conv(x, li, 1); li += 1
    kk = w[fo.conv_name(li) + "/kernel"]
    lg = fo.conv2d_same(x, kk, 1) + torch.from_numpy(w[fo.conv_name(li) + "/bias"]).view(1, -1, 1, 1)
    return lg.permute(0, 2, 3, 1).numpy()

if __name__ == "__main__":
    torch.set_num_threads(8)
    HI = int(sys.argv[1]) if len(sys.argv) > 1 else 11     # bits of a single 16-bit value (11 fp16, 8 bf16)
    SP = 22 if HI == 11 else 16
    w = synth.make_weights(0, 4)
    vol = synth.make_stack(0)
    img = do.rescale_intensity(vol.copy(order="F"), (1, 99))
    fr = np.concatenate([np.transpose(img[:, :, 2:8:2, t], (2, 0, 1)) for t in (0, 12, 25, 37)]).astype(np.float32)[..., None]
    ref64 = fo.build_fcn(fr, w, torch.float64)
    pred = ref64.argmax(-1)
    names = [t[0] for t in fo.layer_table(4)]
    def run(name, a_over={}, w_over={}, t_over={}):
        ab = [SP] * 20; wb = [SP] * 20; tb = [SP] * 5
        for k, v in a_over.items(): ab[k] = v
        for k, v in w_over.items(): wb[k] = v
        for k, v in t_over.items(): tb[k] = v
        report(name, forward(fr, w, ab, wb, tb), ref64, pred)
    run("all split")
    for li in range(20):
        run("act out of %s (conv %d) single" % (names[li], li), a_over={li: HI})
    for l in range(1, 5):
        run("t_%d single" % l, t_over={l: HI})
    for li in range(20):
        run("weights of %s (conv %d) single" % (names[li], li), w_over={li: HI})
