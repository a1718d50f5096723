"""CPU emulation: which storage precision of activations / weights reaches the north_star
agreement bar on the synthetic fixtures?  (rounding after every layer, fp32 accumulate)"""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fcn_oracle as fo, deploy_oracle as do
from ukbb_cardiac_b200 import synth

def rnd(x, dt):
    return x if dt is None else x.to(dt).to(torch.float32)

def forward(img, w, act_dt, w_dt, head_dt=None):
    head_dt = head_dt or act_dt
    n_class = w["conv2d_20/kernel"].shape[-1]
    tab = fo.layer_table(n_class)
    x = torch.from_numpy(np.transpose(img, (0, 3, 1, 2))).float()
    def conv(x, li, s, dt):
        k = rnd(torch.from_numpy(w[fo.conv_name(li) + "/kernel"]), w_dt).numpy()
        y = fo.conv2d_same(x, k, s)
        bn = fo.bn_name(li)
        return rnd(fo.bn_relu(y, w[bn+"/gamma"], w[bn+"/beta"], w[bn+"/moving_mean"], w[bn+"/moving_variance"]), dt)
    li = 0; lv = []
    for l in range(5):
        for b in range(fo.N_BLOCK[l]):
            x = conv(x, li, tab[li][4], act_dt); li += 1
        lv.append(x)
    ups = []
    for l in range(5):
        y = conv(lv[l], li, 1, act_dt); li += 1
        ups.append(y if l == 0 else rnd(fo.transpose_upsample2d(y, 2 ** l), head_dt))
    x = torch.cat(ups, 1)
    x = conv(x, li, 1, head_dt); li += 1
    x = conv(x, li, 1, head_dt); li += 1
    k = torch.from_numpy(w[fo.conv_name(li) + "/kernel"]).numpy()
    lg = fo.conv2d_same(x, k, 1) + torch.from_numpy(w[fo.conv_name(li) + "/bias"]).view(1, -1, 1, 1)
    return lg.permute(0, 2, 3, 1).numpy()

if __name__ == "__main__":
    w = synth.make_weights(0, 4)
    vol = synth.make_stack(0)
    img = do.rescale_intensity(vol.copy(order="F"), (1, 99))
    fr = np.concatenate([np.transpose(img[:, :, 3:6, t], (2, 0, 1)) for t in (0, 20)]).astype(np.float32)[..., None]
    ref = forward(fr, w, None, None)
    pred = ref.argmax(-1)
    srt = np.sort(ref, -1); gap = srt[..., -1] - srt[..., -2]
    print("logit std", ref.std(), "gap median", np.median(gap), "frac gap<0.05:", (gap < 0.05).mean())
    for name, a, wd in [("bf16/bf16", torch.bfloat16, torch.bfloat16), ("fp16/fp16", torch.float16, torch.float16),
                        ("fp16 act / bf16 w", torch.float16, torch.bfloat16), ("bf16 act / fp32 w", torch.bfloat16, None),
                        ("fp32 act / bf16 w", None, torch.bfloat16)]:
        lg = forward(fr, w, a, wd)
        p = lg.argmax(-1)
        print("%-20s agree %.5f  rel err %.4g  dice %s" % (name, (p == pred).mean(), np.abs(lg - ref).max() / np.abs(ref).max(),
              ["%.4f" % fo.categorical_dice(p, pred, k) for k in range(4)]))
