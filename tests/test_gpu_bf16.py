"""GPU parity tests, BF16 tensor-core mode (tcgen05 implicit-GEMM convolutions).

Tolerances (BASELINE.json north_star): >= 99.9 % per-pixel label agreement with the
reference restatement and Dice >= 0.999 per class.  Per-layer unit tests compare every
tensor-core conv instance with the float64 oracle evaluated on the SAME bf16-rounded inputs
and weights, so the only differences are FP32 accumulation order and the final bf16
rounding: |err| <= 2^-7 * |ref| + 2e-3.
"""
import numpy as np
import pytest
import torch

from oracle import deploy_oracle as do
from oracle import fcn_oracle as fo
from ukbb_cardiac_b200 import synth
from ukbb_cardiac_b200 import weights as W
from ukbb_cardiac_b200.fcn import FCNEngine

from gpu_util import from_device_labels, from_device_logits, to_device_layout

pytestmark = pytest.mark.gpu


TDT = {"bf16": torch.bfloat16, "fp16": torch.float16}


def SUB_BATCHES(n, cap=500):
    """forward_bf16 splits n slices into ceil(n / 500) equal sub-batches (conv_tc.cu)."""
    return -(-n // cap)


def bf16_round(a: np.ndarray, dt=torch.bfloat16) -> np.ndarray:
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dt).to(torch.float32).numpy()


def layer_reference(w, li, x_dev_layout: np.ndarray, dt=torch.bfloat16) -> np.ndarray:
    """float64 conv + BN + ReLU of layer li on a device-layout [N, Y, X, C] input."""
    sp = W.layer_table(4)[li]
    x_tf = np.transpose(x_dev_layout, (0, 2, 1, 3)).astype(np.float64)             # [N, X, Y, C]
    k = bf16_round(w[W.conv_name(li) + "/kernel"], dt).astype(np.float64)
    y = fo.conv2d_same(torch.from_numpy(np.transpose(x_tf, (0, 3, 1, 2))), k, sp.stride)
    bn = W.bn_name(li)
    g, b, m, v = (w[bn + "/" + s].astype(np.float64) for s in ("gamma", "beta", "moving_mean", "moving_variance"))
    sc = g / np.sqrt(v + 1e-3)
    y = torch.relu(y * torch.from_numpy(sc).view(1, -1, 1, 1) + torch.from_numpy(b - m * sc).view(1, -1, 1, 1))
    return np.transpose(y.numpy(), (0, 3, 2, 1))                                    # -> [N, Y, X, C]


LAYER_LEVEL = [0, 0, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4, 0, 1, 2, 3, 4, 0, 0]


@pytest.fixture(scope="module", params=["bf16", "fp16"])
def engine(request):
    w = synth.make_weights(0, 4)
    eng = FCNEngine(w, mode=request.param)
    yield eng, w, request.param
    eng.close()


@pytest.mark.parametrize("li", list(range(1, 20)))
def test_tc_layer(engine, li):
    eng, w, mode = engine
    dt = TDT[mode]
    sp = W.layer_table(4)[li]
    lvl = LAYER_LEVEL[li]
    lvl_in = lvl - 1 if sp.stride == 2 else lvl
    n, H, Wd = 3, 32 >> lvl_in, 48 >> lvl_in
    rng = np.random.default_rng(li)
    x = bf16_round(rng.normal(0.0, 1.0, size=(n, H, Wd, sp.cin)), dt)
    out = eng.debug_conv(li, torch.from_numpy(x).to(dt).cuda(), lvl).float().cpu().numpy()
    ref = layer_reference(w, li, x, dt)
    assert out.shape == ref.shape
    err = np.abs(out - ref)
    tol = (2.0 ** -7 if mode == "bf16" else 2.0 ** -10) * np.abs(ref) + (2e-3 if mode == "bf16" else 3e-4)
    assert (err <= tol).all(), "layer %d (%s): max err %g at %s, ref there %g; frac bad %g" % (
        li, sp.role, err.max(), np.unravel_index(err.argmax(), err.shape), ref.flat[err.argmax()], (err > tol).mean())


# Agreement floors on the RANDOM-INIT fixture.  A random network is the adversarial case for a
# 16-bit path: its top-2 logit gap has its mode at 0 (8 % of the pixels have a gap < 0.05 sigma),
# so a 1 % logit error flips a few per cent of the labels whatever kernel computes it (the CPU
# emulation experiments/precision_sim.py reproduces the figures below with plain torch ops).
# The head folds the BN scales into the 16-bit weights (a different, equally valid quantisation): on this
# fixture that moved BF16 from 97.7 % to 98.7 % and FP16 from 99.77 % to 99.64 % (experiments/agree_probe.py).
RANDOM_INIT_FLOOR = {"bf16": (0.97, 0.95), "fp16": (0.995, 0.99)}


@pytest.mark.parametrize("mode", ["bf16", "fp16"])
@pytest.mark.parametrize("n_class", [4, 2, 3, 6])
def test_forward_bf16_small(n_class, mode):
    w = synth.make_weights(0, n_class)
    img = np.random.default_rng(n_class).random((3, 64, 48, 1)).astype(np.float32)
    with FCNEngine(w, mode=mode) as eng:
        labels, logits, _ = eng.forward(to_device_layout(img), want_logits=True)
        torch.cuda.synchronize()
    ref = fo.build_fcn(img, w, torch.float64)
    lg = from_device_logits(logits)
    rel = np.abs(lg - ref).max() / np.abs(ref).max()
    assert rel < (0.05 if mode == "bf16" else 0.008), "%s logits rel err %g" % (mode, rel)
    pred = np.argmax(ref, -1)
    lab = from_device_labels(labels)
    agree = (lab == pred).mean()
    assert agree >= (0.98 if mode == "bf16" else 0.997), agree


@pytest.mark.parametrize("mode", ["bf16", "fp16"])
def test_forward_tc_sa_random_init(mode):
    """Synthetic SA frames through the RANDOM-INIT network: agreement / Dice floors per mode."""
    w = synth.make_weights(0, 4)
    vol = synth.make_stack(0)
    img = do.rescale_intensity(vol.copy(order="F"), (1, 99))
    fr = np.concatenate([np.transpose(img[:, :, :, t], (2, 0, 1)) for t in (0, 20)]).astype(np.float32)[..., None]
    with FCNEngine(w, mode=mode) as eng:
        labels, _, _ = eng.forward(to_device_layout(fr))
        torch.cuda.synchronize()
    _, pred = fo.session_run(fr, w)
    lab = from_device_labels(labels)
    agree = (lab == pred).mean()
    dice = [fo.categorical_dice(lab, pred, k) for k in range(4)]
    print("%s agreement %.5f dice %s" % (mode, agree, dice))
    assert agree >= RANDOM_INIT_FLOOR[mode][0], agree
    assert min(dice) >= RANDOM_INIT_FLOOR[mode][1], dice


@pytest.mark.parametrize("mode", ["bf16", "fp16"])
def test_segment_volume_tc_la(mode):
    w = synth.make_weights(0, 2)
    vol = synth.make_stack(5, (50, 43, 1, 6))
    pred_ref, _ = do.deploy_sequence(vol.copy(order="F"), do.make_runner(w))
    with FCNEngine(w, mode=mode) as eng:
        lab, (vl, vh), counts = eng.segment_volume(vol)
    assert vl == do.percentile_linear(vol, 1) and vh == do.percentile_linear(vol, 99)
    assert (lab == pred_ref).mean() >= (0.98 if mode == "bf16" else 0.997)
    assert counts.sum() == lab.size


def test_fused_and_unfused_paths_agree(monkeypatch):
    """The fused head / halo-reuse kernels and the stage-1 kernels (per-tap TMA, materialised
    concat) implement the same arithmetic: labels agree except at near-ties, logits within the
    16-bit rounding of the two extra intermediate tensors of the unfused path."""
    w = synth.make_weights(0, 4)
    img = np.random.default_rng(11).random((2, 64, 96, 1)).astype(np.float32)
    dev = to_device_layout(img)
    with FCNEngine(w, mode="fp16") as eng:
        l1, g1, _ = eng.forward(dev, want_logits=True)
        torch.cuda.synchronize()
    monkeypatch.setenv("UKBB_HEAD_GATHER", "1")          # head v1: 4-tap gather on CUDA cores
    with FCNEngine(w, mode="fp16") as eng:
        l3, g3, _ = eng.forward(dev, want_logits=True)
        torch.cuda.synchronize()
    monkeypatch.setenv("UKBB_NO_FUSED_HEAD", "1")
    monkeypatch.setenv("UKBB_NO_HALO", "1")
    with FCNEngine(w, mode="fp16") as eng:
        l2, g2, _ = eng.forward(dev, want_logits=True)
        torch.cuda.synchronize()
    for g, l, name in ((g1, l1, "tensor-core upsample head"), (g3, l3, "gather head")):
        rel = float((g - g2).abs().max() / g2.abs().max())
        assert rel < 5e-3, (name, rel)
        assert float((l == l2).float().mean()) >= 0.998, name


@pytest.mark.parametrize("mode", ["bf16", "fp16"])
@pytest.mark.parametrize("shape", [(3, 64, 96), (5, 48, 80), (130, 32, 48), (510, 32, 32)])
def test_side_kernel_bit_identical(monkeypatch, mode, shape):
    """side_tc_kernel (same_dim_l + fc0 column block of levels 1..4 chained in one launch, s_l kept
    on chip) rounds at the same two points as the two conv_tc launches per level it replaces, so
    logits and labels are bit-identical -- including partial last tiles (72 or 30 pixels at level 4)."""
    w = synth.make_weights(0, 4)
    img = np.random.default_rng(shape[0]).random(shape + (1,)).astype(np.float32)
    dev = to_device_layout(img)
    with FCNEngine(w, mode=mode) as eng:
        l1, g1, _ = eng.forward(dev, want_logits=True)
        torch.cuda.synchronize()
        n_side = eng.launch_count
    monkeypatch.setenv("UKBB_NO_SIDE", "1")
    with FCNEngine(w, mode=mode) as eng:
        l2, g2, _ = eng.forward(dev, want_logits=True)
        torch.cuda.synchronize()
        n_plain = eng.launch_count
    assert n_plain - n_side == 7 * SUB_BATCHES(shape[0])            # 8 launches became 1 per sub-batch
    assert torch.equal(g1, g2)
    assert torch.equal(l1, l2)


@pytest.mark.parametrize("mode", ["bf16", "fp16"])
@pytest.mark.parametrize("shape", [(3, 64, 96), (5, 48, 80), (2, 208, 192), (130, 32, 48), (501, 16, 32)])
def test_first_kernel_bit_identical(monkeypatch, mode, shape):
    """conv_first_kernel (conv0_0 on the CUDA cores building the halo patch of conv0_1 in shared
    memory, a0 never written to HBM) does the same FP32 FMA sequence and the same UMMAs as the two
    launches it replaces: logits and labels are bit-identical, including partial border tiles and
    the zero padding of a0 (not of the image) around the picture."""
    w = synth.make_weights(0, 4)
    img = np.random.default_rng(shape[1]).random(shape + (1,)).astype(np.float32)
    dev = to_device_layout(img)
    monkeypatch.setenv("UKBB_FIRST_FP32", "1")           # conv0_0 in FP32 on the CUDA cores (conv_first.cuh)
    with FCNEngine(w, mode=mode) as eng:
        l1, g1, _ = eng.forward(dev, want_logits=True)
        torch.cuda.synchronize()
        n_first = eng.launch_count
    monkeypatch.setenv("UKBB_NO_FIRST", "1")
    with FCNEngine(w, mode=mode) as eng:
        l2, g2, _ = eng.forward(dev, want_logits=True)
        torch.cuda.synchronize()
        n_plain = eng.launch_count
    assert n_plain - n_first == SUB_BATCHES(shape[0])                   # 2 launches became 1 per sub-batch
    assert torch.equal(g1, g2)
    assert torch.equal(l1, l2)


@pytest.mark.parametrize("mode", ["bf16", "fp16"])
@pytest.mark.parametrize("n_class,shape", [(4, (3, 64, 96)), (2, (5, 48, 80)), (3, (2, 208, 192)), (6, (9, 32, 48))])
def test_head_ts_matches_head_tc(monkeypatch, mode, n_class, shape):
    """head_ts_kernel keeps U_l / b0 / A0 / A2 in tensor memory (tcgen05.st + tcgen05.mma with a [tmem] A operand) and
    issues the same UMMAs in the same order on the same 16-bit values as head_tc_kernel, whose A operands live in
    shared memory; only the FP32 class-score epilogue is re-associated (relu(d + s) . w = max(d, -s) . w + s . w with the
    constant folded into the bias).  Logits agree to FP32 rounding, labels everywhere but at exact near-ties."""
    w = synth.make_weights(0, n_class)
    img = np.random.default_rng(shape[2]).random(shape + (1,)).astype(np.float32)
    dev = to_device_layout(img)
    with FCNEngine(w, mode=mode) as eng:
        l1, g1, _ = eng.forward(dev, want_logits=True)
        torch.cuda.synchronize()
    monkeypatch.setenv("UKBB_HEAD_V3", "1")
    with FCNEngine(w, mode=mode) as eng:
        l2, g2, _ = eng.forward(dev, want_logits=True)
        torch.cuda.synchronize()
    rel = float((g1 - g2).abs().max() / g2.abs().max())
    assert rel < 2e-6, rel
    assert float((l1 != l2).float().mean()) < 1e-4


@pytest.mark.parametrize("mode", ["bf16", "fp16"])
@pytest.mark.parametrize("shape", [(3, 64, 96), (5, 48, 80), (2, 208, 192), (130, 32, 48)])
def test_first_tc_kernel_matches_fp32_conv0(monkeypatch, mode, shape):
    """conv_first_tc_kernel runs conv0_0 on the tensor pipe with a hi/lo 16-bit split of the FP32 image and of the FP32
    weights (x.w ~ hi.w_hi + lo.w_hi + hi.w_lo, relative error ~2^-16), then rounds to 16 bit like the FP32 CUDA-core path.
    The two a0 tensors differ only where that 2^-16 error crosses a 16-bit rounding boundary, but a random network
    decorrelates the 16-bit rounding noise of every later layer from such flips (experiments/first_cmp.py: the two
    variants differ from each other by 0.86 % rms in BF16 and from the float64 oracle by 1.04 % / 1.05 %; label agreement
    with the oracle 99.50 % / 99.45 %), so the bound here is the noise floor of the mode, and the oracle comparison of
    test_forward_bf16_small covers the default (tensor-pipe) variant."""
    w = synth.make_weights(0, 4)
    img = np.random.default_rng(shape[1]).random(shape + (1,)).astype(np.float32)
    dev = to_device_layout(img)
    with FCNEngine(w, mode=mode) as eng:
        l1, g1, _ = eng.forward(dev, want_logits=True)
        torch.cuda.synchronize()
        n_first = eng.launch_count
    monkeypatch.setenv("UKBB_NO_FIRST", "1")
    with FCNEngine(w, mode=mode) as eng:
        l2, g2, _ = eng.forward(dev, want_logits=True)
        torch.cuda.synchronize()
        n_plain = eng.launch_count
    assert n_plain - n_first == SUB_BATCHES(shape[0])
    rel = float((g1 - g2).abs().max() / g2.abs().max())
    assert rel < (0.04 if mode == "bf16" else 0.006), rel
    assert float((l1 == l2).float().mean()) >= (0.985 if mode == "bf16" else 0.997)


@pytest.mark.parametrize("mode", ["bf16", "fp16"])
@pytest.mark.parametrize("shape", [(3, 64, 96), (5, 208, 192), (7, 48, 80), (1, 32, 48)])
def test_cluster_multicast_bit_identical(monkeypatch, mode, shape):
    """Levels 3 / 4 (conv_halo_kernel, streamed weights): clusters of two CTAs fetch every weight tile once and multicast it to
    both shared memories; the arithmetic per tile is unchanged, so logits and labels are bit-identical to the single-CTA
    launches -- including odd tile counts, where the second CTA of the last pair walks the weight stream on a dummy tile."""
    w = synth.make_weights(0, 4)
    img = np.random.default_rng(shape[0]).random(shape + (1,)).astype(np.float32)
    dev = to_device_layout(img)
    with FCNEngine(w, mode=mode) as eng:
        l1, g1, _ = eng.forward(dev, want_logits=True)
        torch.cuda.synchronize()
    monkeypatch.setenv("UKBB_NO_CLUSTER", "1")
    with FCNEngine(w, mode=mode) as eng:
        l2, g2, _ = eng.forward(dev, want_logits=True)
        torch.cuda.synchronize()
    assert torch.equal(g1, g2)
    assert torch.equal(l1, l2)
