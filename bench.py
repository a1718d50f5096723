#!/usr/bin/env python
"""Benchmark of the FCN deploy hot path (BASELINE.json metric: SA FCN 192x208 slices/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode fp16x3|fp16x2|bf16x3|fp16|bf16|fp32] [--workload c3|c1|c2|c4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the CPU restatement of the reference loop

Workloads (BASELINE.json configs; the default is the one the metric is quoted on):
    c3  batch of S synthetic short-axis subjects (192x208x10x50 = 500 slices each) per GPU per step      [default, S = 256]
    c1  ONE synthetic SA subject per step, synchronised every step (single-subject latency)
    c2  S pairs of long-axis sequences per step: la_2ch (2 classes) + la_4ch (3 classes), 210x171x1x50 -> padded 224x176
    c4  the per-subject segmentation stage with NIfTI in / NIfTI out (sa + la_2ch + la_4ch through the drop-in CLI's deploy()),
        files on local disk; reported in subjects/s (the host I/O stages are part of the measurement)
    c5  aortic UNet + bidirectional ConvLSTM (network_ao) on one synthetic 240x196x1x100 aortic cine per step (FP32 CUDA-core path)

One step = one pass of the hot path (percentile rescale + pad + build_FCN forward + argmax/crop) over the step's
sequences.  `value` times the device-resident path (inputs already in HBM); `e2e.value` times the public host-buffer
call (pinned host -> H2D -> compute -> D2H of the label volumes) over the same sequences.  One JSON line is printed by rank 0.
The default mode is fp16x2, the fastest tensor-core mode that meets the parity tolerance; `parity` in the line is measured on the
timed inputs against the float32 CPU restatement.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SA = (192, 208, 10, 50)
LA = (210, 171, 1, 50)
POOL = 8                               # distinct synthetic volumes cycled through a batch (8 x 80 MB > 126 MB L2)
DTYPE = {"bf16": "bf16", "fp16": "f16", "fp32": "f32", "fp16x3": "f16x3", "bf16x3": "bf16x3", "fp16x2": "f16+e4m3 (x2)"}


def flops_per_slice(h: int, w: int, n_class: int) -> float:
    """Algorithmic FLOPs of build_FCN on an h x w slice: 2*Hout*Wout*k^2*Cin*Cout per conv + 4-tap bilinear upsampling."""
    nf, nb = [16, 32, 64, 128, 256], [2, 2, 3, 3, 3]
    total, cin, hh, ww, sizes = 0.0, 1, h, w, []
    for l in range(5):
        for b in range(nb[l]):
            s = 2 if (l > 0 and b == 0) else 1
            hh, ww = -(-hh // s), -(-ww // s)
            total += 2.0 * hh * ww * 9 * cin * nf[l]
            cin = nf[l]
        sizes.append((hh, ww))
    for l in range(5):
        total += 2.0 * sizes[l][0] * sizes[l][1] * nf[l] * 32
    total += 2.0 * h * w * (160 * 64 + 64 * 64 + 64 * n_class) + 4 * 2.0 * 4 * h * w * 32
    return total


def head_flops_per_slice(h: int, w: int, n_class: int) -> float:
    """The fused head's share (head_ts_kernel): same_dim0 + 4-tap upsampling + fc0 + fc1 + class scores."""
    return 2.0 * h * w * (16 * 32 + 160 * 64 + 64 * 64 + 64 * n_class) + 4 * 2.0 * 4 * h * w * 32


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "source": "MEASURED_PEAKS.json"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def measured_traffic(mode: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, written by the profiling script from the
    `ncu --set full` capture of the same workload (experiments/ncu_traffic.py -> profiles/r2_head_traffic.json); None if absent."""
    p = os.path.join(ROOT, "profiles", "r2_head_traffic.json")
    if os.path.exists(p):
        d = json.load(open(p)).get(mode)
        if d:
            return d["bytes_per_launch"], d.get("source")
    return None, None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.stop_flag = threading.Event()
        self.rows = []

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([c.strip() for c in line.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_sample(vol: np.ndarray, n_class: int, threads: int):
    """The reference's loop (deploy_network.py:89-116) restated on CPU: global percentile rescale, then ONE forward per time
    frame with batch Z, float32 (PyTorch/oneDNN stands in for TensorFlow-CPU, which is not installable here).
    Returns (slices_per_s, seconds, n_slices, pred (X,Y,Z,T))."""
    import torch
    from oracle import deploy_oracle as do
    from ukbb_cardiac_b200 import synth
    torch.set_num_threads(threads)
    w = synth.make_weights(0, n_class)
    run = do.make_runner(w)
    run(np.zeros((1, 32, 32, 1), np.float32))            # warm oneDNN primitives
    t0 = time.perf_counter()
    pred, _ = do.deploy_sequence(vol.copy(order="F"), run)
    dt = time.perf_counter() - t0
    n = vol.shape[2] * vol.shape[3]
    return n / dt, dt, n, pred


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from ukbb_cardiac_b200 import synth
    threads = os.cpu_count() or 1
    if args.workload == "c5":
        import torch
        from oracle import ao_oracle as ao
        torch.set_num_threads(threads)
        w = synth.make_ao_weights(0)
        frames = 12
        vol = np.asfortranarray(synth.make_ao_stack(0)[..., :frames])
        secs = []
        for _ in range(args.steps):
            t0 = time.perf_counter()
            ao.deploy_sequence(vol, w)
            secs.append(time.perf_counter() - t0)
        value = frames * args.steps / sum(secs)
        sample = "%d-frame aortic cine (240x196 padded to 256x256, window 9) per step, reference loop restated with the UNet evaluated once per frame" % frames
        print(json.dumps({
            "impl": "reference", "metric": "aortic UNet-LSTM 240x196 frames/sec", "value": value, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": "c5: " + WORKLOAD_NAMES["c5"] + " (CPU sample)", "sample": sample},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample,
                             "note": "TF-CPU proxy (PyTorch/oneDNN float32 restatement); TensorFlow unavailable"},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return 0
    la = args.workload == "c2"
    shape = LA if la else (SA[0], SA[1], SA[2], args.ref_frames)
    vol = synth.make_stack(0, shape)
    cpu_reference_sample(synth.make_stack(0, shape[:3] + (1,)), 2 if la else 4, threads)
    secs = []
    for _ in range(args.steps):
        _, dt, n, _ = cpu_reference_sample(vol, 2 if la else 4, threads)
        secs.append(dt)
    n = shape[2] * shape[3]
    value = n * args.steps / sum(secs)
    sample = "%d frames x %d slices of one synthetic %s sequence per step, reference loop (one forward per frame, batch Z)" % (
        shape[3], shape[2], "LA" if la else "SA")
    unit = "slices/s"
    if args.workload == "c4":
        unit = "subjects/s"
        value = value / 600.0                             # one subject = 500 SA + 50 + 50 LA slices; the CPU arm times the network only
        sample += "; converted at 600 slices per subject, file I/O not included"
    line = {
        "impl": "reference", "metric": "SA FCN 192x208 slices/sec", "value": value, "unit": unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAMES[args.workload] + " (CPU sample)", "sample": sample},
        "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": "port", "sample": sample,
                         "note": "TF-CPU proxy (PyTorch/oneDNN float32 restatement); TensorFlow unavailable"},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


WORKLOAD_NAMES = {
    "c3": "batch of synthetic SA subjects (192x208x10x50, 500 slices each) per GPU per step",
    "c1": "one synthetic SA subject (192x208x10x50) per step, synchronised every step (latency)",
    "c2": "long-axis pairs la_2ch (2 classes) + la_4ch (3 classes), synthetic 210x171x1x50 sequences (padded 224x176)",
    "c4": "per-subject segmentation stage, NIfTI in / NIfTI out: sa + la_2ch + la_4ch .nii.gz per subject through deploy()",
    "c5": "aortic UNet + BiConvLSTM (network_ao, UNet-LSTM model, window 9, weight_R 5) on one synthetic 240x196x1x100 cine per step",
}


# ------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default=os.environ.get("UKBB_BENCH_MODE", "fp16x2"), choices=sorted(DTYPE))
    ap.add_argument("--workload", default=os.environ.get("UKBB_BENCH_WORKLOAD", "c3"), choices=sorted(WORKLOAD_NAMES))
    ap.add_argument("--subjects", type=int, default=None, help="sequences (c3: SA subjects, c2: LA pairs, c4: subjects) per GPU per step")
    ap.add_argument("--ref-frames", type=int, default=20, help="frames per step of the reference arm (20 frames = 200 slices, ~5 s of CPU work)")
    ap.add_argument("--cpu-frames", type=int, default=50, help="frames in the cpu_baseline / parity sample (0 = skip); 50 = one whole subject, ~14 s")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.workload == "c4":
        return run_c4(args)
    if args.workload == "c5":
        return run_c5(args)

    import torch
    import torch.distributed as dist
    from ukbb_cardiac_b200 import synth
    from ukbb_cardiac_b200.fcn import FCNEngine, pad16

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL announces its version on STDOUT when the first communicator is created; the contract is ONE JSON line there
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    # ---- the step's sequences: (engine, shape, pool index)
    wl = args.workload
    if wl == "c2":
        S = args.subjects if args.subjects else 64
        engines = {2: FCNEngine(synth.make_weights(0, 2), device=local, mode=args.mode),
                   3: FCNEngine(synth.make_weights(0, 3), device=local, mode=args.mode)}
        shape = LA
        plan = [(nc, s % POOL) for s in range(S) for nc in (2, 3)]
    else:
        S = 1 if wl == "c1" else (args.subjects if args.subjects else (2 if args.mode == "fp32" else 256))
        engines = {4: FCNEngine(synth.make_weights(0, 4), device=local, mode=args.mode)}
        shape = SA
        plan = [(4, s % POOL) for s in range(S)]
    X, Y, Z, T = shape
    nvox = X * Y * Z * T
    pool = min(POOL, S)
    host_pool = [torch.empty(nvox, dtype=torch.float32, pin_memory=True) for _ in range(pool)]
    for i, hp in enumerate(host_pool):
        hp.numpy()[:] = synth.make_stack(100 * rank + i, shape).reshape(-1, order="F")
    dev_pool = [hp.to(dev) for hp in host_pool]
    host_labels = [torch.empty(nvox, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    host_counts = [torch.empty((Z * T, 4), dtype=torch.int64, pin_memory=True) for _ in range(2)]
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    fwd_ev = []
    # device-resident path: the percentile rescale of sequence s + 1 runs on a side stream while the forward of sequence s runs
    # on the main stream (two padded buffers), as ukbb_fcn_segment_host does internally for host buffers
    pre_stream = torch.cuda.Stream(dev)
    (x2, _), (y2, _) = pad16(X), pad16(Y)
    pad = [torch.empty((Z * T, y2, x2), dtype=torch.float32, device=dev) for _ in range(2)]
    vv = [torch.empty(2, dtype=torch.float64, device=dev) for _ in range(2)]
    ev_pre = [torch.cuda.Event() for _ in range(2)]
    ev_fwd = [torch.cuda.Event() for _ in range(2)]
    NS = len(plan)

    def step_device(record=False):
        xp = yp = 0
        pre_stream.wait_stream(stream)
        for s in range(NS + 1):
            if s < NS:
                slot = s & 1
                nc, pi = plan[s]
                with torch.cuda.stream(pre_stream):
                    pre_stream.wait_event(ev_fwd[slot])              # the forward that read pad[slot] two sequences ago
                    _, _, (xp, yp) = engines[nc].preprocess(dev_pool[pi], Z * T, X, Y, out=pad[slot], vlvh=vv[slot])
                    ev_pre[slot].record(pre_stream)
            if s >= 1:
                slot = (s - 1) & 1
                stream.wait_event(ev_pre[slot])
                if record:
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                engines[plan[s - 1][0]].forward(pad[slot], xp, yp, X, Y)
                if record:
                    e1.record(stream)
                    fwd_ev.append((e0, e1))
                ev_fwd[slot].record(stream)
        if wl == "c1":
            stream.synchronize()                                     # latency: the host sees every subject's result before the next

    def step_e2e():
        for s, (nc, pi) in enumerate(plan):
            engines[nc].segment_host_async(host_pool[pi], shape, host_labels[s & 1], None, host_counts[s & 1])
        for e in engines.values():
            e.join()
        if wl == "c1":
            stream.synchronize()

    def timed(fn, steps, **kw):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn(**kw)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    warm = max(args.warmup, 3)
    steps = args.steps if wl != "c1" else max(args.steps, 20)
    for _ in range(warm):
        step_device()
    l0 = sum(e.launch_count for e in engines.values())
    sampler = ClockSampler(local)
    sampler.start()
    ms_dev = timed(step_device, steps)
    launches = sum(e.launch_count for e in engines.values()) - l0
    for _ in range(warm):
        step_e2e()
    for e in engines.values():
        e.sync()
    ms_e2e = timed(step_e2e, steps)
    sampler.stop_flag.set()
    sampler.join(timeout=3)
    # ---- roofline pass (NOT part of `value`): per-forward and per-head-launch CUDA events on the launching stream; the events around
    # the head launch suppress the programmatic-dependent-launch overlap on both sides of it, which is why this is a separate pass
    for e in engines.values():
        e.kernel_timer(True)
        e.kernel_timer_read()
    barrier()
    step_device(record=True)
    barrier()
    head_ms = head_n = 0
    for e in engines.values():
        ms, n = e.kernel_timer_read()
        head_ms += ms; head_n += n
        e.kernel_timer(False)

    slices = NS * Z * T * steps * world
    value = slices / (ms_dev * 1e-3)
    e2e_value = slices / (ms_e2e * 1e-3)
    fwd_ms = sum(a.elapsed_time(b) for a, b in fwd_ev) / max(len(fwd_ev), 1)        # per forward (one sequence)
    peaks = measured_peaks()
    ncs = sorted(engines)
    flop_slice = sum(flops_per_slice(y2, x2, nc) for nc in ncs) / len(ncs)
    head_flop_slice = sum(head_flops_per_slice(y2, x2, nc) for nc in ncs) / len(ncs)
    fwd_tf = flop_slice * Z * T / (fwd_ms * 1e-3) / 1e12
    peak_tf = peaks["bf16_sustained"]
    whole = {"achieved": fwd_tf, "frac": fwd_tf / peak_tf, "frac_of_burst": fwd_tf / peaks["bf16_burst"],
             "frac_of_nominal_2250": fwd_tf / 2250.0, "algorithmic_flop_per_slice": flop_slice, "avg_forward_ms_per_sequence": fwd_ms,
             "note": "algorithmic FLOPs (1x); the x3 modes execute 3 tensor-core products per algorithmic product" if "x3" in args.mode else
             "algorithmic FLOPs (1x); the x2 mode executes 2 tensor-core instructions per algorithmic product" if "x2" in args.mode else None}
    if head_n > 0 and args.mode != "fp32":
        head_avg_ms = head_ms / head_n
        achieved_tf = head_flop_slice * Z * T / (head_avg_ms * 1e-3) / 1e12
        traffic, traffic_src = measured_traffic(args.mode) if wl in ("c3", "c1") else (None, None)
        roofline = {"bound": "tensor", "kernel": "head_ts_kernel (same_dim0 + upsample + fc0 + fc1 + class scores + softmax/argmax/crop; 1 launch per sequence)",
                    "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                    "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read + write)", "traffic_source": traffic_src,
                    "algorithmic_flop_per_launch": head_flop_slice * Z * T, "avg_launch_ms": head_avg_ms, "launches_timed": head_n,
                    "timing": "CUDA events on the launching stream around every launch, in a separate pass after the timed region "
                              "(the events break the dependent-launch overlap around the kernel)",
                    "peak_source": peaks["source"] + " (sustained bf16: the kernel runs inside a long step)",
                    "frac_of_burst": achieved_tf / peaks["bf16_burst"], "frac_of_nominal_2250": achieved_tf / 2250.0,
                    "share_of_forward": head_avg_ms / fwd_ms, "whole_forward": whole}
    else:
        roofline = dict(whole, bound="tensor", kernel="build_FCN forward (all conv launches of one sequence)", peak=peak_tf, unit="TFLOP/s",
                        traffic=None, peak_source=peaks["source"] + " (sustained bf16, of measured)")

    line = None
    if rank == 0:
        cpu = parity = None
        if args.cpu_frames > 0 and world == 1:               # the CPU baseline and the parity sample are measured at N = 1 only
            threads = os.cpu_count() or 1
            nc0, pi0 = plan[0]
            frames = min(args.cpu_frames, T)
            vol0 = host_pool[pi0].numpy().reshape(shape, order="F")
            # GPU labels of the same timed input through the public host-buffer call
            lab, _, _ = engines[nc0].segment_volume(vol0)
            v, dt, n, pred = cpu_reference_sample(vol0 if frames == T else vol0[..., :frames], nc0, threads)
            cpu = {"value": v, "unit": "slices/s", "cores": threads, "kind": "port",
                   "sample": "%d frames x %d slices of timed input 0 (%d classes), reference loop (one forward per frame, batch Z), %.1f s"
                             % (frames, Z, nc0, dt),
                   "note": "TF-CPU proxy (PyTorch/oneDNN float32 restatement); TensorFlow unavailable"}
            if frames == T:                                  # same percentiles only when the whole sequence went through the oracle
                from oracle import fcn_oracle as fo
                dice = [fo.categorical_dice(lab, pred, k) for k in range(nc0)]
                parity = {"agreement": float((lab == pred).mean()), "min_dice": float(min(dice)), "dice": dice,
                          "sample": "timed input 0, all %d slices, labels of the host-buffer call vs the float32 CPU restatement" % (Z * T),
                          "tolerance": ">= 0.999 agreement, Dice >= 0.999 per class (north_star)"}
        line = {
            "metric": "SA FCN 192x208 slices/sec" if wl != "c2" else "LA FCN 224x176 (padded 210x171) slices/sec",
            "value": value, "unit": "slices/s", "n_gpus": world,
            "steps": steps, "warmup": warm, "ms_per_step": ms_dev / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": DTYPE[args.mode], "data": "synthetic",
            "config": {"workload": "%s: %s [%d sequences per GPU per step]" % (wl, WORKLOAD_NAMES[wl], NS),
                       "sequences_per_gpu": NS, "global_sequences": NS * world, "mode": args.mode, "n_class": ncs,
                       "l2_policy": "inputs larger than L2: %d distinct volumes of %.0f MB cycled" % (pool, nvox * 4 / 1e6) if pool * nvox * 4 > 126e6
                                    else "L2 flushed by the activation traffic of every forward (>= 1 GB per sequence)",
                       "pipeline": "rescale of sequence s+1 on a side stream overlaps the forward of sequence s",
                       "weights": "random-init (seed 0), reference TF checkpoint layout", "parallelism": "dp%d" % world},
            "subjects_per_s": value / (Z * T) if wl != "c2" else None,
            "ms_per_sequence": ms_dev / steps / NS,
            "e2e": {"value": e2e_value, "unit": "slices/s", "h2d_bytes_per_step": NS * nvox * 4,
                    "d2h_bytes_per_step": NS * (nvox + Z * T * 4 * 8), "ms_per_step": ms_e2e / steps},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "parity": parity,
            "clocks": sampler.summary(),
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def ao_flops_per_frame(size: int = 256, f0: int = 16, nh: int = 16, n_class: int = 3, window: int = 9) -> float:
    """Algorithmic FLOPs per OUTPUT frame of the UNet-LSTM deploy loop with the UNet evaluated once per frame:
    UNet (2 x n^2 x 9 x cin x cout per conv, transposed convs at 2.25 taps per output pixel) + `window` bidirectional ConvLSTM steps
    (3x3 conv of concat([x, h]) -> 4 nh channels) + the output 1x1 conv."""
    nf = [f0 * 2 ** i for i in range(5)]
    total, cin, n = 0.0, 1, size
    for l in range(5):
        if l > 0:
            n //= 2
        total += 2.0 * n * n * 9 * cin * nf[l] + 2.0 * n * n * 9 * nf[l] * nf[l]
        cin = nf[l]
    for l in range(3, -1, -1):
        n = size >> l
        total += 2.0 * n * n * 2.25 * nf[l + 1] * nf[l] + 2.0 * n * n * 9 * (2 * nf[l]) * nf[l] + 2.0 * n * n * 9 * nf[l] * nf[l]
    lstm = window * 2 * (2.0 * size * size * 9 * (f0 + nh) * 4 * nh) + window * 2.0 * size * size * 2 * nh * n_class
    return total + lstm


def run_c5(args):
    """One synthetic aortic cine (240 x 196 x 1 x 100) per step.  `value`: frames/s with the z-scored, padded cine resident in HBM
    (ukbb_ao_segment); `e2e`: the public host call (AortaEngine.segment_sequence: z-score on the host as the reference does, pad,
    H2D, device call, D2H of the labels).  Ranks run independent replicas of the same workload."""
    import torch
    import torch.distributed as dist
    from ukbb_cardiac_b200 import aorta, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    w = synth.make_ao_weights(0)
    eng = aorta.AortaEngine(w, device=local)
    shape = synth.AO_SHAPE
    X, Y, Z, T = shape
    vol = synth.make_ao_stack(rank, shape)
    size = aorta.IMAGE_SIZE
    x_pre, y_pre = (size - X) // 2, (size - Y) // 2
    img = np.pad(aorta.normalise_intensity(vol, 10.0), ((x_pre, size - X - x_pre), (y_pre, size - Y - y_pre), (0, 0), (0, 0)), 'constant')
    d_img = torch.from_numpy(np.ascontiguousarray(np.transpose(img[:, :, 0, :], (2, 1, 0)), dtype=np.float32)).to(dev)
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    step_dev = lambda: eng.segment_frames(d_img, x_pre, y_pre, X, Y)
    step_e2e = lambda: eng.segment_sequence(vol)
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_dev()
    l0 = eng.launch_count
    sampler = ClockSampler(local)
    sampler.start()
    ms_dev = timed(step_dev, args.steps)
    launches = eng.launch_count - l0
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    sampler.stop_flag.set()
    sampler.join(timeout=3)
    value = T * args.steps * world / (ms_dev * 1e-3)
    e2e_value = T * args.steps * world / (ms_e2e * 1e-3)
    flop = ao_flops_per_frame(size)
    achieved = flop * T / (ms_dev / args.steps * 1e-3) / 1e12
    if rank == 0:
        cpu = parity = None
        if args.cpu_frames > 0 and world == 1:
            import torch as _t
            from oracle import ao_oracle as ao
            _t.set_num_threads(os.cpu_count() or 1)
            frames = 12                                                  # bounded sample: a 12-frame cine at full size
            sub = np.asfortranarray(vol[..., :frames])
            t0 = time.perf_counter()
            pred_ref, _ = ao.deploy_sequence(sub, w)
            dt = time.perf_counter() - t0
            pred, _ = eng.segment_sequence(sub)
            cpu = {"value": frames / dt, "unit": "frames/s", "cores": os.cpu_count() or 1, "kind": "port",
                   "sample": "%d-frame cine of the same size (240x196, window 9), reference loop restated with the UNet evaluated once per frame, %.1f s" % (frames, dt),
                   "note": "TF-CPU proxy (PyTorch/oneDNN float32 restatement); TensorFlow unavailable"}
            parity = {"agreement": float((pred == pred_ref).mean()), "sample": "the same %d-frame cine through the host call vs the float32 CPU restatement" % frames}
        line = {
            "metric": "aortic UNet-LSTM 240x196 frames/sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "c5: " + WORKLOAD_NAMES["c5"], "frames": T, "padded": "%dx%d" % (size, size), "mode": "fp32 (CUDA cores)",
                       "weights": "random-init (seed 0), UNet-LSTM checkpoint variable names", "parallelism": "dp%d (replicas)" % world,
                       "l2_policy": "activations of one cine (> 10 GB) exceed L2 many times over"},
            "sequences_per_s": value / T,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": T * size * size * 4, "d2h_bytes_per_step": T * X * Y, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp32 FMA pipe (CUDA cores): this path has no tensor-core kernels yet", "achieved": achieved, "unit": "TFLOP/s",
                         "peak": 75.0, "frac": achieved / 75.0, "peak_source": "B200 FP32 CUDA-core ceiling (~75 TFLOP/s, SURVEY 8d)", "traffic": None,
                         "algorithmic_flop_per_frame": flop,
                         "note": "algorithmic FLOPs count the ConvLSTM conv over concat([x, h]); the device evaluates conv(x) once per frame (declared shortcut)"},
            "cpu_baseline": cpu, "parity": parity, "clocks": sampler.summary(),
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def c4_parity(subject_dir: str, seqs):
    """Label volumes written by deploy() for one subject (seg_sa / seg_la_2ch / seg4_la_4ch .nii.gz) against the CPU restatement of
    the reference loop on the same input files."""
    from oracle import fcn_oracle as fo
    from ukbb_cardiac_b200 import nifti
    out = {"tolerance": ">= 0.999 agreement, Dice >= 0.999 per class (north_star)", "sample": "subject 0, label files vs the float32 CPU restatement"}
    worst_a, worst_d = 1.0, 1.0
    for name, _, nc in seqs:
        vol = nifti.load(os.path.join(subject_dir, name + ".nii.gz")).get_data()
        seg = [f for f in ("seg_%s.nii.gz" % name, "seg4_%s.nii.gz" % name) if os.path.exists(os.path.join(subject_dir, f))][0]
        lab = nifti.load(os.path.join(subject_dir, seg)).get_data()
        _, _, _, pred = cpu_reference_sample(np.asarray(vol), nc, os.cpu_count() or 1)
        a = float((np.asarray(lab) == pred).mean())
        d = min(fo.categorical_dice(np.asarray(lab), pred, k) for k in range(nc))
        out[name] = {"agreement": a, "min_dice": float(d)}
        worst_a, worst_d = min(worst_a, a), min(worst_d, float(d))
    out["agreement"], out["min_dice"] = worst_a, worst_d
    return out


def run_c4(args):
    """NIfTI in / NIfTI out: S synthetic subjects on local disk (sa + la_2ch + la_4ch .nii.gz each), three deploy() passes per step as
    demo_pipeline.py:63-64,89-96 runs deploy_network.py three times; outputs are deleted between steps (skip-if-exists would
    otherwise turn the next step into a no-op).  Ranks shard the subject list (subject i -> rank i % N)."""
    import io
    import shutil
    import tempfile
    import torch
    import torch.distributed as dist
    from ukbb_cardiac_b200 import deploy, nifti, synth
    from ukbb_cardiac_b200.fcn import FCNEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    S = args.subjects if args.subjects else 8
    root = tempfile.mkdtemp(prefix="ukbb_c4_r%d_" % rank, dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    seqs = (("sa", SA, 4), ("la_2ch", LA, 2), ("la_4ch", LA, 3))
    in_bytes = 0
    for s in range(S):
        d = os.path.join(root, "subj%03d" % s)
        os.makedirs(d)
        for name, shape, _ in seqs:
            img = nifti.Nifti1Image(synth.make_stack(1000 * rank + 10 * s + len(name), shape), np.diag([1.8, 1.8, 10.0, 1.0]))
            nifti.save(img, os.path.join(d, name + ".nii.gz"))
            in_bytes += os.path.getsize(os.path.join(d, name + ".nii.gz"))
    engines = {nc: FCNEngine(synth.make_weights(0, nc), device=local, mode=args.mode) for _, _, nc in seqs}

    def clean():
        for s in range(S):
            d = os.path.join(root, "subj%03d" % s)
            for f in os.listdir(d):
                if f not in ("sa.nii.gz", "la_2ch.nii.gz", "la_4ch.nii.gz"):
                    os.remove(os.path.join(d, f))

    stages = {}

    def step():
        for name, _, nc in seqs:
            flags = deploy.parse_flags(["--seq_name", name, "--data_dir", root, "--mode", args.mode])
            sink = io.StringIO()
            deploy.deploy(flags, engine=engines[nc], out=sink, stage_times=stages)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    warm = max(args.warmup, 1)
    parity = None
    try:
        for _ in range(warm):
            step(); clean()
        stages.clear()
        l0 = sum(e.launch_count for e in engines.values())
        sampler = ClockSampler(local)
        sampler.start()
        secs = 0.0
        out_bytes = 0
        for _ in range(args.steps):
            barrier()
            t0 = time.perf_counter()
            step()
            barrier()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            secs += dt
            out_bytes = sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(root) for f in fs) - in_bytes
            if _ == args.steps - 1 and rank == 0 and world == 1 and args.cpu_frames > 0:
                parity = c4_parity(os.path.join(root, "subj000"), seqs)       # the label FILES of subject 0 vs the CPU restatement
            clean()
        sampler.stop_flag.set()
        sampler.join(timeout=3)
        launches = sum(e.launch_count for e in engines.values()) - l0
    finally:
        shutil.rmtree(root, ignore_errors=True)
    value = S * args.steps * world / secs
    if rank == 0:
        total = sum(stages.values()) or 1.0
        line = {
            "metric": "per-subject segmentation stage, NIfTI in / NIfTI out", "value": value, "unit": "subjects/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": DTYPE[args.mode], "data": "synthetic",
            "config": {"workload": "c4: %s [%d subjects per GPU per step, files under %s]" % (WORKLOAD_NAMES["c4"], S, os.path.dirname(root)),
                       "subjects_per_gpu": S, "mode": args.mode, "slices_per_subject": 600, "label_dtype": "float64 (reference format)",
                       "timing": "wall clock around the three deploy() passes incl. .nii.gz decode / encode, max over ranks", "parallelism": "dp%d" % world},
            "slices_per_s": value * 600,
            "e2e": {"value": value, "unit": "subjects/s", "h2d_bytes_per_step": S * 4 * (np.prod(SA) + 2 * np.prod(LA)).item(),
                    "d2h_bytes_per_step": S * (np.prod(SA) + 2 * np.prod(LA)).item(), "file_bytes_in_per_step": in_bytes, "file_bytes_out_per_step": out_bytes},
            "host_stage_seconds": {k: round(v, 3) for k, v in sorted(stages.items())},
            "host_stage_bound": max(stages, key=stages.get) if stages else None,
            "gpu_launches": int(launches), "roofline": None, "cpu_baseline": None, "parity": parity, "clocks": sampler.summary(),
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
