#!/usr/bin/env python
"""Benchmark of the FCN deploy hot path (BASELINE.json metric: SA FCN 192x208 slices/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode bf16|fp32] [--subjects S]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the CPU restatement of the reference loop

One step = one pass of the hot path (percentile rescale + pad + build_FCN forward +
argmax/crop) over a batch of S synthetic short-axis subjects (192x208x10x50 = 500 slices
each) per GPU.  `value` times the device-resident path (inputs already in HBM);
`e2e.value` times the public host-buffer call (pinned host -> H2D -> compute -> D2H of the
label volumes) over the same batch.  One JSON line is printed by rank 0.
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
FLOP_PER_SLICE = 3.13737216e9          # algorithmic FLOPs, SA 192x208, 4 classes
# dominant kernel = the fused head (head_ts_kernel, ~30 % of the step): same_dim0 40.89 + 4-tap upsample 40.89 + fc0 817.89 +
# fc1 327.16 + class scores 20.45 MFLOP per slice by SURVEY 8(d)'s per-layer count
HEAD_FLOP_PER_SLICE = (40.89 + 40.89 + 817.89 + 327.16 + 20.45) * 1e6
# dram__bytes_read.sum + dram__bytes_write.sum of ONE head_ts_kernel launch (500 slices) from the ncu --set full capture
# summarised in profiles/r1_ncu_full_final_summary.txt (a number taken under the profiler is evidence, not a bench value)
HEAD_DRAM_BYTES_PER_LAUNCH = 1.516e9
POOL = 8                               # distinct synthetic subjects cycled through a batch


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "source": "MEASURED_PEAKS.json"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


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
def cpu_reference_sample(frames: int, threads: int):
    """The reference's loop (deploy_network.py:89-116) restated on CPU: global percentile
    rescale, then ONE forward per time frame with batch Z=10, float32 (PyTorch/oneDNN stands in
    for TensorFlow-CPU, which is not installable here).  Bounded sample: `frames` frames of one
    synthetic SA subject.  Returns (slices_per_s, seconds, n_slices)."""
    import torch
    from oracle import deploy_oracle as do
    from ukbb_cardiac_b200 import synth
    torch.set_num_threads(threads)
    w = synth.make_weights(0, 4)
    vol = synth.make_stack(0, (SA[0], SA[1], SA[2], frames))
    run = do.make_runner(w)
    run(np.zeros((1, 32, 32, 1), np.float32))            # warm oneDNN primitives
    t0 = time.perf_counter()
    pred, _ = do.deploy_sequence(vol, run)
    dt = time.perf_counter() - t0
    n = SA[2] * frames
    return n / dt, dt, n


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    frames = args.ref_frames
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_reference_sample(1, threads)
    vals, secs = [], []
    for _ in range(args.steps):
        v, dt, n = cpu_reference_sample(frames, threads)
        vals.append(v); secs.append(dt)
    value = (SA[2] * frames * args.steps) / sum(secs)
    sample = "%d frames x 10 slices of one synthetic SA subject per step, reference loop (batch Z per frame)" % frames
    line = {
        "impl": "reference", "metric": "SA FCN 192x208 slices/sec", "value": value, "unit": "slices/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "SA 192x208x10x50 subjects, FCN deploy (CPU sample)", "sample": sample},
        "cpu_baseline": {"value": value, "unit": "slices/s", "cores": threads, "kind": "port", "sample": sample,
                         "note": "TF-CPU proxy (PyTorch/oneDNN float32 restatement); TensorFlow unavailable"},
        "e2e": {"value": value, "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default=os.environ.get("UKBB_BENCH_MODE", "fp16x3"), choices=["fp16x3", "bf16x3", "bf16", "fp16", "fp32"])
    ap.add_argument("--subjects", type=int, default=None, help="SA subjects per GPU per step (default 256 bf16, 2 fp32)")
    ap.add_argument("--ref-frames", type=int, default=20, help="frames per step of the reference arm (20 frames = 200 slices, ~5 s of CPU work)")
    ap.add_argument("--cpu-frames", type=int, default=50, help="frames in the cpu_baseline sample (0 = skip); 50 = one whole subject, ~14 s")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from ukbb_cardiac_b200 import synth
    from ukbb_cardiac_b200.fcn import FCNEngine

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
    S = args.subjects if args.subjects else (2 if args.mode == "fp32" else 256)
    X, Y, Z, T = SA
    nvox = X * Y * Z * T

    eng = FCNEngine(synth.make_weights(0, 4), device=local, mode=args.mode)
    pool = min(POOL, S)
    host_pool = [torch.empty(nvox, dtype=torch.float32, pin_memory=True) for _ in range(pool)]
    for i, hp in enumerate(host_pool):
        hp.numpy()[:] = synth.make_stack(100 * rank + i).reshape(-1, order="F")
    dev_pool = [hp.to(dev) for hp in host_pool]                 # 8 x 80 MB > 126 MB L2
    host_labels = [torch.empty(nvox, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    host_counts = [torch.empty((Z * T, 4), dtype=torch.int64, pin_memory=True) for _ in range(2)]
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    fwd_ev = []
    # device-resident path: the percentile rescale of subject s + 1 runs on a side stream while the forward of subject s runs
    # on the main stream (two padded buffers), as ukbb_fcn_segment_host does internally for host buffers
    pre_stream = torch.cuda.Stream(dev)
    x2, y2 = (X + 15) // 16 * 16, (Y + 15) // 16 * 16
    pad = [torch.empty((Z * T, y2, x2), dtype=torch.float32, device=dev) for _ in range(2)]
    vv = [torch.empty(2, dtype=torch.float64, device=dev) for _ in range(2)]
    ev_pre = [torch.cuda.Event() for _ in range(2)]
    ev_fwd = [torch.cuda.Event() for _ in range(2)]

    def step_device(record=False):
        xp = yp = 0
        pre_stream.wait_stream(stream)
        for s in range(S + 1):
            if s < S:
                slot = s & 1
                with torch.cuda.stream(pre_stream):
                    pre_stream.wait_event(ev_fwd[slot])              # the forward that read pad[slot] two subjects ago
                    _, _, (xp, yp) = eng.preprocess(dev_pool[s % pool], Z * T, X, Y, out=pad[slot], vlvh=vv[slot])
                    ev_pre[slot].record(pre_stream)
            if s >= 1:
                slot = (s - 1) & 1
                stream.wait_event(ev_pre[slot])
                if record:
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                eng.forward(pad[slot], xp, yp, X, Y)
                if record:
                    e1.record(stream)
                    fwd_ev.append((e0, e1))
                ev_fwd[slot].record(stream)

    def step_e2e():
        for s in range(S):
            eng.segment_host_async(host_pool[s % pool], SA, host_labels[s & 1], None, host_counts[s & 1])
        eng.join()

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

    for _ in range(max(args.warmup, 3)):
        step_device()
    eng.kernel_timer(True)
    eng.kernel_timer_read()
    l0 = eng.launch_count
    sampler = ClockSampler(local)
    sampler.start()
    ms_dev = timed(step_device, args.steps, record=True)
    head_ms, head_n = eng.kernel_timer_read()
    eng.kernel_timer(False)
    launches = eng.launch_count - l0
    for _ in range(max(args.warmup, 3)):
        step_e2e()
    eng.sync()
    ms_e2e = timed(step_e2e, args.steps)
    sampler.stop_flag.set()
    sampler.join(timeout=3)
    eng.sync()

    slices = S * Z * T * args.steps * world
    value = slices / (ms_dev * 1e-3)
    e2e_value = slices / (ms_e2e * 1e-3)
    fwd_ms = sum(a.elapsed_time(b) for a, b in fwd_ev) / max(len(fwd_ev), 1)        # per 500-slice forward
    peaks = measured_peaks()
    fwd_tf = FLOP_PER_SLICE * Z * T / (fwd_ms * 1e-3) / 1e12
    peak_tf = peaks["bf16_sustained"]
    if head_n > 0 and args.mode != "fp32":
        head_avg_ms = head_ms / head_n
        achieved_tf = HEAD_FLOP_PER_SLICE * Z * T / (head_avg_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "head_ts_kernel (same_dim0 + upsample + fc0 + fc1 + class scores + softmax/argmax/crop; 1 launch per subject)",
                    "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                    "traffic": HEAD_DRAM_BYTES_PER_LAUNCH, "traffic_unit": "bytes per launch (ncu dram read + write)",
                    "algorithmic_flop_per_launch": HEAD_FLOP_PER_SLICE * Z * T, "avg_launch_ms": head_avg_ms, "launches_timed": head_n,
                    "timing": "CUDA events on the launching stream around every launch inside the timed region",
                    "peak_source": peaks["source"] + " (sustained bf16: the kernel runs inside a long step)",
                    "frac_of_burst": achieved_tf / peaks["bf16_burst"], "frac_of_nominal_2250": achieved_tf / 2250.0,
                    "share_of_forward": head_avg_ms / fwd_ms,
                    "whole_forward": {"achieved": fwd_tf, "frac": fwd_tf / peak_tf, "frac_of_burst": fwd_tf / peaks["bf16_burst"],
                                      "frac_of_nominal_2250": fwd_tf / 2250.0, "algorithmic_flop_per_slice": FLOP_PER_SLICE,
                                      "avg_forward_ms_per_subject": fwd_ms}}
    else:
        roofline = {"bound": "tensor", "kernel": "build_FCN forward (all conv launches of one 500-slice subject)",
                    "achieved": fwd_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": fwd_tf / peak_tf,
                    "traffic": None, "peak_source": peaks["source"] + " (sustained bf16, of measured)",
                    "frac_of_burst": fwd_tf / peaks["bf16_burst"], "frac_of_nominal_2250": fwd_tf / 2250.0,
                    "algorithmic_flop_per_slice": FLOP_PER_SLICE, "avg_forward_ms_per_subject": fwd_ms}

    line = None
    if rank == 0:
        cpu = None
        if args.cpu_frames > 0 and world == 1:               # the CPU baseline is measured at N = 1 only
            threads = os.cpu_count() or 1
            v, dt, n = cpu_reference_sample(args.cpu_frames, threads)
            cpu = {"value": v, "unit": "slices/s", "cores": threads, "kind": "port",
                   "sample": "%d frames x 10 slices of one synthetic SA subject, reference loop (batch Z per frame), %.1f s"
                             % (args.cpu_frames, dt),
                   "note": "TF-CPU proxy (PyTorch/oneDNN float32 restatement); TensorFlow unavailable"}
        line = {
            "metric": "SA FCN 192x208 slices/sec", "value": value, "unit": "slices/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"bf16": "bf16", "fp16": "f16", "fp32": "f32", "fp16x3": "f16x3", "bf16x3": "bf16x3"}[args.mode], "data": "synthetic",
            "config": {"workload": "%d synthetic SA subjects (192x208x10x50, 500 slices each) per GPU per step" % S,
                       "subjects_per_gpu": S, "global_subjects": S * world, "mode": args.mode, "n_class": 4,
                       "l2_policy": "inputs larger than L2: %d distinct 80 MB volumes cycled" % pool,
                       "pipeline": "rescale of subject s+1 on a side stream overlaps the forward of subject s",
                       "weights": "random-init (seed 0), reference TF checkpoint layout", "parallelism": "dp%d" % world},
            "subjects_per_s": value / (Z * T),
            "e2e": {"value": e2e_value, "unit": "slices/s", "h2d_bytes_per_step": S * nvox * 4,
                    "d2h_bytes_per_step": S * (nvox + Z * T * 4 * 8), "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "clocks": sampler.summary(),
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
