"""Drop-in replacement of ``common/deploy_network.py`` (reference lines cited inline).

Same command line (``tf.app.flags`` syntax: ``--flag value``, ``--flag=value``, boolean
``--flag`` / ``--noflag`` / ``--flag=false``), same directory contract
(``<data_dir>/<subject>/<seq>.nii.gz`` in; ``seg_<seq>.nii.gz``, ``<seq>_ED/ES.nii.gz``,
``seg_<seq>_ED/ES.nii.gz`` out), same skip-if-exists resume rule and the same stdout lines.
What changed is underneath: the graph is not imported from ``.meta`` -- it IS build_FCN,
implemented in libukbb_fcn.so -- the weights are read from the same checkpoint files
(``<model_path>.index`` / ``.data-00000-of-00001``) without TensorFlow, and a whole cine
sequence is segmented by ONE device call instead of one ``sess.run`` per time frame.

Extra flags (all optional, defaults keep the reference behaviour):
  --mode {fp16x2,fp16x3,bf16x3,fp16,bf16,fp32}   arithmetic of the conv layers.  Default fp16x2: split-operand tensor-core mode
                            (FP16 main product + one FP8 correction product per K step, FP32 accumulate) -- the fastest
                            mode that meets the parity tolerance (>= 99.9 % label agreement, Dice >= 0.999 vs the float32
                            reference); fp16x3 (hi + lo FP16 pairs, three products) is ~9 % slower and ~4x tighter on the
                            logits.  fp16 / bf16 are faster
                            but do not meet it on random-init weights; fp32 is the CUDA-core exactness mode.
  --gpus N                  shard the sorted subject list over N GPUs, subject i -> GPU i % N
                            (one worker process per GPU, no device collective)
  --label_dtype {float64,uint8}  dtype of the saved label volumes (reference: float64)
"""
from __future__ import annotations

import math
import os
import subprocess
import sys
import threading
import time
from typing import Dict, List, Optional

import numpy as np

from . import nifti

SEQ_NAMES = ("sa", "la_2ch", "la_4ch")
MODE_NAMES = ("fp16x2", "fp16x3", "bf16x3", "fp16", "bf16", "fp32")


# ----------------------------------------------------------------------------- flags
class Flags:
    seq_name = "sa"                 # deploy_network.py:26-28
    data_dir = "ukbb_cardiac_demo"  # :29-31
    model_path = ""                 # :32-34
    process_seq = True              # :35-36
    save_seg = True                 # :37-38
    seg4 = False                    # :39-40
    mode = "fp16x2"
    gpus = 1
    label_dtype = "float64"
    shard_index = 0
    num_shards = 1


_BOOL = {"process_seq", "save_seg", "seg4"}
_INT = {"gpus", "shard_index", "num_shards"}


def _parse_bool(v: str) -> bool:
    if v.lower() in ("true", "t", "1", "yes", "y"):
        return True
    if v.lower() in ("false", "f", "0", "no", "n"):
        return False
    raise SystemExit("flag value %r is not a boolean" % v)


def parse_flags(argv: List[str]) -> Flags:
    """absl/tf.app.flags-style parsing; unknown flags are ignored like TF-1's wrapper does."""
    f = Flags()
    known = {k for k in vars(Flags) if not k.startswith("_")}
    i = 0
    while i < len(argv):
        a = argv[i]
        i += 1
        if not a.startswith("-"):
            continue
        name = a.lstrip("-")
        val: Optional[str] = None
        if "=" in name:
            name, val = name.split("=", 1)
        if name in _BOOL:
            setattr(f, name, True if val is None else _parse_bool(val))
            continue
        if name.startswith("no") and name[2:] in _BOOL and val is None:
            setattr(f, name[2:], False)
            continue
        if name not in known:
            continue
        if val is None:
            if i >= len(argv):
                raise SystemExit("flag --%s needs a value" % name)
            val = argv[i]
            i += 1
        setattr(f, name, int(val) if name in _INT else val)
    if f.seq_name not in SEQ_NAMES:
        raise SystemExit("flag --seq_name=%s: value should be one of <%s>" % (f.seq_name, "|".join(SEQ_NAMES)))
    if f.mode not in MODE_NAMES:
        raise SystemExit("flag --mode=%s: value should be one of <%s>" % (f.mode, "|".join(MODE_NAMES)))
    if f.label_dtype not in ("float64", "uint8"):
        raise SystemExit("flag --label_dtype=%s: value should be one of <float64|uint8>" % f.label_dtype)
    return f


# ----------------------------------------------------------------------------- host logic
def seg_prefix(flags: Flags) -> str:
    """deploy_network.py:62-65: 'seg4_' for the 4-chamber model of la_4ch, else 'seg_'."""
    return "seg4" if (flags.seq_name == "la_4ch" and flags.seg4) else "seg"


def es_frame_from_counts(counts: np.ndarray, seq_name: str, seg4: bool) -> int:
    """deploy_network.py:125-131 from the per-slice class counts [T, Z, C] the classifier emits:
    ES = argmin_t (sa, la_4ch+seg4) / argmax_t (otherwise) of the class-1 voxel count."""
    c1 = counts[:, :, 1].sum(axis=1)
    if seq_name == "sa" or (seq_name == "la_4ch" and seg4):
        return int(np.argmin(c1))
    return int(np.argmax(c1))


def clip_like_reference(frame: np.ndarray, vl: float, vh: float) -> np.ndarray:
    """The reference clips its input array in place (image_utils.py:73-75), so the
    <seq>_ED/ES.nii.gz it writes (deploy_network.py:144-146) are CLIPPED images."""
    out = np.array(frame, copy=True)
    out[out < np.float64(vl)] = vl
    out[out > np.float64(vh)] = vh
    return out


def shard(items: List[str], index: int, count: int) -> List[str]:
    """Deterministic subject sharding: subject i of the sorted list -> shard i % count,
    so a re-run with the same --gpus resumes with the same assignment."""
    return [s for i, s in enumerate(items) if i % count == index]


def _as_float32(image: np.ndarray) -> np.ndarray:
    # UK Biobank volumes are written as float32 (data/biobank_utils.py:314)
    return image if image.dtype == np.float32 else image.astype(np.float32)


def rescale_intensity_native(image: np.ndarray, thres=(1.0, 99.0)):
    """image_utils.py:70-77 verbatim, for volumes whose dtype is NOT float32: the reference clips its input array IN PLACE in the
    file's native dtype, so for an integer volume the thresholds are truncated on assignment (vl = 12.7 is stored as 12) before the
    float32 rescale, and a float64 volume (scl_slope set) is clipped in float64.  The device path assumes float32 voxels, so these
    volumes take the reference's own arithmetic on the host and only the network runs on the device.  Returns (rescaled array,
    the clipped input = what the reference saves as <seq>_ED/ES.nii.gz, vl, vh)."""
    val_l, val_h = np.percentile(image, thres)
    image2 = image
    image2[image < val_l] = val_l
    image2[image > val_h] = val_h
    return (image2.astype(np.float32) - val_l) / (val_h - val_l), image2, val_l, val_h


class _SyncEngine:
    """Adapter for engines that only offer the synchronous `segment_volume` (tests' stub): same ticket protocol as FCNEngine."""

    def __init__(self, engine):
        self.engine = engine

    def host_buffer(self, nbytes):
        return np.empty(nbytes, dtype=np.uint8)

    def release_host_buffer(self, arr):
        pass

    def submit_volume(self, image):
        return {"res": self.engine.segment_volume(image)}

    def collect_volume(self, ticket):
        return ticket["res"]

    def release_ticket(self, ticket):
        ticket["res"] = None


class _StageClock:
    """Thread-seconds spent in each host stage of the pipeline (decode / device wait / encode + write)."""

    def __init__(self, sink):
        self.sink = sink if sink is not None else {}
        self.lock = threading.Lock()

    def add(self, stage: str, seconds: float):
        with self.lock:
            self.sink[stage] = self.sink.get(stage, 0.0) + seconds


def deploy(flags: Flags, engine=None, out=sys.stdout, stage_times: Optional[Dict[str, float]] = None) -> int:
    """The body of deploy_network.py:43-225 for one shard of the subject list.

    The --process_seq branch is a three-stage pipeline over subjects: reader threads inflate <seq>.nii.gz straight into pinned host
    memory (member-parallel for files written by nifti.save, one file per thread otherwise), the main thread submits one
    asynchronous device call per subject (H2D, rescale, forward, D2H overlap between consecutive subjects), and writer threads build
    and deflate the output volumes.  Each subject's stdout block is emitted in subject order, with the reference's lines."""
    def say(s):
        print(s, file=out, flush=True)

    if engine is None:
        from .fcn import FCNEngine
        engine = FCNEngine.from_checkpoint(flags.model_path, device=0, mode=flags.mode)
    say("Start deployment on the data set ...")                        # :51
    start_time = time.time()
    data_list = shard(sorted(os.listdir(flags.data_dir)), flags.shard_index, flags.num_shards)   # :55
    processed_list, table_time = [], []
    prefix = seg_prefix(flags)
    label_dt = np.float64 if flags.label_dtype == "float64" else np.uint8
    clock = _StageClock(stage_times)

    if flags.process_seq:
        from concurrent.futures import ThreadPoolExecutor
        eng = engine if hasattr(engine, "submit_volume") else _SyncEngine(engine)
        todo = []
        for data in data_list:
            data_dir = os.path.join(flags.data_dir, data)
            seg_name = "{0}/{1}_{2}.nii.gz".format(data_dir, prefix, flags.seq_name)
            image_name = "{0}/{1}.nii.gz".format(data_dir, flags.seq_name)
            skip = os.path.exists(seg_name)                            # :62-67
            todo.append((data, data_dir, image_name, skip))
        n_readers = max(1, min(4, (os.cpu_count() or 2) // 2))
        readers = ThreadPoolExecutor(max_workers=n_readers)
        writers = ThreadPoolExecutor(max_workers=3)
        LOOKAHEAD = n_readers + 1

        def read_job(path):
            t0 = time.time()
            bufs = []

            def alloc(nbytes):
                bufs.append(eng.host_buffer(nbytes))
                return bufs[-1]
            try:
                nim = nifti.load(path, alloc=alloc)
            except Exception:
                for b_ in bufs:
                    eng.release_host_buffer(b_)
                raise
            clock.add("decode", time.time() - t0)
            return nim, bufs

        def write_job(data_dir, nim, image, bufs, ticket, labels, vl, vh, k):
            t0 = time.time()
            try:
                # uint8 labels are widened to the reference's float64 volume (:92) chunk by chunk inside the gzip workers
                nim2 = nifti.Nifti1Image(labels, nim.affine)
                nim2.header["pixdim"] = nim.header["pixdim"]           # :137
                nifti.save(nim2, "{0}/{1}_{2}.nii.gz".format(data_dir, prefix, flags.seq_name), dtype=label_dt, label_data=True)
                for fr in ("ED", "ES"):
                    frame = clip_like_reference(image[:, :, :, k[fr]], vl, vh)
                    nifti.save(nifti.Nifti1Image(np.asfortranarray(frame), nim.affine),
                               "{0}/{1}_{2}.nii.gz".format(data_dir, flags.seq_name, fr))   # :144-146
                    nifti.save(nifti.Nifti1Image(np.asfortranarray(labels[:, :, :, k[fr]]), nim.affine),
                               "{0}/{1}_{2}_{3}.nii.gz".format(data_dir, prefix, flags.seq_name, fr), dtype=label_dt, label_data=True)   # :147-151
            finally:
                eng.release_ticket(ticket)
                for b_ in bufs:
                    eng.release_host_buffer(b_)
                clock.add("encode_write", time.time() - t0)

        reads = {}

        def schedule_reads(upto):
            for j in range(len(todo)):
                if j >= upto:
                    break
                if j not in reads:
                    _, _, img, skip = todo[j]
                    reads[j] = None if (skip or not os.path.exists(img)) else readers.submit(read_job, img)

        pending = []            # subjects submitted to the device, oldest first: (lines, data, data_dir, nim, image, bufs, ticket, t_submit)
        saves = []

        def finish(entry):
            lines, data, data_dir, nim, image, bufs, ticket, t_submit = entry
            t0 = time.time()
            labels, (vl, vh), counts = eng.collect_volume(ticket)      # :89-116 in one device call
            clock.add("device_wait", time.time() - t0)
            seg_time = time.time() - t_submit
            lines.append("  Segmentation time = {:3f}s".format(seg_time))       # :119
            table_time.append(seg_time)
            processed_list.append(data)
            k = {"ED": 0, "ES": es_frame_from_counts(counts, flags.seq_name, flags.seg4)}   # :125-131
            lines.append("  ED frame = {:d}, ES frame = {:d}".format(k["ED"], k["ES"]))
            if flags.save_seg:
                lines.append("  Saving segmentation ...")              # :135
                while len(saves) >= 3:                                 # bounded backlog of label volumes in host memory
                    saves.pop(0).result()
                saves.append(writers.submit(write_job, data_dir, nim, image, bufs, ticket, labels, vl, vh, k))
            else:
                eng.release_ticket(ticket)
                for b_ in bufs:
                    eng.release_host_buffer(b_)
            for ln in lines:
                say(ln)

        try:
            for idx, (data, data_dir, image_name, skip) in enumerate(todo):
                schedule_reads(idx + LOOKAHEAD)
                lines = [data]                                         # :59
                fut = reads.pop(idx)
                if skip:
                    while pending:
                        finish(pending.pop(0))
                    say(data)
                    continue
                if fut is None:
                    while pending:
                        finish(pending.pop(0))
                    say(data)
                    say("  Directory {0} does not contain an image with file "
                        "name {1}. Skip.".format(data_dir, os.path.basename(image_name)))       # :73-76
                    continue
                lines.append("  Reading {} ...".format(image_name))    # :79
                nim, bufs = fut.result()
                image = nim.get_data()
                if image.ndim != 4:
                    for b_ in bufs:
                        eng.release_host_buffer(b_)
                    while pending:
                        finish(pending.pop(0))
                    for ln in lines:
                        say(ln)
                    say("  Error: {0} is not a 4-D sequence (shape {1}). Skip.".format(image_name, image.shape))
                    continue
                lines.append("  Segmenting full sequence ...")         # :85
                if image.dtype != np.float32 and hasattr(engine, "segment_rescaled"):
                    # native-dtype volume: the reference's host arithmetic for the rescale (see rescale_intensity_native), synchronous
                    while pending:
                        finish(pending.pop(0))
                    start_seg_time = time.time()
                    image2, clipped, _, _ = rescale_intensity_native(np.array(image, order="F"))
                    labels, counts = engine.segment_rescaled(image2)
                    seg_time = time.time() - start_seg_time
                    lines.append("  Segmentation time = {:3f}s".format(seg_time))
                    table_time.append(seg_time)
                    processed_list.append(data)
                    k = {"ED": 0, "ES": es_frame_from_counts(counts, flags.seq_name, flags.seg4)}
                    lines.append("  ED frame = {:d}, ES frame = {:d}".format(k["ED"], k["ES"]))
                    if flags.save_seg:
                        lines.append("  Saving segmentation ...")
                        nim2 = nifti.Nifti1Image(labels, nim.affine)
                        nim2.header["pixdim"] = nim.header["pixdim"]
                        nifti.save(nim2, "{0}/{1}_{2}.nii.gz".format(data_dir, prefix, flags.seq_name), dtype=label_dt, label_data=True)
                        for fr in ("ED", "ES"):                        # the clipped frames keep the file's dtype, like orig_image in the reference
                            nifti.save(nifti.Nifti1Image(np.asfortranarray(clipped[:, :, :, k[fr]]), nim.affine),
                                       "{0}/{1}_{2}.nii.gz".format(data_dir, flags.seq_name, fr))
                            nifti.save(nifti.Nifti1Image(np.asfortranarray(labels[:, :, :, k[fr]]), nim.affine),
                                       "{0}/{1}_{2}_{3}.nii.gz".format(data_dir, prefix, flags.seq_name, fr), dtype=label_dt, label_data=True)
                    for b_ in bufs:
                        eng.release_host_buffer(b_)
                    for ln in lines:
                        say(ln)
                    continue
                image = _as_float32(image)
                t_submit = time.time()
                ticket = eng.submit_volume(image)
                pending.append((lines, data, data_dir, nim, image, bufs, ticket, t_submit))
                if len(pending) > 1:
                    finish(pending.pop(0))
            while pending:
                finish(pending.pop(0))
            for f in saves:
                f.result()
        finally:
            readers.shutdown(wait=True)
            writers.shutdown(wait=True)
    else:
        for data in data_list:
            say(data)
            data_dir = os.path.join(flags.data_dir, data)
            if os.path.exists("{0}/{1}_{2}.nii.gz".format(data_dir, prefix, flags.seq_name)):
                continue                                               # :62-67 applies to both branches
            names = {fr: "{0}/{1}_{2}.nii.gz".format(data_dir, flags.seq_name, fr) for fr in ("ED", "ES")}
            if not os.path.exists(names["ED"]) or not os.path.exists(names["ES"]):
                say("  Directory {0} does not contain an image with "
                    "file name {1} or {2}. Skip.".format(data_dir, os.path.basename(names["ED"]),
                                                         os.path.basename(names["ES"])))    # :156-161
                continue
            for fr in ("ED", "ES"):
                say("  Reading {} ...".format(names[fr]))              # :168
                nim = nifti.load(names[fr])
                image = _as_float32(nim.get_data())
                squeeze = image.ndim == 2                              # :172-173
                say("  Segmenting {} frame ...".format(fr))            # :175
                start_seg_time = time.time()
                labels, _, _ = engine.segment_volume(image)            # :179-200
                seg_time = time.time() - start_seg_time
                say("  Segmentation time = {:3f}s".format(seg_time))   # :203
                table_time += [seg_time]
                processed_list += [data]
                if flags.save_seg:
                    say("  Saving segmentation ...")
                    # the reference saves the int32 `pred:0` array itself in this branch (:199-213)
                    pred = labels.astype(np.int32 if flags.label_dtype == "float64" else np.uint8, order="F")
                    if squeeze:
                        pred = pred.reshape(pred.shape[0], pred.shape[1], 1, order="F")   # reference keeps the Z axis it added
                    nim2 = nifti.Nifti1Image(pred, nim.affine)
                    nim2.header["pixdim"] = nim.header["pixdim"]
                    nifti.save(nim2, "{0}/{1}_{2}_{3}.nii.gz".format(data_dir, prefix, flags.seq_name, fr))   # :210-216

    if flags.process_seq:
        say("Average segmentation time = {:.3f}s per sequence".format(np.mean(table_time) if table_time else float("nan")))
    else:
        say("Average segmentation time = {:.3f}s per frame".format(np.mean(table_time) if table_time else float("nan")))
    process_time = time.time() - start_time
    n = len(processed_list)
    # the reference divides by len(processed_list) and dies with ZeroDivisionError when nothing
    # was processed (:225); a resumed run with nothing left to do is not an error here
    say("Including image I/O, CUDA resource allocation, "
        "it took {:.3f}s for processing {:d} subjects ({:.3f}s per subjects).".format(
            process_time, n, process_time / n if n else float("nan")))
    return 0


def worker_devices(n: int, env=None) -> List[str]:
    """Physical device of worker r: the r-th entry of the caller's CUDA_VISIBLE_DEVICES when one is set (a mask such as
    '4,5,6,7' must keep meaning those GPUs), else r."""
    env = os.environ if env is None else env
    mask = [d.strip() for d in env.get("CUDA_VISIBLE_DEVICES", "").split(",") if d.strip()]
    if mask:
        if n > len(mask):
            raise SystemExit("--gpus %d but CUDA_VISIBLE_DEVICES lists %d device(s)" % (n, len(mask)))
        return mask[:n]
    return [str(r) for r in range(n)]


def run_workers(flags: Flags, argv: List[str]) -> int:
    """One worker process per GPU, each pinned with CUDA_VISIBLE_DEVICES (demo_pipeline.py:25 style).  Every worker's output is
    drained by its own thread while it runs (a worker must never block on a full pipe: cohort-scale runs print far more than
    the 64 KB a pipe holds) and is replayed in rank order at the end."""
    devices = worker_devices(flags.gpus)
    procs, logs, threads = [], [], []
    for r in range(flags.gpus):
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=devices[r])
        cmd = [sys.executable, "-m", "ukbb_cardiac_b200.deploy"] + argv + ["--shard_index", str(r), "--num_shards", str(flags.gpus), "--gpus", "1"]
        p = subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        buf: List[str] = []
        t = threading.Thread(target=lambda p=p, buf=buf: buf.extend(p.stdout), daemon=True)
        t.start()
        procs.append(p); logs.append(buf); threads.append(t)
    rc = 0
    for p, buf, t in zip(procs, logs, threads):
        p.wait()
        t.join()
        sys.stdout.write("".join(buf))
        rc = rc or p.returncode
    sys.stdout.flush()
    return rc


def main(argv: Optional[List[str]] = None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    flags = parse_flags(argv)
    if flags.gpus > 1 and flags.num_shards == 1:
        return run_workers(flags, argv)
    return deploy(flags)


if __name__ == "__main__":
    sys.exit(main())
