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
  --mode {bf16,fp16,fp32}   arithmetic of the conv layers (default bf16 tensor cores)
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


# ----------------------------------------------------------------------------- flags
class Flags:
    seq_name = "sa"                 # deploy_network.py:26-28
    data_dir = "ukbb_cardiac_demo"  # :29-31
    model_path = ""                 # :32-34
    process_seq = True              # :35-36
    save_seg = True                 # :37-38
    seg4 = False                    # :39-40
    mode = "bf16"
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
    if f.mode not in ("bf16", "fp16", "fp32"):
        raise SystemExit("flag --mode=%s: value should be one of <bf16|fp16|fp32>" % f.mode)
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


class _Prefetcher:
    """Decode the next subject's .nii.gz on a host thread while the GPU works on the current one."""

    def __init__(self, paths: List[Optional[str]]):
        self.paths = paths
        self.results: Dict[int, object] = {}
        self.cv = threading.Condition()
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        for i, p in enumerate(self.paths):
            try:
                r = nifti.load(p) if p is not None else None
            except Exception as e:   # surfaced to the consumer
                r = e
            with self.cv:
                self.results[i] = r
                self.cv.notify_all()
                while len(self.results) > 2:          # bounded look-ahead
                    self.cv.wait(0.05)

    def get(self, i: int):
        with self.cv:
            while i not in self.results:
                self.cv.wait()
            r = self.results.pop(i)
            self.cv.notify_all()
        if isinstance(r, Exception):
            raise r
        return r


def deploy(flags: Flags, engine=None, out=sys.stdout) -> int:
    """The body of deploy_network.py:43-225 for one shard of the subject list."""
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

    if flags.process_seq:
        todo = []
        for data in data_list:
            data_dir = os.path.join(flags.data_dir, data)
            seg_name = "{0}/{1}_{2}.nii.gz".format(data_dir, prefix, flags.seq_name)
            image_name = "{0}/{1}.nii.gz".format(data_dir, flags.seq_name)
            skip = os.path.exists(seg_name)                            # :62-67
            todo.append((data, data_dir, image_name, skip))
        pre = _Prefetcher([None if (skip or not os.path.exists(img)) else img for _, _, img, skip in todo])
        for idx, (data, data_dir, image_name, skip) in enumerate(todo):
            say(data)                                                  # :59
            nim = pre.get(idx)
            if skip:
                continue
            if nim is None:
                say("  Directory {0} does not contain an image with file "
                    "name {1}. Skip.".format(data_dir, os.path.basename(image_name)))       # :73-76
                continue
            say("  Reading {} ...".format(image_name))                 # :79
            image = _as_float32(nim.get_data())
            if image.ndim != 4:
                say("  Error: {0} is not a 4-D sequence (shape {1}). Skip.".format(image_name, image.shape))
                continue
            say("  Segmenting full sequence ...")                      # :85
            start_seg_time = time.time()
            labels, (vl, vh), counts = engine.segment_volume(image)    # :89-116 in one device call
            seg_time = time.time() - start_seg_time
            say("  Segmentation time = {:3f}s".format(seg_time))       # :119
            table_time += [seg_time]
            processed_list += [data]
            k = {"ED": 0, "ES": es_frame_from_counts(counts, flags.seq_name, flags.seg4)}   # :125-131
            say("  ED frame = {:d}, ES frame = {:d}".format(k["ED"], k["ES"]))
            if flags.save_seg:
                say("  Saving segmentation ...")                       # :135
                pred = labels.astype(label_dt, order="F")
                nim2 = nifti.Nifti1Image(pred, nim.affine)
                nim2.header["pixdim"] = nim.header["pixdim"]           # :137
                nifti.save(nim2, "{0}/{1}_{2}.nii.gz".format(data_dir, prefix, flags.seq_name))
                for fr in ("ED", "ES"):
                    frame = clip_like_reference(image[:, :, :, k[fr]], vl, vh)
                    nifti.save(nifti.Nifti1Image(np.asfortranarray(frame), nim.affine),
                               "{0}/{1}_{2}.nii.gz".format(data_dir, flags.seq_name, fr))   # :144-146
                    nifti.save(nifti.Nifti1Image(np.asfortranarray(pred[:, :, :, k[fr]]), nim.affine),
                               "{0}/{1}_{2}_{3}.nii.gz".format(data_dir, prefix, flags.seq_name, fr))   # :147-151
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


def main(argv: Optional[List[str]] = None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    flags = parse_flags(argv)
    if flags.gpus > 1 and flags.num_shards == 1:
        # one worker process per GPU, each pinned with CUDA_VISIBLE_DEVICES (demo_pipeline.py:25 style)
        procs = []
        for r in range(flags.gpus):
            env = dict(os.environ, CUDA_VISIBLE_DEVICES=str(r))
            cmd = [sys.executable, "-m", "ukbb_cardiac_b200.deploy"] + argv + ["--shard_index", str(r), "--num_shards", str(flags.gpus), "--gpus", "1"]
            procs.append(subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        rc = 0
        for r, p in enumerate(procs):
            outp, _ = p.communicate()
            sys.stdout.write(outp)
            rc = rc or p.returncode
        return rc
    return deploy(flags)


if __name__ == "__main__":
    sys.exit(main())
