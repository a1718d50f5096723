"""Drop-in replacement of ``common/deploy_network_ao.py`` for the UNet-LSTM aortic model (reference lines cited inline).

Same flags (``tf.app.flags`` syntax), same directory contract (``<data_dir>/<subject>/ao.nii.gz`` in, ``seg_ao.nii.gz`` out as the
int32 label volume the reference writes, ``deploy_network_ao.py:189-193``) and the same stdout lines.  The graph is the UNet +
bidirectional ConvLSTM of ``common/network_ao.py`` implemented in libukbb_fcn.so; a whole cine is segmented by one device call.
Only ``--model UNet-LSTM`` with ``--time_step 1`` (the defaults) is implemented; anything else fails loudly.
"""
from __future__ import annotations

import os
import sys
import time
from typing import List, Optional

import numpy as np

from . import nifti
from .deploy import _parse_bool


class Flags:
    time_step = 1                   # deploy_network_ao.py:26-27
    seq_name = "ao"                 # :28-29
    model = "UNet-LSTM"             # :30-31
    data_dir = "Biobank_ao/validation"   # :32-35
    model_path = ""                 # :36-38
    process_seq = True              # :39-40
    save_seg = True                 # :41-42
    z_score = True                  # :43-45
    weight_R = 5                    # :46-47
    weight_r = 0.1                  # :48-49


_BOOL = {"process_seq", "save_seg", "z_score"}
_INT = {"time_step", "weight_R"}
_FLOAT = {"weight_r"}


def parse_flags(argv: List[str]) -> Flags:
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
        setattr(f, name, int(val) if name in _INT else float(val) if name in _FLOAT else val)
    if f.seq_name != "ao":
        raise SystemExit("flag --seq_name=%s: value should be one of <ao>" % f.seq_name)
    if f.model not in ("UNet", "UNet-LSTM", "Temporal-UNet"):
        raise SystemExit("flag --model=%s: value should be one of <UNet|UNet-LSTM|Temporal-UNet>" % f.model)
    return f


def deploy(flags: Flags, engine=None, out=sys.stdout) -> int:
    def say(s):
        print(s, file=out, flush=True)

    if flags.model != "UNet-LSTM" or flags.time_step != 1:
        raise SystemExit("only --model UNet-LSTM with --time_step 1 (the reference's defaults) is implemented by this engine")
    if engine is None:
        from .aorta import AortaEngine
        engine = AortaEngine.from_checkpoint(flags.model_path, device=0)
    say("Start evaluating on the test set ...")                       # :62
    start_time = time.time()
    data_list = sorted(os.listdir(flags.data_dir))                     # :66
    processed_list = []
    for data in data_list:
        say(data)                                                      # :70
        data_dir = os.path.join(flags.data_dir, data)
        if not flags.process_seq:
            say("UNet-LSTM does not support frame-wise segmentation. "
                "Please use the -process_seq flag.")                   # :203-205
            return 0
        image_name = "{0}/{1}.nii.gz".format(data_dir, flags.seq_name)
        if not os.path.exists(image_name):
            say("  Directory {0} does not contain an image with file name {1}. "
                "Skip.".format(data_dir, os.path.basename(image_name)))      # :77-80
            continue
        say("  Reading {} ...".format(image_name))                     # :83
        nim = nifti.load(image_name)
        image = nim.get_data()
        say("  Segmenting full sequence ...")                          # :91
        start_seg_time = time.time()
        pred, _ = engine.segment_sequence(image, weight_R=flags.weight_R, weight_r=flags.weight_r, z_score=flags.z_score)    # :94-186
        if flags.save_seg:
            say("  Saving segmentation ...")                           # :190
            nim2 = nifti.Nifti1Image(np.asfortranarray(pred), nim.affine)
            nim2.header["pixdim"] = nim.header["pixdim"]
            nifti.save(nim2, "{0}/seg_{1}.nii.gz".format(data_dir, flags.seq_name), label_data=True)      # :191-193
        seg_time = time.time() - start_seg_time
        say("  Segmentation time = {:3f}s".format(seg_time))          # :196
        processed_list += [data]
    process_time = time.time() - start_time
    n = len(processed_list)
    say("Including image I/O, CUDA resource allocation, "
        "it took {:.3f}s for processing {:d} subjects ({:.3f}s per subjects).".format(
            process_time, n, process_time / n if n else float("nan")))
    return 0


def main(argv: Optional[List[str]] = None) -> int:
    return deploy(parse_flags(list(sys.argv[1:] if argv is None else argv)))


if __name__ == "__main__":
    sys.exit(main())
