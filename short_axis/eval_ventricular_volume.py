#!/usr/bin/env python3
"""Clinical measures of the short-axis segmentation, one CSV row per subject.

Command line and output are those of the reference's script of the same name (`--data_dir DIR --output_csv FILE`;
subject folders holding `sa.nii.gz` and `seg_sa.nii.gz`; columns LVEDV ... RVEF), so `demo_pipeline.py` can call it unchanged.
The arithmetic lives in `ukbb_cardiac_b200.volumes` (shared with the deploy path, where the per-frame class counts come
from the device instead of a second pass over the label volume); NIfTI files are read with the repository's own reader."""
import argparse
import pathlib
import sys

import pandas as pd

REPO = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from ukbb_cardiac_b200 import nifti, volumes  # noqa: E402


def subject_row(folder: pathlib.Path):
    """Measures of one subject folder, or None when either file is missing (the reference skips such folders silently)."""
    image_path, label_path = folder / "sa.nii.gz", folder / "seg_sa.nii.gz"
    if not (image_path.is_file() and label_path.is_file()):
        return None
    header = nifti.load(str(image_path)).header
    counts = volumes.frame_counts_from_labels(nifti.load(str(label_path)).get_data())
    return volumes.table_row(volumes.ventricular_volumes(counts, header["pixdim"], int(header["dim"][4])))


def main(argv=None) -> int:
    cli = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    cli.add_argument("--data_dir", metavar="dir_name", default="", required=True)
    cli.add_argument("--output_csv", metavar="csv_name", default="", required=True)
    opts = cli.parse_args(argv)
    rows, names = [], []
    for folder in sorted(pathlib.Path(opts.data_dir).iterdir(), key=lambda q: q.name):
        row = subject_row(folder) if folder.is_dir() else None
        if row is not None:
            print(folder.name)
            rows.append(row)
            names.append(folder.name)
    pd.DataFrame(rows, index=names, columns=volumes.COLUMNS).to_csv(opts.output_csv)
    return 0


if __name__ == "__main__":
    sys.exit(main())
