#!/usr/bin/env python3
"""Drop-in for the reference's short_axis/eval_ventricular_volume.py (same flags, same CSV): clinical measures of every
subject directory that holds sa.nii.gz and seg_sa.nii.gz.  Reads the label volume with the repository's NIfTI reader
(nibabel is not required) and counts voxels per frame; inside the deploy pipeline the same numbers come from the device
(`FCNEngine.segment_volume(...)[2]`, `ukbb_cardiac_b200.volumes.ventricular_volumes`)."""
import argparse
import os
import sys

import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ukbb_cardiac_b200 import nifti, volumes  # noqa: E402

if __name__ == '__main__':
    parser = argparse.ArgumentParser()
    parser.add_argument('--data_dir', metavar='dir_name', default='', required=True)
    parser.add_argument('--output_csv', metavar='csv_name', default='', required=True)
    args = parser.parse_args()

    table, processed_list = [], []
    for data in sorted(os.listdir(args.data_dir)):
        data_dir = os.path.join(args.data_dir, data)
        image_name = '{0}/sa.nii.gz'.format(data_dir)
        seg_name = '{0}/seg_sa.nii.gz'.format(data_dir)
        if os.path.exists(image_name) and os.path.exists(seg_name):
            print(data)
            nim = nifti.load(image_name)
            seg = nifti.load(seg_name).get_data()
            val = volumes.ventricular_volumes(volumes.frame_counts_from_labels(seg), nim.header['pixdim'], int(nim.header['dim'][4]))
            table += [volumes.table_row(val)]
            processed_list += [data]
    df = pd.DataFrame(table, index=processed_list, columns=volumes.COLUMNS)
    df.to_csv(args.output_csv)
