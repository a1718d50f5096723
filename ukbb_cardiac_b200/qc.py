"""Quality-control gates on label maps: the checks that decide which subjects go on to the
wall-thickness / strain / atrial-volume stages of the reference pipeline.

  sa_pass_quality_control      common/cardiac_utils.py:77-136   (short_axis/eval_wall_thickness.py:43, eval_strain_sax.py:46)
  la_pass_quality_control      common/cardiac_utils.py:137-166  (long_axis/eval_strain_lax.py:46)
  atrium_pass_quality_control  common/cardiac_utils.py:1616-1652 (long_axis/eval_atrial_volume.py:66,102)
  aorta_pass_quality_control   common/cardiac_utils.py:1739-1795 (aortic/eval_aortic_area.py:68)

Same names, arguments, printed messages and verdicts as the reference.  The per-slice, per-class connected-component
statistics (area, number of components above the pixel threshold, largest component, area kept by remove_small_cc) come from
ONE launch of ``ukbb_cc_stats`` on the device label volume -- e.g. all 50 frames x 2 atria of a long-axis sequence at once, where
the reference labels every frame with scikit-image on the host.  Only the "epicardium" mask of the single mid-cavity slice
(union of two derived masks, image_utils.py:227-249 semantics incl. tie-breaking) is built on the host.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Sequence

import numpy as np
import torch

from . import _lib, nifti

PIXEL_THRES = 10


def cc_stats(labels_nyx: torch.Tensor, classes: Sequence[int], connectivity: int = 1, thres: int = PIXEL_THRES) -> np.ndarray:
    """labels_nyx: cuda uint8 [N, Y, X].  Returns int32 [N, len(classes), 6] = (area, n_cc, n_cc with area > thres, largest area,
    first pixel of the largest component, area kept by remove_small_cc(thres))."""
    assert labels_nyx.is_cuda and labels_nyx.dtype == torch.uint8 and labels_nyx.dim() == 3 and labels_nyx.is_contiguous()
    lib = _lib.load()
    n, y, x = labels_nyx.shape
    out = torch.empty((n, len(classes), 6), dtype=torch.int32, device=labels_nyx.device)
    cls = (C.c_int * len(classes))(*[int(c) for c in classes])
    with torch.cuda.device(labels_nyx.device):
        _lib.check(lib.ukbb_cc_stats(labels_nyx.data_ptr(), n, x, y, cls, len(classes), connectivity, thres, out.data_ptr(),
                                     torch.cuda.current_stream(labels_nyx.device).cuda_stream))
    return out.cpu().numpy()


def _device_slices(seg_xyz: np.ndarray, device) -> torch.Tensor:
    """(X, Y, N) label array -> cuda uint8 [N, Y, X] (NIfTI memory order of each slice)."""
    a = np.ascontiguousarray(np.transpose(np.asarray(seg_xyz), (2, 1, 0))).astype(np.uint8)
    return torch.from_numpy(a).to(device)


def _largest_cc(binary: np.ndarray) -> np.ndarray:
    """image_utils.py:227-238 on the host (first label wins ties), for the one derived mask per subject."""
    from scipy import ndimage
    cc, n_cc = ndimage.label(binary)
    if n_cc == 0:
        return cc == -1
    areas = np.bincount(cc.ravel(), minlength=n_cc + 1)[1:]
    return cc == (int(np.argmax(areas)) + 1)          # argmax returns the first maximum = the reference's strict '>' scan


def _remove_small_cc(binary: np.ndarray, thres: int = PIXEL_THRES) -> np.ndarray:
    from scipy import ndimage
    cc, n_cc = ndimage.label(binary)
    areas = np.bincount(cc.ravel(), minlength=n_cc + 1)
    keep = areas >= thres
    keep[0] = False
    return keep[cc]


def _load(seg):
    if isinstance(seg, str):
        return np.asarray(nifti.load(seg).get_data()), seg
    return np.asarray(seg), "segmentation"


def sa_pass_quality_control(seg_sa_name, device=0) -> bool:
    """Quality control for short-axis image segmentation (cardiac_utils.py:77-136).  `seg_sa_name`: file name (as in the
    reference) or the (X, Y, Z) label array itself."""
    seg_sa, name = _load(seg_sa_name)
    seg_sa = seg_sa[:, :, :, 0] if seg_sa.ndim == 4 else seg_sa
    Z = seg_sa.shape[2]
    label = {'LV': 1, 'Myo': 2, 'RV': 3}
    st = cc_stats(_device_slices(seg_sa, torch.device("cuda", device)), list(label.values()))      # [Z, 3, 6]
    area = st[:, :, 0].astype(np.int64)
    for i, l_name in enumerate(label):                                                             # criterion 1 (:89-95)
        if area[:, i].sum() < PIXEL_THRES:
            print('{0}: The segmentation for class {1} is smaller than {2} pixels. '
                  'It does not pass the quality control.'.format(name, l_name, PIXEL_THRES))
            return False
    z_pos = [z for z in range(Z) if area[z, 0] >= PIXEL_THRES and area[z, 1] >= PIXEL_THRES]       # criterion 2 (:99-118)
    slice_thres = 6
    if len(z_pos) < slice_thres:
        print('{0}: The segmentation has less than {1} slices. '
              'It does not pass the quality control.'.format(name, slice_thres))
        return False
    if len(z_pos) != (max(z_pos) - min(z_pos) + 1):
        print('{0}: There is missing segmentation between the slices. '
              'It does not pass the quality control.'.format(name))
        return False
    cz = float((np.arange(Z) * area[:, 0]).sum()) / float(area[:, 0].sum())                        # criterion 3 (:121-135)
    z = int(round(cz))
    seg_z = seg_sa[:, :, z]
    endo = _largest_cc(seg_z == label['LV'])
    myo = _remove_small_cc(seg_z == label['Myo'])
    epi = _largest_cc(endo | myo)
    rv_area = int(st[z, 2, 3])                                                                     # sum(get_largest_cc(rv)) = largest area
    if epi.sum() < PIXEL_THRES or rv_area < PIXEL_THRES:
        print('{0}: Can not find LV epi or RV to determine the AHA '
              'coordinate system.'.format(name))
        return False
    return True


def la_pass_quality_control(seg_la_name, device=0) -> bool:
    """Quality control for long-axis image segmentation (cardiac_utils.py:137-166)."""
    seg, name = _load(seg_la_name)
    seg = seg[:, :, :, 0] if seg.ndim == 4 else seg
    seg_z = seg[:, :, 0]
    label = {'LV': 1, 'Myo': 2, 'RV': 3, 'LA': 4, 'RA': 5}
    st = cc_stats(_device_slices(seg[:, :, :1], torch.device("cuda", device)), list(label.values()))[0]     # [5, 6]
    for i, l_name in enumerate(label):
        if st[i, 0] < PIXEL_THRES:
            print('{0}: The segmentation for class {1} is smaller than {2} pixels. '
                  'It does not pass the quality control.'.format(name, l_name, PIXEL_THRES))
            return False
    endo_area, myo_area = int(st[0, 3]), int(st[1, 5])            # sum(get_largest_cc(endo)), sum(remove_small_cc(myo))
    epi = _largest_cc(_largest_cc(seg_z == label['LV']) | _remove_small_cc(seg_z == label['Myo']))
    if endo_area < PIXEL_THRES or myo_area < PIXEL_THRES or epi.sum() < PIXEL_THRES:
        print('{0}: Can not find LV endo, myo or epi to extract the long-axis '
              'myocardial contour.'.format(name))
        return False
    return True


def atrium_pass_quality_control(label: np.ndarray, label_dict: Dict[str, int], device=0) -> bool:
    """Quality control for atrial volume estimation (cardiac_utils.py:1616-1652): label (X, Y, 1, T).  All T frames and all
    atria are analysed by one device launch (8-connected components in the slice plane = skimage connectivity 2 for Z = 1)."""
    label = np.asarray(label)
    if label.shape[2] != 1:
        raise ValueError("atrium_pass_quality_control expects single-slice long-axis label maps (X, Y, 1, T)")
    T = label.shape[3]
    st = cc_stats(_device_slices(label[:, :, 0, :], torch.device("cuda", device)), list(label_dict.values()), connectivity=2)   # [T, L, 6]
    for i, l_name in enumerate(label_dict):
        A = st[:, i, 0].astype(np.int64)
        for t in range(T):                                                                        # criterion 1 (:1619-1627)
            if A[t] == 0:
                print('The area of {0} is 0 at time frame {1}.'.format(l_name, t))
                return False
        for t in range(T):                                                                        # criterion 2 (:1629-1643)
            if st[t, i, 2] >= 2:
                print('The segmentation has at least two connected components with more than {0} pixels '
                      'at time frame {1}.'.format(PIXEL_THRES, t))
                return False
        for t in range(T):                                                                        # criterion 3 (:1645-1651)
            ratio = A[t] / float(A[t - 1])
            if ratio >= 2 or ratio <= 0.5:
                print('There is abrupt change of area at time frame {0}.'.format(t))
                return False
    return True


def aorta_pass_quality_control(image: np.ndarray, seg: np.ndarray, device=0) -> bool:
    """Quality control for aortic segmentation (cardiac_utils.py:1739-1795): image, seg (X, Y, 1, T).  Areas, fragment counts
    (8-connected) and the area ratios of all T frames and both vessels come from one device launch; the intensity criterion
    (max within the vessel vs the ED mean) is a masked reduction of the image on the host."""
    image, seg = np.asarray(image), np.asarray(seg)
    if seg.shape[2] != 1:
        raise ValueError("aorta_pass_quality_control expects single-slice aortic label maps (X, Y, 1, T)")
    T = seg.shape[3]
    st = cc_stats(_device_slices(seg[:, :, 0, :], torch.device("cuda", device)), [1, 2], connectivity=2)    # [T, 2, 6]
    for i, (l_name, l) in enumerate([('AAo', 1), ('DAo', 2)]):
        A = st[:, i, 0].astype(np.int64)
        for t in range(T):                                                                        # criterion 1 (:1742-1749)
            if A[t] == 0:
                print('The area of {0} is 0 at time frame {1}.'.format(l_name, t))
                return False
        mean_intensity_ED = image[:, :, :, 0][seg[:, :, :, 0] == l].mean()                          # criterion 2 (:1751-1763)
        for t in range(T):
            if np.max(image[:, :, :, t][seg[:, :, :, t] == l]) / mean_intensity_ED >= 3:
                print('The image becomes very noisy at time frame {0}.'.format(t))
                return False
        for t in range(T):                                                                        # criterion 3 (:1765-1780)
            if st[t, i, 2] >= 2:
                print('The segmentation has at least two connected components with more than {0} pixels '
                      'at time frame {1}.'.format(PIXEL_THRES, t))
                return False
        for t in range(T):                                                                        # criterion 4 (:1782-1788)
            ratio = A[t] / float(A[t - 1])
            if ratio >= 2 or ratio <= 0.5:
                print('There is abrupt change of area at time frame {0}.'.format(t))
                return False
        if np.max(A) / np.min(A) >= 2:                                                            # criterion 5 (:1790-1794)
            print('There is large change of area between maximum and minimum areas.')
            return False
    return True
