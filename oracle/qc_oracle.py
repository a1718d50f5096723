"""CPU restatement of the reference's segmentation quality-control gates (TEST INFRASTRUCTURE ONLY).

Follows
  * ``common/image_utils.py:227-249``    get_largest_cc / remove_small_cc (``scipy.ndimage.measurements.label``, default
    structure = face connectivity; the FIRST label wins ties of the largest area; components with area < thres are removed)
  * ``common/cardiac_utils.py:77-136``   sa_pass_quality_control
  * ``common/cardiac_utils.py:137-166``  la_pass_quality_control
  * ``common/cardiac_utils.py:1739-1795`` aorta_pass_quality_control (same labelling call)
  * ``common/cardiac_utils.py:1616-1652`` atrium_pass_quality_control (``skimage.measure.label(mask, connectivity=2)`` on an
    (X, Y, Z) array: neighbours sharing a face or an edge; scikit-image is absent here, the same labelling is obtained from
    ``scipy.ndimage.label`` with ``generate_binary_structure(3, 2)``)
Pinned by ``tests/golden/qc_reference.npz``: outputs of the reference's OWN functions (imported from /root/reference with its
unavailable imports stubbed; ``tests/golden/make_golden.py``) for the two helpers and the sa / la gates; the atrium gate's
labelling call goes to scikit-image in the reference and is therefore pinned only through this restatement.
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage


def get_largest_cc(binary: np.ndarray) -> np.ndarray:
    cc, n_cc = ndimage.label(binary)
    max_n, max_area = -1, 0
    for n in range(1, n_cc + 1):
        area = np.sum(cc == n)
        if area > max_area:
            max_area, max_n = area, n
    return cc == max_n


def remove_small_cc(binary: np.ndarray, thres: int = 10) -> np.ndarray:
    cc, n_cc = ndimage.label(binary)
    binary2 = np.copy(binary)
    for n in range(1, n_cc + 1):
        if np.sum(cc == n) < thres:
            binary2[cc == n] = 0
    return binary2


def sa_pass_quality_control(seg_sa: np.ndarray) -> bool:
    X, Y, Z = seg_sa.shape[:3]
    label = {'LV': 1, 'Myo': 2, 'RV': 3}
    for l in label.values():
        if np.sum(seg_sa == l) < 10:
            return False
    z_pos = []
    for z in range(Z):
        seg_z = seg_sa[:, :, z]
        if np.sum(seg_z == label['LV']) < 10 or np.sum(seg_z == label['Myo']) < 10:
            continue
        z_pos += [z]
    if len(z_pos) < 6:
        return False
    if len(z_pos) != (np.max(z_pos) - np.min(z_pos) + 1):
        return False
    _, _, cz = [np.mean(x) for x in np.nonzero(seg_sa == label['LV'])]
    seg_z = seg_sa[:, :, int(round(cz))]
    endo = get_largest_cc((seg_z == label['LV']).astype(np.uint8)).astype(np.uint8)
    myo = remove_small_cc((seg_z == label['Myo']).astype(np.uint8)).astype(np.uint8)
    epi = get_largest_cc((endo | myo).astype(np.uint8)).astype(np.uint8)
    rv = get_largest_cc((seg_z == label['RV']).astype(np.uint8)).astype(np.uint8)
    return not (np.sum(epi) < 10 or np.sum(rv) < 10)


def la_pass_quality_control(seg: np.ndarray) -> bool:
    seg_z = seg[:, :, 0]
    label = {'LV': 1, 'Myo': 2, 'RV': 3, 'LA': 4, 'RA': 5}
    for l in label.values():
        if np.sum(seg_z == l) < 10:
            return False
    endo = get_largest_cc((seg_z == label['LV']).astype(np.uint8)).astype(np.uint8)
    myo = remove_small_cc((seg_z == label['Myo']).astype(np.uint8)).astype(np.uint8)
    epi = get_largest_cc((endo | myo).astype(np.uint8)).astype(np.uint8)
    return not (np.sum(endo) < 10 or np.sum(myo) < 10 or np.sum(epi) < 10)


def atrium_pass_quality_control(label: np.ndarray, label_dict) -> bool:
    structure = ndimage.generate_binary_structure(3, 2)
    for l in label_dict.values():
        T = label.shape[3]
        for t in range(T):
            if np.sum(label[:, :, :, t] == l) == 0:
                return False
        for t in range(T):
            cc, n_cc = ndimage.label(label[:, :, :, t] == l, structure=structure)
            count_cc = sum(1 for i in range(1, n_cc + 1) if np.sum(cc == i) > 10)
            if count_cc >= 2:
                return False
        A = np.sum(label == l, axis=(0, 1, 2))
        for t in range(T):
            ratio = A[t] / float(A[t - 1])
            if ratio >= 2 or ratio <= 0.5:
                return False
    return True


def aorta_pass_quality_control(image: np.ndarray, seg: np.ndarray) -> bool:
    """cardiac_utils.py:1739-1795 (skimage.measure.label(connectivity=2) restated with scipy, see the module docstring)."""
    structure = ndimage.generate_binary_structure(3, 2)
    for l in (1, 2):
        T = seg.shape[3]
        for t in range(T):
            if np.sum(seg[:, :, :, t] == l) == 0:
                return False
        mean_intensity_ED = image[:, :, :, 0][seg[:, :, :, 0] == l].mean()
        for t in range(T):
            if np.max(image[:, :, :, t][seg[:, :, :, t] == l]) / mean_intensity_ED >= 3:
                return False
        for t in range(T):
            cc, n_cc = ndimage.label(seg[:, :, :, t] == l, structure=structure)
            if sum(1 for i in range(1, n_cc + 1) if np.sum(cc == i) > 10) >= 2:
                return False
        A = np.sum(seg == l, axis=(0, 1, 2))
        for t in range(T):
            ratio = A[t] / float(A[t - 1])
            if ratio >= 2 or ratio <= 0.5:
                return False
        if np.max(A) / np.min(A) >= 2:
            return False
    return True
