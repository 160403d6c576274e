"""Geometry helpers that decide WHAT gets stored/gathered (host integer math).

Mirror of the reference's src/visualization_utils.py:93-125 (crop_image, extract_coords,
extract_roi).  The display helpers of that file (:16-90) are out of scope (interactive plotting).
"""
from __future__ import annotations

import numpy as np


def crop_image(img, xmin, ymin, xmax, ymax):
    """reference: visualization_utils.py:93-98 (window clamped to the image)."""
    h, w = img.shape[0:2]
    ymin, ymax = [max(0, min(int(v), h)) for v in (ymin, ymax)]
    xmin, xmax = [max(0, min(int(v), w)) for v in (xmin, xmax)]
    return img[ymin:ymax, xmin:xmax]


def mask_bbox(mask):
    """(row_min, row_max, col_min, col_max) of the True pixels of a 2-D mask."""
    rows = np.flatnonzero(np.any(mask, axis=1))
    cols = np.flatnonzero(np.any(mask, axis=0))
    if rows.size == 0:
        raise ValueError("extract_coords: empty mask")  # reference: np.min of an empty array raises too
    return int(rows[0]), int(rows[-1]), int(cols[0]), int(cols[-1])


def coords_from_bbox(bbox, margin):
    """extract_coords expressed on the mask's bounding box (row_min, row_max, col_min, col_max) -- all the
    reference's formula uses (visualization_utils.py:101-112).  NOTE the reference SHIFTS the box by `margin`
    (rows up, columns right) and keeps extent max-min; it does not expand it.  Kept as is."""
    rmin, rmax, cmin, cmax = bbox
    ymin, xmin = rmin - margin, cmin + margin
    ymax, xmax = rmax - margin, cmax + margin
    h = max(ymax - ymin, margin)
    w = max(xmax - xmin, margin)
    return xmin, ymin, xmin + w, ymin + h


def extract_coords(mask, margin):
    """reference: visualization_utils.py:101-112."""
    return coords_from_bbox(mask_bbox(mask), margin)


def roi_window(img_hw, mask, margin=1):
    """(xmin, ymin, xmax, ymax) that extract_roi crops from an array of spatial shape img_hw,
    clamped like crop_image.  reference: visualization_utils.py:115-125."""
    return roi_window_from_bbox(img_hw, mask.shape[0:2], mask_bbox(mask), margin)


def roi_window_from_bbox(img_hw, mask_hw, bbox, margin=1):
    """roi_window given only the mask's shape and bounding box (row_min, row_max, col_min, col_max)."""
    xmin, ymin, xmax, ymax = coords_from_bbox(bbox, margin)
    if tuple(img_hw) != tuple(mask_hw):
        hs = img_hw[0] / mask_hw[0]
        ws = img_hw[1] / mask_hw[1]
        xmin, ymin, xmax, ymax = [int(v) for v in (xmin * ws, ymin * hs, xmax * ws, ymax * hs)]
        h = max(ymax - ymin, margin)
        w = max(xmax - xmin, margin)
        xmax = xmin + w
        ymax = ymin + h
    H, W = img_hw
    ymin, ymax = [max(0, min(v, H)) for v in (ymin, ymax)]
    xmin, xmax = [max(0, min(v, W)) for v in (xmin, xmax)]
    return xmin, ymin, xmax, ymax


def extract_roi(img, mask, margin=1):
    """reference: visualization_utils.py:115-125."""
    xmin, ymin, xmax, ymax = roi_window(img.shape[0:2], mask, margin)
    return img[ymin:ymax, xmin:xmax]


def crop_window(mask_3d):
    """Square crop window of generate_features (tfds_dense_descriptor.py:257-263), unclamped:
    half-side 2*max(bbox w, bbox h) about the (shifted) bbox centre of the union mask."""
    bigger = mask_3d if mask_3d.ndim == 2 else np.any(mask_3d, axis=-1)   # union over slices (:257)
    return crop_window_from_bbox(mask_bbox(bigger))


def crop_window_from_bbox(bbox):
    """crop_window given the union mask's bounding box (row_min, row_max, col_min, col_max)."""
    xmin, ymin, xmax, ymax = coords_from_bbox(bbox, margin=2)
    crop_size = max(xmax - xmin, ymax - ymin) * 2
    xmid, ymid = int(xmin + (xmax - xmin) / 2), int(ymin + (ymax - ymin) / 2)
    return xmid - crop_size, ymid - crop_size, xmid + crop_size, ymid + crop_size


# ---------------------------------------------------------------------------------------------- HU -> RGB (non-MedSAM CT input)
_AIR, _LUNG, _FAT = (0, 0, 0), (194, 105, 82), (194, 166, 115)
_SOFT_LO, _SOFT_HI, _BONE = (102, 0, 0), (153, 0, 0), (255, 255, 255)
#: (low, low_closed, high, high_closed, colour_at_ramp_start, colour_at_ramp_end or None, ramp_start, ramp_end) -- the nine HU
#: intervals of the reference (visualization_utils.py:146-184).  The soft-tissue plateau [40, 80] is ramped over (80, 400) in the
#: reference (:172-175), i.e. slightly EXTRAPOLATED below its first colour; kept as is.
_HU_SEGMENTS = (
    (-np.inf, False, -1000, True, _AIR, None, 0, 1),
    (-1000, False, -600, False, _AIR, _LUNG, -1000, -600),
    (-600, True, -400, True, _LUNG, None, 0, 1),
    (-400, False, -100, False, _LUNG, _FAT, -400, -100),
    (-100, True, -60, True, _FAT, None, 0, 1),
    (-60, False, 40, False, _FAT, _SOFT_LO, -60, 40),
    (40, True, 80, True, _SOFT_LO, _SOFT_HI, 80, 400),
    (80, False, 400, False, _SOFT_HI, _BONE, 80, 400),
    (400, True, np.inf, False, _BONE, None, 0, 1),
)


def hu_to_rgb_vectorized(hu_matrix):
    """reference: visualization_utils.py:128-186 -- tissue colour map of a CT in Hounsfield units, (…) -> (…, 3) uint8; the
    extraction divides it by 255 for the non-MedSAM backbones (tfds_dense_descriptor.py:445).  Same arithmetic as the reference
    (ramp ratio in the input's dtype, colours mixed in float64, truncation into an integer image), table-driven."""
    hu = np.asarray(hu_matrix)
    rgb = np.zeros(hu.shape + (3,), dtype=int)
    for lo, lo_closed, hi, hi_closed, c_a, c_b, r0, r1 in _HU_SEGMENTS:
        above = (hu >= lo) if lo_closed else (hu > lo)
        below = (hu <= hi) if hi_closed else (hu < hi)
        sel = above & below
        if c_b is None:
            rgb[sel] = np.array(c_a)
        else:
            ratios = (hu[sel] - r0) / (r1 - r0)
            rgb[sel] = np.array(c_a) * (1 - ratios[..., None]) + np.array(c_b) * ratios[..., None]
    return rgb.astype(np.uint8)
