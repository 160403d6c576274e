"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference's mask gathers.

Oracle for SURVEY.md section 8(a) rows G1 (feature-token gather), G2 (voxel point cloud),
G3 (3-D sinusoidal positional encoding) and the V3 geometry helpers.  Every function cites
the reference lines it follows.  Pinned against the live reference by
``tests/test_oracle_vs_reference.py`` (when /root/reference is present) and against the
frozen vectors in ``tests/golden/`` (always).  Nothing in the product package imports this.

The restatements are written as explicit index arithmetic (not as a transliteration of the
reference's meshgrid/flatten/fancy-index code) so that they document the exact contract the
CUDA kernels implement: see SURVEY.md Appendix A.
"""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------- G3
def positional_encoding_3d(x, y, z, D, scale=10000):
    """reference: src/train_models.py:30-44 (float64)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    z = np.asarray(z, dtype=np.float64)
    n = x.shape[0]
    enc = np.zeros((n, D), dtype=np.float64)
    third, two_third = D // 3, 2 * D // 3
    for i in range(D // 6):
        e = scale ** (6 * i / D)
        for base, v in ((0, x), (third, y), (two_third, z)):
            enc[:, 2 * i + base] = np.sin(v / e)
            enc[:, 2 * i + 1 + base] = np.cos(v / e)
    return enc


# ----------------------------------------------------------------------------- mask resize
def nearest_index_map(n_out: int, n_in: int) -> np.ndarray:
    """Source index for every output index of an order-0 resize with pixel-centre sampling.

    reference: src/train_models.py:151 ``resize(mask, shape, order=0)``; skimage >= 0.19 maps
    this to ``scipy.ndimage.zoom(..., order=0, mode='mirror', grid_mode=True)``.  ndimage
    evaluates output index ``o`` at input coordinate ``c = (o + 0.5) * (n_in / n_out) - 0.5``
    (the ratio is rounded to a double first, then multiplied) and order 0 picks
    ``floor(c + 0.5)``.  Checked element-for-element against scipy for all sizes 1..69.
    """
    ratio = np.float64(n_in) / np.float64(n_out)
    o = np.arange(n_out, dtype=np.float64)
    c = (o + 0.5) * ratio - 0.5
    idx = np.floor(c + 0.5).astype(np.int64)
    return np.clip(idx, 0, n_in - 1)


def resize_mask_nearest(mask: np.ndarray, out_hw) -> np.ndarray:
    """reference: src/train_models.py:151 (bool in, bool out)."""
    h, w = int(out_hw[0]), int(out_hw[1])
    r = nearest_index_map(h, mask.shape[0])
    c = nearest_index_map(w, mask.shape[1])
    return np.asarray(mask)[np.ix_(r, c)].astype(bool)


# ----------------------------------------------------------------------------- G1
def token_gather(features, masks, spatial_res, noise=(0.0, 0.0, 0.0), feature_dim=None):
    """Feature-token gather of ``PETCTDataset3D._get_features`` for arch == 'transformer'.

    reference: src/train_models.py:143-182.

    features : list of S arrays (h, w, D)       (HDF5 ``<pid>/features/<i>``)
    masks    : list of S bool arrays (h_m, w_m) (HDF5 ``<pid>/masks/<i>``)
    Returns a dict with
      tokens   (n_sel, D) float64  = gathered features + PE/4         (:180)
      raw      (n_sel, D)          = gathered features only
      flat     (n_sel,)  int64     flat index n = a*(w*S) + b*S + k   (Appendix A1)
      src      (n_sel, 3) int32    (slice k, row a, col b) of each token
      xyz      (n_sel, 3) float64  the reference's centred physical coordinates (:169-176)
    """
    S = len(features)
    h, w, D = features[0].shape
    if feature_dim is None:
        feature_dim = D
    # :151  nearest resize of every pixel mask to the feature grid; :161 stack to (h, w, S)
    m = np.stack([resize_mask_nearest(mk, (h, w)) for mk in masks], axis=-1)  # (h, w, S)
    f = np.stack(features, axis=2)                                           # (h, w, S, D)  (:159,163)
    h_orig, w_orig = masks[-1].shape[0:2]                                     # :162 (LAST slice)
    n = np.arange(h * w * S, dtype=np.int64)
    # :166-168 meshgrid(indexing='xy') then flatten -> arrays of shape (w, h, S): Appendix A2
    xi = (n // S) % h
    yi = n // (h * S)
    zi = n % S
    x = (xi / w) * w_orig * spatial_res[0]                                    # :169  (w_new = w)
    y = (yi / h) * h_orig * spatial_res[1]                                    # :170  (h_new = h)
    z = zi * spatial_res[2]                                                   # :171
    mflat = m.reshape(-1)                                                     # :173
    x = (x - x.mean() + noise[0])[mflat]                                      # :174
    y = (y - y.mean() + noise[1])[mflat]
    z = (z - z.mean() + noise[2])[mflat]
    pe = positional_encoding_3d(x, y, z, D=feature_dim)                       # :178
    raw = f.reshape(-1, feature_dim)[mflat, :]
    tokens = raw + pe / 4                                                     # :180
    flat = n[mflat]
    src = np.stack([flat % S, flat // (w * S), (flat // S) % w], axis=1).astype(np.int32)
    return dict(tokens=tokens, raw=raw, flat=flat, src=src, xyz=np.stack([x, y, z], axis=1))


def conv_features(features, masks):
    """arch == 'conv' branch of ``_get_features``: features * resized mask, returned as
    (D, S, h, w).  reference: src/train_models.py:153-159."""
    out = []
    for f, mk in zip(features, masks):
        h, w = f.shape[0:2]
        out.append(f * resize_mask_nearest(mk, (h, w))[..., None])
    return np.transpose(np.stack(out, axis=0), (3, 0, 1, 2))


# ----------------------------------------------------------------------------- G2
def voxel_pointcloud(img, mask, spatial_res):
    """Columns of ``to_pointcloud_df``.  reference: src/create_pointcloud_dataframe.py:15-31.

    Returns dict of flat arrays over all H*W*S voxels: x, y, z (f64), raw, mask, mask_box.
    ``meshgrid`` default indexing='xy' gives arrays of shape (W, H, S), so after flatten
    x[n] = (n // S) % H, y[n] = n // (H*S), z[n] = n % S while raw/mask are flattened
    row-major over (H, W, S)  (Appendix A6: aligned only when H == W).
    """
    H, W, S = img.shape
    n = np.arange(H * W * S, dtype=np.int64)
    x = ((n // S) % H) * spatial_res[0]                                       # :16-20
    y = (n // (H * S)) * spatial_res[1]                                       # :21
    z = (n % S) * spatial_res[2]                                              # :22
    raw = np.asarray(img).reshape(-1)                                         # :23
    mk = np.asarray(mask).reshape(-1)                                         # :24
    sel = mk > 0                                                              # :26
    if sel.any():
        lo = [x[sel].min(), y[sel].min(), z[sel].min()]
        hi = [x[sel].max(), y[sel].max(), z[sel].max()]
        box = ((x >= lo[0]) & (x <= hi[0]) & (y >= lo[1]) & (y <= hi[1])
               & (z >= lo[2]) & (z <= hi[2]))                                 # :27-30
    else:
        box = np.zeros_like(sel)  # pandas: min/max of empty -> NaN -> all comparisons False
    return dict(x=x, y=y, z=z, raw=raw, mask=mk, mask_box=box)


def voxel_pointcloud_box(img, mask, spatial_res):
    """What the caller keeps: rows with mask_box, coordinates centred by the kept rows' mean.
    reference: src/create_pointcloud_dataframe.py:78-81."""
    pc = voxel_pointcloud(img, mask, spatial_res)
    keep = pc["mask_box"]
    out = {k: v[keep] for k, v in pc.items()}
    out["flat"] = np.nonzero(keep)[0].astype(np.int64)
    for k in ("x", "y", "z"):
        if out[k].size:
            out[k] = out[k] - out[k].mean()
    return out


# ----------------------------------------------------------------------------- V3 geometry
def crop_image(img, xmin, ymin, xmax, ymax):
    """reference: src/visualization_utils.py:93-98."""
    h, w = img.shape[0:2]
    ymin, ymax = [max(0, min(v, h)) for v in (ymin, ymax)]
    xmin, xmax = [max(0, min(v, w)) for v in (xmin, xmax)]
    return img[ymin:ymax, xmin:xmax]


def extract_coords(mask, margin):
    """reference: src/visualization_utils.py:101-112 -- the box is SHIFTED by the margin
    (up and to the right), not expanded (Appendix A5)."""
    rows, cols = np.nonzero(mask)
    ymin = int(rows.min()) - margin
    xmin = int(cols.min()) + margin
    ymax = int(rows.max()) - margin
    xmax = int(cols.max()) + margin
    h = max(ymax - ymin, margin)
    w = max(xmax - xmin, margin)
    return xmin, ymin, xmin + w, ymin + h


def extract_roi(img, mask, margin=1):
    """reference: src/visualization_utils.py:115-125."""
    xmin, ymin, xmax, ymax = extract_coords(mask, margin)
    if img.shape[0:2] != mask.shape[0:2]:
        hs = img.shape[0] / mask.shape[0]
        ws = img.shape[1] / mask.shape[1]
        xmin, ymin, xmax, ymax = [int(v) for v in (xmin * ws, ymin * hs, xmax * ws, ymax * hs)]
        h = max(ymax - ymin, margin)
        w = max(xmax - xmin, margin)
        xmax = xmin + w
        ymax = ymin + h
    return crop_image(img, xmin, ymin, xmax, ymax)


def crop_window(mask_3d):
    """Square crop window of ``generate_features``.
    reference: src/tfds_dense_descriptor.py:257-263.  Returns (xmin, ymin, xmax, ymax), unclamped."""
    bigger = np.sum(mask_3d, axis=-1) > 0
    xmin, ymin, xmax, ymax = extract_coords(bigger, margin=2)
    crop_size = max(xmax - xmin, ymax - ymin) * 2
    xmid, ymid = int(xmin + (xmax - xmin) / 2), int(ymin + (ymax - ymin) / 2)
    return xmid - crop_size, ymid - crop_size, xmid + crop_size, ymid + crop_size


def generate_features(descriptor_fn, img_3d, mask_3d):
    """reference: src/tfds_dense_descriptor.py:242-284 with the per-slice backbone call
    (:277) abstracted as ``descriptor_fn(img2d) -> (h_f, w_f, D)``."""
    bigger = np.sum(mask_3d, axis=-1) > 0
    win = crop_window(mask_3d)
    img_3d = crop_image(img_3d, *win)
    mask_3d = crop_image(mask_3d, *win)
    bigger = crop_image(bigger, *win)
    feats, masks = [], []
    for s in range(img_3d.shape[2]):
        f = descriptor_fn(img_3d[:, :, s])
        feats.append(extract_roi(f, bigger))
        masks.append(extract_roi(mask_3d[:, :, s] > 0, bigger))
    return feats, masks


def apply_window_ct(ct, width, level):
    """reference: src/tfds_dense_descriptor.py:231-233, 287-303."""
    lo, hi = level - width / 2, level + width / 2
    return np.clip((ct - lo) / (hi - lo), 0, 1)
