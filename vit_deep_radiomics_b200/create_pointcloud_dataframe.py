"""Raw-voxel point cloud -- drop-in for the reference's src/create_pointcloud_dataframe.py.

``to_pointcloud_df`` keeps the reference signature and returns the same DataFrame (all voxels with
x, y, z, raw, mask, mask_box).  ``pointcloud_box`` is what the reference's caller keeps
(:78-81: rows inside the mask's bounding box, coordinates centred) computed on the GPU:
a bounding-box reduction kernel followed by a closed-form box gather (libvdr G2 kernels).
"""
from __future__ import annotations

import numpy as np
import pandas as pd
import torch

from . import ops


def _to_dev(img, mask, device):
    img_t = torch.as_tensor(np.ascontiguousarray(img, dtype=np.float32)).to(device)
    m = np.ascontiguousarray(mask)
    m = (m > 0).view(np.uint8)
    return img_t, torch.as_tensor(m).to(device)


def _gather_box(img, mask, spatial_res, device):
    H, W, S = img.shape
    res = np.asarray(spatial_res, dtype=np.float64)
    img_t, mask_t = _to_dev(img, mask, device)
    bbox = ops.voxel_bbox(mask_t)
    bb = bbox.cpu().numpy().astype(np.int64)                       # 24-byte D2H: sizes the output
    if bb[1] < bb[0]:                                              # empty mask: no rows (pandas: all False)
        z = np.zeros(0)
        return bb, dict(flat=np.zeros(0, np.int64), raw=np.zeros(0, np.float32), mask=np.zeros(0, bool), x=z, y=z, z=z)
    # the reference compares PHYSICAL coordinates (index * res, :26-30); for res > 0 that is the same
    # predicate as the index-space box.  A zero resolution collapses an axis -> every index passes.
    lim = [H, W, S]      # xi < H, yi < W (xy-meshgrid convention), zi < S
    for a in range(3):
        if res[a] == 0:
            bb[2 * a], bb[2 * a + 1] = 0, lim[a] - 1
    if (res == 0).any():
        bbox = torch.as_tensor(bb.astype(np.int32)).to(device)
    cap = int((bb[1] - bb[0] + 1) * (bb[3] - bb[2] + 1) * (bb[5] - bb[4] + 1))
    flat, raw, mk, count = ops.voxel_gather(img_t, mask_t, bbox, cap)
    flat = flat.cpu().numpy().astype(np.int64)
    q = flat // S
    out = dict(flat=flat, raw=raw.cpu().numpy(), mask=mk.cpu().numpy().astype(bool),
               x=(q % H) * res[0], y=(q // H) * res[1], z=(flat % S) * res[2])   # :16-22 (xy meshgrid quirk)
    return bb, out


def pointcloud_box(img, mask, spatial_res, device="cuda:0", centre=True) -> pd.DataFrame:
    """Rows of ``to_pointcloud_df(...)`` with mask_box == True, in the same order, with x/y/z centred by the
    mean of the kept rows when ``centre`` (reference caller, create_pointcloud_dataframe.py:78-81)."""
    _, o = _gather_box(img, mask, spatial_res, device)
    df = pd.DataFrame({"x": o["x"], "y": o["y"], "z": o["z"], "raw": o["raw"], "mask": o["mask"]})
    df["mask_box"] = True
    if centre and len(df):
        df[["x", "y", "z"]] = df[["x", "y", "z"]] - df[["x", "y", "z"]].mean(axis=0)
    return df


def patient_pointcloud(img_raw, mask_raw, label, spatial_res, modality, dataset_name, patient_id, device="cuda:0") -> pd.DataFrame:
    """One (patient, modality) of the reference script's loop body (create_pointcloud_dataframe.py:67-82): the voxels inside the
    mask's bounding box with their raw and normalised values (CT: lung window 800 / 40 mapped to 0..1, PET: divided by the
    maximum), the metadata columns, coordinates centred on the kept rows.  Only the box is materialised (device G2 kernels);
    the reference builds the table of ALL voxels first and filters it.
    Columns, in the reference's order: x, y, z, raw, mask, mask_box, modality, norm, dataset, patient_id, label."""
    from .tfds_dense_descriptor import apply_window_ct
    img_raw = np.asarray(img_raw)
    _, o = _gather_box(img_raw, mask_raw, spatial_res, device)
    norm = apply_window_ct(img_raw, width=800, level=40) if modality == "ct" else img_raw / img_raw.max()
    df = pd.DataFrame({"x": o["x"], "y": o["y"], "z": o["z"], "raw": o["raw"], "mask": o["mask"]})
    df["mask_box"] = True
    df["modality"] = modality
    df["norm"] = norm.flatten()[o["flat"]]
    df["dataset"] = dataset_name.replace("_dataset", "")
    df["patient_id"] = patient_id
    df["label"] = label
    if len(df):
        df[["x", "y", "z"]] = df[["x", "y", "z"]] - df[["x", "y", "z"]].mean(axis=0)
    return df


def to_pointcloud_df(img, mask, label, spatial_res, device="cuda:0") -> pd.DataFrame:
    """reference: create_pointcloud_dataframe.py:15-31 (same columns, same row order, all voxels).
    The mask_box column comes from the device bounding-box kernel; x/y/z are index * spatial_res."""
    img, mask = np.asarray(img), np.asarray(mask)
    H, W, S = img.shape
    res = np.asarray(spatial_res, dtype=np.float64)
    bb = _bbox_only(mask, device)
    n = np.arange(H * W * S, dtype=np.int64)
    q = n // S
    xi, yi, zi = q % H, q // H, n % S
    df = pd.DataFrame()
    df["x"] = xi * res[0]
    df["y"] = yi * res[1]
    df["z"] = zi * res[2]
    df["raw"] = img.reshape(-1)
    df["mask"] = mask.reshape(-1)
    if bb[1] < bb[0]:
        df["mask_box"] = False
    else:
        lo = [bb[0] * res[0], bb[2] * res[1], bb[4] * res[2]]
        hi = [bb[1] * res[0], bb[3] * res[1], bb[5] * res[2]]
        df["mask_box"] = ((df["x"] >= lo[0]) & (df["x"] <= hi[0]) & (df["y"] >= lo[1]) & (df["y"] <= hi[1])
                          & (df["z"] >= lo[2]) & (df["z"] <= hi[2]))
    return df


def _bbox_only(mask, device):
    m = np.ascontiguousarray(np.asarray(mask) > 0).view(np.uint8)
    return ops.voxel_bbox(torch.as_tensor(m).to(device)).cpu().numpy().astype(np.int64)
