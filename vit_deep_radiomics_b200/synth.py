"""Seeded synthetic inputs for BASELINE.json's configs: CT volumes with an
ellipsoid tumour mask, patient point clouds, k-fold files in the reference's YAML schema.
No dataset or checkpoint is reachable offline, so every test and bench line uses these."""
from __future__ import annotations

import numpy as np

CONFIGS = {
    # name: (model, H, W, S, ellipsoid centre, radii)
    "C1": dict(model="vit_s16", shape=(224, 224, 8), centre=(112, 112, 4), radii=(40, 32, 3)),
    "C2": dict(model="vit_b16", shape=(512, 512, 120), centre=(256, 256, 60), radii=(96, 80, 40)),
    "C4": dict(model="vit_l14", shape=(224, 224, 16), centre=(112, 112, 8), radii=(48, 40, 5)),
    # tiny variant for smoke tests (same code path, ViT-T/16)
    "T0": dict(model="vit_t16", shape=(64, 64, 4), centre=(32, 30, 2), radii=(12, 9, 1.5)),
}


def ellipsoid_mask(shape, centre, radii) -> np.ndarray:
    ax = [np.arange(n, dtype=np.float64) for n in shape]
    g = np.meshgrid(*ax, indexing="ij", sparse=True)
    d = sum(((g[i] - centre[i]) / radii[i]) ** 2 for i in range(3))
    return d <= 1.0


def ct_volume(shape, seed: int) -> np.ndarray:
    """HU = clip(N(-300, 350), -1024, 1500) -> window (width 800, level 40) -> float32 in 0..1."""
    rng = np.random.default_rng(seed)
    hu = np.clip(rng.normal(-300.0, 350.0, size=shape), -1024, 1500)
    lo, hi = 40 - 800 / 2, 40 + 800 / 2
    return np.clip((hu - lo) / (hi - lo), 0, 1).astype(np.float32)


def make_case(name: str, seed: int | None = None):
    """(img (H,W,S) f32 in 0..1, mask (H,W,S) bool, spatial_res (3,), model name)."""
    c = CONFIGS[name]
    if seed is None:
        seed = 1234 + sorted(CONFIGS).index(name)
    img = ct_volume(c["shape"], seed)
    mask = ellipsoid_mask(c["shape"], c["centre"], c["radii"])
    return img, mask, np.array([0.8, 0.8, 0.8]), c["model"]


def point_cloud_patients(n_patients=200, d=256, n_range=(512, 4096), seed=1236):
    """C3: per patient a (n, d) float32 token cloud ~ N(0,1) + 0.1*label and a Bernoulli(0.3) label."""
    rng = np.random.default_rng(seed)
    labels = (rng.random(n_patients) < 0.3).astype(np.int64)
    sizes = rng.integers(n_range[0], n_range[1] + 1, n_patients)
    ids = [f"syn_{i:03d}" for i in range(n_patients)]

    def cloud(i):
        r = np.random.default_rng(seed * 1000 + i)
        return (r.standard_normal((int(sizes[i]), d)) + 0.1 * labels[i]).astype(np.float32)

    return ids, labels, sizes, cloud


def kfold_yaml_dict(ids, labels, modality="ct", dataset="stanford", n_splits=5):
    """Folds in the reference's parameters_kfold.yaml schema (kfold_patients.<mod>.<dataset>.<k>.{train,test},
    int fold keys), split like src/split_patients.py:22-43: StratifiedKFold(5, shuffle=True, random_state=42)."""
    from sklearn.model_selection import StratifiedKFold
    skf = StratifiedKFold(n_splits=n_splits, shuffle=True, random_state=42)
    ids = np.asarray(ids)
    folds = {}
    for k, (tr, te) in enumerate(skf.split(ids, labels)):
        folds[k] = {"train": ids[tr].tolist(), "test": ids[te].tolist()}
    return {"kfold_patients": {modality: {dataset: folds}}}
