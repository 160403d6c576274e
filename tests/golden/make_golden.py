"""Freeze golden vectors by running the UNMODIFIED reference (via oracle/ref_shim.py).

Run once in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

The reference ships no tests, fixtures or golden vectors, so these
files are the pin for the oracle: outputs of the reference's own functions on seeded
synthetic inputs.  The .npz files are small and committed; this script is how they were made.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def ellipsoid_mask(shape, centre, radii):
    zz = np.indices(shape).astype(np.float64)
    d = sum(((zz[i] - centre[i]) / radii[i]) ** 2 for i in range(3))
    return d <= 1.0


def golden_gather():
    tm = ref_shim.load_reference("train_models")

    class DS(tm.PETCTDataset3D):  # bypass __init__ (needs a dataframe); _get_features is unmodified
        def __init__(self, D, arch):
            self.feature_dim, self.arch = D, arch

    rng = np.random.default_rng(20241018)
    cases = {}
    specs = [  # S, h, w, hm, wm, D, density
        ("square", 5, 6, 6, 20, 20, 12, 0.3),
        ("ragged", 7, 18, 15, 61, 47, 48, 0.25),
        ("wide", 3, 4, 9, 13, 30, 24, 0.5),
        ("probe", 6, 9, 7, 30, 22, 256, 0.2),
        ("empty", 4, 5, 5, 16, 16, 12, 0.0),
        ("full", 2, 3, 4, 9, 12, 6, 1.1),
        ("single", 1, 8, 8, 32, 32, 18, 0.4),
    ]
    for name, S, h, w, hm, wm, D, dens in specs:
        feats = [rng.standard_normal((h, w, D)).astype(np.float32) for _ in range(S)]
        masks = [rng.random((hm, wm)) < dens for _ in range(S)]
        res = rng.uniform(0.5, 1.5, 3)
        noise = rng.uniform(-5, 5, 3)
        path = f"golden_{name}.h5"
        ref_shim.H5_FILES.pop(path, None)
        ref_shim.put_feature_file(path, "pid", feats, masks)
        out_t = DS(D, "transformer")._get_features(path, "pid", list(range(S)), 0, "None", noise, res)
        out_c = DS(D, "conv")._get_features(path, "pid", list(range(S)), 0, "None", noise, res)
        cases[f"{name}__features"] = np.stack(feats)
        cases[f"{name}__masks"] = np.stack(masks)
        cases[f"{name}__res"] = res
        cases[f"{name}__noise"] = noise
        cases[f"{name}__out_transformer"] = out_t
        cases[f"{name}__out_conv"] = out_c.astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "gather_g1.npz"), **cases)
    print("gather_g1.npz", len(specs), "cases")


def golden_pointcloud():
    pc = ref_shim.load_reference("create_pointcloud_dataframe")
    rng = np.random.default_rng(7)
    cases = {}
    for name, shape, c, r in [("sq", (16, 16, 6), (8, 7, 3), (4, 3, 2)),
                              ("rect", (12, 20, 5), (6, 11, 2), (3, 6, 1.5)),
                              ("edge", (10, 10, 4), (0, 9, 0), (3, 3, 1)),
                              ("empty", (6, 6, 3), (3, 3, 1), (0.1, 0.1, 0.1))]:
        img = rng.normal(-300, 350, shape).astype(np.float32)
        mask = ellipsoid_mask(shape, c, r)
        if name == "empty":
            mask[:] = False
        res = rng.uniform(0.5, 2.0, 3)
        df = pc.to_pointcloud_df(img, mask, 1, res)
        cases[f"{name}__img"] = img
        cases[f"{name}__mask"] = mask
        cases[f"{name}__res"] = res
        for col in ("x", "y", "z", "raw", "mask", "mask_box"):
            cases[f"{name}__{col}"] = df[col].values
    np.savez_compressed(os.path.join(OUT, "pointcloud_g2.npz"), **cases)
    print("pointcloud_g2.npz")


def golden_geometry():
    vu = ref_shim.load_reference("visualization_utils")
    td = ref_shim.load_reference("tfds_dense_descriptor")
    tm = ref_shim.load_reference("train_models")
    rng = np.random.default_rng(11)
    cases = {}
    # extract_coords / extract_roi
    masks, coords1, coords2, roi_shapes = [], [], [], []
    for _ in range(12):
        H, W = 48, 40
        m = np.zeros((H, W), bool)
        r0, c0 = rng.integers(0, H - 4), rng.integers(0, W - 4)
        r1, c1 = rng.integers(r0 + 1, H), rng.integers(c0 + 1, W)
        m[r0:r1, c0:c1] = True
        masks.append(m)
        coords1.append([int(v) for v in vu.extract_coords(m, 1)])
        coords2.append([int(v) for v in vu.extract_coords(m, 2)])
        img = np.zeros((H // 4, W // 4, 3))
        roi_shapes.append(list(vu.extract_roi(img, m).shape[:2]) + list(vu.extract_roi(m, m).shape[:2]))
    cases["geo__masks"] = np.stack(masks)
    cases["geo__coords_m1"] = np.array(coords1)
    cases["geo__coords_m2"] = np.array(coords2)
    cases["geo__roi_shapes"] = np.array(roi_shapes)

    # generate_features with a stand-in backbone (4x4 average pooling -> 5 channels)
    def fake_descriptor(model, img):
        h, w = img.shape[0] // 4, img.shape[1] // 4
        f = img[:h * 4, :w * 4].reshape(h, 4, w, 4).mean(axis=(1, 3))
        return np.stack([f * (k + 1) for k in range(5)], axis=-1)

    td.get_dense_descriptor = fake_descriptor

    class M:
        model_name = "medsam"

    img3 = rng.random((96, 96, 6))
    m3 = ellipsoid_mask((96, 96, 6), (50, 44, 3), (9, 13, 2.5))
    fl, ml = td.generate_features(M(), img3, m3, "golden")
    cases["gen__img"] = img3
    cases["gen__mask"] = m3
    cases["gen__features"] = np.stack(fl)
    cases["gen__masks"] = np.stack(ml)
    # PE + window
    x, y, z = rng.uniform(-40, 40, (3, 9))
    for D in (12, 256, 384):
        cases[f"pe__{D}"] = tm.positional_encoding_3d(x, y, z, D)
    cases["pe__xyz"] = np.stack([x, y, z])
    hu = rng.normal(-300, 350, (8, 8))
    cases["win__hu"] = hu
    cases["win__out"] = td.apply_window_ct(hu, 800, 40)
    np.savez_compressed(os.path.join(OUT, "geometry.npz"), **cases)
    print("geometry.npz")


def golden_classifier():
    ma = ref_shim.load_reference("models_archs")
    tm = ref_shim.load_reference("train_models")
    torch.manual_seed(99)
    d, ff, heads, layers = 64, 128, 1, 2   # head_dim 64 like every real config
    model = ma.TransformerNoduleClassifier(d, ff, heads, 2, layers).eval()
    # make every parameter non-trivial (LN weights/biases are 1/0 at init)
    with torch.no_grad():
        for n_, p in model.named_parameters():
            if "norm" in n_:
                p.add_(0.05 * torch.randn_like(p))
            elif n_.endswith("bias"):
                p.add_(0.02 * torch.randn_like(p))
    x = torch.randn(1, 37, d)
    y = torch.tensor([[0.0, 1.0]])
    crit = tm.FocalLoss(alpha=torch.tensor([0.25, 0.75]), gamma=2)
    logits, cls = model(x)
    loss = crit(torch.squeeze(logits), torch.squeeze(y))
    loss.backward()
    cases = {"x": x.numpy(), "y": y.numpy(), "logits": logits.detach().numpy(),
             "cls": cls.detach().numpy(), "loss": loss.detach().numpy(),
             "cfg": np.array([d, ff, heads, layers])}
    for n_, p in model.named_parameters():
        cases["param__" + n_] = p.detach().numpy()
        cases["grad__" + n_] = p.grad.detach().numpy()
    # focal loss table
    lg = torch.randn(6, 2) * 2
    tg = F_onehot = torch.eye(2)[torch.tensor([0, 1, 1, 0, 1, 0])]
    cases["focal__logits"] = lg.numpy()
    cases["focal__targets"] = tg.numpy()
    cases["focal__loss_alpha"] = crit(lg, tg).numpy()
    cases["focal__loss_noalpha"] = tm.FocalLoss(gamma=2)(lg, tg).numpy()
    cases["focal__loss_single"] = np.array([crit(lg[i], tg[i]).item() for i in range(6)])
    np.savez_compressed(os.path.join(OUT, "classifier_small.npz"), **cases)
    print("classifier_small.npz")


def golden_bimodal():
    """TransformerNoduleBimodalClassifier (models_archs.py:38-124), eval mode: outputs and gradients for both modalities and
    for each modality alone; loss = sum of the three logit heads' focal losses + 0.1 * sum(petct_cls) so every path has gradient."""
    ma = ref_shim.load_reference("models_archs")
    tm = ref_shim.load_reference("train_models")
    torch.manual_seed(123)
    d = 64
    model = ma.TransformerNoduleBimodalClassifier(d, 2, 2, 1, 1, 1, 2, 2).eval()
    with torch.no_grad():
        for n_, p in model.named_parameters():
            if "norm" in n_:
                p.add_(0.05 * torch.randn_like(p))
            elif n_.endswith("bias"):
                p.add_(0.02 * torch.randn_like(p))
    x_ct, x_pet = torch.randn(1, 29, d), torch.randn(1, 41, d)
    y = torch.tensor([1.0, 0.0])
    crit = tm.FocalLoss(alpha=torch.tensor([0.25, 0.75]), gamma=2)
    cases = {"x_ct": x_ct.numpy(), "x_pet": x_pet.numpy(), "y": y.numpy(), "cfg": np.array([d, 2, 2, 1, 1, 1, 2, 2])}
    for n_, p in model.named_parameters():
        cases["param__" + n_] = p.detach().numpy()
    for mode, (a, b) in {"both": (x_ct, x_pet), "ct": (x_ct, None), "pet": (None, x_pet)}.items():
        model.zero_grad()
        lg, z, lg_ct, lg_pet = model(a, b)
        loss = crit(torch.squeeze(lg), y) + crit(torch.squeeze(lg_ct), y) + crit(torch.squeeze(lg_pet), y) + 0.1 * z.sum()
        loss.backward()
        cases[f"{mode}__logits"], cases[f"{mode}__z"] = lg.detach().numpy(), z.detach().numpy()
        cases[f"{mode}__logits_ct"], cases[f"{mode}__logits_pet"] = lg_ct.detach().numpy(), lg_pet.detach().numpy()
        cases[f"{mode}__loss"] = loss.detach().numpy()
        for n_, p in model.named_parameters():
            if p.grad is not None:
                cases[f"{mode}__grad__" + n_] = p.grad.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, "bimodal_small.npz"), **cases)
    print("bimodal_small.npz")


def golden_crossmodal():
    """CrossModalFocalLoss (train_models.py:332-378): batch and single-sample values plus logit gradients, for the default
    constructor and for the training script's setting (:593-596)."""
    tm = ref_shim.load_reference("train_models")
    torch.manual_seed(77)
    lx, lc, lp = (torch.randn(6, 2) * 2 for _ in range(3))
    tg = torch.eye(2)[torch.tensor([1, 0, 1, 1, 0, 0])]
    cases = {"lx": lx.numpy(), "lc": lc.numpy(), "lp": lp.numpy(), "targets": tg.numpy()}
    for tag, kw in (("default", {}), ("train", dict(alpha=torch.tensor([0.25, 0.75]), gamma_unimodal=2.0, gamma_bimodal=1.0, beta=0.6))):
        crit = tm.CrossModalFocalLoss(**kw)
        a, b, c = (t.clone().requires_grad_(True) for t in (lx, lc, lp))
        loss = crit(a, b, c, tg)
        loss.backward()
        cases[f"{tag}__loss"] = loss.detach().numpy()
        for n_, t in (("gx", a), ("gc", b), ("gp", c)):
            cases[f"{tag}__{n_}"] = t.grad.numpy()
        cases[f"{tag}__single"] = np.array([crit(lx[i], lc[i], lp[i], tg[i]).item() for i in range(6)])
    np.savez_compressed(os.path.join(OUT, "crossmodal_loss.npz"), **cases)
    print("crossmodal_loss.npz")


if __name__ == "__main__":
    assert ref_shim.reference_available(), "needs /root/reference"
    only = sys.argv[1:]
    for fn in (golden_gather, golden_pointcloud, golden_geometry, golden_classifier, golden_bimodal, golden_crossmodal):
        if not only or fn.__name__ in only:
            fn()
