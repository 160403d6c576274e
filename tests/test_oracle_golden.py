"""The oracle restatements against the golden vectors frozen from the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import classifier_fp32 as C
from oracle import gather_np as G


@pytest.fixture(scope="module")
def g1(golden_dir):
    return np.load(os.path.join(golden_dir, "gather_g1.npz"))


CASES = ["square", "ragged", "wide", "probe", "empty", "full", "single"]


@pytest.mark.parametrize("name", CASES)
def test_token_gather_bit_exact(g1, name):
    feats, masks = list(g1[f"{name}__features"]), list(g1[f"{name}__masks"])
    o = G.token_gather(feats, masks, g1[f"{name}__res"], g1[f"{name}__noise"])
    want = g1[f"{name}__out_transformer"]
    assert o["tokens"].shape == want.shape
    assert np.array_equal(o["tokens"], want)            # float64, bit-exact
    # the integer contract: flat order ascending, (slice,row,col) consistent with it
    S, h, w = len(feats), feats[0].shape[0], feats[0].shape[1]
    assert np.all(np.diff(o["flat"]) > 0)
    assert np.array_equal(o["flat"], o["src"][:, 1].astype(np.int64) * (w * S) + o["src"][:, 2] * S + o["src"][:, 0])


@pytest.mark.parametrize("name", CASES)
def test_conv_branch(g1, name):
    out = G.conv_features(list(g1[f"{name}__features"]), list(g1[f"{name}__masks"]))
    assert np.array_equal(out.astype(np.float32), g1[f"{name}__out_conv"])


@pytest.mark.parametrize("name", ["sq", "rect", "edge", "empty"])
def test_voxel_pointcloud(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "pointcloud_g2.npz"))
    o = G.voxel_pointcloud(g[f"{name}__img"], g[f"{name}__mask"], g[f"{name}__res"])
    for col in ("x", "y", "z", "raw", "mask", "mask_box"):
        assert np.array_equal(o[col], g[f"{name}__{col}"]), col


def test_geometry(golden_dir):
    g = np.load(os.path.join(golden_dir, "geometry.npz"))
    for i, m in enumerate(g["geo__masks"]):
        assert list(G.extract_coords(m, 1)) == g["geo__coords_m1"][i].tolist()
        assert list(G.extract_coords(m, 2)) == g["geo__coords_m2"][i].tolist()
        img = np.zeros((m.shape[0] // 4, m.shape[1] // 4, 3))
        got = list(G.extract_roi(img, m).shape[:2]) + list(G.extract_roi(m, m).shape[:2])
        assert got == g["geo__roi_shapes"][i].tolist()

    def fake(img):
        h, w = img.shape[0] // 4, img.shape[1] // 4
        f = img[:h * 4, :w * 4].reshape(h, 4, w, 4).mean(axis=(1, 3))
        return np.stack([f * (k + 1) for k in range(5)], axis=-1)

    fl, ml = G.generate_features(fake, g["gen__img"], g["gen__mask"])
    assert np.array_equal(np.stack(fl), g["gen__features"]) and np.array_equal(np.stack(ml), g["gen__masks"])
    x, y, z = g["pe__xyz"]
    for D in (12, 256, 384):
        assert np.array_equal(G.positional_encoding_3d(x, y, z, D), g[f"pe__{D}"])
    assert sorted(np.flatnonzero(np.abs(G.positional_encoding_3d(x, y, z, 256)).sum(0) == 0).tolist()) == [84, 169, 254, 255]
    assert np.array_equal(G.apply_window_ct(g["win__hu"], 800, 40), g["win__out"])


def test_nearest_map_matches_scipy():
    from scipy import ndimage
    for n_in in range(1, 40):
        for n_out in range(1, 40):
            ref = ndimage.zoom(np.arange(n_in), n_out / n_in, order=0, mode="mirror", grid_mode=True)
            assert np.array_equal(ref, G.nearest_index_map(n_out, n_in)), (n_in, n_out)


def test_classifier_forward_backward_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "classifier_small.npz"))
    d, ff, heads, layers = g["cfg"].tolist()
    sd = {k[len("param__"):]: torch.tensor(g[k], requires_grad=True) for k in g.files if k.startswith("param__")}
    x, y = torch.tensor(g["x"]), torch.tensor(g["y"])
    logits, cls = C.classifier_forward(sd, x, heads, layers)
    assert torch.allclose(logits, torch.tensor(g["logits"]), atol=2e-6, rtol=1e-5)
    assert torch.allclose(cls, torch.tensor(g["cls"]), atol=2e-6, rtol=1e-5)
    loss = C.focal_loss(logits[0], y[0], gamma=2.0, alpha=torch.tensor([0.25, 0.75]))
    assert abs(loss.item() - float(g["loss"])) < 1e-6
    loss.backward()
    for k, p in sd.items():
        want = torch.tensor(g["grad__" + k])
        assert torch.allclose(p.grad, want, atol=3e-6, rtol=1e-4), k


def test_focal_loss_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "classifier_small.npz"))
    lg, tg = torch.tensor(g["focal__logits"]), torch.tensor(g["focal__targets"])
    a = torch.tensor([0.25, 0.75])
    assert abs(C.focal_loss(lg, tg, 2.0, a).item() - float(g["focal__loss_alpha"])) < 1e-6
    assert abs(C.focal_loss(lg, tg, 2.0, None).item() - float(g["focal__loss_noalpha"])) < 1e-6
    single = [C.focal_loss(lg[i], tg[i], 2.0, a).item() for i in range(6)]
    assert np.allclose(single, g["focal__loss_single"], atol=1e-6)


def test_bimodal_oracle_against_golden(golden_dir):
    """The bimodal restatement reproduces the reference's frozen outputs AND gradients (tests/golden/bimodal_small.npz)."""
    import torch
    from oracle import classifier_fp32 as C
    g = np.load(os.path.join(golden_dir, "bimodal_small.npz"))
    d, _, _, h_ct, h_pet, l_ct, l_pet, _ = (int(v) for v in g["cfg"])
    x_ct, x_pet, y = torch.from_numpy(g["x_ct"]), torch.from_numpy(g["x_pet"]), torch.from_numpy(g["y"])
    alpha = torch.tensor([0.25, 0.75])
    for mode, (a, b) in {"both": (x_ct, x_pet), "ct": (x_ct, None), "pet": (None, x_pet)}.items():
        sd = {k[7:]: torch.from_numpy(g[k]).clone().requires_grad_(True) for k in g.files if k.startswith("param__")}
        lg, z, lg_ct, lg_pet = C.bimodal_forward(sd, a, b, h_ct, h_pet, l_ct, l_pet)
        loss = C.focal_loss(lg[0], y, 2.0, alpha) + C.focal_loss(lg_ct[0], y, 2.0, alpha) + C.focal_loss(lg_pet[0], y, 2.0, alpha) + 0.1 * z.sum()
        loss.backward()
        assert np.allclose(lg.detach().numpy(), g[f"{mode}__logits"], atol=1e-5)
        assert np.allclose(z.detach().numpy().reshape(-1), g[f"{mode}__z"].reshape(-1), atol=1e-5)
        assert abs(float(loss) - float(g[f"{mode}__loss"])) < 1e-5
        for k in g.files:
            if k.startswith(f"{mode}__grad__"):
                assert np.allclose(sd[k[len(mode) + 8:]].grad.numpy(), g[k], atol=2e-5), k
