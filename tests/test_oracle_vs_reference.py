"""Oracle restatements against the LIVE reference (imported unmodified through oracle/ref_shim.py).
Only runs where /root/reference exists (the build container); the golden tests cover the GPU box."""
import numpy as np
import pytest
import torch

from oracle import classifier_fp32 as C
from oracle import gather_np as G
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference sources not present")


def test_get_features_random_cases():
    tm = ref_shim.load_reference("train_models")

    class DS(tm.PETCTDataset3D):
        def __init__(self, D):
            self.feature_dim, self.arch = D, "transformer"

    rng = np.random.default_rng(5)
    for _ in range(12):
        S, h, w = rng.integers(1, 9), rng.integers(2, 20), rng.integers(2, 20)
        hm, wm = rng.integers(h, 5 * h), rng.integers(w, 5 * w)
        D = int(rng.choice([6, 12, 30, 256]))
        feats = [rng.standard_normal((h, w, D)).astype(np.float32) for _ in range(S)]
        masks = [rng.random((hm, wm)) < 0.3 for _ in range(S)]
        res, noise = rng.uniform(0.4, 2, 3), rng.uniform(-5, 5, 3)
        ref_shim.H5_FILES.pop("t.h5", None)
        ref_shim.put_feature_file("t.h5", "p", feats, masks)
        want = DS(D)._get_features("t.h5", "p", list(range(S)), 0, "None", noise, res)
        got = G.token_gather(feats, masks, res, noise, D)["tokens"]
        assert np.array_equal(want, got)


def test_pointcloud_and_geometry():
    pc = ref_shim.load_reference("create_pointcloud_dataframe")
    vu = ref_shim.load_reference("visualization_utils")
    rng = np.random.default_rng(6)
    for shape in [(8, 8, 3), (6, 10, 4), (11, 7, 5)]:
        img = rng.standard_normal(shape).astype(np.float32)
        mask = rng.random(shape) < 0.05
        res = rng.uniform(0.5, 2, 3)
        df = pc.to_pointcloud_df(img, mask, 1, res)
        o = G.voxel_pointcloud(img, mask, res)
        for col in ("x", "y", "z", "raw", "mask", "mask_box"):
            assert np.array_equal(df[col].values, o[col]), (shape, col)
    for _ in range(100):
        H, W = rng.integers(20, 80, 2)
        m = np.zeros((H, W), bool)
        r0, c0 = rng.integers(0, H - 3), rng.integers(0, W - 3)
        m[r0:rng.integers(r0 + 1, H), c0:rng.integers(c0 + 1, W)] = True
        for mg in (1, 2):
            assert tuple(int(v) for v in vu.extract_coords(m, mg)) == G.extract_coords(m, mg)


def test_classifier_against_reference_module():
    ma = ref_shim.load_reference("models_archs")
    torch.manual_seed(3)
    model = ma.TransformerNoduleClassifier(128, 256, 2, 2, 2).eval()
    x = torch.randn(1, 50, 128)
    with torch.no_grad():
        want_l, want_c = model(x)
        got_l, got_c = C.classifier_forward(dict(model.state_dict()), x, 2, 2)
    assert torch.allclose(want_l, got_l, atol=1e-5) and torch.allclose(want_c, got_c, atol=1e-5)


def test_bimodal_classifier_against_reference_module():
    """oracle/classifier_fp32.bimodal_forward vs the UNMODIFIED TransformerNoduleBimodalClassifier (eval mode), all three input modes."""
    ma = ref_shim.load_reference("models_archs")
    torch.manual_seed(4)
    model = ma.TransformerNoduleBimodalClassifier(128, 2, 3, 2, 2, 2, 1, 2).eval()
    x_ct, x_pet = torch.randn(1, 30, 128), torch.randn(1, 45, 128)
    sd = dict(model.state_dict())
    with torch.no_grad():
        for a, b in ((x_ct, x_pet), (x_ct, None), (None, x_pet)):
            want = model(a, b)
            got = C.bimodal_forward(sd, a, b, 2, 2, 2, 1)
            for w, g in zip(want, got):
                assert torch.allclose(w.reshape(g.shape), g, atol=1e-5)
