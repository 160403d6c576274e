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


_dataset_table = ref_shim.make_dataset_table


def test_prepare_df_and_label_encoder_match_reference():
    """train_models.py:416-448 run unmodified (under the pandas < 2 indexing shim it was written for) against the port."""
    import warnings

    import pandas as pd
    from vit_deep_radiomics_b200 import train_models as ours
    tm = ref_shim.load_reference("train_models")
    df = _dataset_table()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with ref_shim.legacy_series_getitem():
            want = tm.prepare_df(df.copy())
    got = ours.prepare_df(df.copy())
    pd.testing.assert_frame_equal(want, got, check_dtype=False)
    assert "divisor" not in df.columns                                  # the port does not modify its argument
    assert got[got.modality == "ct"].groupby("patient_id_new")["slice"].nunique().max() == 14      # 13 + 1: inclusive bounds
    assert got[(got.patient_id == "P3") & (got.modality == "ct")]["patient_id_new"].unique().tolist() == ["P3:0"]   # 9 slices, window 8
    a, b = tm.get_label_encoder(got), ours.get_label_encoder(got)
    x = np.array([[0], [1], [1]])
    assert np.array_equal(a.transform(x).toarray(), b.transform(x).toarray())
    assert [ours.find_divisor(n, m) for n, m in ((40, "ct"), (5, "chest"), (9, "pet"), (1, "pet"))] == \
           [int(tm.find_divisor(n, m)) for n, m in ((40, "ct"), (5, "chest"), (9, "pet"), (1, "pet"))]


@pytest.mark.parametrize("augment", [False, True])
def test_dataset_sampling_matches_reference(augment):
    """PETCTDataset3D (train_models.py:47-141): the same seeded draws from NumPy's global generator select the same window,
    flip / angle, slice crop, noise and scale, so every item carries bit-identical token sequences (gather = the oracle here;
    the device gather is checked against the same oracle in tests/test_gpu_gather.py), labels and patient ids."""
    import warnings

    from vit_deep_radiomics_b200 import train_models as ours
    tm = ref_shim.load_reference("train_models")
    D = 12
    df = ours.prepare_df(_dataset_table(seed=3, D=D))
    enc = ours.get_label_encoder(df)

    def oracle_gather(feats, masks, res, noise, d):
        return G.token_gather(feats, masks, res, noise, d)["tokens"]

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = tm.PETCTDataset3D(df.copy(), enc, "ct.h5", "pet.h5", use_augmentation=augment, feature_dim=D, arch="transformer")
        mine = ours.PETCTDataset3D(df.copy(), enc, "ct.h5", "pet.h5", use_augmentation=augment, feature_dim=D, arch="transformer",
                                   store=ref_shim.H5_FILES, gather=oracle_gather)
        assert len(ref) == len(mine) > 0
        for epoch in range(2):
            for i in range(len(ref)):
                np.random.seed(100 * epoch + i)
                a, next_a = ref[i], np.random.random()
                np.random.seed(100 * epoch + i)
                b, next_b = mine[i], np.random.random()
                assert next_a == next_b                               # the same number of draws from the global generator
                assert a[3] == b[3] and torch.equal(a[2], b[2])
                assert a[0].shape == b[0].shape and a[1].shape == b[1].shape
                assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
