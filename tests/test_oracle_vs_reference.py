"""Oracle restatements against the LIVE reference (imported unmodified through oracle/ref_shim.py).
Only runs where /root/reference exists (the build container); the golden tests cover the GPU box."""
import os

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


def _feature_rows(rng, pid, dataset, label, n):
    return [dict(patient_id=pid, dataset=dataset, label=label, modality=m, slice=s_, feature_id=s_, flip=f, angle=a,
                 spatial_res=(1.0, 1.0, 2.0))
            for m in ("ct", "pet") for (f, a) in (("None", 0), ("H", 90), (None, 0)) for s_ in range(n)]


def test_split_patients_matches_reference_script(tmp_path, monkeypatch):
    """src/split_patients.py executed unmodified (runpy) on a temporary project directory against the function port."""
    import runpy
    import sys
    import types

    import pandas as pd
    import yaml
    from vit_deep_radiomics_b200 import split_patients as sp
    rng = np.random.default_rng(0)
    rows = []
    for ds, npat in (("stanford", 23), ("santa_maria", 17)):
        for i in range(npat):
            rows += _feature_rows(rng, f"{ds[:2]}{i:03d}", ds, int(rng.random() < 0.4), 2)
    df = pd.DataFrame(rows)
    (tmp_path / "data" / "features").mkdir(parents=True)
    (tmp_path / "conf").mkdir()
    df.to_parquet(tmp_path / "data" / "features" / "petct.parquet")
    fake = types.ModuleType("config_manager")
    fake.get_project_dir = lambda *a, **k: str(tmp_path)
    monkeypatch.setitem(sys.modules, "config_manager", fake)
    runpy.run_path(os.path.join(ref_shim.REFERENCE_SRC, "split_patients.py"), run_name="split_patients_ref")
    want = yaml.safe_load((tmp_path / "conf" / "parameters_kfold.yaml").read_text())
    os.remove(tmp_path / "conf" / "parameters_kfold.yaml")
    assert sp.main(str(tmp_path)).endswith("parameters_kfold.yaml")
    got = yaml.safe_load((tmp_path / "conf" / "parameters_kfold.yaml").read_text())
    assert got == want and set(got["kfold_patients"]) == {"ct", "pet"}
    folds = got["kfold_patients"]["ct"]["stanford"]
    assert sorted(folds) == [0, 1, 2, 3, 4] and all(not set(f["train"]) & set(f["test"]) for f in folds.values())


def test_merge_dataframe_features_matches_reference_script(tmp_path, monkeypatch):
    """src/merge_dataframe_features.py executed unmodified (runpy, cwd = a temporary src/ directory) against the port."""
    import runpy

    import pandas as pd
    from vit_deep_radiomics_b200 import merge_dataframe_features as mf
    rng = np.random.default_rng(1)
    feat = tmp_path / "data" / "features"
    for ds, pids in (("stanford_dataset", ("a", "b", "c")), ("santa_maria_dataset", ("x", "y"))):
        (feat / ds).mkdir(parents=True)
        for pid in pids:
            pd.DataFrame(_feature_rows(rng, pid, ds, 1, 3)).to_parquet(feat / ds / f"{pid}.parquet")
    (tmp_path / "src").mkdir()
    monkeypatch.chdir(tmp_path / "src")
    runpy.run_path(os.path.join(ref_shim.REFERENCE_SRC, "merge_dataframe_features.py"), run_name="__main__")
    want = pd.read_parquet(feat / "petct.parquet")
    os.remove(feat / "petct.parquet")
    got = mf.merge_features(os.path.join("..", "data", "features"))
    pd.testing.assert_frame_equal(want, got)
    assert (got["augmentation"] == ~((got["flip"] == "None") & (got["angle"] == 0))).all() and (got["flip"] == "None").any()
    assert pd.read_parquet(mf.main(os.path.join("..", "data", "features"))).equals(want)


def test_hu_to_rgb_and_flip_rotate_match_reference():
    """visualization_utils.hu_to_rgb_vectorized (:128-186) bit for bit on every input dtype the HDF5 volumes may have (the
    ramp ratio follows the input's dtype), interval edges included; flip_image / rotate_image (:306-350) on a small volume."""
    from vit_deep_radiomics_b200 import tfds_dense_descriptor as tdd, visualization_utils as V
    vu = ref_shim.load_reference("visualization_utils")
    ref = ref_shim.load_reference("tfds_dense_descriptor")
    rng = np.random.default_rng(3)
    edges = np.array([-1500, -1000, -999.5, -600, -599, -400, -399, -100, -99, -60, -59, 40, 41, 80, 81, 399, 400, 401, 1500, 0])
    for dt in (np.float32, np.float64, np.int16, np.int32):
        a = np.concatenate([rng.uniform(-1200, 800, 4000), edges]).astype(dt).reshape(-1, 4)
        want, got = vu.hu_to_rgb_vectorized(a), V.hu_to_rgb_vectorized(a)
        assert got.dtype == want.dtype == np.uint8 and got.shape == a.shape + (3,) and np.array_equal(got, want), dt
    img = rng.random((12, 10, 3)).astype(np.float32)
    mask = rng.random((12, 10, 3)) < 0.2
    for flip in (None, "horizontal", "vertical"):
        a, b = ref.flip_image(img, mask, flip), tdd.flip_image(img, mask, flip)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        for angle in tdd.AUG_ANGLES:
            c, d = ref.rotate_image(a[0], a[1], angle), tdd.rotate_image(b[0], b[1], angle)
            assert np.array_equal(c[0], d[0]) and np.array_equal(c[1], d[1]), (flip, angle)
    hu = rng.uniform(-1100, 600, (6, 5, 2)).astype(np.float32)
    assert np.array_equal(tdd.normalize_volume(hu, "ct", "medsam"), ref.apply_window_ct(hu, width=800, level=40))
    assert np.array_equal(tdd.normalize_volume(hu, "ct", "dinov2"), vu.hu_to_rgb_vectorized(hu) / 255.0)
    assert np.array_equal(tdd.normalize_volume(np.abs(hu), "pet", "medsam"), np.abs(hu) / np.abs(hu).max())


def test_report_text_and_param_count_match_reference(capsys):
    """print_classification_report (:185-218) yields the same text for the same report dict; get_number_of_params (:450-453)."""
    from vit_deep_radiomics_b200 import train_models as tm
    ref = ref_shim.load_reference("train_models")
    rng = np.random.default_rng(1)
    y_true = [np.array([int(v)]) for v in rng.integers(0, 2, 40)]
    y_score = [np.array([[1 - p, p]]) for p in rng.random(40)]
    pids = [np.array([f"p{int(v)}"]) for v in rng.integers(0, 7, 40)]
    rep = tm.split_report(y_true, y_score, pids, 0.1234567, 3, 7, "test")
    want = ref.print_classification_report(dict(rep))
    got = tm.print_classification_report(dict(rep), echo=False)
    capsys.readouterr()
    assert got == want
    lin = torch.nn.Sequential(torch.nn.Linear(5, 3), torch.nn.Linear(3, 2))
    lin[0].bias.requires_grad_(False)
    assert tm.get_number_of_params(lin) == int(ref.get_number_of_params(lin)) == 15 + 6 + 2


def test_patient_pointcloud_matches_reference_loop_body(monkeypatch):
    """create_pointcloud_dataframe.py:67-82 (inline in the reference's __main__): restated here step by step on the unmodified
    to_pointcloud_df / apply_window_ct, against patient_pointcloud with the device box gather replaced by the oracle's."""
    from vit_deep_radiomics_b200 import create_pointcloud_dataframe as cp
    pc = ref_shim.load_reference("create_pointcloud_dataframe")
    ref = ref_shim.load_reference("tfds_dense_descriptor")
    rng = np.random.default_rng(12)

    def oracle_box(img, mask, spatial_res, device):
        o = G.voxel_pointcloud(np.asarray(img, np.float32), mask, spatial_res)
        keep = o["mask_box"]
        flat = np.flatnonzero(keep)
        return None, dict(flat=flat, raw=o["raw"][keep], mask=o["mask"][keep], x=o["x"][keep], y=o["y"][keep], z=o["z"][keep])

    monkeypatch.setattr(cp, "_gather_box", oracle_box)
    for modality in ("ct", "pet"):
        img = (rng.uniform(-1000, 600, (9, 7, 5)) if modality == "ct" else rng.uniform(0, 9, (9, 7, 5))).astype(np.float32)
        mask = np.zeros((9, 7, 5), bool)
        mask[2:6, 1:5, 1:4] = rng.random((4, 4, 3)) < 0.6
        mask[3, 2, 2] = True
        res = np.array([0.8, 0.8, 1.5])
        # the reference's loop body
        df = pc.to_pointcloud_df(img, mask, 1, res)
        df["modality"] = modality
        norm = ref.apply_window_ct(img, width=800, level=40) if modality == "ct" else img / img.max()
        df["norm"] = norm.flatten()
        df["dataset"] = "stanford_dataset".replace("_dataset", "")
        df["patient_id"] = "p7"
        df = df[df["mask_box"]].copy()
        df["label"] = 1
        df.reset_index(drop=True, inplace=True)
        df[["x", "y", "z"]] = df[["x", "y", "z"]] - df[["x", "y", "z"]].mean(axis=0)
        got = cp.patient_pointcloud(img, mask, 1, res, modality, "stanford_dataset", "p7", device="cpu")
        assert list(got.columns) == list(df.columns) and len(got) == len(df) > 0
        for col in df.columns:
            assert np.array_equal(got[col].values, df[col].values), (modality, col)


def test_hdf5_handoff_executed_against_the_h5py_stand_in(capsys):
    """V5 (tfds_dense_descriptor.py:142-165, :353-362): h5py is not in this image, so the HDF5 hand-off is EXECUTED against
    ref_shim's dict-backed ``h5py.File`` stand-in -- our save_features / get_voxels and the unmodified reference's write and read
    the same stand-in; the stores must match key for key, dtype for dtype, byte for byte, and the trainer's reader
    (PETCTDataset3D._get_features through the same File API, store=None) must return the reference's tokens from the file
    save_features wrote."""
    ref = ref_shim.load_reference("tfds_dense_descriptor")          # installs the h5py stand-in into sys.modules
    tm_ref = ref_shim.load_reference("train_models")
    from vit_deep_radiomics_b200 import tfds_dense_descriptor as tdd
    from vit_deep_radiomics_b200 import train_models as tm
    rng = np.random.default_rng(17)
    D = 12
    feats = [rng.standard_normal((5, 4, D)).astype(np.float32) for _ in range(7)]
    masks = [rng.random((17, 13)) < 0.4 for _ in range(7)]
    for path in ("ours.h5", "ref.h5"):
        ref_shim.H5_FILES.pop(path, None)
    # a stale group of the same patient must be replaced, another patient kept (:153-155)
    tdd.save_features("ours.h5", feats[:2], masks[:2], "P9")
    ref.save_features("ref.h5", feats[:2], masks[:2], "P9")
    tdd.save_features("ours.h5", feats[:3], masks[:3], "P1")
    ref.save_features("ref.h5", feats[:3], masks[:3], "P1")
    tdd.save_features("ours.h5", feats, masks, "P1")
    ref.save_features("ref.h5", feats, masks, "P1")
    assert capsys.readouterr().out.count("already exists") == 2      # both implementations announce the overwrite
    ours, want = ref_shim.H5_FILES["ours.h5"], ref_shim.H5_FILES["ref.h5"]
    assert sorted(ours) == sorted(want) and len(ours) == 2 * 7 + 2 * 2
    for k in want:
        assert ours[k].dtype == want[k].dtype and ours[k].shape == want[k].shape and ours[k].tobytes() == want[k].tobytes(), k
    assert ours["P1/features/6"].dtype == np.float32 and ours["P1/masks/6"].dtype == bool
    # the trainer reads that file back through the File API (no `store` shortcut): reference tokens, bit for bit

    class RefDS(tm_ref.PETCTDataset3D):
        def __init__(self):
            self.feature_dim, self.arch = D, "transformer"

    ds = tm.PETCTDataset3D.__new__(tm.PETCTDataset3D)
    ds.store, ds.feature_dim, ds.arch, ds.device = None, D, "transformer", "cpu"
    ds.gather = lambda f, m, r, n, d: G.token_gather(f, m, r, n, d)["tokens"]
    res, noise = np.array([0.8, 0.8, 1.5]), np.array([0.5, -1.0, 2.0])
    ids = [1, 2, 3, 5]
    want_tok = RefDS()._get_features("ref.h5", "P1", ids, 0, "None", noise, res)
    got_tok = ds._get_features("ours.h5", "P1", ids, 0, "None", noise, res)
    assert np.array_equal(got_tok.numpy(), want_tok.astype(np.float32))
    # input side: get_voxels on <pid>_<mod>/img_exam/<k>, mask_exam/<k> (slices named by integers, stacked in numeric order)
    store = ref_shim.H5_FILES.setdefault("in.h5", {})
    store.clear()
    for k in (10, 2, 0, 1, 11):
        store[f"A7_ct/img_exam/{k}"] = rng.random((6, 5)).astype(np.float32)
        store[f"A7_ct/mask_exam/{k}"] = rng.random((6, 5)) < 0.3
    img, mask, sr = tdd.get_voxels("in.h5", "A7", "ct")
    img_r, mask_r, sr_r = ref.get_voxels("in.h5", "A7", "ct")
    assert img.shape == (6, 5, 5) and np.array_equal(sr, sr_r)
    assert np.array_equal(img, img_r) and np.array_equal(mask, mask_r) and img.dtype == img_r.dtype and mask.dtype == mask_r.dtype
    for j, k in enumerate((0, 1, 2, 10, 11)):                        # numeric slice order, not the key (string) order
        assert np.array_equal(img[:, :, j], store[f"A7_ct/img_exam/{k}"]) and np.array_equal(mask[:, :, j], store[f"A7_ct/mask_exam/{k}"])
