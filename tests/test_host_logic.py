"""Host-side mirror of the reference interface: geometry, config, synthetic data, PE, loss (CPU)."""
import os

import numpy as np
import pytest
import torch
import yaml

from oracle import gather_np as G
from vit_deep_radiomics_b200 import config_manager, ops, synth
from vit_deep_radiomics_b200 import train_models as tm
from vit_deep_radiomics_b200 import visualization_utils as vu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_geometry_matches_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "geometry.npz"))
    for i, m in enumerate(g["geo__masks"]):
        assert list(vu.extract_coords(m, 1)) == g["geo__coords_m1"][i].tolist()
        assert list(vu.extract_coords(m, 2)) == g["geo__coords_m2"][i].tolist()
        img = np.zeros((m.shape[0] // 4, m.shape[1] // 4, 3))
        got = list(vu.extract_roi(img, m).shape[:2]) + list(vu.extract_roi(m, m).shape[:2])
        assert got == g["geo__roi_shapes"][i].tolist()
    assert vu.crop_window(g["gen__mask"]) == G.crop_window(g["gen__mask"])
    with pytest.raises(ValueError):
        vu.extract_coords(np.zeros((4, 4), bool), 1)


def test_survey_probe_values():
    # rows 5..8, cols 10..15, margin 2 -> (12, 3, 17, 6)
    m = np.zeros((20, 20), bool)
    m[5:9, 10:16] = True
    assert vu.extract_coords(m, 2) == (12, 3, 17, 6)
    assert vu.crop_image(np.zeros((10, 12)), -3, 2, 40, 7).shape == (5, 12)


def test_config_manager_reads_reference_schema():
    cfg = config_manager.load_conf(project_dir=ROOT)
    t = cfg["models"]["transformer"]
    assert (t["learning_rate"], t["feature_dim"], t["batch_size"], t["virtual_batch_size"], t["num_epochs"], t["patience"]) == \
        (0.0005, 256, 1, 32, 50, 15)
    for mod in ("ct", "pet", "chest"):
        assert t[mod] == {"num_layers": 2, "num_heads": 4, "mlp_ratio": 4}
    assert config_manager.get_project_dir(os.path.join(ROOT, "tests")) == ROOT
    ref = "/root/reference/conf/parameters_models.yaml"
    if os.path.isfile(ref):
        assert cfg["models"] == yaml.safe_load(open(ref))["models"]


def test_kfold_schema_and_split(tmp_path):
    ids, labels, sizes, cloud = synth.point_cloud_patients(50, d=64, n_range=(8, 16))
    kf = synth.kfold_yaml_dict(ids, labels)
    folds = kf["kfold_patients"]["ct"]["stanford"]
    assert sorted(folds) == [0, 1, 2, 3, 4]                       # int fold keys like the reference file
    alltest = sorted(sum((folds[k]["test"] for k in folds), []))
    assert alltest == sorted(ids)
    for k in folds:
        assert not set(folds[k]["train"]) & set(folds[k]["test"])
    (tmp_path / "conf").mkdir()
    (tmp_path / "conf" / "parameters_kfold.yaml").write_text(yaml.safe_dump(kf))
    (tmp_path / "conf" / "parameters_models.yaml").write_text(open(os.path.join(ROOT, "conf", "parameters_models.yaml")).read())
    cfg = config_manager.load_conf(project_dir=tmp_path)
    assert cfg["kfold_patients"]["ct"]["stanford"][0]["test"] == folds[0]["test"] and "models" in cfg
    assert cloud(3).shape == (sizes[3], 64) and np.array_equal(cloud(3), cloud(3))


def test_positional_encoding_and_maps(golden_dir):
    g = np.load(os.path.join(golden_dir, "geometry.npz"))
    x, y, z = g["pe__xyz"]
    for D in (12, 256, 384):
        assert np.array_equal(tm.positional_encoding_3d(x, y, z, D), g[f"pe__{D}"])
    for n_in in range(1, 30):
        for n_out in range(1, 30):
            assert np.array_equal(ops.nearest_index_map(n_out, n_in), G.nearest_index_map(n_out, n_in))
    # grid means reproduce the reference's x.mean() exactly (same array, same numpy reduction)
    h, w, S, hm, wm, res = 7, 5, 3, 30, 22, (0.8, 0.7, 1.25)
    n = np.arange(h * w * S)
    xr = ((n // S) % h / w) * wm * res[0]
    assert ops.grid_means(h, w, S, hm, wm, res)[0] == xr.mean()


def test_focal_loss_matches_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "classifier_small.npz"))
    lg, tg = torch.tensor(g["focal__logits"]), torch.tensor(g["focal__targets"])
    crit = tm.FocalLoss(alpha=torch.tensor([0.25, 0.75]), gamma=2)
    assert abs(crit(lg, tg).item() - float(g["focal__loss_alpha"])) < 1e-6
    assert abs(tm.FocalLoss(gamma=2)(lg, tg).item() - float(g["focal__loss_noalpha"])) < 1e-6
    assert np.allclose([crit(lg[i], tg[i]).item() for i in range(6)], g["focal__loss_single"], atol=1e-6)


def test_crossmodal_focal_loss_matches_golden(golden_dir):
    """CrossModalFocalLoss (reference train_models.py:332-378): value, per-sample values and logit gradients."""
    g = np.load(os.path.join(golden_dir, "crossmodal_loss.npz"))
    tg = torch.tensor(g["targets"])
    for tag, kw in (("default", {}), ("train", dict(alpha=torch.tensor([0.25, 0.75]), gamma_unimodal=2.0, gamma_bimodal=1.0, beta=0.6))):
        crit = tm.CrossModalFocalLoss(**kw)
        a, b, c = (torch.tensor(g[k]).requires_grad_(True) for k in ("lx", "lc", "lp"))
        loss = crit(a, b, c, tg)
        loss.backward()
        assert abs(loss.item() - float(g[f"{tag}__loss"])) < 1e-6
        for n_, t in (("gx", a), ("gc", b), ("gp", c)):
            assert np.allclose(t.grad.numpy(), g[f"{tag}__{n_}"], atol=1e-6)
        single = [crit(a[i], b[i], c[i], tg[i]).item() for i in range(6)]
        assert np.allclose(single, g[f"{tag}__single"], atol=1e-6)
    crit = tm.make_criterion("crossmodal", "cpu")
    assert (crit.gamma_bimodal, crit.gamma_unimodal, crit.beta) == (1.0, 2.0, 0.6) and crit.alpha.tolist() == [0.25, 0.75]
    assert isinstance(tm.make_criterion("focal", "cpu"), tm.FocalLoss)


def test_build_model_bimodal_takes_ct_from_modality_b():
    """reference train_models.py:455-472: CT encoder <- cfg[modality_b], PET encoder <- cfg[modality_a]."""
    cfg = {"models": {"transformer": {"feature_dim": 128,
                                      "pet": {"mlp_ratio": 2, "num_heads": 2, "num_layers": 1},
                                      "ct": {"mlp_ratio": 4, "num_heads": 2, "num_layers": 3}}}}
    m = tm.build_model(cfg, "transformer", "petct", "pet", "ct")
    assert len(m.transformer_encoder_ct.layers) == 3 and len(m.transformer_encoder_pet.layers) == 1
    assert m.transformer_encoder_ct.layers[0].linear1.out_features == 512
    assert m.transformer_encoder_pet.layers[0].linear1.out_features == 256
    if os.path.isdir("/root/reference/src"):
        from oracle import ref_shim
        ref = ref_shim.load_reference("train_models").build_model(cfg, "transformer", "petct", "pet", "ct")
        assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == {k: tuple(v.shape) for k, v in m.state_dict().items()}
        m.load_state_dict(ref.state_dict())


def test_synth_cases_are_seeded():
    a = synth.make_case("T0")
    b = synth.make_case("T0")
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[0].dtype == np.float32
    assert a[0].min() >= 0 and a[0].max() <= 1 and a[1].any()
    img, mask, res, name = synth.make_case("C1")
    assert img.shape == (224, 224, 8) and name == "vit_s16"


def test_cli_flags_match_reference():
    from vit_deep_radiomics_b200 import tfds_dense_descriptor as tdd
    a = tdd.build_arg_parser().parse_args(["-mn", "vit_b16", "-mp", "x.pth", "-d", "d", "-f", "f", "-h5", "h", "-df", "c", "-mod", "chest"])
    assert (a.model_name, a.model_path, a.dataset_path, a.feature_folder, a.hdf5_path, a.df_path, a.modality) == \
        ("vit_b16", "x.pth", "d", "f", "h", "c", "chest")
    b = tm.build_arg_parser().parse_args(["-a", "transformer", "-d", "stanford", "-b", "vit", "-m", "ct", "-gpu", "1", "-l", "focal", "-e", "e"])
    assert (b.arch, b.dataset, b.backbone, b.modality, b.gpu, b.loss, b.experiment) == ("transformer", "stanford", "vit", "ct", 1, "focal", "e")
    assert tdd.build_arg_parser().parse_args(["-mn", "medsam"]).model_name == "medsam"     # the reference's default backbone


def test_sam_encoder_host_side():
    """Key names / geometry of the MedSAM encoder drop-in (no device work: the constructor itself needs CUDA)."""
    from vit_deep_radiomics_b200 import sam_encoder
    from oracle import sam_fp32
    cfg = sam_encoder.SAM_CONFIGS["medsam"]
    assert cfg == sam_fp32.SAM_CONFIGS["medsam"] and (cfg["dim"], cfg["depth"], cfg["heads"], cfg["out_chans"]) == (768, 12, 12, 256)
    tiny = sam_encoder.SAM_CONFIGS["sam_tiny"]
    a, b = sam_encoder.init_sam_state_dict(tiny, (256, 256), seed=3), sam_fp32.init_sam_state_dict(tiny, (256, 256), seed=3)
    assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)
    assert a["blocks.0.attn.rel_pos_h"].shape == (27, 64) and a["blocks.1.attn.rel_pos_h"].shape == (31, 64)
    assert sam_encoder.SAM_CONFIGS == sam_fp32.SAM_CONFIGS
    assert abs(sam_encoder.sam_flops_per_slice(cfg, (64, 64)) / 1e9 - 941.727) < 0.01
    full = {"image_encoder." + k: v for k, v in a.items()} | {"mask_decoder.x": torch.zeros(1)}
    assert sam_encoder._strip_prefix(full, "image_encoder.").keys() == a.keys()
    r = sam_encoder._fit_rel_pos(a["blocks.0.attn.rel_pos_h"], 16)                      # 27 -> 31 rows, linear
    assert r.shape == (31, 64) and torch.allclose(r, sam_fp32.rel_pos_rows(16, a["blocks.0.attn.rel_pos_h"])[(torch.arange(31) - 15).clamp(min=0), (15 - torch.arange(31)).clamp(min=0)])


def test_build_model_and_state_dict_keys():
    cfg = config_manager.load_conf(project_dir=ROOT)
    model = tm.build_model(cfg, "transformer", "ct")
    keys = set(model.state_dict().keys())
    want = {"cls_token", "norm.weight", "norm.bias", "classifier.dense1.weight", "classifier.dense1.bias",
            "classifier.dense2.weight", "classifier.dense2.bias"}
    for i in range(2):
        p = f"transformer_encoder.layers.{i}."
        want |= {p + s for s in ("self_attn.in_proj_weight", "self_attn.in_proj_bias", "self_attn.out_proj.weight",
                                 "self_attn.out_proj.bias", "linear1.weight", "linear1.bias", "linear2.weight",
                                 "linear2.bias", "norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias")}
    assert keys == want
    assert sum(p.numel() for p in model.parameters()) == 1_712_898
    if os.path.isdir("/root/reference/src"):
        from oracle import ref_shim
        ma = ref_shim.load_reference("models_archs")
        ref = ma.TransformerNoduleClassifier(256, 1024, 4, 2, 2)
        assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == {k: tuple(v.shape) for k, v in model.state_dict().items()}
        model.load_state_dict(ref.state_dict())                          # .pth interchange


def test_epoch_policy_and_split_report():
    """Epoch-end bookkeeping of the reference loop (train_models.py:727-810): patient-weighted report, target metric
    test_auc^2 * sqrt(test_f1), checkpoint when >= the fold mean, stop when the first best epoch is `patience` epochs back."""
    hist = [dict(epoch=0, test_auc=0.60, test_f1=0.50), dict(epoch=1, test_auc=0.70, test_f1=0.64),
            dict(epoch=2, test_auc=0.62, test_f1=0.50), dict(epoch=3, test_auc=0.70, test_f1=0.64)]
    save, stop, target = tm.epoch_policy(hist[:2], patience=2)
    assert save and not stop and abs(target - 0.49 * 0.8) < 1e-12
    save, stop, _ = tm.epoch_policy(hist[:3], patience=2)
    assert not save and not stop                      # below the mean; best epoch 1 is one epoch back
    save, stop, _ = tm.epoch_policy(hist, patience=2)
    assert save and stop                              # ties count as improvement, but the FIRST best epoch (1) is 2 back
    assert tm.get_sampler_weights(np.array(["a", "b", "a", "c", "a"])) == [1 / 3, 1, 1 / 3, 1, 1 / 3]
    if os.path.isdir("/root/reference/src"):
        from oracle import ref_shim
        ref = ref_shim.load_reference("train_models")
        assert ref.get_sampler_weights(np.array(["a", "b", "a", "c", "a"])) == tm.get_sampler_weights(np.array(["a", "b", "a", "c", "a"]))
    y_true = [np.array([0]), np.array([1]), np.array([1]), np.array([0]), np.array([1])]
    y_score = [np.array([[0.8, 0.2]]), np.array([[0.3, 0.7]]), np.array([[0.6, 0.4]]), np.array([[0.4, 0.6]]), np.array([[0.1, 0.9]])]
    pids = [np.array(["p1"]), np.array(["p2"]), np.array(["p2"]), np.array(["p3"]), np.array(["p4"])]
    rep = tm.split_report(y_true, y_score, pids, loss=0.25, kfold=1, epoch=4, split="test")
    assert rep["split"] == "test" and rep["kfold"] == 1 and rep["epoch"] == 4 and rep["loss"] == 0.25
    # weights 1, .5, .5, 1, 1: positives 2 (weighted), of which 1.5 predicted positive; negatives 2, of which 1 predicted negative
    assert abs(rep["1"]["recall"] - 0.75) < 1e-12 and abs(rep["0"]["recall"] - 0.5) < 1e-12 and abs(rep["accuracy"] - 0.625) < 1e-12
    assert 0.0 <= rep["ROC AUC"] <= 1.0


def test_extract_patient_features_augmentation_table():
    """The per-patient extraction loop (tfds_dense_descriptor.py:452-488) with a stub backbone: 3 flips x 4 angles, the
    metadata table the trainer reads, including the reference's all-True `augmentation` column (:486)."""
    from vit_deep_radiomics_b200 import tfds_dense_descriptor as tdd
    rng = np.random.default_rng(2)
    S = 3
    img = rng.random((16, 16, S)).astype(np.float32)
    mask = np.zeros((16, 16, S), bool)
    mask[5:9, 6:11] = True
    seen = []

    def stub(model, img_3d, mask_3d, tqdm_text, display):
        seen.append((img_3d.copy(), mask_3d.copy(), tqdm_text))
        return ([np.full((2, 2, 4), float(len(seen)), np.float32)] * img_3d.shape[2], [np.ones((2, 2), bool)] * img_3d.shape[2])

    res = np.array([0.8, 0.8, 0.8])
    df, feats, masks = tdd.extract_patient_features(None, img, mask, "p01", 1, "stanford_dataset", "ct", res, generate=stub)
    assert len(seen) == 12 and len(feats) == len(masks) == 12 * S == len(df)
    assert list(df.columns) == ["feature_id", "slice", "angle", "flip", "patient_id", "label", "dataset", "modality", "augmentation", "spatial_res"]
    assert df["feature_id"].tolist() == list(range(12 * S)) and df["slice"].tolist() == list(range(S)) * 12
    assert df["angle"].tolist() == [a for _ in range(3) for a in (0, 45, 90, 135) for _ in range(S)]
    flips = df["flip"].tolist()            # pandas stores the None of the first four volumes as a missing value
    assert all(f is None or f != f for f in flips[:4 * S]) and flips[4 * S:] == [f for f in ("horizontal", "vertical") for _ in range(4 * S)]
    assert df["augmentation"].all() and (df["dataset"] == "stanford").all() and (df["patient_id"] == "p01").all()
    assert all(np.array_equal(r, res) for r in df["spatial_res"]) and seen[0][2] == "ct p01"
    assert np.array_equal(seen[0][0], img) and np.array_equal(seen[4][0], img[:, ::-1]) and np.array_equal(seen[8][0], img[::-1])
    assert feats[0][0, 0, 0] == 1.0 and feats[-1][0, 0, 0] == 12.0        # concatenated in (flip, angle, slice) order
    # the reference's expression for the augmentation column, evaluated literally
    import pandas as pd
    lit = np.logical_not(np.logical_and(df["flip"] is None, df["angle"] == 0))
    assert np.array_equal(np.asarray(lit), df["augmentation"].values)


def test_run_fold_epoch_loop_files_and_policy(tmp_path, monkeypatch):
    """The fold loop of the training script (train_models.py:562-810) with a stub classifier on the CPU (the real one needs
    CUDA; tests/test_gpu_pipeline.py runs it): accumulation / evaluation passes, the per-epoch metric files, checkpoints by the
    target-metric rule, early stopping by patience."""
    import json
    from oracle import gather_np as G
    from oracle import ref_shim
    D = 12
    df = tm.prepare_df(ref_shim.make_dataset_table(seed=9, D=D))
    enc = tm.get_label_encoder(df)
    cfg = config_manager.load_conf(project_dir=ROOT)
    cfg["models"]["transformer"].update(feature_dim=D, patience=1, virtual_batch_size=3)

    class Stub(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(D, 2)

        def forward(self, x):
            cls = x.mean(1)
            return self.lin(cls), cls

    steps = []
    monkeypatch.setattr(tm, "build_model", lambda *a, **k: Stub())
    monkeypatch.setattr(torch.optim.AdamW, "step", lambda self, *a, **k: steps.append(1))
    torch.manual_seed(0)
    np.random.seed(0)
    hist = tm.run_fold(cfg, "transformer", "ct", df[df.patient_id.isin(["P1", "P2"])].reset_index(drop=True),
                       df[df.patient_id.isin(["P3", "P4"])].reset_index(drop=True), enc, "ct.h5", "pet.h5", str(tmp_path), kfold=0,
                       device="cpu", store=ref_shim.H5_FILES, num_epochs=4,
                       gather=lambda f, m, r, n, d: G.token_gather(f, m, r, n, d)["tokens"])
    # the optimizer never moves (step patched out): every epoch has the same test metrics -> the first epoch stays the first
    # maximum, so patience 1 stops the fold after its second epoch
    assert [h["epoch"] for h in hist] == [0, 1]
    for e in (0, 1):
        for split in ("train", "test"):
            rep = json.load(open(tmp_path / f"{split}_metrics_{e}.json"))
            assert rep["split"] == split and rep["epoch"] == e and rep["kfold"] == 0 and 0.0 <= rep["ROC AUC"] <= 1.0
            assert np.isfinite(rep["loss"]) and "macro avg" in rep
    assert (tmp_path / "model_epoch_0000.pth").exists() and (tmp_path / "model_epoch_0001.pth").exists()   # target == the fold's mean
    assert hist[0]["test_loss"] == hist[1]["test_loss"] and np.isfinite(hist[0]["train_loss"])
    # optimizer steps: every min(virtual_batch, items) items and at the last item of each training pass
    n_items = len(tm.PETCTDataset3D(df[df.patient_id.isin(["P1", "P2"])].reset_index(drop=True), enc, "ct.h5", "pet.h5",
                                    use_augmentation=True, feature_dim=D, arch="transformer", store=ref_shim.H5_FILES, gather=lambda *a: None))
    assert len(steps) == 2 * -(-n_items // 3)


def test_extraction_main_walks_the_metadata_table(tmp_path, monkeypatch):
    """The extraction script's HDF5 branch (tfds_dense_descriptor.py:364-490) with the backbone, the HDF5 reader and writer
    stubbed: which (patient, modality) pairs are processed, labels from the `egfr` column, output names, resume by parquet."""
    import pandas as pd
    from vit_deep_radiomics_b200 import tfds_dense_descriptor as tdd
    meta = pd.DataFrame({"patient_id": ["a1", "a2", "b1", "b2"], "egfr": ["Mutant", "Wildtype", "Mutant", "Wildtype"],
                         "dataset": ["stanford", "stanford", "santa_maria", "santa_maria"], "has_petct": [True, False, True, True]})
    csv = tmp_path / "meta.csv"
    meta.to_csv(csv, index=False)
    calls, saved = [], []
    rng = np.random.default_rng(0)

    def voxels(ds_path, patient_id, modality):
        m = np.zeros((12, 12, 2), bool)
        m[4:8, 3:9] = True
        return rng.random((12, 12, 2)).astype(np.float32), m, np.array([0.8, 0.8, 0.8])

    def gen(model, img_3d, mask_3d, tqdm_text, display=False):
        calls.append(tqdm_text)
        return [np.zeros((2, 2, 4), np.float32)] * 2, [np.ones((2, 2), bool)] * 2

    monkeypatch.setattr(tdd, "load_model", lambda name, path=None: ("model", name, path))
    monkeypatch.setattr(tdd, "get_voxels", voxels)
    monkeypatch.setattr(tdd, "generate_features", gen)
    monkeypatch.setattr(tdd, "save_features", lambda f, feats, masks, pid: saved.append((os.path.basename(f), pid, len(feats))))
    out = tmp_path / "features"
    argv = ["-mn", "medsam", "-mp", "w.pth", "-f", str(out), "-h5", "vol.h5", "-df", str(csv), "-mod", "ct"]
    tdd.main(argv)
    files = sorted(str(p.relative_to(out)) for p in out.rglob("*.parquet"))
    assert files == ["santa_maria_dataset/b1_ct.parquet", "santa_maria_dataset/b1_pet.parquet", "santa_maria_dataset/b2_ct.parquet",
                     "santa_maria_dataset/b2_pet.parquet", "stanford_dataset/a1_ct.parquet", "stanford_dataset/a1_pet.parquet"]
    assert len(calls) == 6 * 12 and len(saved) == 6 and all(n == 24 for _, _, n in saved)
    assert {f for f, _, _ in saved} == {"features_masks_ct.hdf5", "features_masks_pet.hdf5"}
    df = pd.read_parquet(out / "stanford_dataset" / "a1_ct.parquet")
    assert (df["label"] == 1).all() and (df["dataset"] == "stanford").all() and (df["modality"] == "ct").all() and len(df) == 24
    assert (pd.read_parquet(out / "santa_maria_dataset" / "b2_pet.parquet")["label"] == 0).all()
    n = len(calls)
    tdd.main(argv)                                   # every parquet exists: nothing is recomputed (:424)
    assert len(calls) == n
