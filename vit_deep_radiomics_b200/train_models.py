"""Gather + classifier training -- drop-in for the hot-path parts of the reference's src/train_models.py.

Kept entry points: ``positional_encoding_3d`` (:30-44), ``PETCTDataset3D._get_features`` semantics
(:143-182, here ``get_features`` on device), ``FocalLoss`` (:381-405), ``build_model`` (:455-486),
``get_y_true_and_pred`` (:283-311), the gradient-accumulation train step (:652-688) and the CLI flags
(:500-515); the callers on the data side of the path: ``prepare_df`` sliding windows (:416-448),
``get_label_encoder`` (:492-499) and ``PETCTDataset3D`` with its augmentation sampling (:48-141).
The epoch-end bookkeeping is kept as functions (``split_report``, ``epoch_policy``: report JSON, checkpoint and
early-stopping rule of :727-810); plotting is outside the path.
"""
from __future__ import annotations

import argparse
import os

import numpy as np
import pandas as pd
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .models_archs import TransformerNoduleBimodalClassifier, TransformerNoduleClassifier


def positional_encoding_3d(x, y, z, D, scale=10000):
    """reference: train_models.py:30-44 (host, float64).  The device gather computes the same encoding
    in its epilogue (ops.mask_gather(pe=...)); this is the API-compatible host version."""
    x, y, z = np.asarray(x, np.float64), np.asarray(y, np.float64), np.asarray(z, np.float64)
    enc = np.zeros((x.shape[0], D))
    i = np.arange(D // 6)
    div = np.array([scale ** (6 * k / D) for k in range(D // 6)])
    for base, v in ((0, x), (D // 3, y), (2 * D // 3, z)):
        arg = v[:, None] / div[None, :]
        enc[:, 2 * i + base] = np.sin(arg)
        enc[:, 2 * i + 1 + base] = np.cos(arg)
    return enc


def get_features(features, masks, spatial_res, noise=(0.0, 0.0, 0.0), feature_dim=None, arch="transformer",
                 device="cuda:0", add_pe=True):
    """Device version of ``PETCTDataset3D._get_features`` (train_models.py:143-182) for data already in
    memory: ``features`` = S arrays (h, w, D) (or one (S,h,w,D) tensor), ``masks`` = S pixel masks (hm, wm).
    Returns the (n_sel, D) float32 token sequence (features[mask] + PE/4) on the device."""
    if arch != "transformer":
        raise NotImplementedError("arch='conv' (Conv3d classifier) is outside the hot path")
    f = features if isinstance(features, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(np.stack(features, 0)))
    m = masks if isinstance(masks, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(np.stack(masks, 0)).astype(np.uint8))
    f, m = f.to(device), m.to(device)
    if f.dtype not in (torch.float32, torch.bfloat16):
        f = f.float()
    pe = dict(res=spatial_res, noise=noise, scale=0.25) if add_pe else None
    tokens, src, count = ops.mask_gather(f.contiguous(), m.contiguous(), pe=pe)
    n = int(count.item())
    return tokens[:n]


def find_divisor(slice_count, modality):
    """reference: train_models.py:408-413 -- slices per sample window: 13 for CT-like series, 2 for PET, at most the series."""
    return int(np.clip(13 if modality in ("ct", "chest") else 2, 1, slice_count))


def prepare_df(df, modality_a="pet", modality_b="ct"):
    """reference: train_models.py:416-448.  Adds ``divisor`` (window size of the row's patient / modality) and
    ``patient_id_new``; every CT-like series is expanded into overlapping windows -- window ``i`` holds the rows with
    ``i <= slice <= i + divisor`` (divisor + 1 slice positions: the reference's bounds are inclusive) for
    ``i in range(n_unique_slices - divisor)`` and is named ``<patient>:<i>``; series with no complete window disappear.
    PET rows keep ``<patient>:<ceil(slice / divisor)>``.  CT windows first (patients in order of appearance), then PET.
    (The reference body indexes a row Series by position, which pandas >= 2 rejects; same table, label-based here.)"""
    df = df.copy()
    last_slice = df.groupby(["patient_id", "modality"])["slice"].max()
    divisor = {key: find_divisor(n, key[1]) for key, n in last_slice.items()}
    df["divisor"] = [divisor[key] for key in zip(df["patient_id"], df["modality"])]
    df["patient_id_new"] = [f"{pid}:{int(np.ceil(sl / dv))}" for pid, sl, dv in zip(df["patient_id"], df["slice"], df["divisor"])]
    df_pet, df_ct = df[df["modality"] == modality_a], df[df["modality"] == modality_b]
    windows = []
    for patient_id in df_ct["patient_id"].unique():
        rows = df_ct[df_ct["patient_id"] == patient_id]
        window = int(rows["divisor"].max())
        for i in range(len(rows["slice"].unique()) - window):
            part = rows[(rows["slice"] >= i) & (rows["slice"] <= i + window)].copy()
            part["patient_id_new"] = f"{patient_id}:{i}"
            windows.append(part)
    out = pd.concat(windows + [df_pet], axis=0) if windows else df_pet
    return out.reset_index(drop=True)


def get_label_encoder(df):
    """reference: train_models.py:492-499 -- one-hot encoder over the sorted label values (label map = identity on them)."""
    from sklearn.preprocessing import OneHotEncoder
    names = sorted(df["label"].unique())
    enc = OneHotEncoder(handle_unknown="ignore")
    enc.fit(np.array(names).reshape(-1, 1))
    return enc


class PETCTDataset3D(torch.utils.data.Dataset):
    """reference: train_models.py:47-141 -- one item = (CT tokens, PET tokens, one-hot label, patient id) of a sample window.

    Same constructor, table handling and -- call for call -- the same draws from NumPy's global generator (position noise,
    scale noise, flip / angle, window id, random slice crop), so a seeded run visits the same samples as the reference.
    The token sequences come from the device gather (``get_features``: bit-identical to ``_get_features``, :143-182) and
    stay on the GPU.  ``store`` selects where slice features and masks are read from: ``None`` = the HDF5 files the
    reference writes (needs h5py), or a mapping ``{path: {"<pid>/features/<id>": array, "<pid>/masks/<id>": array}}``;
    ``gather`` replaces the token gather (tests inject the CPU oracle)."""

    def __init__(self, dataframe, label_encoder, hdf5_ct_path, hdf5_pet_path, modality_a="pet", modality_b="ct",
                 use_augmentation=False, feature_dim=256, arch="conv", device="cuda:0", store=None, gather=None):
        if arch != "transformer":
            raise NotImplementedError("arch='conv' (Conv3d classifier) is outside the hot path")
        self.modality_a, self.modality_b = modality_a, modality_b
        self.slice_per_modality = dataframe.groupby(["patient_id", "modality"])["slice"].max()
        ct = dataframe[dataframe["modality"] == modality_b].reset_index(drop=True)
        pet = dataframe[dataframe["modality"] == modality_a].reset_index(drop=True)
        if use_augmentation:
            # one row per patient carrying its HIGHEST window index, repeated so that an epoch has about as many items as windows
            n_windows = ct["patient_id_new"].nunique()
            t = ct.copy()
            t["patient_id_new_int"] = t["patient_id_new"].str.split(":").str[-1].astype(int)
            t = t.sort_values(by="patient_id_new_int", ascending=False)
            t = t.groupby(["patient_id"])[["modality", "dataset", "label", "patient_id_new", "patient_id_new_int"]].first().reset_index()
            repeat = np.clip(np.ceil(n_windows / t.shape[0]), 2, 8)
            self.dataframe = pd.DataFrame(np.repeat(t.values, repeat, axis=0), columns=t.columns)
        else:
            self.dataframe = ct.groupby(["patient_id_new"])[["modality", "dataset", "label", "patient_id"]].first().reset_index()
        self.use_augmentation = use_augmentation
        self.flip_angles = dataframe.groupby(["flip", "angle"], as_index=False).size()[["flip", "angle"]]
        self.df_ct = ct.set_index(["patient_id_new", "angle", "flip"]).sort_index()
        self.df_pet = pet.set_index(["patient_id", "angle", "flip"]).sort_index()
        self.hdf5_ct_path, self.hdf5_pet_path = hdf5_ct_path, hdf5_pet_path
        self.label_encoder, self.feature_dim, self.arch = label_encoder, feature_dim, arch
        self.device, self.store, self.gather = device, store, gather

    def __len__(self):
        return len(self.dataframe)

    def __getitem__(self, idx):
        noise_val = 10
        sample = self.dataframe.iloc[idx]
        window_id, patient_id, label = sample.patient_id_new, sample.patient_id, sample.label
        noise = np.random.random(3) * noise_val - noise_val / 2          # drawn in both modes, like the reference
        scale_noise = np.random.uniform(0.85, 1.15)
        if self.use_augmentation:
            [[flip, angle]] = self.flip_angles.sample(n=1).values
            top = sample.patient_id_new_int
            window_id = f"{patient_id}:{np.random.randint(0, top) if top > 0 else top}"
        else:
            flip, angle, noise, scale_noise = "None", 0, noise * 0, 1.0
        ct_rows = self.df_ct.loc[(window_id, angle, flip)]
        ct_slices = ct_rows["slice"].values
        first, last = ct_slices.argmin(), ct_slices.argmax()              # [first, last): the last slice is left out (:115)
        if self.use_augmentation and len(ct_slices) > 7:                   # random slice crop of 7 .. n-1 slices
            window = int(np.random.randint(7, len(ct_slices), 1)[0])
            first = int(np.random.randint(0, len(ct_slices) - window, 1)[0])
            last = first + window
        res = np.abs(ct_rows["spatial_res"].values[0]) * scale_noise
        features_ct = self._get_features(self.hdf5_ct_path, patient_id, ct_rows["feature_id"].values[first:last], angle, flip, noise, res)
        # the same relative extent of the PET series
        rel = ct_slices[first:last] / self.slice_per_modality.loc[(patient_id, self.modality_b)]
        pet_last = self.slice_per_modality[patient_id, self.modality_a]
        lo, hi = max(0, int(rel.min() * pet_last)), min(pet_last, int(rel.max() * pet_last))
        pet_rows = self.df_pet.loc[(patient_id, angle, flip)]
        res = np.abs(pet_rows["spatial_res"].values[0]) * scale_noise
        ids = pet_rows[np.logical_and(pet_rows["slice"] >= lo, pet_rows["slice"] <= hi)]["feature_id"].values
        features_pet = self._get_features(self.hdf5_pet_path, patient_id, ids, angle, flip, noise, res)
        onehot = self.label_encoder.transform(np.array(label).reshape(-1, 1)).toarray()
        return features_ct, features_pet, torch.as_tensor(onehot, dtype=torch.float32), patient_id

    def _read(self, path, patient_id, feature_ids):
        if self.store is not None:
            f = self.store[str(path)]
            return ([np.asarray(f[f"{patient_id}/features/{i}"]) for i in feature_ids],
                    [np.asarray(f[f"{patient_id}/masks/{i}"]) for i in feature_ids])
        from .tfds_dense_descriptor import _h5py
        with _h5py().File(path, "r") as f:
            return ([f[f"{patient_id}/features/{i}"][()] for i in feature_ids],
                    [f[f"{patient_id}/masks/{i}"][()] for i in feature_ids])

    def _get_features(self, hdf5_path, patient_id, feature_ids, angle, flip, noise, spatial_res):
        """reference: :143-182 -- (n_sel, feature_dim) f32 tokens = features[mask] + PE/4 (``angle`` / ``flip`` select the
        rows upstream; the stored arrays are already transformed)."""
        feats, masks = self._read(hdf5_path, patient_id, feature_ids)
        if not feats:
            raise ValueError("need at least one array to stack")      # the reference's failure for an empty window
        if self.gather is not None:
            return torch.as_tensor(self.gather(feats, masks, spatial_res, noise, self.feature_dim), dtype=torch.float32)
        return get_features(feats, masks, spatial_res, noise=noise, feature_dim=self.feature_dim, device=self.device)


class FocalLoss(nn.Module):
    """reference: train_models.py:381-405 -- sum over samples of -alpha_c (1-p_c)^gamma log p_c at the
    true class c = argmax(one-hot target).  (2-element tensors: host-launched elementwise math.)"""

    def __init__(self, gamma=2, alpha=None):
        super().__init__()
        self.gamma = gamma
        self.weight = alpha

    def forward(self, inputs, targets):
        if inputs.dim() == 1:
            inputs, targets = inputs.unsqueeze(0), targets.unsqueeze(0)
        cls_idx = torch.argmax(targets, dim=1)
        logpt = F.log_softmax(inputs, dim=1)
        logpt = (1 - torch.exp(logpt)) ** self.gamma * logpt
        return F.nll_loss(logpt, cls_idx, self.weight, reduction="sum")


class CrossModalFocalLoss(nn.Module):
    """reference: train_models.py:332-378 -- bimodal focal term on the fused logits plus two unimodal
    focal terms whose modulating factor uses the harmonic mean of the two unimodal probabilities:
    beta * L_petct + (1-beta) * (L_ct + L_pet), each a class-weighted MEAN nll (FocalLoss sums)."""

    def __init__(self, gamma_bimodal=0, gamma_unimodal=2, alpha=None, beta=0.5):
        super().__init__()
        self.gamma_bimodal, self.gamma_unimodal = gamma_bimodal, gamma_unimodal
        self.alpha, self.beta, self.eps = alpha, beta, 1e-8

    def forward(self, inputs_petct, inputs_ct, inputs_pet, targets):
        if inputs_petct.dim() == 1:
            inputs_petct, inputs_ct, inputs_pet, targets = (t.unsqueeze(0) for t in (inputs_petct, inputs_ct, inputs_pet, targets))
        cls_idx = torch.argmax(targets, dim=1)
        lp_x, lp_c, lp_p = (F.log_softmax(t, dim=1) for t in (inputs_petct, inputs_ct, inputs_pet))
        nll = lambda lp: F.nll_loss(lp, cls_idx, self.alpha, reduction="mean")          # noqa: E731
        loss_x = nll((1 - torch.exp(lp_x)) ** self.gamma_bimodal * lp_x)
        pt_c, pt_p = torch.exp(lp_c), torch.exp(lp_p)
        pt_mean = (2 * pt_c * pt_p) / (pt_c + pt_p + self.eps)
        loss_c = nll((1 - pt_mean * pt_c) ** self.gamma_unimodal * lp_c)
        loss_p = nll((1 - pt_mean * pt_p) ** self.gamma_unimodal * lp_p)
        return self.beta * loss_x + (1 - self.beta) * (loss_c + loss_p)


def make_criterion(loss_func, device):
    """reference: train_models.py:591-598 -- class weights (0.25, 0.75); 'crossmodal' = gamma 2 / 1, beta 0.6."""
    alpha = torch.tensor([0.25, 0.75], device=device)
    if loss_func == "crossmodal":
        return CrossModalFocalLoss(alpha=alpha, gamma_unimodal=2.0, gamma_bimodal=1.0, beta=0.6)
    return FocalLoss(alpha=alpha, gamma=2.0)


def get_y_true_and_pred(y_true, y_pred, cpu=False):
    """reference: train_models.py:283-311."""
    y_true, y_pred = torch.squeeze(y_true), torch.squeeze(y_pred)
    assert y_pred.size() == y_true.size()
    if y_true.dim() == 1:
        y_pred, y_true = y_pred.unsqueeze(0), y_true.unsqueeze(0)
    y_score = F.softmax(y_pred, dim=1)
    y_true = torch.argmax(y_true, dim=1)
    if cpu:
        y_true, y_score = y_true.detach().cpu().numpy(), y_score.detach().cpu().numpy()
    return y_true, y_score


def get_sampler_weights(train_labels):
    """reference: train_models.py:313-328 -- 1 / (count of the element's value) for every element."""
    values, counts = np.unique(train_labels, return_counts=True)
    per_value = dict(zip(values, counts))
    return [1 / per_value[v] for v in train_labels]


def get_number_of_params(model):
    """reference: train_models.py:450-453 -- number of trainable parameters."""
    return int(sum(p.numel() for p in model.parameters() if p.requires_grad))


#: the scalar entries a split report carries next to sklearn's per-class rows (reference: train_models.py:196-197)
REPORT_GLOBALS = ("accuracy", "ROC AUC", "kfold", "loss", "epoch", "split")


def print_classification_report(report, global_metrics=None, echo=True):
    """reference: train_models.py:185-218 -- the text form of a split report: one line with the scalar metrics, then sklearn's
    per-class table, both rendered by pandas (the string the reference keeps in its metrics table, :782-783).
    Same text as the reference for the same dict; ``echo=False`` skips the print."""
    names = list(global_metrics) if global_metrics is not None else list(REPORT_GLOBALS)
    table = pd.DataFrame(report).round(3)
    n_cols = len(table.index)                         # precision / recall / f1-score / support
    table = table.T.astype(str)
    support = table.loc["macro avg"].iloc[-1]
    for name in names:                                # a scalar fills its whole column: keep one copy under 'f1-score'
        value = table.loc[name].iloc[-2]
        table.loc[name] = [" "] * (n_cols - 2) + [value, support]
    per_class = table.loc[[r for r in table.index if r not in names]]
    scalars = table.loc[names].T[-2:-1]
    scalars.index = ["   "]
    text = f"\n{scalars}\n\n{per_class}\n\n"
    if echo:
        print(text)
    return text


def split_report(y_true, y_score, patient_ids, loss, kfold, epoch, split):
    """The per-split report the epoch loop writes to ``<split>_metrics_<epoch>.json`` (reference: train_models.py:727-768):
    sklearn's classification_report at threshold 0.5 on the positive-class score plus 'ROC AUC', 'kfold', 'loss', 'epoch',
    'split', every sample weighted by 1 / (items of its patient) so that patients, not windows, count equally."""
    from sklearn.metrics import classification_report, roc_auc_score
    y_true = np.concatenate([np.atleast_1d(v) for v in y_true], axis=0)
    score = np.concatenate([np.atleast_2d(v) for v in y_score], axis=0)[:, 1]
    weights = get_sampler_weights(np.concatenate([np.atleast_1d(np.array(v)) for v in patient_ids], axis=0))
    report = classification_report(y_true, (score >= 0.5) * 1, output_dict=True, zero_division=0, sample_weight=weights)
    report.update({"ROC AUC": roc_auc_score(y_true, score, sample_weight=weights), "kfold": kfold, "loss": loss, "epoch": epoch,
                   "split": split})
    return report


def epoch_policy(history, patience):
    """Checkpoint / early-stopping decision at the end of an epoch (reference: train_models.py:786-810).  ``history`` = the
    fold's per-epoch records so far (dicts or DataFrame rows with 'epoch', 'test_auc', 'test_f1'), the current epoch last.
    target = test_auc^2 * sqrt(test_f1); a checkpoint is written when the current target is at least the fold's mean;
    training stops when the FIRST epoch that reached the fold's maximum lies ``patience`` or more epochs back.
    Returns (save_checkpoint, stop, target_metric of the current epoch)."""
    df = pd.DataFrame(list(history)) if not isinstance(history, pd.DataFrame) else history.copy()
    df["target_metric"] = df["test_auc"] * df["test_auc"] * np.sqrt(df["test_f1"])
    df["is_improvement"] = df["target_metric"] >= df["target_metric"].max()
    df = df.sort_values(by="epoch", ascending=True).reset_index(drop=True)
    epoch = df["epoch"].iloc[-1]
    since = epoch - df.iloc[df["is_improvement"].argmax()]["epoch"]
    save = bool(df["target_metric"].iloc[-1] >= df["target_metric"].mean())
    return save, bool(since >= patience), float(df["target_metric"].iloc[-1])


def build_model(cfg, arch, modality, modality_a="pet", modality_b="ct", num_classes=2):
    """reference: train_models.py:455-486.  'petct'/'petchest' build the bimodal classifier with the CT
    encoder configured from ``modality_b`` and the PET encoder from ``modality_a`` (:459-472)."""
    cfg_model = cfg["models"][arch]
    feature_dim = cfg_model["feature_dim"]
    if modality in ("petct", "petchest"):
        ct, pet = cfg_model[modality_b], cfg_model[modality_a]
        return TransformerNoduleBimodalClassifier(feature_dim, ct["mlp_ratio"], pet["mlp_ratio"], ct["num_heads"], pet["num_heads"],
                                                  ct["num_layers"], pet["num_layers"], num_classes=num_classes)
    if arch == "conv":
        raise NotImplementedError("Conv3d classifier is out of scope (north star names the transformer)")
    m = cfg_model[modality]
    return TransformerNoduleClassifier(input_dim=feature_dim, dim_feedforward=int(feature_dim * m["mlp_ratio"]),
                                       num_heads=m["num_heads"], num_classes=num_classes, num_layers=m["num_layers"])


def make_optimizer(model, cfg, arch="transformer"):
    """reference: train_models.py:600-601 -- AdamW(lr, wd 0.01) + cosine annealing to 1e-4 over 0.8*epochs."""
    c = cfg["models"][arch]
    opt = torch.optim.AdamW(model.parameters(), lr=c["learning_rate"], weight_decay=0.01)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=int(c["num_epochs"] * 0.8), eta_min=1e-4)
    return opt, sched


def train_epoch(model, samples, criterion, optimizer, virtual_batch_size=32, grad_sync=None, cuda_graphs=False):
    """One pass of the reference's accumulation loop (train_models.py:652-688).

    samples: iterable of (tokens (n, d) f32 CUDA, one-hot label (C,) f32 CUDA); for the bimodal model
    (tokens_ct, tokens_pet, label), and ``criterion`` may be CrossModalFocalLoss (:668-672: it takes
    outputs[0], outputs[2], outputs[3]).
    Loss is divided by iters_to_accumulate = min(virtual_batch, len(samples)) (:655,674); the optimizer
    steps every iters_to_accumulate samples and at the last sample (:685-687).  ``grad_sync`` (optional
    callable) is invoked right before each optimizer step: the data-parallel gradient all-reduce.
    ``cuda_graphs``: forward + loss + backward of a sample run as one CUDA graph per token count (per (CT, PET) pair of counts for
    the bimodal model; graph_step.GraphedTrainStep: captured the second time a length is seen, same kernels in the same order as
    the eager step).
    Returns (mean loss, list of softmax scores)."""
    from .distributed import zero_grads
    samples = list(samples)
    iters = min(virtual_batch_size, len(samples))
    model.train()
    step = None
    if cuda_graphs and isinstance(model, (TransformerNoduleClassifier, TransformerNoduleBimodalClassifier)):
        from .graph_step import graphed_step
        step = graphed_step(model, criterion)
    zero_grads(model, optimizer)
    total, scores = 0.0, []
    for i, sample in enumerate(samples):
        label = sample[-1]
        if step is not None:
            loss, logits = step(sample[0] if len(sample) == 2 else tuple(sample[:-1]), label, 1.0 / iters)
        else:
            outputs = model(*(t.unsqueeze(0) for t in sample[:-1]))
            logits = outputs[0]
            if isinstance(criterion, CrossModalFocalLoss):
                loss = criterion(torch.squeeze(logits), torch.squeeze(outputs[2]), torch.squeeze(outputs[3]), label) / iters
            else:
                loss = criterion(torch.squeeze(logits), label) / iters
            loss.backward()
        total += float(loss.item()) * iters
        scores.append(torch.softmax(logits.detach(), dim=1)[0].cpu().numpy())
        if (i + 1) % iters == 0 or (i + 1) == len(samples):
            if grad_sync is not None:
                grad_sync(model)
            optimizer.step()
            zero_grads(model, optimizer)
    return total / max(len(samples), 1), scores


def _forward_item(model, item, modality, device):
    """One dataset item through the model as the reference's loops do (train_models.py:656-667): which token sequences go in
    depends on the modality flag.  Returns (outputs, one-hot label on the device, patient id)."""
    ct, pet, onehot, patient_id = item
    label = torch.squeeze(torch.as_tensor(onehot)).to(device)
    if modality in ("petct", "petchest"):
        outputs = model(ct.to(device).unsqueeze(0), pet.to(device).unsqueeze(0))
    elif modality == "pet":
        outputs = model(pet.to(device).unsqueeze(0))
    else:                                     # 'ct' / 'chest'
        outputs = model(ct.to(device).unsqueeze(0))
    return outputs, label, patient_id


def _item_loss(criterion, outputs, label):
    if isinstance(criterion, CrossModalFocalLoss):
        return criterion(torch.squeeze(outputs[0]), torch.squeeze(outputs[2]), torch.squeeze(outputs[3]), label)
    return criterion(torch.squeeze(outputs[0]), label)


def run_epoch(model, dataset, order, criterion, modality, device, optimizer=None, virtual_batch_size=32, grad_sync=None,
              rank=0, world=1, cuda_graphs=False):
    """One pass over ``dataset`` in ``order`` (batch size 1, as the reference's loaders, :639-640).  With ``optimizer`` it is
    the training loop (:648-688: loss / iters_to_accumulate, step every iters_to_accumulate items and at the last one),
    without it the evaluation loop (:689-718, no gradients).

    Data parallel (``world`` > 1): an accumulation window is the same ``iters_to_accumulate`` consecutive
    items of ``order`` on every rank; rank r takes the window's items r, r + world, ...; the loss keeps its GLOBAL 1 / iters
    scale (:674), so ``grad_sync`` (``distributed.allreduce_grads``: one sum over the ranks) right before each optimizer step
    yields the single-process gradients.  ``order`` must be the same list on every rank.  Labels / scores / ids / the loss
    are gathered, so every rank returns the whole epoch's records.
    ``cuda_graphs`` (training pass): each sample's forward + loss + backward replays a CUDA graph kept per token
    count (per pair of counts for the two clouds of the bimodal model) (graph_step.GraphedTrainStep); lengths seen for the first time run eagerly.
    Returns (mean loss, y_true list, y_score list, patient ids)."""
    import torch.distributed as dist
    from .distributed import zero_grads
    train = optimizer is not None
    order = list(order)
    iters = min(virtual_batch_size, len(order)) if train else max(len(order), 1)
    model.train(train)
    if train:
        zero_grads(model, optimizer)
    total, y_true, y_score, pids = 0.0, [], [], []
    step = None
    if train and cuda_graphs and isinstance(model, (TransformerNoduleClassifier, TransformerNoduleBimodalClassifier)):
        from .graph_step import graphed_step
        step = graphed_step(model, criterion)
    with torch.enable_grad() if train else torch.no_grad():
        for w0 in range(0, len(order), iters):
            for idx in order[w0:w0 + iters][rank::world]:
                if step is not None:
                    ct, pet, onehot, pid = dataset[int(idx)]
                    label = torch.squeeze(torch.as_tensor(onehot)).to(device)
                    if modality in ("petct", "petchest"):
                        loss, logits = step((ct.to(device), pet.to(device)), label, 1.0 / iters)
                    else:
                        loss, logits = step((pet if modality == "pet" else ct).to(device), label, 1.0 / iters)
                    outputs = (logits,)
                else:
                    outputs, label, pid = _forward_item(model, dataset[int(idx)], modality, device)
                    loss = _item_loss(criterion, outputs, label) / (iters if train else 1)
                yt, ys = get_y_true_and_pred(y_true=label, y_pred=outputs[0], cpu=True)
                y_true.append(yt)
                y_score.append(ys)
                pids.append(np.array([pid]))
                total += float(loss.item()) * (iters if train else 1)           # :681, :716
                if train and step is None:
                    loss.backward()
            if train:
                if grad_sync is not None:
                    grad_sync(model)
                optimizer.step()
                zero_grads(model, optimizer)
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, (total, y_true, y_score, pids))
        total = sum(p[0] for p in parts)
        y_true, y_score, pids = ([x for p in parts for x in p[k]] for k in (1, 2, 3))
    return total / max(len(order), 1), y_true, y_score, pids


def run_fold(cfg, arch, modality, df_train, df_test, label_encoder, hdf5_ct_path, hdf5_pet_path, save_dir, kfold,
             loss_func="focal", device="cuda:0", modality_a="pet", modality_b="ct", store=None, num_epochs=None, grad_sync=None,
             gather=None, rank=0, world=1, cuda_graphs=False):
    """One fold of the reference's training script (train_models.py:562-810): model / criterion / AdamW + cosine schedule from
    the YAML, train and test datasets (augmentation on / off), per epoch a shuffled training pass, an evaluation pass, the
    scheduler step, the patient-weighted reports written to ``<split>_metrics_<epoch>.json``, a checkpoint when the target
    metric is at least the fold's mean and early stopping after ``patience`` epochs without a new maximum.
    The loss plot (plotly, :797-798) is not produced.  Returns the fold's history (one dict per epoch).
    ``world`` > 1 (one process per GPU, process group initialised): the fold is data parallel over the virtual batch
    (``run_epoch``); rank 0's initial weights and epoch orders are broadcast, rank 0 alone writes files."""
    import json
    import torch.distributed as dist
    from .distributed import allreduce_grads
    from .models_archs import save_checkpoint
    if rank == 0:
        os.makedirs(save_dir, exist_ok=True)
    cfg_model = cfg["models"][arch]
    if torch.device(device).type == "cuda":
        torch.cuda.set_device(device)        # libvdr launches on the current device / its current stream (ops._req checks)
    model = build_model(cfg, arch, modality, modality_a, modality_b, num_classes=2).to(device)
    if world > 1:
        with torch.no_grad():
            for p_ in model.parameters():
                dist.broadcast(p_, 0)        # in place on the parameter itself: bumps its version counter, so the cached bf16 operand copies refresh
        grad_sync = grad_sync or allreduce_grads
    criterion = make_criterion(loss_func, device)
    optimizer, scheduler = make_optimizer(model, cfg, arch)
    kw = dict(label_encoder=label_encoder, hdf5_ct_path=hdf5_ct_path, hdf5_pet_path=hdf5_pet_path, modality_a=modality_a,
              modality_b=modality_b, feature_dim=cfg_model["feature_dim"], arch=arch, device=device, store=store, gather=gather)
    train_ds = PETCTDataset3D(df_train, use_augmentation=True, **kw)
    test_ds = PETCTDataset3D(df_test, use_augmentation=False, **kw)
    history = []
    for epoch in range(num_epochs if num_epochs is not None else cfg_model["num_epochs"]):
        order = [torch.randperm(len(train_ds)).tolist()]                                 # DataLoader(shuffle=True), :639
        if world > 1:
            dist.broadcast_object_list(order, 0)
        tr_loss, yt, ys, pid = run_epoch(model, train_ds, order[0], criterion, modality, device, optimizer,
                                         cfg_model["virtual_batch_size"], grad_sync, rank, world, cuda_graphs=cuda_graphs)
        te_loss, yt2, ys2, pid2 = run_epoch(model, test_ds, range(len(test_ds)), criterion, modality, device, rank=rank, world=world)
        scheduler.step()
        train_report = split_report(yt, ys, pid, tr_loss, kfold, epoch, "train")
        test_report = split_report(yt2, ys2, pid2, te_loss, kfold, epoch, "test")
        for name, rep in (("train", train_report), ("test", test_report)):
            if rank == 0:
                with open(os.path.join(save_dir, f"{name}_metrics_{epoch}.json"), "w") as fh:
                    json.dump(rep, fh)
        history.append(dict(kfold=kfold, epoch=epoch, train_loss=tr_loss, test_loss=te_loss, train_auc=train_report["ROC AUC"],
                            test_auc=test_report["ROC AUC"], train_f1=train_report["macro avg"]["f1-score"],
                            test_f1=test_report["macro avg"]["f1-score"],
                            train_report=print_classification_report(train_report, echo=rank == 0).replace("\n", "<br>").replace(" ", "  "),
                            test_report=print_classification_report(test_report, echo=rank == 0).replace("\n", "<br>").replace(" ", "  ")))
        save, stop, _ = epoch_policy(history, cfg_model["patience"])
        if save and rank == 0:
            save_checkpoint(model, save_dir, epoch)
        if stop:
            if rank == 0:
                print(f"Early stopping triggered after {epoch + 1} epochs")
            break
    return history


def main(argv=None):
    """The reference's training entry point (train_models.py:500-810) on the libvdr classifier: same flags, same relative
    paths (../data/features/features_masks_<modality>.hdf5, ../data/features/petct.parquet, ../models/<experiment>/...).
    Launched with torchrun (one process per GPU) every fold trains data parallel: NCCL all-reduce of the gradients per optimizer step."""
    args = build_arg_parser().parse_args(argv)
    from .config_manager import load_conf
    from .distributed import init_distributed
    rank, world = init_distributed()                       # under torchrun: data parallel over the virtual batch
    device = f"cuda:{int(os.environ.get('LOCAL_RANK', '0'))}" if world > 1 else f"cuda:{args.gpu}"
    torch.cuda.set_device(device)
    modality_a, modality_b = "pet", ("chest" if "chest" in args.modality else "ct")
    hdf5_pet = os.path.join("..", "data", "features", f"features_masks_{modality_a}.hdf5")
    hdf5_ct = os.path.join("..", "data", "features", f"features_masks_{modality_b}.hdf5")
    models_dir = os.path.join("..", "models", args.experiment, f"{args.backbone}_{args.arch}_{args.dataset}")
    cfg = load_conf()
    df = pd.read_parquet(os.path.join("..", "data", "features", "petct.parquet"))
    df["flip"] = df["flip"].astype(str)
    df.reset_index(drop=True, inplace=True)
    df = prepare_df(df, modality_a, modality_b)
    encoder = get_label_encoder(df)
    folds = cfg["kfold_patients"][modality_b][args.dataset]
    histories = {}
    for kfold in folds:
        split = folds[kfold]
        df_train = df[df["patient_id"].isin(split["train"])].reset_index(drop=True)
        df_test = df[df["patient_id"].isin(split["test"])].reset_index(drop=True)
        save_dir = os.path.join(models_dir, args.modality, f"kfold_{kfold}")
        histories[kfold] = run_fold(cfg, args.arch, args.modality, df_train, df_test, encoder, hdf5_ct, hdf5_pet, save_dir, kfold,
                                    loss_func=args.loss, device=device, modality_a=modality_a, modality_b=modality_b, rank=rank, world=world,
                                    cuda_graphs=getattr(args, "cuda_graphs", False))
    return histories


def build_arg_parser():
    """Same flags as the reference CLI (train_models.py:500-515)."""
    p = argparse.ArgumentParser(description="Train the point-cloud transformer for lung nodule classification")
    p.add_argument("-a", "--arch", type=str, default="transformer")
    p.add_argument("-d", "--dataset", type=str, default="stanford")
    p.add_argument("-b", "--backbone", type=str, default="medsam")
    p.add_argument("-m", "--modality", type=str, default="petchest")
    p.add_argument("-gpu", "--gpu", type=int, default=0)
    p.add_argument("-l", "--loss", type=str, default="focal")
    p.add_argument("-e", "--experiment", type=str, default="petct")
    # not in the reference: replay each training sample as a CUDA graph per cloud length (unimodal models; graph_step.py)
    p.add_argument("--cuda_graphs", action="store_true")
    return p


if __name__ == "__main__":
    main()
