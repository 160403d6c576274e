"""Patient-level stratified k-fold split -- the reference's src/split_patients.py:16-43 as functions.

The reference is a script: it reads data/features/petct.parquet, and for every modality and dataset writes the train / test
patient ids of five stratified folds (``StratifiedKFold(n_splits=5, shuffle=True, random_state=42)`` over one label per
patient) to conf/parameters_kfold.yaml under ``kfold_patients`` -- the file train_models.py reads its folds from (:559-566).
"""
from __future__ import annotations

import os

import pandas as pd
import yaml

from .config_manager import get_project_dir


def kfold_patients(df_all: pd.DataFrame, n_splits: int = 5, seed: int = 42) -> dict:
    """{modality: {dataset: {fold: {'train': [...], 'test': [...]}}}} with the reference's ordering: patients sorted by id
    (groupby), label = the first row's, one splitter per modality re-used across its datasets."""
    from sklearn.model_selection import StratifiedKFold
    out = {modality: {} for modality in df_all["modality"].unique()}
    for modality in out:
        skf = StratifiedKFold(n_splits=n_splits, shuffle=True, random_state=seed)
        df = df_all[df_all["modality"] == modality].reset_index(drop=True)
        for dataset in df["dataset"].unique():
            first_label = df[df["dataset"] == dataset].groupby(["patient_id"])["label"].first()
            patients, labels = first_label.index.to_list(), first_label.to_list()
            out[modality][dataset] = {
                fold: {"train": [patients[i] for i in tr], "test": [patients[i] for i in te]}
                for fold, (tr, te) in enumerate(skf.split(patients, labels))}
    return out


def main(project_dir: str | None = None) -> str:
    project_dir = project_dir or get_project_dir()
    df_all = pd.read_parquet(os.path.join(project_dir, "data", "features", "petct.parquet"))
    path = os.path.join(project_dir, "conf", "parameters_kfold.yaml")
    with open(path, "w") as f:
        yaml.dump({"kfold_patients": kfold_patients(df_all)}, f)
    return path


if __name__ == "__main__":
    main()
