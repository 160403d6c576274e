"""Merge the per-patient feature tables into data/features/petct.parquet -- the reference's
src/merge_dataframe_features.py:12-30 as a function (same column handling: ``flip`` as str, ``augmentation`` = not the
identity copy, fresh index)."""
from __future__ import annotations

import os

import numpy as np
import pandas as pd

DATASETS = ("santa_maria_dataset", "stanford_dataset")


def merge_features(feature_dir: str, datasets=DATASETS) -> pd.DataFrame:
    parts = []
    for dataset in datasets:
        d = os.path.join(feature_dir, dataset)
        if os.path.exists(d):
            parts += [pd.read_parquet(os.path.join(d, fn)) for fn in os.listdir(d)]
    df = pd.concat(parts)
    df["flip"] = df["flip"].astype(str)
    df["augmentation"] = np.logical_not(np.logical_and(df["flip"] == "None", df["angle"] == 0))
    return df.reset_index(drop=True)


def main(feature_dir: str = os.path.join("..", "data", "features")) -> str:
    out = os.path.join(feature_dir, "petct.parquet")
    merge_features(feature_dir).to_parquet(out)
    return out


if __name__ == "__main__":
    main()
