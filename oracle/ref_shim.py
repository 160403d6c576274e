"""TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference modules from /root/reference/src.

The reference (larosi/vit-deep-radiomics) is pure Python but depends on packages that are
absent from this image (h5py, plotly, skimage, tensorflow_datasets, segment_anything).
This shim injects minimal stand-ins for those into ``sys.modules`` and then imports the
reference's own modules so that their functions can be run as ground truth:

  * ``h5py.File``  -> dict-backed in-memory store (enough for ``_get_features``,
    ``train_models.py:146-150`` and ``save_features``, ``tfds_dense_descriptor.py:151-165``)
  * ``skimage.transform.resize`` -> scipy restatement (nearest for bool / order 0;
    bilinear without anti-aliasing otherwise).  skimage itself cannot be diffed here.
  * ``skimage.color.gray2rgb`` -> channel stack
  * plotly / tensorflow_datasets / segment_anything / other skimage sub-modules -> empty stubs

It only works where ``/root/reference`` exists (the build container).  It is used by
``tests/golden/make_golden.py`` (to freeze golden vectors) and by the ``not gpu`` tests that
re-check the oracle restatements against the live reference.  Nothing in the product package
imports this file.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np

REFERENCE_SRC = os.environ.get("VDR_REFERENCE_SRC", "/root/reference/src")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "train_models.py"))


# ----------------------------------------------------------------------------- h5py stand-in
class _Dataset:
    def __init__(self, data):
        self._data = np.asarray(data)

    def __getitem__(self, key):
        if key == ():
            return self._data.copy()
        return self._data[key]

    @property
    def shape(self):
        return self._data.shape


class _Group:
    def __init__(self, store, prefix):
        self._store = store
        self._prefix = prefix.strip("/")

    def _full(self, key):
        key = str(key).strip("/")
        return f"{self._prefix}/{key}" if self._prefix else key

    def __contains__(self, key):
        full = self._full(key)
        return any(k == full or k.startswith(full + "/") for k in self._store)

    def __getitem__(self, key):
        full = self._full(key)
        if full in self._store:
            return _Dataset(self._store[full])
        if any(k.startswith(full + "/") for k in self._store):
            return _Group(self._store, full)
        raise KeyError(full)

    def __delitem__(self, key):
        full = self._full(key)
        for k in [k for k in self._store if k == full or k.startswith(full + "/")]:
            del self._store[k]

    def keys(self):
        pre = self._prefix + "/" if self._prefix else ""
        out = []
        for k in self._store:
            if k.startswith(pre):
                head = k[len(pre):].split("/")[0]
                if head not in out:
                    out.append(head)
        return out

    def create_group(self, name):
        return _Group(self._store, self._full(name))

    def create_dataset(self, name, data=None, **_kw):
        self._store[self._full(name)] = np.array(data)
        return _Dataset(self._store[self._full(name)])


#: in-memory "files": path -> {dataset path -> ndarray}
H5_FILES: dict[str, dict[str, np.ndarray]] = {}


class _File(_Group):
    def __init__(self, path, mode="r", **_kw):
        store = H5_FILES.setdefault(str(path), {})
        super().__init__(store, "")

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def close(self):
        pass


# ----------------------------------------------------------------------------- skimage stand-ins
def _resize(image, output_shape, order=None, mode="reflect", anti_aliasing=None, **_kw):
    """Restatement of skimage.transform.resize for the two ways the reference calls it.

    * bool input / order=0 (train_models.py:151): nearest-neighbour with pixel-centre
      ("grid_mode") sampling, no anti-aliasing; returns bool for bool input.
    * float input, default order (tfds_dense_descriptor.py:42,44): order-1 spline with
      grid_mode and mirror boundary.  (skimage also Gaussian-prefilters when down-scaling;
      the synthetic configs never down-scale, so that branch is not restated.)
    """
    from scipy import ndimage

    image = np.asarray(image)
    output_shape = tuple(int(s) for s in output_shape)
    out_full = output_shape + image.shape[len(output_shape):]
    zoom = [o / i for o, i in zip(out_full, image.shape)]
    if image.dtype == bool:
        out = ndimage.zoom(image.astype(np.uint8), zoom, order=0, mode="mirror", grid_mode=True)
        return out.astype(bool)
    if order is None:
        order = 1
    if order == 0:
        return ndimage.zoom(image, zoom, order=0, mode="mirror", grid_mode=True)
    img = image.astype(np.float64)
    if all(abs(z - 1.0) < 1e-12 for z in zoom):
        return img
    return ndimage.zoom(img, zoom, order=order, mode="mirror", grid_mode=True)


def _gray2rgb(image):
    image = np.asarray(image)
    return np.stack([image] * 3, axis=-1)


def _install_stubs():
    def mod(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    try:
        import h5py  # noqa: F401  (real one wins if it ever appears)
    except Exception:
        mod("h5py", File=_File)
    try:
        import skimage  # noqa: F401
    except Exception:
        sk = mod("skimage")
        sk.io = mod("skimage.io")
        sk.transform = mod("skimage.transform", resize=_resize)
        sk.color = mod("skimage.color", gray2rgb=_gray2rgb)
        sk.filters = mod("skimage.filters", threshold_otsu=lambda x: float(np.mean(x)))
        sk.segmentation = mod("skimage.segmentation", mark_boundaries=lambda img, lab, **k: img)
    try:
        import plotly  # noqa: F401
    except Exception:
        pl = mod("plotly")
        pl.graph_objs = mod("plotly.graph_objs")
        pl.subplots = mod("plotly.subplots", make_subplots=lambda *a, **k: None)
        pl.express = mod("plotly.express")
    for name in ("tensorflow_datasets", "segment_anything"):
        try:
            importlib.import_module(name)
        except Exception:
            mod(name, sam_model_registry={})


_REF_MODULES: dict[str, types.ModuleType] = {}


def load_reference(name: str):
    """Import reference module ``name`` (e.g. 'train_models') unmodified from /root/reference/src."""
    if name in _REF_MODULES:
        return _REF_MODULES[name]
    if not reference_available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_SRC}")
    _install_stubs()
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    try:
        m = importlib.import_module(name)
    finally:
        # keep REFERENCE_SRC on the path: reference modules import each other by bare name
        pass
    _REF_MODULES[name] = m
    return m


class legacy_series_getitem:
    """Context manager for running reference code written for pandas < 2: ``series[0]`` on a Series with a non-integer index
    falls back to position 0 (the behaviour the reference's prepare_df relies on, train_models.py:424).  Test-only."""

    def __enter__(self):
        import pandas as pd
        self._pd, self._orig = pd, pd.Series.__getitem__
        orig = self._orig

        def getitem(ser, key):
            try:
                return orig(ser, key)
            except KeyError:
                if isinstance(key, (int, np.integer)) and not pd.api.types.is_integer_dtype(ser.index):
                    return ser.iloc[key]
                raise
        pd.Series.__getitem__ = getitem
        return self

    def __exit__(self, *exc):
        self._pd.Series.__getitem__ = self._orig
        return False


def put_feature_file(path: str, patient_id: str, features: list, masks: list):
    """Fill an in-memory 'HDF5' file with the layout save_features() writes
    (tfds_dense_descriptor.py:156-165): <pid>/features/<i>, <pid>/masks/<i>."""
    store = H5_FILES.setdefault(str(path), {})
    for i, (f, m) in enumerate(zip(features, masks)):
        store[f"{patient_id}/features/{i}"] = np.asarray(f)
        store[f"{patient_id}/masks/{i}"] = np.asarray(m)


def make_dataset_table(seed=0, D=12):
    """A three-patient CT + PET table in the layout the reference's parquet has (one row per stored slice and per
    flip / angle copy), with the slice features and masks in the in-memory 'HDF5' files of ref_shim."""
    import pandas as pd
    rng = np.random.default_rng(seed)
    rows = []
    H5_FILES.clear()
    for pid, nct, npet, lab in (("P1", 20, 6, 0), ("P2", 15, 4, 1), ("P3", 9, 3, 0), ("P4", 30, 9, 1)):
        for mod, n, path, res in (("ct", nct, "ct.h5", (1.0, 1.0, 2.0)), ("pet", npet, "pet.h5", (4.0, 4.0, 3.0))):
            feats, masks, fid = [], [], 0
            for flip, angle in (("None", 0), ("H", 90), ("V", 45)):
                for s_ in range(n):
                    rows.append(dict(patient_id=pid, modality=mod, slice=s_, feature_id=fid, flip=flip, angle=angle, label=lab,
                                     dataset="d", spatial_res=res))
                    fid += 1
                    feats.append(rng.standard_normal((4, 5, D)).astype(np.float32))
                    masks.append(rng.random((16, 20)) < 0.4)
            put_feature_file(path, pid, feats, masks)
    return pd.DataFrame(rows)
