"""The committed golden vectors are what the UNMODIFIED reference produces today: tests/golden/make_golden.py is re-run into a
temporary directory and every array is compared with the committed file (integers / booleans / NumPy float64 exactly; torch float32
outputs to 1e-6, they were bit-identical when this test was written).  Needs /root/reference (this container only)."""
import importlib.util
import os
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def test_make_golden_reproduces_the_committed_fixtures(tmp_path):
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference checkout not present")
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.OUT = str(tmp_path)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for fn in (mod.golden_gather, mod.golden_pointcloud, mod.golden_geometry, mod.golden_classifier, mod.golden_bimodal,
                   mod.golden_crossmodal):
            fn()
    made = sorted(os.listdir(tmp_path))
    assert made == ["bimodal_small.npz", "classifier_small.npz", "crossmodal_loss.npz", "gather_g1.npz", "geometry.npz", "pointcloud_g2.npz"]
    n = 0
    for f in made:
        new, old = np.load(tmp_path / f, allow_pickle=True), np.load(os.path.join(GOLDEN, f), allow_pickle=True)
        assert sorted(new.files) == sorted(old.files), f
        for k in new.files:
            x, y = new[k], old[k]
            assert x.shape == y.shape and x.dtype == y.dtype, (f, k)
            if x.dtype == np.float32:
                assert np.allclose(x, y, rtol=1e-6, atol=1e-6, equal_nan=True), (f, k)
            elif x.dtype.kind == "f":
                assert np.array_equal(x, y, equal_nan=True), (f, k)
            else:
                assert np.array_equal(x, y), (f, k)
            n += 1
    assert n > 300
