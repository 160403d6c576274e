"""The resize restatement (oracle/resize_np.py) against scipy.ndimage, the library skimage.transform.resize delegates to."""
import numpy as np
import pytest
from scipy import ndimage as ndi

from oracle import resize_np


@pytest.mark.parametrize("ih,iw,oh,ow", [(37, 53, 64, 64), (200, 180, 512, 512), (64, 64, 48, 40), (300, 260, 128, 224), (96, 96, 96, 96)])
def test_resize_matches_scipy_pipeline(ih, iw, oh, ow):
    rng = np.random.default_rng(ih * 1000 + ow)
    img = rng.random((ih, iw))
    ref = img
    if oh < ih or ow < iw:
        ref = ndi.gaussian_filter(ref, (max(0.0, (ih / oh - 1) / 2), max(0.0, (iw / ow - 1) / 2)), mode="mirror")
    ref = ndi.zoom(ref, (oh / ih, ow / iw), order=1, mode="mirror", grid_mode=True)
    got = resize_np.resize(img, (oh, ow))
    assert got.shape == (oh, ow)
    assert np.abs(got - ref).max() < 1e-12


def test_resize_identity_and_channels():
    rng = np.random.default_rng(5)
    img = rng.random((40, 56, 3))
    assert np.array_equal(resize_np.resize(img, (40, 56)), img)
    out = resize_np.resize(img, (80, 70))
    for c in range(3):
        assert np.allclose(out[..., c], resize_np.resize(img[..., c], (80, 70)), atol=1e-14)
    assert out.min() >= img.min() - 1e-12 and out.max() <= img.max() + 1e-12      # convex weights: the clip is a no-op
