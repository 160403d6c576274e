"""TEST INFRASTRUCTURE ONLY (oracle): NumPy restatement of the image resize inside the reference's ``prepare_image``
(src/tfds_dense_descriptor.py:40-44: ``skimage.transform.resize(img, (1024, 1024))`` / ``(896, 896)`` on a float image,
default arguments).

scikit-image is not installed in this image and the reference does not vendor it, so this restates the published
algorithm of skimage >= 0.19 (``skimage/transform/_warps.py: resize``): for a float image, ``order=1``, ``mode='reflect'``
(-> ndimage mode 'mirror'), ``anti_aliasing=True`` exactly when some axis shrinks, per-axis
``sigma = max(0, (in/out - 1) / 2)``, ``ndi.gaussian_filter(image, sigma, mode='mirror')`` followed by
``ndi.zoom(filtered, out/in, order=1, mode='mirror', grid_mode=True)`` and a clip to the input range (a no-op for these
convex weights).  PINNED against scipy.ndimage (which is what skimage calls; present here) in tests/test_oracle_resize.py;
parity with skimage itself is UNPINNED (package absent).  Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np


def mirror_index(i: np.ndarray, n: int) -> np.ndarray:
    """scipy.ndimage 'mirror' extension (d c b | a b c d | c b a): period 2n - 2."""
    if n == 1:
        return np.zeros_like(i)
    p = 2 * n - 2
    i = np.mod(i, p)
    return np.where(i >= n, p - i, i)


def linear_taps(n_in: int, n_out: int):
    """Source indices and weight of output sample o on one axis: coordinate (o + 0.5) * n_in / n_out - 0.5."""
    o = np.arange(n_out, dtype=np.float64)
    x = (o + 0.5) * (n_in / n_out) - 0.5
    f = np.floor(x)
    return mirror_index(f.astype(np.int64), n_in), mirror_index(f.astype(np.int64) + 1, n_in), x - f


def gaussian_weights(sigma: float, truncate: float = 4.0):
    r = int(truncate * sigma + 0.5)
    x = np.arange(-r, r + 1, dtype=np.float64)
    w = np.exp(-0.5 * (x / sigma) ** 2)
    return w / w.sum(), r


def gaussian_axis(img: np.ndarray, sigma: float, axis: int) -> np.ndarray:
    if sigma <= 1e-15:
        return img
    w, r = gaussian_weights(sigma)
    if r == 0:
        return img
    n = img.shape[axis]
    idx = mirror_index(np.arange(n)[:, None] + np.arange(-r, r + 1)[None, :], n)      # (n, taps)
    moved = np.moveaxis(img, axis, 0)
    out = np.tensordot(w, moved[idx], axes=([0], [1]))                                  # sum over taps
    return np.moveaxis(out, 0, axis)


def resize(img: np.ndarray, out_hw) -> np.ndarray:
    """float image (H, W[, C]) -> (OH, OW[, C]) float64, as skimage.transform.resize with default arguments."""
    img = np.asarray(img, dtype=np.float64)
    ih, iw = img.shape[0:2]
    oh, ow = int(out_hw[0]), int(out_hw[1])
    if oh < ih or ow < iw:                                 # anti_aliasing=None -> True when any axis shrinks
        img = gaussian_axis(img, max(0.0, (iw / ow - 1) / 2), 1)
        img = gaussian_axis(img, max(0.0, (ih / oh - 1) / 2), 0)
    y0, y1, wy = linear_taps(ih, oh)
    x0, x1, wx = linear_taps(iw, ow)
    shape_x = (1, ow) + (1,) * (img.ndim - 2)
    shape_y = (oh, 1) + (1,) * (img.ndim - 2)
    wx, wy = wx.reshape(shape_x), wy.reshape(shape_y)
    top = img[y0][:, x0] * (1 - wx) + img[y0][:, x1] * wx
    bot = img[y1][:, x0] * (1 - wx) + img[y1][:, x1] * wx
    return top * (1 - wy) + bot * wy
