"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of what ``scipy.ndimage.rotate(x, angle, axes=(0, 1), reshape=False, mode='nearest')``
(cubic spline) computes, operation for operation: the algorithm the device kernels of ``csrc/augment.cu`` implement.

reference: src/tfds_dense_descriptor.py:328-350 (rotate_image) calls exactly that scipy function; scipy is third-party code that
IS present in this image (1.18.1), so this restatement is PINNED: ``tests/test_oracle_rotate.py`` checks it bit for bit against
scipy itself (float64 / float32 / bool / uint8 inputs, square and non-square planes).  It exists so that every step of the CUDA
path has a readable counterpart (scipy's own steps live in compiled C: ni_splines.c, ni_interpolation.c):

  1. pad every plane by 12 pixels of edge values                         (ndimage._interpolation._prepad_for_spline_filter)
  2. cubic prefilter along axis 0, then axis 1                           (apply_filter, 'reflect' initialisers for mode 'nearest')
  3. cc = ((offset + i*m0) + j*m1) + 12 per axis, unmapped; start = floor(cc) - 1; cubic weights; taps clamped;
     t = sum_a sum_b (c[a, b] * w0[a]) * w1[b]                           (NI_GeometricTransform)
  4. float output = cast(t); bool output = (unsigned char)t; uint8 output = trunc(t + 0.5) for t > 0 else 0
"""
from __future__ import annotations

import math

import numpy as np

#: sqrt(3) - 2 as constant-folded into scipy's binary (2 ulp away from ``math.sqrt(3.0) - 2.0``)
POLE = float.fromhex("-0x1.126145e9ecd56p-2")
NPAD = 12


def prefilter_lines(c: np.ndarray) -> np.ndarray:
    """Cubic spline prefilter along axis 0 of a float64 array (any trailing shape), scipy's arithmetic order."""
    c = np.array(c, dtype=np.float64, copy=True)
    n, z = c.shape[0], POLE
    gain = (1.0 - 1.0 / z) * (1.0 - z)
    c *= gain
    z_n = math.pow(z, n)
    c0 = c[0].copy()
    acc = c[n - 1] * z_n + c0
    z_i = z
    for i in range(1, n):
        acc = acc + (c[n - 1 - i] * z_n + c[i]) * z_i
        z_i *= z
    c[0] = (z * acc) / (1.0 - z_n * z_n) + c0
    for i in range(1, n):
        c[i] = c[i - 1] * z + c[i]
    c[n - 1] = (z / (z - 1.0)) * c[n - 1]
    for i in range(n - 2, -1, -1):
        c[i] = (c[i + 1] - c[i]) * z
    return c


def _weights(cc):
    fl = np.floor(cc)
    y = cc - fl
    z = 1.0 - y
    w0 = (z * (z * z)) / 6.0
    w1 = (((y + -2.0) * (y * y)) * 3.0 + 4.0) / 6.0
    w2 = (((z + -2.0) * (z * z)) * 3.0 + 4.0) / 6.0
    w3 = ((1.0 - w0) - w1) - w2
    return fl.astype(np.int64) - 1, (w0, w1, w2, w3)


def rotate_plane_f64(plane: np.ndarray, angle) -> np.ndarray:
    """The interpolated values t (float64) of one (H, W) plane, before the output conversion."""
    from scipy import special
    H, W = plane.shape
    c, s = special.cosdg(angle), special.sindg(angle)
    rot = np.array([[c, s], [-s, c]])
    shp = np.array([H, W])
    off = (shp - 1) / 2 - rot @ ((shp - 1) / 2)
    f = prefilter_lines(np.pad(plane, NPAD, mode="edge").astype(np.float64))
    f = np.ascontiguousarray(prefilter_lines(np.ascontiguousarray(f.T)).T)
    HP, WP = f.shape
    ii, jj = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    cc0 = ((off[0] + ii * rot[0, 0]) + jj * rot[0, 1]) + NPAD
    cc1 = ((off[1] + ii * rot[1, 0]) + jj * rot[1, 1]) + NPAD
    s0, w0 = _weights(cc0)
    s1, w1 = _weights(cc1)
    t = np.zeros((H, W))
    for a in range(4):
        ia = np.clip(s0 + a, 0, HP - 1)
        for b in range(4):
            t = t + (f[ia, np.clip(s1 + b, 0, WP - 1)] * w0[a]) * w1[b]
    return t


def rotate(volume: np.ndarray, angle) -> np.ndarray:
    """== scipy.ndimage.rotate(volume, angle, axes=(0, 1), reshape=False, mode='nearest') for float32 / float64 / bool / uint8
    volumes (H, W, ...)."""
    v = np.asarray(volume)
    flat = v.reshape(v.shape[0], v.shape[1], -1)
    out = np.empty(flat.shape, dtype=v.dtype)
    for p in range(flat.shape[2]):
        t = rotate_plane_f64(flat[:, :, p], angle)
        if v.dtype == bool:
            out[:, :, p] = np.abs(t) >= 1.0                     # C cast double -> unsigned char truncates toward zero
        elif v.dtype == np.uint8:
            out[:, :, p] = np.where(t > 0, np.minimum(t + 0.5, 255.0), 0.0).astype(np.uint8)
        else:
            out[:, :, p] = t.astype(v.dtype)
    return out.reshape(v.shape)
