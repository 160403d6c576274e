"""The NumPy restatement of scipy's cubic rotation (oracle/rotate_np.py: the algorithm csrc/augment.cu implements) against
scipy.ndimage.rotate itself -- bit for bit.  scipy is what the reference's rotate_image calls (tfds_dense_descriptor.py:328-350)."""
import numpy as np
import pytest
from scipy import ndimage

from oracle import rotate_np


@pytest.mark.parametrize("shape", [(40, 40, 2), (37, 53, 3), (64, 48, 1)])
@pytest.mark.parametrize("angle", [45, 90, 135])
def test_rotate_restatement_is_bit_identical_to_scipy(shape, angle):
    rng = np.random.default_rng(shape[1] + angle)
    img32 = rng.random(shape).astype(np.float32)
    img64 = rng.random(shape)
    blob = np.zeros(shape, bool)
    blob[shape[0] // 4: shape[0] // 2 + 5, shape[1] // 3: shape[1] // 3 + 11] = True
    noise = rng.random(shape) < 0.4
    kw = dict(axes=(0, 1), reshape=False, mode="nearest")
    for x in (img32, img64, blob, noise, blob.astype(np.uint8), (noise * 255).astype(np.uint8)):
        want = ndimage.rotate(x, angle, **kw)
        got = rotate_np.rotate(x, angle)
        assert got.dtype == want.dtype and np.array_equal(got, want), (x.dtype, shape, angle)


def test_prefilter_restatement_is_bit_identical_to_scipy():
    rng = np.random.default_rng(0)
    for n in (24, 25, 88, 536):
        x = rng.random((n, 5))
        assert np.array_equal(rotate_np.prefilter_lines(x), ndimage.spline_filter1d(x, 3, axis=0, output=np.float64, mode="nearest"))


def test_reference_bool_mask_is_the_truncated_interpolant():
    """What the reference's ``rotate(mask, ...) > 0`` really keeps for a bool mask: voxels whose cubic interpolant is >= 1 --
    at 90 degrees that is decided by the last bit of a sum that is 1 up to rounding."""
    m = np.zeros((64, 64, 1), bool)
    m[20:40, 25:45] = True
    r = ndimage.rotate(m, 90, axes=(0, 1), reshape=False, mode="nearest")
    t = ndimage.rotate(m.astype(np.float64), 90, axes=(0, 1), reshape=False, mode="nearest")
    assert np.array_equal(r, np.abs(t) >= 1.0) and 0 < r.sum() < m.sum()
