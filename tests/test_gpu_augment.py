"""-m gpu: the offline augmentation on the device (csrc/augment.cu) against the host functions the
reference runs (flip_image + rotate_image = scipy.ndimage.rotate, tfds_dense_descriptor.py:306-350): bit-identical masks AND
bit-identical float32 images, for every (flip, angle) of the reference's grid (:463-466)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _host(img, mask, flip, angle):
    from vit_deep_radiomics_b200 import tfds_dense_descriptor as tdd
    i, m = tdd.flip_image(img, mask, flip)
    return tdd.rotate_image(i, m, angle)


@pytest.mark.parametrize("shape", [(40, 40, 3), (37, 53, 4), (96, 64, 5)])
def test_flip_rotate_volume_bit_identical_to_host(cuda, shape):
    from vit_deep_radiomics_b200 import ops, tfds_dense_descriptor as tdd
    rng = np.random.default_rng(shape[0])
    img = rng.random(shape).astype(np.float32)
    mask = np.zeros(shape, bool)
    mask[shape[0] // 4: shape[0] // 2 + 7, shape[1] // 3: shape[1] // 3 + 13, 1:] = True
    mask |= rng.random(shape) < 0.02
    img_d = torch.from_numpy(img).to(cuda)
    mask_d = torch.from_numpy(mask.view(np.uint8)).to(cuda)
    for flip in tdd.AUG_FLIPS:
        for angle in tdd.AUG_ANGLES:
            want_i, want_m = _host(img, mask, flip, angle)
            got_i = ops.flip_rotate_volume(img_d, flip, angle, kind="image").cpu().numpy()
            got_m = ops.flip_rotate_volume(mask_d, flip, angle, kind="mask_bool").cpu().numpy().astype(bool)
            assert np.array_equal(got_m, want_m), (flip, angle, int((got_m != want_m).sum()))
            assert got_i.dtype == np.float32 and np.array_equal(got_i, want_i.astype(np.float32)), (flip, angle, float(np.abs(got_i - want_i).max()))


def test_rotate_uint8_mask_and_rgb_planes(cuda):
    """uint8 masks go through scipy's rounding output conversion (not the bool truncation); an (H, W, S, 3) colour volume (the
    dinov2 pre-processing, hu_to_rgb_vectorized) rotates plane by plane."""
    from scipy import ndimage
    from vit_deep_radiomics_b200 import ops
    rng = np.random.default_rng(3)
    m8 = (rng.random((50, 44, 3)) < 0.3).astype(np.uint8)
    want = ndimage.rotate(m8, 45, axes=(0, 1), reshape=False, mode="nearest") > 0
    got = ops.flip_rotate_volume(torch.from_numpy(m8).to(cuda), None, 45, kind="mask_u8").cpu().numpy().astype(bool)
    assert np.array_equal(got, want)
    rgb = rng.random((33, 41, 2, 3)).astype(np.float32)
    want = np.clip(ndimage.rotate(rgb, 135, axes=(0, 1), reshape=False, mode="nearest"), 0, 1)
    got = ops.flip_rotate_volume(torch.from_numpy(rgb).to(cuda), None, 135, kind="image").cpu().numpy()
    assert np.array_equal(got, want)


def test_extract_patient_features_device_augmentation_equals_host_loop(cuda):
    """extract_patient_features with the augmentation on the device == the reference's loop with host flips / scipy rotations:
    identical metadata table, identical masks (bit-exact) and identical descriptors (same kernels on bit-identical inputs)."""
    from vit_deep_radiomics_b200 import synth, tfds_dense_descriptor as tdd
    img, mask, res, name = synth.make_case("T0")
    model = tdd.load_model(name, img_hw=img.shape[:2], device=cuda, seed=7)
    df_h, f_h, m_h = tdd.extract_patient_features(model, img, mask, "p0", 1, "stanford_dataset", "ct", res, device_augment=False)
    df_d, f_d, m_d = tdd.extract_patient_features(model, img, mask, "p0", 1, "stanford_dataset", "ct", res, device_augment=True)
    assert df_h.drop(columns=["spatial_res"]).equals(df_d.drop(columns=["spatial_res"])) and len(f_h) == len(f_d) == 12 * img.shape[2]
    for a, b in zip(m_h, m_d):
        assert a.shape == b.shape and np.array_equal(a, b)
    for a, b in zip(f_h, f_d):
        assert a.shape == b.shape and np.array_equal(a, b)


def test_pipelined_augmentation_equals_copy_by_copy_device_loop(cuda):
    """The pipelined device loop (masks + plans first, ROI read-backs on a copy stream behind the next backbone) returns what the
    copy-by-copy device calls return, and its arrays are caller-owned: a second patient does not overwrite the first one's."""
    from vit_deep_radiomics_b200 import ops, synth, tfds_dense_descriptor as tdd
    img, mask, res, name = synth.make_case("C1")
    model = tdd.load_model(name, img_hw=img.shape[:2], device=cuda, seed=11)
    img_d = torch.from_numpy(np.ascontiguousarray(img, dtype=np.float32)).to(cuda)
    mask_d = torch.from_numpy(np.ascontiguousarray(mask).view(np.uint8)).to(cuda)
    grid = [(f, a) for f in tdd.AUG_FLIPS for a in tdd.AUG_ANGLES]
    got = tdd._augmented_features_device(model, img_d, mask_d, "mask_bool", grid)
    keep = [(np.array(f[0]), np.array(m[0])) for f, m in got]
    again = tdd._augmented_features_device(model, torch.flip(img_d, dims=(2,)).contiguous(), mask_d, "mask_bool", grid)
    assert len(got) == len(again) == 12
    for k, (flip, angle) in enumerate(grid):
        i = ops.flip_rotate_volume(img_d, flip, angle, kind="image")
        m = ops.flip_rotate_volume(mask_d, flip, angle, kind="mask_bool")
        want_f, want_m = tdd.generate_features_device(model, i, m)
        assert len(want_f) == len(got[k][0]) == img.shape[2]
        for a, b in zip(want_f, got[k][0]):
            assert a.shape == b.shape and np.array_equal(a, b)
        for a, b in zip(want_m, got[k][1]):
            assert a.dtype == b.dtype == np.bool_ and np.array_equal(a, b)
        assert np.array_equal(keep[k][0], got[k][0][0]) and np.array_equal(keep[k][1], got[k][1][0])
