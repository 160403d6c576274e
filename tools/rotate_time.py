"""Device time of the 24 whole-volume transforms of one patient's augmentation grid (image + mask per (flip, angle)), C2 volume."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from vit_deep_radiomics_b200 import ops, synth, tfds_dense_descriptor as tdd

img, mask, res, name = synth.make_case("C2", seed=1240)
img_d = torch.as_tensor(img).cuda()
mask_d = torch.as_tensor(np.ascontiguousarray(mask).view(np.uint8)).cuda()
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for flip in tdd.AUG_FLIPS:
        for angle in tdd.AUG_ANGLES:
            ops.flip_rotate_volume(img_d, flip, angle, kind="image")
            ops.flip_rotate_volume(mask_d, flip, angle, kind="mask_bool")
    e1.record()
    torch.cuda.synchronize()
    print(f"rep {rep}: {e0.elapsed_time(e1):.2f} ms for 24 transforms (18 rotations)", flush=True)
