"""Wall time of one patient of the offline augmentation loop (extract_patient_features, device path) on the C2 volume."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from vit_deep_radiomics_b200 import synth, tfds_dense_descriptor as tdd

img, mask, res, name = synth.make_case("C2", seed=1240)
model = tdd.load_model(name, img_hw=img.shape[:2], device="cuda:0", seed=1234)
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    df, feats, masks = tdd.extract_patient_features(model, img, mask, "p0", 1, "synthetic_dataset", "CT", res, device_augment=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"rep {rep}: {dt:.3f} s per patient ({len(feats)} maps, {sum(f.nbytes for f in feats) / 1e6:.0f} MB of descriptors)", flush=True)
