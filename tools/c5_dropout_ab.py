"""A/B of the C5 pipeline record (bench.bench_pipeline) with the classifier's train-mode dropout off / on, twice each, in one process."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from vit_deep_radiomics_b200 import models_archs  # noqa: E402

D = bench.Dist()
orig = models_archs.TransformerNoduleClassifier.__init__
for tag, p in (("off", 0.0), ("on", None), ("off", 0.0), ("on", None)):
    def patched(self, *a, _p=p, **k):
        orig(self, *a, **k)
        if _p is not None:
            models_archs.set_dropout(self, _p, _p)
    models_archs.TransformerNoduleClassifier.__init__ = patched
    r = bench.bench_pipeline(D, patients=4, cpu_slices=1)
    print(f"dropout {tag}: {r['value']:.2f} patients/s, {r['ms_per_patient']:.2f} ms per patient, {r['gpu_launches']} launches", flush=True)
