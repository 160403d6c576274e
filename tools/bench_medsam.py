"""MedSAM image-encoder throughput: B gray 1024 x 1024 slices resident on the device ->
(B, 64, 64, 256) descriptors.  Prints slices/s, model TFLOP/s and the per-kernel-kind split of one profiled pass
(CUDA events around every launch, ops.PROFILE).  Usage: python tools/bench_medsam.py [B] [steps]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_deep_radiomics_b200 import _C, ops, tfds_dense_descriptor as tdd  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    dev = torch.device("cuda:0")
    model = tdd.load_model("medsam", device=dev, seed=1)
    x = torch.rand(B, 1024, 1024, device=dev)
    strides = (x.stride(0), 0, x.stride(1), x.stride(2))
    for _ in range(3):
        model.forward_tokens(x, strides, B)
    torch.cuda.synchronize()
    n0 = _C.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        model.forward_tokens(x, strides, B)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    launches = (_C.launch_count() - n0) // steps
    ops.PROFILE = []
    model.forward_tokens(x, strides, B)
    torch.cuda.synchronize()
    agg = {}
    for kind, work, a, b, label in ops.PROFILE:
        key = label.split(" ")[0] + (" windows" if "windows" in label else " global" if "N4096" in label or "64x64" in label else "")
        t, w = agg.get(key, (0.0, 0.0))
        agg[key] = (t + a.elapsed_time(b), w + work)
    ops.PROFILE = None
    split = {k: dict(ms=round(t, 3), tflops=round(w / t / 1e9, 1)) for k, (t, w) in sorted(agg.items(), key=lambda kv: -kv[1][0])}
    print(json.dumps(dict(metric="MedSAM encoder slices/s (1024x1024, resident)", value=round(B / ms * 1e3, 1), batch=B,
                          ms_per_batch=round(ms, 3), model_tflops=round(model.flops_per_slice() * B / ms / 1e9, 1),
                          launches_per_batch=launches, split=split)))


if __name__ == "__main__":
    main()
