"""Per-op device time of one classifier training step (fwd + bwd) with and without train-mode dropout.
Usage: python tools/cls_dropout_profile.py [n] [d]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_deep_radiomics_b200 import models_archs, ops  # noqa: E402
from vit_deep_radiomics_b200.train_models import FocalLoss  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5051
d = int(sys.argv[2]) if len(sys.argv) > 2 else 768
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.randn(1, n, d, device=dev)
y = torch.eye(2, device=dev)[1]
crit = FocalLoss(alpha=torch.tensor([0.25, 0.75], device=dev), gamma=2)
for p in (0.0, 0.1):
    model = models_archs.set_dropout(models_archs.TransformerNoduleClassifier(d, 4 * d, d // 64, 2, 2).to(dev), p, p).train()

    def step():
        model.zero_grad()
        logits, _ = model(x)
        crit(torch.squeeze(logits), y).backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        step()
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 5 * 1e3
    ops.PROFILE = []
    step()
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    agg = {}
    for kind, work, a, b, label in prof:
        t = agg.setdefault(label.split(" M")[0] if kind == "gemm" else label, [0, 0.0])
        t[0] += 1
        t[1] += a.elapsed_time(b)
    print(f"p={p}: {e0.elapsed_time(e1) / 5:.2f} ms per step on the device, {wall:.2f} ms wall; profiled ops:")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:8]:
        print(f"    {k:40s} x{v[0]:3d} {v[1]:8.3f} ms")
