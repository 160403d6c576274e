"""Where a classifier training step goes: launches per sample, GPU time vs wall time with / without the per-sample loss read-back."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_deep_radiomics_b200 import _C, synth
from vit_deep_radiomics_b200.models_archs import TransformerNoduleClassifier
from vit_deep_radiomics_b200.train_models import FocalLoss
dev = torch.device("cuda:0")
D = int(os.environ.get("CLS_D", "256"))
NR = (int(os.environ.get("CLS_NMIN", "512")), int(os.environ.get("CLS_NMAX", "4096")))
ids, labels, sizes, cloud = synth.point_cloud_patients(16, d=D, n_range=NR, seed=1236)
torch.manual_seed(0)
model = TransformerNoduleClassifier(D, 4 * D, D // 64, 2, 2).to(dev)
crit = FocalLoss(alpha=torch.tensor([0.25, 0.75], device=dev), gamma=2)
data = [(torch.from_numpy(cloud(i)).to(dev), torch.eye(2, device=dev)[int(labels[i])]) for i in range(16)]
def run(sync):
    for x, y in data:
        logits, _ = model(x.unsqueeze(0))
        loss = crit(torch.squeeze(logits), y) / 32
        loss.backward()
        if sync:
            float(loss.detach())
for sync in (True, False):
    run(sync); torch.cuda.synchronize()
    l0 = _C.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(3): run(sync)
    e1.record(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    n = 3 * len(data)
    print(f"sync={sync}: wall {1e3*dt/n:.2f} ms/sample, gpu {e0.elapsed_time(e1)/n:.2f} ms/sample, libvdr launches/sample {(_C.launch_count()-l0)/n:.0f}, mean tokens {sum(sizes)/len(sizes):.0f}")
# forward only
with torch.no_grad():
    for x, y in data: model(x.unsqueeze(0))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        for x, y in data: model(x.unsqueeze(0))
    torch.cuda.synchronize(); print(f"forward only: {1e3*(time.perf_counter()-t0)/48:.2f} ms/sample")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run(False); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=60))
