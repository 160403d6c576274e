"""Bimodal training step (TransformerNoduleBimodalClassifier as conf/parameters_models.yaml builds it for 'petct'): eager op-by-op vs one
CUDA graph per (CT, PET) pair of token counts.  samples/s over 3 epochs of 32 synthetic patients (after 2 warm-up epochs)."""
import sys, time
import torch
sys.path.insert(0, ".")
from vit_deep_radiomics_b200.config_manager import load_conf
from vit_deep_radiomics_b200 import train_models as tm

dev = torch.device("cuda:0")
import os
cfg = load_conf(project_dir=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
gen = torch.Generator().manual_seed(3)
d = cfg["models"]["transformer"]["feature_dim"]
data = [(torch.randn(int(a), d, generator=gen).to(dev), torch.randn(int(b), d, generator=gen).to(dev), torch.eye(2)[i % 2].to(dev))
        for i, (a, b) in enumerate(zip(torch.randint(512, 4096, (32,), generator=gen), torch.randint(64, 1024, (32,), generator=gen)))]
crit = tm.make_criterion("crossmodal", dev)
for graphs in (False, True):
    torch.manual_seed(0)
    model = tm.build_model(cfg, "transformer", "petct").to(dev)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
    for _ in range(2):
        tm.train_epoch(model, data, crit, opt, virtual_batch_size=32, cuda_graphs=graphs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        tm.train_epoch(model, data, crit, opt, virtual_batch_size=32, cuda_graphs=graphs)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"cuda_graphs={graphs}: {3 * len(data) / dt:.0f} samples/s ({dt / (3 * len(data)) * 1e3:.2f} ms per sample)", flush=True)
