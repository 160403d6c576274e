import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_deep_radiomics_b200 import ops
dev = torch.device("cuda:0")
B, h = 120, 12
for N in (1024, 1025, 1040, 1152):
    qkv = torch.randn(B * N, 3 * h * 64, device=dev).bfloat16()
    out = torch.empty(B * N, h * 64, device=dev, dtype=torch.bfloat16)
    for _ in range(3): ops.flash_attn(qkv, B, N, h, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.flash_attn(qkv, B, N, h, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"N={N}: {ms:.3f} ms  {4.0*B*h*N*N*64/ms/1e9:.0f} TF", flush=True)
