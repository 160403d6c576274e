"""A/B timing of the attention forward (B 120, 12 heads): VDR_ATTN_DBG / VDR_ATTN_V5 / VDR_ATTN_V6 select the variant."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_deep_radiomics_b200 import ops
dev = torch.device("cuda:0")
B, h = 120, 12
Ns = [int(a) for a in sys.argv[1:]] or [1024, 1025]
for N in Ns:
    qkv = torch.randn(B * N, 3 * h * 64, device=dev).bfloat16()
    out = torch.empty(B * N, h * 64, device=dev, dtype=torch.bfloat16)
    for _ in range(3): ops.flash_attn(qkv, B, N, h, out=out)
    ts = []
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): ops.flash_attn(qkv, B, N, h, out=out)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 20)
    tag = " ".join(f"{k[9:]}={v}" for k, v in os.environ.items() if k.startswith("VDR_ATTN_"))
    print(f"[{tag}] N={N}: " + " ".join(f"{t:.3f}" for t in ts), flush=True)
