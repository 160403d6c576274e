"""Time the four GEMM shapes of a ViT-B/16 layer at config C2 (M = 120 x 1025 tokens), each with its epilogue."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_deep_radiomics_b200 import ops
dev = torch.device("cuda:0")
M = 123000
for (N, K, epi) in [(2304, 768, "bias"), (768, 768, "residual"), (3072, 768, "gelu"), (768, 3072, "residual")]:
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16(); w = (torch.randn(N, K, device=dev) * 0.05).bfloat16(); b = torch.randn(N, device=dev)
    r = torch.randn(M, N, device=dev).bfloat16() if epi == "residual" else None
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3): ops.gemm(a, w, b, epilogue=epi, residual=r, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.gemm(a, w, b, epilogue=epi, residual=r, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"dbg={os.environ.get('VDR_GEMM_DBG','0')} 1cta={os.environ.get('VDR_GEMM_1CTA','-')} M={M} N={N} K={K} {epi}: {ms:.3f} ms {2*M*N*K/ms/1e9:.0f} TF", flush=True)
