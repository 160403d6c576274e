import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_deep_radiomics_b200 import ops
dev = torch.device("cuda:0")
for (M, N, K) in [(123000, 2304, 768), (123000, 768, 3072)]:
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16(); w = (torch.randn(N, K, device=dev) * 0.05).bfloat16(); b = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3): ops.gemm(a, w, b, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.gemm(a, w, b, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"dbg={os.environ.get('VDR_GEMM_DBG','0')} 1cta={os.environ.get('VDR_GEMM_1CTA','-')} M={M} N={N} K={K}: {ms:.3f} ms {2*M*N*K/ms/1e9:.0f} TF", flush=True)
