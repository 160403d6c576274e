"""Debug: per-tile timeline of CTA 0 of the GEMM kernel (producer / MMA / epilogue timestamps, ns).
Needs a trace build: `make -C vit_deep_radiomics_b200/csrc EXTRA=-DVDR_GEMM_TRACE` (the stamps are compiled out otherwise)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_deep_radiomics_b200 import _C, ops
dev = torch.device("cuda:0")
L = _C.lib()
L.vdr_debug_set_gemm_trace.argtypes = [ctypes.c_void_p]
for (M, N, K, epi) in [(123000, 2304, 768, "bias"), (123000, 3072, 768, "gelu"), (123000, 768, 768, "residual")]:
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16(); w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    b = torch.randn(N, device=dev); r = torch.randn(M, N, device=dev).bfloat16() if epi == "residual" else None
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.gemm(a, w, b, epilogue=epi, residual=r, out=out)
    buf = torch.zeros(64 * 16, dtype=torch.int64, device=dev)
    L.vdr_debug_set_gemm_trace(buf.data_ptr())
    ops.gemm(a, w, b, epilogue=epi, residual=r, out=out)
    torch.cuda.synchronize()
    L.vdr_debug_set_gemm_trace(None)
    t = buf.cpu().view(64, 16).numpy()
    t0 = t[0, 5]
    print(f"== M={M} N={N} K={K} {epi}: columns = mma_free mma_first_stage mma_last_issue epi_ready epi_done tma_first tma_last (us since first TMA)")
    for i in range(12):
        print(i, " ".join(f"{(int(v) - int(t0)) / 1e3:8.2f}" for v in t[i, :7]), "| epi warp0 c0: ld_done math_done buf_free stored; c1: ...", " ".join(f"{(int(v) - int(t[i, 3])) / 1e3:6.2f}" for v in t[i, 8:16]))
