"""Debug: run-to-run determinism of the in-place residual GEMM with row statistics (pair and single-CTA configurations)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_deep_radiomics_b200 import ops
dev = torch.device("cuda:0")
REPS = int(os.environ.get("REPS", "100"))
for (M, N, K) in [(40000, 768, 768), (123000, 768, 768), (123000, 768, 3072), (2050, 768, 3072)]:
    torch.manual_seed(1)
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16(); w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    b = torch.randn(N, device=dev) * 0.1; x0 = torch.randn(M, N, device=dev).bfloat16()
    ref = None
    bad_x = bad_s = nan_s = 0
    for rep in range(REPS):
        x = x0.clone(); st = torch.full((N // 64, M, 2), float("nan"), device=dev)
        ops.gemm(a, w, b, epilogue="residual", residual=x, out=x, stats_out=st)
        torch.cuda.synchronize()
        if ref is None:
            ref = (x, st)
            nan_s = int(torch.isnan(st).sum())
            continue
        if not torch.equal(x, ref[0]):
            bad_x += 1
            if bad_x == 1:
                idx = (x != ref[0]).nonzero()
                print("  x differs at", idx.shape[0], "elements; first", idx[:4].tolist(), "last", idx[-2:].tolist())
        if not torch.equal(st, ref[1]):
            bad_s += 1
            if bad_s == 1:
                idx = (st != ref[1]).nonzero()
                print("  st differs at", idx.shape[0], "entries; first", idx[:4].tolist(), "nan", int(torch.isnan(st).sum()))
    print(f"M{M} N{N} K{K}: x mismatches {bad_x}/{REPS - 1}, stats mismatches {bad_s}/{REPS - 1}, NaN left in first stats {nan_s}")
