"""Cost of train-mode dropout per kernel (classifier shapes of config C5: n = 5052 tokens, d = 768, 12 heads, ff = 3072).
Usage: python tools/dropout_time.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_deep_radiomics_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for (N, h, ff) in ((5052, 12, 3072), (2300, 4, 1024)):
    d = h * 64
    qkv = (torch.randn(N, 3 * d, device=dev) * 0.5).bfloat16()
    drop = ops.Drop(1, 2, 0.1)
    out, lse = ops.flash_attn(qkv, 1, N, h, return_lse=True)
    do = torch.randn(N, d, device=dev).bfloat16()
    a = (torch.randn(N, d, device=dev) * 0.5).bfloat16()
    w = (torch.randn(d, d, device=dev) * 0.05).bfloat16()
    b = torch.randn(d, device=dev)
    z = torch.randn(N, ff, device=dev).bfloat16()
    rows = [
        ("attn fwd", lambda: ops.flash_attn(qkv, 1, N, h, return_lse=True), lambda: ops.flash_attn(qkv, 1, N, h, return_lse=True, drop=drop)),
        ("attn bwd", lambda: ops.flash_attn_bwd(qkv, out, do, lse, 1, N, h), lambda: ops.flash_attn_bwd(qkv, out, do, lse, 1, N, h, drop=drop)),
        ("gemm residual", lambda: ops.gemm(a, w, b, epilogue="residual", residual=a), lambda: ops.gemm(a, w, b, epilogue="residual", residual=a, drop=drop)),
        ("gelu fwd", lambda: ops.gelu(z), lambda: ops.gelu(z, drop=drop)),
        ("gelu bwd", lambda: ops.gelu_bwd(z, z), lambda: ops.gelu_bwd(z, z, drop=drop)),
        ("dropout_apply", None, lambda: ops.dropout_apply(a, drop)),
    ]
    print(f"N={N} heads={h} ff={ff}")
    for name, f0, f1 in rows:
        t0 = timeit(f0) if f0 else float("nan")
        print(f"  {name:14s} p=0 {t0:8.3f} ms   p=0.1 {timeit(f1):8.3f} ms")
