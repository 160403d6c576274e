import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_deep_radiomics_b200 import ops
cuda = torch.device("cuda:0")
B, H, W, d, p = int(os.environ.get("PE_B", "3")), 512, 512, 768, 16
imgs = torch.rand((B, H, W), device=cuda).bfloat16()
gh, gw = H // p, W // p
Np, N, K = gh * gw, gh * gw + 1, 3 * p * p
w_pe = (torch.randn(d, K, device=cuda) * 0.05).bfloat16()
bias, pos = torch.randn(d, device=cuda), torch.randn(N, d, device=cuda)
x = torch.full((B * N, d), 7.0, device=cuda, dtype=torch.bfloat16)
ops.patch_embed(imgs, w_pe, bias, pos, p, out=x)
torch.cuda.synchronize()
src = imgs.float()
A = torch.empty(B * Np, K, device=cuda, dtype=torch.bfloat16)
ops.im2col_patches(src, (src.stride(0), 0, src.stride(1), src.stride(2)), B, H, W, p, out=A)
x_ref = torch.full((B * N, d), 7.0, device=cuda, dtype=torch.bfloat16)
ops.gemm(A, w_pe, bias, epilogue="residual", residual=pos, out=x_ref, k=K, out_group=(Np, N, 1), res_mod=(Np, 1))
torch.cuda.synchronize()
print("equal:", torch.equal(x, x_ref), "max diff", (x.float() - x_ref.float()).abs().max().item())
