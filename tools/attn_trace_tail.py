"""Debug: timeline of one trailing-row CTA of the attention kernel (library built with -DVDR_ATTN_TRACE; us since CTA entry)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_deep_radiomics_b200 import _C, ops
dev = torch.device("cuda:0")
L = _C.lib()
L.vdr_debug_set_attn_trace.argtypes = [ctypes.c_void_p]
B, N, h = 120, 1025, 12
qkv = torch.randn(B * N, 3 * h * 64, device=dev).bfloat16()
out = torch.empty(B * N, h * 64, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    ops.flash_attn(qkv, B, N, h, out=out)
buf = torch.zeros(2 * 16 * 16, dtype=torch.int64, device=dev)      # [0, 256): the tile CTA's stamps, [256, 512): the trailing-row CTA's
L.vdr_debug_set_attn_trace(buf.data_ptr())
ops.flash_attn(qkv, B, N, h, out=out)
torch.cuda.synchronize()
L.vdr_debug_set_attn_trace(None)
t = buf.cpu()[256:].view(16, 16).numpy()
t0 = int(t[15, 0])
f = lambda v: f"{(int(v) - t0) / 1e3:7.2f}" if v else "      -"
print("CTA: entry barriers_ready loop_done(warp 0) all_warps_done merged:", " ".join(f(v) for v in t[15, :5]))
print("block | warp 0: loop_top stage_landed S+softmax_done PV_done+released | producer: issued")
for j in range(9):
    print(j, " ".join(f(v) for v in t[j, :4]), "|", f(t[j, 8]))
