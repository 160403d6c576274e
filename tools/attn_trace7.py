"""Debug: timeline of the third and fourth tiles (global key blocks 16..31) of CTA 100 of the persistent attention kernel (us).
Needs a library built with EXTRA=-DVDR_ATTN_TRACE."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_deep_radiomics_b200 import _C, ops
dev = torch.device("cuda:0")
L = _C.lib()
L.vdr_debug_set_attn_trace.argtypes = [ctypes.c_void_p]
B, N, h = 120, int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 12
qkv = torch.randn(B * N, 3 * h * 64, device=dev).bfloat16()
out = torch.empty(B * N, h * 64, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    ops.flash_attn(qkv, B, N, h, out=out)
buf = torch.zeros(16 * 16, dtype=torch.int64, device=dev)
L.vdr_debug_set_attn_trace(buf.data_ptr())
ops.flash_attn(qkv, B, N, h, out=out)
torch.cuda.synchronize()
L.vdr_debug_set_attn_trace(None)
t = buf.cpu().view(16, 16).numpy()
t0 = int(t[0, 0])
print("g | softmax: top s_ready s_loaded math_done o_wait_done p_stored [tile end: pv_done epilogue_done] | S-issuer: top sfree_seen s_issued | PV-issuer: top pready_seen pv_issued")
for j in range(16):
    f = lambda v: f"{(int(v) - t0) / 1e3:7.2f}" if v else "      -"
    print(f"{j + 16:2d}", " ".join(f(v) for v in t[j, :8]), "|", " ".join(f(v) for v in t[j, 8:11]), "|", " ".join(f(v) for v in t[j, 11:14]), "| v_landed", f(t[j, 14]), "k_next_landed", f(t[j, 15]))
