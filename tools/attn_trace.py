"""Debug: per-kv-block timeline of one CTA of the attention kernel (ns)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_deep_radiomics_b200 import _C, ops
dev = torch.device("cuda:0")
L = _C.lib()
L.vdr_debug_set_attn_trace.argtypes = [ctypes.c_void_p]
B, N, h = 120, int(sys.argv[1]) if len(sys.argv) > 1 else 1025, 12
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
t0 = int(t[15, 0]) if t[15, 0] else int(t[0, 0])
if os.environ.get("VDR_TRACE_CLOCK"):   # library built with -DVDR_ATTN_TRACE_CLOCK: SM cycles instead of ns
    print("cycles since the first stamp | softmax: loop_top s_ready s_loaded(+sfree) math_done o_wait_done p_stored(+pready) | S issuer: top sfree_seen s_next_issued | PV issuer: top pready_seen pv_issued")
    for j in range(9):
        print(j, " ".join(f"{int(v) - t0:7d}" for v in t[j, :6]), "|", " ".join(f"{int(v) - t0:7d}" for v in t[j, 8:14]))
    sys.exit(0)
print("softmax: loop_top s_ready s_loaded math_done o_wait_done p_stored | issuer: top sfree_seen s_next_issued before_pready pready_seen pv_issued (us)")
print("v6 columns: softmax loop_top s_ready s_loaded math_done p_stored - | issuer: top k_ready o_ready s_issued p_seen v_ready")
print("CTA (row 15): entry setup_done last_P_stored last_PV_done rows_stored cta_end:", " ".join(f"{(int(v) - t0) / 1e3:7.2f}" for v in t[15, :6]))
for j in range(9):
    print(j, " ".join(f"{(int(v) - t0) / 1e3:7.2f}" for v in t[j, :6]), "|", " ".join(f"{(int(v) - t0) / 1e3:7.2f}" for v in t[j, 8:14]))
