"""Phase timing of the fused gather kernel (vdr_debug_set_gather_trace): %globaltimer stamps of the first / last block.
Usage: python tools/gather_trace.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_deep_radiomics_b200 import _C, ops  # noqa: E402

dev = torch.device("cuda:0")
S, gh, gw, D = 120, 32, 32, 768
tok = torch.randn(S * (gh * gw + 1), D, device=dev)
rng = np.random.default_rng(0)
cases = {"dense": (torch.ones(512, 512, S, dtype=torch.uint8, device=dev), None, None),
         "c2-like roi": (torch.from_numpy((rng.random((512, 512, S)) < 0.35).astype(np.uint8)).to(dev), (10, 22, 11, 21), (160, 352, 176, 336))}
trace = torch.zeros(16, dtype=torch.int64, device=dev)
flush = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
names = ["table", "count", "sync1", "ranks", "sync2", "emit"]
for name, (mask, froi, mroi) in cases.items():
    for pe in (dict(res=(0.8, 0.8, 0.8)), None):
        call = lambda: ops.mask_gather(tok, mask, grid=(S, gh, gw, gh * gw + 1, 1), feat_roi=froi, mask_roi=mroi, mask_layout="hws", pe=pe)  # noqa: E731
        for _ in range(3):
            call()
        _C.lib().vdr_debug_set_gather_trace(trace.data_ptr())
        rows = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = call()
            e1.record()
            torch.cuda.synchronize()
            t = trace.cpu().numpy()
            rows.append((e0.elapsed_time(e1) * 1e3, np.diff(t[0:7]) / 1e3, np.diff(t[8:15]) / 1e3, (t[14] - t[0]) / 1e3))
        _C.lib().vdr_debug_set_gather_trace(None)
        ev = np.median([r[0] for r in rows])
        first = np.median([r[1] for r in rows], axis=0)
        last = np.median([r[2] for r in rows], axis=0)
        print(f"{name:12s} pe={'on ' if pe else 'off'} n={int(out[2].item()):6d} events {ev:7.1f} us | kernel span {np.median([r[3] for r in rows]):7.1f} us")
        print("   block 0   : " + "  ".join(f"{n} {v:6.1f}" for n, v in zip(names, first)))
        print("   last block: " + "  ".join(f"{n} {v:6.1f}" for n, v in zip(names, last)))
