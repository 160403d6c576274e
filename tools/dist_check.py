"""Multi-GPU check (run under torchrun, NCCL): the point-cloud table assembled from slice-sharded extraction with one
variable-length all-gather is bit-identical on every rank to the 1-GPU table; DDP gradients match accumulation."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_deep_radiomics_b200 import ops, synth, tfds_dense_descriptor as tdd  # noqa: E402
from vit_deep_radiomics_b200.distributed import all_gather_table, init_distributed, shard_range  # noqa: E402

rank, world = init_distributed("nccl")
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
img, mask, res, name = synth.make_case("C1")            # ViT-S/16, 224x224x8
H, W, S = img.shape
model = tdd.load_model(name, img_hw=(H, W), device=dev, seed=7)
full = tdd.extract_point_cloud(model, img, mask, res, add_pe=False)      # every rank computes the 1-GPU table
# slice-sharded: this rank's contiguous slice range through the same backbone, gathered with the volume-level ROI
lo, hi = shard_range(S, rank, world)
plan = full["plan"]
img_dev = torch.as_tensor(img[:, :, lo:hi].copy()).to(dev)
tok = tdd._forward_volume(model, img_dev, plan)
mask_dev = torch.as_tensor(np.ascontiguousarray(mask[:, :, lo:hi]).view(np.uint8)).to(dev)
gh, gw = model.grid
t, src, cnt = ops.mask_gather(tok, mask_dev, grid=(hi - lo, gh, gw, model.n_tokens, 1), feat_roi=plan["feat_roi"],
                              mask_roi=tdd._shift_roi(plan["mask_roi"], plan["crop"]), mask_layout="hws")
n = int(cnt.item())
keys = src[:n].clone()
keys[:, 0] += lo                                          # global slice index
# canonical key order of the reference table is (row, col, slice): reorder columns for the lexicographic sort
k2 = torch.stack([keys[:, 1], keys[:, 2], keys[:, 0]], 1)
gk, gr = all_gather_table(k2, t[:n])
ref_k = torch.stack([full["src"][:, 1], full["src"][:, 2], full["src"][:, 0]], 1).to(dev)
ok_keys = torch.equal(gk, ref_k)
ok_rows = torch.equal(gr, full["tokens"].to(dev))
flag = torch.tensor([int(ok_keys and ok_rows)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"dist_check world={world}: table rows {gk.shape[0]} keys_equal={ok_keys} rows_bit_identical={ok_rows} all_ranks_ok={bool(flag.item())}")
dist.destroy_process_group()
