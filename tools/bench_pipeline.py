"""End-to-end pipeline throughput (BASELINE.json config C5): per patient, a synthetic 512x512x120 CT volume goes through
ViT-B/16 dense-descriptor extraction, the tumour-mask gather (+ 3-D positional encoding) and ONE training step of the
point-cloud transformer classifier (forward, focal loss, backward; AdamW every virtual batch of 32 patients, NCCL gradient
all-reduce across ranks).  The classifier takes the backbone's 768-wide descriptors directly (feature_dim: 768,
num_heads 12, mlp_ratio 4 in the YAML schema) -- the reference's 256 comes from MedSAM's neck, which plain ViT-B has not.

    python tools/bench_pipeline.py [--patients 8]
    torchrun --nproc-per-node N tools/bench_pipeline.py

Prints one JSON line (rank 0): patients/s and slices/s over all ranks (weak scaling: every rank its own patients), device
time between CUDA events, max over ranks; the host CPU does the same on a bounded sample through the oracle port."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_deep_radiomics_b200 import synth, tfds_dense_descriptor as tdd  # noqa: E402
from vit_deep_radiomics_b200.distributed import allreduce_grads, init_distributed  # noqa: E402
from vit_deep_radiomics_b200.models_archs import TransformerNoduleClassifier  # noqa: E402
from vit_deep_radiomics_b200.train_models import FocalLoss  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--patients", type=int, default=8, help="patients per rank in the timed region")
    ap.add_argument("--config", type=str, default="C2")
    ap.add_argument("--cpu-slices", type=int, default=8, dest="cpu_slices")
    ap.add_argument("--model", type=str, default=None, help="backbone override, e.g. medsam: the reference's default pipeline "
                    "(SAM ViT-B encoder on the crop window resized to 1024^2 -> 256-wide descriptors -> the shipped classifier config)")
    args = ap.parse_args()
    rank, world = init_distributed("nccl")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    img, mask, res, name = synth.make_case(args.config, seed=1238 + rank)
    H, W, S = img.shape
    if args.model:
        name = args.model
    backbone = tdd.load_model(name, img_hw=None if args.model else (H, W), device=dev, seed=1234)
    D = backbone.feature_dim
    torch.manual_seed(0)
    clf = TransformerNoduleClassifier(D, 4 * D, D // 64, 2, 2).to(dev)      # D = 256 (medsam): exactly conf/parameters_models.yaml
    if world > 1:
        for p in clf.parameters():
            torch.distributed.broadcast(p.data, 0)
    opt = torch.optim.AdamW(clf.parameters(), lr=5e-4, weight_decay=0.01)
    crit = FocalLoss(alpha=torch.tensor([0.25, 0.75], device=dev), gamma=2)
    img_pin = torch.as_tensor(img).pin_memory()
    mask_pin = torch.as_tensor(np.ascontiguousarray(mask).view(np.uint8)).pin_memory()
    ex = tdd.PointCloudExtractor(backbone)
    window = max(1, 32 // world)                         # the reference's virtual batch of 32, split over the ranks
    labels = [torch.eye(2, device=dev)[i % 2] for i in range(2)]

    def run(k):
        opt.zero_grad()
        last = None
        for i, out in enumerate(ex.run([(img_pin, mask_pin, res)] * k, to_host=False)):
            n = int(out["count"].item())                 # the point cloud's size is data-dependent: one 4-byte read-back
            logits, _ = clf(out["tokens"][:n].unsqueeze(0))
            loss = crit(torch.squeeze(logits), labels[i % 2]) / 32
            loss.backward()
            last = loss
            if (i + 1) % window == 0 or i + 1 == k:
                allreduce_grads(clf)
                opt.step()
                opt.zero_grad()
        return n, float(last.detach())

    run(2)                                               # warm-up (allocations, first launches)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n_tok, loss = run(args.patients)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
        torch.distributed.barrier()
    if rank != 0:
        return
    if args.model:       # no host-CPU leg for the override (the fp32 SAM oracle needs minutes per 1024^2 slice batch)
        print(json.dumps({
            "metric": f"patients/sec end-to-end: {name} extraction + mask gather + classifier fwd/bwd", "value": world * args.patients / (ms / 1e3),
            "unit": "patients/s", "slices_per_s": world * args.patients * S / (ms / 1e3), "n_gpus": world, "patients_per_rank": args.patients,
            "ms_per_patient": ms / args.patients, "tokens_per_patient": n_tok, "scaling": "weak", "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{name} ({backbone.img_hw[0]}x{backbone.img_hw[1]} encoder input, {backbone.grid[0]}x{backbone.grid[1]}x{D} descriptors) over a "
                                   f"{H}x{W}x{S} volume -> point cloud -> 2-layer transformer classifier (d {D}, {D // 64} heads) training step per "
                                   "patient; H2D of every volume inside the timed region"}, "loss": loss}))
        return
    # CPU baseline: the oracle port (fp32 ViT + NumPy gather + fp32 classifier fwd/bwd) on a bounded sample of slices
    from oracle import classifier_fp32 as C, gather_np, vit_fp32
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = vit_fp32.VIT_CONFIGS[name]
    w = vit_fp32.init_weights(cfg, (H, W), seed=1234)
    s0 = S // 2 - args.cpu_slices // 2
    x = torch.from_numpy(np.ascontiguousarray(np.moveaxis(img[:, :, s0:s0 + args.cpu_slices], -1, 0)))[:, None].expand(-1, 3, -1, -1).contiguous()
    t0 = time.perf_counter()
    with torch.no_grad():
        dense = vit_fp32.vit_forward(w, cfg, x).numpy()
    o = gather_np.token_gather([dense[i] for i in range(dense.shape[0])], [mask[:, :, s0 + i] for i in range(dense.shape[0])], res)
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in clf.state_dict().items()}
    lg, _ = C.classifier_forward(sd, torch.from_numpy(o["tokens"].astype(np.float32))[None], D // 64, 2)
    C.focal_loss(lg[0], torch.eye(2)[0], 2.0, torch.tensor([0.25, 0.75])).backward()
    cpu_dt = time.perf_counter() - t0
    print(json.dumps({
        "metric": "patients/sec end-to-end: ViT-B/16 extraction + mask gather + classifier fwd/bwd", "value": world * args.patients / (ms / 1e3),
        "unit": "patients/s", "slices_per_s": world * args.patients * S / (ms / 1e3), "n_gpus": world, "patients_per_rank": args.patients,
        "ms_per_patient": ms / args.patients, "tokens_per_patient": n_tok, "scaling": "weak", "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"C5: {name} extraction over a {H}x{W}x{S} volume -> point cloud -> 2-layer transformer classifier "
                               f"(d {D}, {D // 64} heads) training step per patient; H2D of every volume inside the timed region"},
        "loss": loss,
        "cpu_baseline": {"value": (args.cpu_slices / S) / cpu_dt, "unit": "patients/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": f"{args.cpu_slices} of {S} slices through the fp32 oracle ViT + NumPy gather + oracle classifier fwd/bwd "
                                   f"({o['flat'].size} tokens), {cpu_dt:.2f} s wall, scaled by slices"}}))


if __name__ == "__main__":
    main()
