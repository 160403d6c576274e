"""Classifier training throughput: point-cloud transformer fwd + bwd + AdamW with the
reference's accumulation rule, data-parallel over the ranks of one box (NCCL gradient all-reduce).

    python tools/bench_classifier.py [--samples 64] [--cpu-samples 4]
    torchrun --nproc-per-node N tools/bench_classifier.py

Prints one JSON line (rank 0): samples/s on the GPU(s), the unmodified-architecture fp32 oracle on the host CPU
(same clouds, bounded sample), and the loss trajectory check (the loss must fall)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_deep_radiomics_b200 import synth  # noqa: E402
from vit_deep_radiomics_b200.distributed import allreduce_grads, init_distributed  # noqa: E402
from vit_deep_radiomics_b200.models_archs import TransformerNoduleClassifier  # noqa: E402
from vit_deep_radiomics_b200.train_models import FocalLoss  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=64)       # per optimizer window: 2 virtual batches of 32
    ap.add_argument("--cpu-samples", type=int, default=4, dest="cpu_samples")
    ap.add_argument("--epochs", type=int, default=3)
    args = ap.parse_args()
    rank, world = init_distributed("nccl")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    ids, labels, sizes, cloud = synth.point_cloud_patients(args.samples, d=256, n_range=(512, 4096), seed=1236)
    torch.manual_seed(0)
    model = TransformerNoduleClassifier(256, 1024, 4, 2, 2).to(dev)
    if world > 1:   # same initial weights on every rank
        for p in model.parameters():
            torch.distributed.broadcast(p.data, 0)
    opt = torch.optim.AdamW(model.parameters(), lr=5e-4, weight_decay=0.01)
    crit = FocalLoss(alpha=torch.tensor([0.25, 0.75], device=dev), gamma=2)
    mine = list(range(rank, args.samples, world))
    data = [(torch.from_numpy(cloud(i)).to(dev), torch.eye(2, device=dev)[int(labels[i])]) for i in mine]
    virtual = 32
    iters_global = min(virtual, args.samples)
    per_rank_window = max(1, virtual // world)

    def epoch():
        tot = 0.0
        opt.zero_grad()
        for k, (x, y) in enumerate(data):
            logits, _ = model(x.unsqueeze(0))
            loss = crit(torch.squeeze(logits), y) / iters_global          # train_models.py:674
            loss.backward()
            tot += float(loss.detach()) * iters_global
            if (k + 1) % per_rank_window == 0 or k + 1 == len(data):       # :685
                allreduce_grads(model)
                opt.step()
                opt.zero_grad()
        if world > 1:   # the loss trajectory is over ALL samples of the epoch, not rank 0's share
            t = torch.tensor([tot, float(len(data))], device=dev, dtype=torch.float64)
            torch.distributed.all_reduce(t)
            return float(t[0] / t[1])
        return tot / len(data)

    losses = [epoch()]                                                     # warm-up epoch (also first loss)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    t0 = time.perf_counter()
    for _ in range(args.epochs):
        losses.append(epoch())
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    dt = time.perf_counter() - t0
    if rank != 0:
        return
    # CPU baseline: fp32 restatement of the same architecture (oracle), fwd + bwd, all host cores, bounded sample
    from oracle import classifier_fp32 as C
    torch.set_num_threads(os.cpu_count() or 1)
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    t1 = time.perf_counter()
    for i in range(args.cpu_samples):
        x = torch.from_numpy(cloud(i))[None]
        lg, _ = C.classifier_forward(sd, x, 4, 2)
        C.focal_loss(lg[0], torch.eye(2)[int(labels[i])], 2.0, torch.tensor([0.25, 0.75])).backward()
    cpu_dt = time.perf_counter() - t1
    tokens = int(sum(sizes))
    print(json.dumps({"metric": "classifier training samples/s (fwd+bwd+AdamW, virtual batch 32)", "value": args.samples * args.epochs / dt,
                      "unit": "samples/s", "n_gpus": world, "tokens_per_s": tokens * args.epochs / dt, "mean_tokens": tokens / args.samples,
                      "loss_by_epoch": losses, "loss_decreases": bool(losses[-1] < losses[0]),
                      "cpu_baseline": {"value": args.cpu_samples / cpu_dt, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                                       "sample": f"first {args.cpu_samples} clouds, fp32 oracle fwd+bwd"}}))


if __name__ == "__main__":
    main()
