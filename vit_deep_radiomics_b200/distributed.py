"""Multi-GPU plumbing: one process per GPU (torchrun), NCCL over NVLink for the two real exchange
steps of the path; gloo on CPU for the tests.

  * extraction: patients / slices are independent -> sharded with NO data-path collective; the
    per-rank point-cloud tables are assembled with one variable-length all-gather
    (counts first, then rows) -- ``all_gather_table``;
  * classifier training: data parallel over the 32-sample virtual batch -> one flat-bucket
    all-reduce(sum) of the gradients right before each optimizer step -- ``allreduce_grads``
    (the loss is pre-divided by the GLOBAL accumulation count, train_models.py:674, so the
    reduce op is a plain sum).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_distributed(backend: str | None = None):
    """Initialise from torchrun's env (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).  Returns (rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of n_items (slices of one volume): rank-order concatenation of the
    per-rank tables is then already in canonical (patient, slice, row, col) order."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_modulo(n_items: int, rank: int, world: int):
    """Patient i -> rank i mod world."""
    return list(range(rank, n_items, world))


def all_gather_table(keys: torch.Tensor, rows: torch.Tensor, sort: bool = True):
    """Variable-length all-gather of a point-cloud table.

    keys (n, k) int32/int64 -- e.g. (patient, slice, row, col); rows (n, D) payload.
    Every rank receives the concatenation of all ranks' tables; with ``sort`` the result is put in
    canonical lexicographic key order so it is bit-identical to the 1-GPU table whatever the sharding.
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        all_keys, all_rows = keys, rows
    else:
        world = dist.get_world_size()
        n = torch.tensor([keys.shape[0]], dtype=torch.int64, device=keys.device)
        counts = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(counts, n)                                   # exchange 1: row counts
        counts = [int(c.item()) for c in counts]
        cap = max(max(counts), 1)
        kpad = torch.zeros((cap, keys.shape[1]), dtype=keys.dtype, device=keys.device)
        rpad = torch.zeros((cap, rows.shape[1]), dtype=rows.dtype, device=rows.device)
        kpad[: keys.shape[0]] = keys
        rpad[: rows.shape[0]] = rows
        kall = torch.empty((world * cap, keys.shape[1]), dtype=keys.dtype, device=keys.device)
        rall = torch.empty((world * cap, rows.shape[1]), dtype=rows.dtype, device=rows.device)
        dist.all_gather_into_tensor(kall, kpad)                      # exchange 2: padded rows
        dist.all_gather_into_tensor(rall, rpad)
        sel = torch.cat([torch.arange(r * cap, r * cap + c, device=keys.device) for r, c in enumerate(counts)])
        all_keys, all_rows = kall[sel], rall[sel]
    if sort and all_keys.shape[0] > 1:
        order = torch.arange(all_keys.shape[0], device=all_keys.device)
        for col in range(all_keys.shape[1] - 1, -1, -1):              # stable LSD radix over key columns
            order = order[torch.sort(all_keys[order, col], stable=True).indices]
        all_keys, all_rows = all_keys[order], all_rows[order]
    return all_keys, all_rows


def allreduce_grads(model: torch.nn.Module):
    """One all-reduce(sum) of all gradients as a single flat fp32 bucket (1,712,898 elements = 6.85 MB
    for the reference's classifier): latency-bound, so one launch instead of one per tensor."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    params = [p for p in model.parameters() if p.requires_grad]
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    off = 0
    for p in params:
        n = p.numel()
        p.grad.copy_(flat[off:off + n].view_as(p))
        off += n
