"""Multi-GPU plumbing: one process per GPU (torchrun), NCCL over NVLink for the two real exchange
steps of the path; gloo on CPU for the tests.

  * extraction: patients are independent -> sharded by contiguous patient range with NO data-path
    collective during compute; the point-cloud table is assembled by ``PointCloudTable``:
    every patient's row count is known before its backbone runs (``ops.mask_count`` on the uploaded
    mask), ONE small all-gather exchanges the counts, a device-side exclusive scan turns them into row
    offsets, and the gather kernel of each patient writes its rows directly at that offset of the
    preallocated table (``vdr_mask_gather_table`` reads the offset from device memory) -- so the table
    buffer IS the all-gather buffer and the final exchange is one in-place variable-length all-gather
    (grouped NCCL broadcasts of each rank's row range).  Row order = (patient, slice-major candidate
    order) by construction: bit-identical to the single-GPU table, no sort.
  * classifier training: data parallel over the 32-sample virtual batch -> one all-reduce(sum) per
    optimizer step of a persistent flat fp32 bucket whose slices ARE the parameters' ``.grad`` tensors
    (``GradBucket``; no concatenation, no copy back).  The loss is pre-divided by the GLOBAL
    accumulation count (train_models.py:674), so the reduce op is a plain sum.
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


def _active() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of n_items (patients, or slices of one volume): rank-order concatenation of the
    per-rank tables is then already in canonical (patient, slice, row, col) order."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_modulo(n_items: int, rank: int, world: int):
    """Patient i -> rank i mod world."""
    return list(range(rank, n_items, world))


def variable_all_gather_(buf: torch.Tensor, ranges, rank: int) -> None:
    """In-place variable-length all-gather over dim 0 of ``buf``, a preallocated buffer that exists on every rank:
    ``ranges[r] = (lo, hi)`` are the consecutive row ranges of the ranks; rank r's range already holds its data.
    NCCL: ONE launch -- equal sizes: ``ncclAllGather`` in place (send buffer = the rank's slot of the receive buffer);
    unequal: ``dist.all_gather`` on the slot views = ncclGroupStart / one ncclBroadcast per rank / ncclGroupEnd, input aliasing
    the rank's own slot.  gloo (CPU tests): one broadcast per non-empty range."""
    world = dist.get_world_size()
    if len(ranges) != world:
        raise ValueError("one row range per rank expected")
    sizes = [hi - lo for lo, hi in ranges]
    if any(ranges[r][1] != ranges[r + 1][0] for r in range(world - 1)):
        raise ValueError("row ranges must be consecutive")
    if sum(sizes) == 0:
        return
    views = [buf[lo:hi] for lo, hi in ranges]
    if dist.get_backend() == "nccl" and min(sizes) > 0:
        if len(set(sizes)) == 1:
            dist.all_gather_into_tensor(buf[ranges[0][0]:ranges[-1][1]], views[rank])
        else:
            dist.all_gather(views, views[rank])
        return
    for r, v in enumerate(views):
        if v.numel() > 0:
            dist.broadcast(v, src=r)


def all_gather_table(keys: torch.Tensor, rows: torch.Tensor, sort: bool = True):
    """Variable-length all-gather of an arbitrary per-rank table (host-level API; the extraction path uses
    ``PointCloudTable``, whose buffer the gather kernels fill in place).

    keys (n, k) int32/int64 -- e.g. (patient, slice, row, col); rows (n, D) payload.
    Every rank receives the rank-order concatenation of all ranks' tables (counts exchange -> preallocated output -> in-place
    all-gather); with ``sort`` (needed only when the sharding is not by contiguous ranges) the result is put in canonical
    lexicographic key order so it is bit-identical to the 1-GPU table whatever the sharding.
    """
    if not _active():
        all_keys, all_rows = keys, rows
    else:
        world, rank = dist.get_world_size(), dist.get_rank()
        n = torch.tensor([keys.shape[0]], dtype=torch.int64, device=keys.device)
        counts = torch.empty(world, dtype=torch.int64, device=keys.device)
        dist.all_gather_into_tensor(counts, n)                       # exchange 1: row counts
        counts = counts.tolist()
        offs = [0]
        for c in counts:
            offs.append(offs[-1] + c)
        all_keys = torch.empty((offs[-1], keys.shape[1]), dtype=keys.dtype, device=keys.device)
        all_rows = torch.empty((offs[-1], rows.shape[1]), dtype=rows.dtype, device=rows.device)
        all_keys[offs[rank]:offs[rank + 1]] = keys
        all_rows[offs[rank]:offs[rank + 1]] = rows
        ranges = [(offs[r], offs[r + 1]) for r in range(world)]
        variable_all_gather_(all_keys, ranges, rank)                 # exchange 2: the rows, in place
        variable_all_gather_(all_rows, ranges, rank)
    if sort and all_keys.shape[0] > 1:
        order = torch.arange(all_keys.shape[0], device=all_keys.device)
        for col in range(all_keys.shape[1] - 1, -1, -1):              # stable LSD radix over key columns
            order = order[torch.sort(all_keys[order, col], stable=True).indices]
        all_keys, all_rows = all_keys[order], all_rows[order]
    return all_keys, all_rows


class PointCloudTable:
    """The point-cloud table of ``n_patients`` patients sharded over the ranks by contiguous patient range
    (reference: the per-patient loop tfds_dense_descriptor.py:421 + merge_dataframe_features.py:12-30 build it on one host).

    Per step:   ``count(i, ...)`` for every local patient (device int64 counts, from the mask alone)
                -> ``exchange_counts()`` (one all-gather of the padded count vectors, device exclusive scan -> row offsets;
                   the counts also go to the host asynchronously: NCCL needs the range sizes later)
                -> per local patient ``ops.mask_gather(..., table=self.slot(i))``: rows land at the patient's offset
                -> ``all_gather()``: one in-place variable-length all-gather of each rank's contiguous row range.
    ``tokens`` (cap, D) f32 / ``src`` (cap, 4) int32 = (patient, slice, row, col); rows [0, total) are valid afterwards and
    identical on every rank and to the single-GPU table."""

    def __init__(self, n_patients: int, D: int, cap_rows: int, device, rank: int | None = None, world: int | None = None):
        self.world = world if world is not None else (dist.get_world_size() if _active() else 1)
        self.rank = rank if rank is not None else (dist.get_rank() if _active() else 0)
        self.n_patients, self.D, self.cap = int(n_patients), int(D), int(cap_rows)
        self.lo, self.hi = shard_range(self.n_patients, self.rank, self.world)
        self.p_max = -(-self.n_patients // self.world)                 # padded patients per rank (count vector length)
        dev = torch.device(device)
        self.tokens = torch.empty((self.cap, self.D), dtype=torch.float32, device=dev)
        self.src = torch.empty((self.cap, 4), dtype=torch.int32, device=dev)
        self.counts_local = torch.zeros(self.p_max, dtype=torch.int64, device=dev)
        self.counts_all = torch.zeros(self.world * self.p_max, dtype=torch.int64, device=dev)
        self.offsets = torch.zeros(self.world * self.p_max + 1, dtype=torch.int64, device=dev)
        pin = dev.type == "cuda"
        self.counts_host = torch.zeros(self.world * self.p_max, dtype=torch.int64, pin_memory=pin)
        self._counts_ready = torch.cuda.Event() if pin else None
        self.total = 0

    def local_patients(self):
        return range(self.lo, self.hi)

    def _slot_index(self, patient: int) -> int:
        """Position of a patient in the padded (rank-major) count / offset vectors."""
        for r in range(self.world):
            lo, hi = shard_range(self.n_patients, r, self.world)
            if lo <= patient < hi:
                return r * self.p_max + (patient - lo)
        raise IndexError(patient)

    def count_out(self, patient: int) -> torch.Tensor:
        """int64[1] view that receives the row count of a LOCAL patient (pass as ``out=`` to ops.mask_count)."""
        if not self.lo <= patient < self.hi:
            raise IndexError(f"patient {patient} is not local to rank {self.rank}")
        i = patient - self.lo
        return self.counts_local[i:i + 1]

    def exchange_counts(self) -> None:
        from . import ops
        if self.world > 1:
            dist.all_gather_into_tensor(self.counts_all, self.counts_local)
        else:
            self.counts_all.copy_(self.counts_local)
        if self.counts_all.is_cuda:
            ops.exclusive_scan_i64(self.counts_all, out=self.offsets)
            self.counts_host.copy_(self.counts_all, non_blocking=True)     # consumed by all_gather(), after the backbones are enqueued
            self._counts_ready.record()
        else:                                                              # CPU (gloo tests of the host logic)
            self.offsets[0] = 0
            torch.cumsum(self.counts_all, 0, out=self.offsets[1:])
            self.counts_host.copy_(self.counts_all)

    def slot(self, patient: int) -> dict:
        """``table=`` argument of ops.mask_gather for a local patient."""
        i = self._slot_index(patient)
        return dict(tokens=self.tokens, src=self.src, row_offset=self.offsets[i:i + 1], patient=int(patient))

    def rank_ranges(self):
        """[(row_lo, row_hi)] per rank, from the host copy of the counts."""
        if self._counts_ready is not None:
            self._counts_ready.synchronize()
        c = self.counts_host.view(self.world, self.p_max).sum(1).tolist()
        out, run = [], 0
        for n in c:
            out.append((run, run + int(n)))
            run += int(n)
        return out

    def all_gather(self) -> int:
        """Exchange the rows; returns the table's total row count."""
        ranges = self.rank_ranges()
        self.total = ranges[-1][1]
        if self.total > self.cap:
            raise ValueError(f"point-cloud table overflow: {self.total} rows, capacity {self.cap}")
        if self.world > 1:
            variable_all_gather_(self.tokens, ranges, self.rank)
            variable_all_gather_(self.src, ranges, self.rank)
        return self.total

    def bytes_exchanged(self) -> int:
        """Payload bytes one rank RECEIVES in all_gather() (everything but its own range)."""
        lo, hi = self.rank_ranges()[self.rank]
        return (self.total - (hi - lo)) * (self.D * 4 + 16)


class GradBucket:
    """All gradients of a model as ONE persistent flat fp32 buffer; every parameter's ``.grad`` is a view into it, so the
    backward pass accumulates straight into the all-reduce buffer (1,712,898 elements = 6.85 MB for the reference's
    classifier): no concatenation before and no copy back after the collective, and zeroing is one memset."""

    def __init__(self, model: torch.nn.Module):
        self.params = [p for p in model.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("model has no trainable parameters")
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.attach()

    def attach(self) -> None:
        """(Re-)point every ``.grad`` at its slice; a gradient that lives elsewhere (``zero_grad(set_to_none=True)`` then a
        backward) is first copied in."""
        for p, v in zip(self.params, self.views):
            g = p.grad
            if g is None or g.data_ptr() != v.data_ptr():
                if g is not None:
                    v.copy_(g)
                p.grad = v

    def zero(self) -> None:
        self.attach()
        self.flat.zero_()

    def allreduce(self) -> None:
        self.attach()
        if _active():
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)


def grad_bucket(model: torch.nn.Module) -> GradBucket:
    b = model.__dict__.get("_vdr_grad_bucket")
    if b is None or b.params[0].device != next(model.parameters()).device:
        b = model.__dict__["_vdr_grad_bucket"] = GradBucket(model)
    return b


def allreduce_grads(model: torch.nn.Module):
    """One all-reduce(sum) of all gradients through the model's persistent flat bucket (``GradBucket``): latency-bound, so one
    launch instead of one per tensor."""
    if not _active():
        return
    grad_bucket(model).allreduce()


def zero_grads(model: torch.nn.Module, optimizer=None) -> None:
    """zero_grad of the training loops (train_models.py:653,687): one memset of the flat bucket when the model has one (the
    views must survive, so not ``set_to_none``), else the optimizer's own."""
    b = model.__dict__.get("_vdr_grad_bucket")
    if b is not None:
        b.zero()
    elif optimizer is not None:
        optimizer.zero_grad()
    else:
        model.zero_grad()
