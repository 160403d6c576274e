"""One training sample of the point-cloud classifier (unimodal or bimodal) -- forward, focal loss / iters_to_accumulate,
backward -- as a CUDA graph.

The reference trains at batch size 1 over variable-length clouds (src/train_models.py:652-688): per sample ~120 kernels of a few
microseconds each, so the step is bound by how fast the host can issue them.  A captured graph replays the whole sample with one
launch.  The token count is baked into a graph (grids, tensor maps), so graphs are kept per count -- per (CT count, PET count)
pair for the bimodal model: a cloud length is captured the
second time it is seen (the first pass runs eagerly and initialises every per-shape kernel attribute) and replayed from then on --
a training set is a fixed list of patients visited every epoch, and 180 GB of HBM hold the activation pools of hundreds of
lengths (LRU-evicted beyond ``max_bytes``).

What makes the replay equal to the eager step:
  * gradients accumulate into the model's persistent flat bucket (distributed.GradBucket): same addresses every time;
  * the bf16 operand copies of the weights are refreshed IN PLACE after an optimizer step (classifier_kernels.refresh_cache);
  * train-mode dropout draws from (seed + *seed_offset): the graph itself increments the device counter, so every replay uses
    fresh masks although its launch parameters are frozen (include/vdr.h, vdr_dropout.seed_offset).
"""
from __future__ import annotations

from collections import OrderedDict

import torch

from . import classifier_kernels as ck
from .distributed import grad_bucket


class GraphedTrainStep:
    def __init__(self, model, criterion, max_bytes: int = 48 << 30, min_repeats: int = 2):
        self.model, self.criterion = model, criterion
        self.max_bytes, self.min_repeats = int(max_bytes), int(min_repeats)
        self.graphs: OrderedDict = OrderedDict()          # token count -> dict(graph, x, y, scale, loss, logits, cls, bytes)
        self.seen: dict = {}
        self.bytes = 0
        self.replays = self.eager = 0
        dev = next(model.parameters()).device
        self.device = dev
        self.seed_offset = torch.zeros(1, dtype=torch.int64, device=dev)
        self.base_seed = ck.new_seed()
        grad_bucket(model)                                 # .grad = views of one flat buffer, from now on

    # ------------------------------------------------------------------ eager reference path (also the first visit of a length)
    def _loss(self, outputs, label):
        """(loss, logits) of one sample's outputs: the bimodal criterion takes the fused and the two single-modality logits
        (train_models.py:668-672), the unimodal one the logits."""
        from .train_models import CrossModalFocalLoss
        logits = outputs[0]
        if isinstance(self.criterion, CrossModalFocalLoss):
            return self.criterion(torch.squeeze(logits), torch.squeeze(outputs[2]), torch.squeeze(outputs[3]), label), logits
        return self.criterion(torch.squeeze(logits), label), logits

    def _eager(self, xs, label, inv_iters):
        outputs = self.model(*(x.unsqueeze(0) for x in xs))
        loss, logits = self._loss(outputs, label)
        loss = loss * inv_iters
        loss.backward()
        self.eager += 1
        return loss.detach(), logits.detach()

    def _capture(self, key, d: int, classes: int):
        dev = self.device
        e = dict(xs=[torch.zeros(n, d, dtype=torch.float32, device=dev) for n in key], y=torch.zeros(classes, dtype=torch.float32, device=dev),
                 scale=torch.ones((), dtype=torch.float32, device=dev))
        ck.refresh_cache()
        grad_bucket(self.model).attach()
        before = torch.cuda.memory_reserved(dev)               # the graph's private pool is reserved, not 'allocated', once capture ends
        g = torch.cuda.CUDAGraph()
        model = self.model
        model._drop_override = lambda: ck.DropCfg(self.base_seed, *model._drop_rates(), seed_offset=self.seed_offset)
        try:
            with torch.cuda.graph(g):
                self.seed_offset.add_(1)
                outputs = model(*(x.unsqueeze(0) for x in e["xs"]))
                loss, logits = self._loss(outputs, e["y"])
                loss = loss * e["scale"]
                loss.backward()
                e["loss"], e["logits"] = loss.detach(), logits.detach()
        finally:
            model._drop_override = None
        e["graph"] = g
        e["bytes"] = max(0, torch.cuda.memory_reserved(dev) - before)
        self.bytes += e["bytes"]
        self.graphs[key] = e
        while self.bytes > self.max_bytes and len(self.graphs) > 1:
            _, old = self.graphs.popitem(last=False)
            self.bytes -= old["bytes"]
        return e

    def __call__(self, x, label: torch.Tensor, inv_iters: float = 1.0):
        """x (n, d) f32 CUDA -- or a tuple (x_ct, x_pet) for the bimodal model -- label (classes,) one-hot f32 CUDA.  Accumulates the
        sample's gradients of loss * inv_iters into the parameters' .grad; returns (loss, logits) -- device tensors that stay valid
        until the next call of the same length(s)."""
        xs = tuple(x) if isinstance(x, (tuple, list)) else (x,)
        key = tuple(int(t.shape[0]) for t in xs)
        e = self.graphs.get(key)
        if e is None:
            if len(self.seen) > 8192:          # augmented datasets crop at random: lengths rarely repeat, do not count them forever
                self.seen.clear()
            c = self.seen[key] = self.seen.get(key, 0) + 1
            if c < self.min_repeats or not torch.is_grad_enabled():
                return self._eager(xs, label, inv_iters)
            e = self._capture(key, int(xs[0].shape[1]), int(label.numel()))
        else:
            self.graphs.move_to_end(key)
        ck.refresh_cache()
        for dst, src in zip(e["xs"], xs):
            dst.copy_(src, non_blocking=True)
        e["y"].copy_(label, non_blocking=True)
        e["scale"].fill_(inv_iters)
        e["graph"].replay()
        self.replays += 1
        return e["loss"], e["logits"]


def graphed_step(model, criterion) -> GraphedTrainStep:
    """The model's cached GraphedTrainStep for this criterion (graphs survive across epochs)."""
    s = model.__dict__.get("_vdr_graphed_step")
    if s is None or s.criterion is not criterion or s.device != next(model.parameters()).device:
        s = model.__dict__["_vdr_graphed_step"] = GraphedTrainStep(model, criterion)
    return s
