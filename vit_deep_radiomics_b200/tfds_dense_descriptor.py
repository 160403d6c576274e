"""ViT dense-descriptor extraction -- drop-in for the reference's src/tfds_dense_descriptor.py.

Same entry points (``load_model``, ``prepare_image``, ``get_dense_descriptor``,
``generate_features``, ``save_features``, ``apply_window_ct``, ``flip_image``, ``rotate_image``,
``get_voxels``) and the same CLI flags; the backbone forward runs in libvdr.so (tcgen05 GEMMs,
fused attention, warp-shuffle LayerNorm) on batches of slices instead of one slice per call
with a host round trip (reference hot loop, :271-283).

``extract_point_cloud`` is the fused device path for one patient: crop -> batched ViT forward ->
ROI crop -> tumour-mask gather (+ positional encoding), i.e. L2 -> L4a of SURVEY.md without
the HDF5 hand-off; it returns exactly what ``PETCTDataset3D._get_features`` would compute from
the files ``generate_features`` + ``save_features`` write.
"""
from __future__ import annotations

import argparse
import os

import numpy as np
import torch

from . import ops
from .visualization_utils import (crop_image, crop_window, crop_window_from_bbox, extract_roi, mask_bbox, roi_window,
                                  roi_window_from_bbox)
from .sam_encoder import SAM_CONFIGS, DinoV2PatchEmbed, SamImageEncoder
from .vit import VIT_CONFIGS, ViTBackbone

# --------------------------------------------------------------------------------------------- model
def load_model(model_name, model_path=None, img_hw=None, device="cuda:0", seed=1234):
    """reference: tfds_dense_descriptor.py:51-67.  ``model_name``: the reference's 'medsam' (SAM ViT-B image encoder,
    sam_encoder.SamImageEncoder) and 'dinov2' (patch embedding only, as the reference uses it), plus the plain-ViT
    backbones of BASELINE.json: 'vit_s16', 'vit_b16', 'vit_l14'.  ``model_path`` (optional) is a torch state-dict
    (segment_anything / DINOv2 / timm key names); without it the weights are seeded random (no checkpoints exist offline)."""
    sd = None
    if model_path is not None:
        if not os.path.isfile(model_path):
            raise FileNotFoundError(f"load_model: checkpoint {model_path!r} not found (random weights only on request: model_path=None)")
        sd = torch.load(model_path, map_location="cpu")
        sd = sd.get("state_dict", sd) if isinstance(sd, dict) else sd
    else:
        import warnings
        warnings.warn(f"load_model({model_name!r}): no model_path -- the backbone gets SEEDED RANDOM weights (seed {seed}); "
                      "descriptors from it are only good for tests and benchmarks", stacklevel=2)
    if model_name in SAM_CONFIGS:
        # load_medsam (:91-107): sam_model_registry['vit_b'](model_path); only model.image_encoder is ever called (:123).
        # prepare_image feeds it 1024 x 1024 (:42); other sizes need a pos_embed of that grid.
        model = SamImageEncoder(model_name, img_hw=img_hw or (1024, 1024), state_dict=sd, device=device, seed=seed)
        model.model_name = model_name
        return model
    if model_name == "dinov2":
        # load_dinov2 (:70-88) + model.patch_embed (:128): ViT-S/14 patch embedding only; RGB inputs are resized to 896^2 (:44)
        model = DinoV2PatchEmbed(img_hw=img_hw or (896, 896), state_dict=sd, device=device, seed=seed)
        return model
    if model_name not in VIT_CONFIGS:
        raise ValueError(f"unknown model_name {model_name!r}")
    if img_hw is None:
        img_hw = (518, 518) if VIT_CONFIGS[model_name]["patch"] == 14 else (512, 512)
    model = ViTBackbone(model_name, img_hw=img_hw, state_dict=sd, device=device, seed=seed)
    model.model_name = model_name
    return model


def prepare_image(img, size=None, device="cuda:0"):
    """reference: tfds_dense_descriptor.py:30-48 -- gray2rgb, resize, HWC -> NCHW, float32, to GPU.
    Returns a (1, 3, h, w) CUDA tensor.  The reference resizes to 1024^2 / 896^2 with skimage.transform.resize; here
    ``size`` is the backbone input and the resize (same algorithm: anti-aliasing Gaussian when shrinking + order-1
    resampling with mirrored borders) runs on the device in f32, rounded to bf16 like every backbone input."""
    img = np.asarray(img)
    if size is not None and tuple(size) != tuple(img.shape[0:2]):
        planes = img[..., None] if img.ndim < 3 else img                      # (H, W, C): channels take the slice axis
        vol = torch.as_tensor(np.ascontiguousarray(planes, dtype=np.float32)).to(device)
        out = ops.volume_to_slices(vol, (0, img.shape[0], 0, img.shape[1]), out_hw=size).float()   # (C, h, w)
        if out.shape[0] == 1:
            out = out.expand(3, -1, -1)                                        # gray2rgb (:41)
        return out[None].contiguous()
    if img.ndim < 3:
        img = np.stack([img] * 3, axis=-1)          # gray2rgb (:41)
    t = torch.as_tensor(np.ascontiguousarray(img.transpose((2, 0, 1))[None]), dtype=torch.float32)
    return t.to(device)


def get_dense_descriptor(model, img):
    """reference: tfds_dense_descriptor.py:110-139.  img (N, M[, CH]) in 0..1 ->
    features (N//patch, M//patch, feature_dim) float32 numpy (patch tokens only, HWC)."""
    x = prepare_image(img, size=model.img_hw, device=model.device)
    feats = model.dense_descriptors(x)
    return feats[0].cpu().numpy()


# --------------------------------------------------------------------------------------------- volume level
def _plan_from_bbox(model, H, W, bbox):
    """Host-side index math of generate_features (:257-267, :278-279) from the union mask's bounding box
    (row_min, row_max, col_min, col_max): crop window, feature ROI, pixel-mask ROI.  Pure integer geometry.
    Returns None when the box is not inside its own crop window (then the cropped mask must be inspected)."""
    rmin, rmax, cmin, cmax = bbox
    x0, y0, x1, y1 = crop_window_from_bbox(bbox)
    y0c, y1c = [max(0, min(v, H)) for v in (y0, y1)]
    x0c, x1c = [max(0, min(v, W)) for v in (x0, x1)]
    if not (y0c <= rmin and rmax < y1c and x0c <= cmin and cmax < x1c):
        return None
    ch, cw = y1c - y0c, x1c - x0c          # a window of another size than the backbone input is resized on the device (:40-44)
    bbox_c = (rmin - y0c, rmax - y0c, cmin - x0c, cmax - x0c)            # union mask after crop_image (:267)
    gh, gw = model.grid
    fx0, fy0, fx1, fy1 = roi_window_from_bbox((gh, gw), (ch, cw), bbox_c, margin=1)   # extract_roi(features, bigger_mask)
    mx0, my0, mx1, my1 = roi_window_from_bbox((ch, cw), (ch, cw), bbox_c, margin=1)   # extract_roi(mask, bigger_mask)
    return dict(crop=(y0c, y1c, x0c, x1c), feat_roi=(fy0, fy1, fx0, fx1), mask_roi=(my0, my1, mx0, mx1))


def _plan(model, mask_3d):
    """Plan from a host mask (H, W, S)."""
    H, W = mask_3d.shape[0:2]
    bigger = np.any(mask_3d, axis=-1)                                    # == (np.sum(mask_3d, -1) > 0), :257
    plan = _plan_from_bbox(model, H, W, mask_bbox(bigger))
    if plan is not None:
        return plan
    # rare: the (shifted) crop window cuts the mask -> use the cropped union mask itself, as the reference does
    x0, y0, x1, y1 = crop_window(bigger)
    y0c, y1c = [max(0, min(v, H)) for v in (y0, y1)]
    x0c, x1c = [max(0, min(v, W)) for v in (x0, x1)]
    bigger_c = bigger[y0c:y1c, x0c:x1c]
    ch, cw = bigger_c.shape
    gh, gw = model.grid
    fx0, fy0, fx1, fy1 = roi_window((gh, gw), bigger_c, margin=1)
    mx0, my0, mx1, my1 = roi_window((ch, cw), bigger_c, margin=1)
    return dict(crop=(y0c, y1c, x0c, x1c), feat_roi=(fy0, fy1, fx0, fx1), mask_roi=(my0, my1, mx0, mx1))


def _forward_volume(model, img_dev, plan, max_batch=None):
    """Batched backbone forward over all slices of a device-resident volume (H, W, S[, 3]) f32.
    Gray volumes are staged slice-major (and resized to the backbone input when the crop window has another size) by one
    kernel; RGB volumes are read in place through element strides (no host transpose, no NCHW copy)."""
    y0, y1, x0, x1 = plan["crop"]
    S = img_dev.shape[2]
    if img_dev.dim() == 3:
        if img_dev.is_contiguous() and (max_batch is None or max_batch >= S):
            return model.forward_volume(img_dev, plan["crop"])      # coalesced slice staging (+ resize) -> TMA im2col GEMM
        step = S if max_batch is None else max_batch
        outs = []
        for s0 in range(0, S, step):
            chunk = img_dev[:, :, s0:s0 + min(step, S - s0)].contiguous()
            outs.append(model.forward_volume(chunk, plan["crop"]).clone())
        return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)
    if (y1 - y0, x1 - x0) != tuple(model.img_hw):
        raise NotImplementedError("RGB volumes whose crop window differs from the backbone input: resize slice by slice with "
                                  "prepare_image(img, size=model.img_hw)")
    view = img_dev[y0:y1, x0:x1]
    strides = (view.stride(2), view.stride(3), view.stride(0), view.stride(1))
    if max_batch is None or max_batch >= S:
        return model.forward_tokens(view, strides, S)
    outs = []
    for s0 in range(0, S, max_batch):
        b = min(max_batch, S - s0)
        outs.append(model.forward_tokens(view[:, :, s0:s0 + b], strides, b).clone())
    return torch.cat(outs, dim=0)


def generate_features(model, img_3d, mask_3d, tqdm_text="", display=False, max_batch=None):
    """reference: tfds_dense_descriptor.py:242-284.  Returns (features_list, mask_list): per slice the
    ROI-cropped descriptor map (h_f, w_f, D) float32 and the ROI-cropped boolean pixel mask."""
    img_3d, mask_3d = np.asarray(img_3d), np.asarray(mask_3d)
    plan = _plan(model, mask_3d)
    img_dev = torch.as_tensor(np.ascontiguousarray(img_3d, dtype=np.float32)).to(model.device)
    tok = _forward_volume(model, img_dev, plan, max_batch)
    S, d, off = img_3d.shape[2], model.feature_dim, model.token_offset
    gh, gw = model.grid
    fy0, fy1, fx0, fx1 = plan["feat_roi"]
    feats = tok.view(S, model.n_tokens, d)[:, off:, :].reshape(S, gh, gw, d)[:, fy0:fy1, fx0:fx1, :].cpu().numpy()
    y0, y1, x0, x1 = plan["crop"]
    my0, my1, mx0, mx1 = plan["mask_roi"]
    mask_c = mask_3d[y0:y1, x0:x1]
    features_list = [feats[s] for s in range(S)]
    mask_list = [(mask_c[:, :, s] > 0)[my0:my1, mx0:mx1] for s in range(S)]
    return features_list, mask_list


_BUILTIN_GENERATE_FEATURES = generate_features


def generate_features_device(model, img_dev, mask_dev, max_batch=None):
    """``generate_features`` for a volume that already lives on the device ((H, W, S) f32 image, (H, W, S) uint8 mask): the
    union-mask bounding box comes from a device reduction (24-byte read-back), the ROI crops of the descriptors and of the masks
    are read back as the reference returns them (lists of per-slice arrays)."""
    ex = getattr(model, "_extractor", None)
    if ex is None:
        ex = model._extractor = PointCloudExtractor(model)
    plan = ex.plan_patient(mask_dev)
    tok = _forward_volume(model, img_dev, plan, max_batch)
    S, d, off = img_dev.shape[2], model.feature_dim, model.token_offset
    gh, gw = model.grid
    fy0, fy1, fx0, fx1 = plan["feat_roi"]
    feats = tok.view(S, model.n_tokens, d)[:, off:, :].reshape(S, gh, gw, d)[:, fy0:fy1, fx0:fx1, :].cpu().numpy()
    y0, y1, x0, x1 = plan["crop"]
    my0, my1, mx0, mx1 = plan["mask_roi"]
    mask_c = mask_dev[y0 + my0:y0 + my1, x0 + mx0:x0 + mx1].cpu().numpy() > 0
    return [feats[s] for s in range(S)], [np.ascontiguousarray(mask_c[:, :, s]) for s in range(S)]


class _RoiReadback:
    """Read-back ring of the augmentation loop: the ROI crop of every augmented copy (descriptors (S, h_f, w_f, D) f32 and pixel
    masks (h_m, w_m, S) u8) is packed on the main stream into a device slot, copied into pinned staging on a copy stream and
    turned into caller-owned NumPy arrays by one worker thread, while the main stream is already in the backbone of the next
    copy.  Results come back in submission order from ``finish``."""

    def __init__(self, device, slots):
        import queue
        import threading
        self.device = device
        self.stream = torch.cuda.Stream(device=device)
        self.slots = slots                      # persistent across patients (pinned staging is expensive to allocate)
        for slot in slots:
            slot.setdefault("cap_f", 0), slot.setdefault("cap_m", 0)
            slot["idle"] = threading.Semaphore(1)
        self.jobs = queue.Queue()
        self.results = []
        self.error = None
        self.n = 0
        self.worker = threading.Thread(target=self._drain, daemon=True)
        self.worker.start()

    def _drain(self):
        while True:
            job = self.jobs.get()
            if job is None:
                return
            slot, copied, fshape, mshape = job
            try:
                copied.synchronize()                                        # (releases the GIL)
                nf, nm = int(np.prod(fshape)), int(np.prod(mshape))
                feats = np.array(slot["pin_f"][:nf].numpy().reshape(fshape))      # caller-owned copies: the staging is reused
                masks = np.array(slot["pin_m"][:nm].numpy().reshape(mshape))
                self.results.append((feats, masks))
            except BaseException as e:   # surfaced by finish()
                self.error = e
                self.results.append(None)
            finally:
                slot["idle"].release()

    def submit(self, feat_view, mask_view):
        """feat_view / mask_view: (strided) device views, valid on the current stream until the next backbone launch."""
        slot = self.slots[self.n % len(self.slots)]
        self.n += 1
        slot["idle"].acquire()                                              # the worker has emptied this slot's staging
        nf, nm = feat_view.numel(), mask_view.numel()
        if slot["cap_f"] < nf:
            slot.update(cap_f=nf, dev_f=torch.empty(nf, dtype=torch.float32, device=self.device),
                        pin_f=torch.empty(nf, dtype=torch.float32).pin_memory())
        if slot["cap_m"] < nm:
            slot.update(cap_m=nm, dev_m=torch.empty(nm, dtype=torch.uint8, device=self.device),
                        pin_m=torch.empty(nm, dtype=torch.uint8).pin_memory())
        main = torch.cuda.current_stream(self.device)
        # (the slot's previous D2H was synchronised by the worker before it released the slot: dev_f / dev_m are free)
        slot["dev_f"][:nf].view(feat_view.shape).copy_(feat_view)
        slot["dev_m"][:nm].view(mask_view.shape).copy_(mask_view)
        packed = torch.cuda.Event()
        packed.record(main)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(packed)
            slot["pin_f"][:nf].copy_(slot["dev_f"][:nf], non_blocking=True)
            slot["pin_m"][:nm].copy_(slot["dev_m"][:nm], non_blocking=True)
            copied = torch.cuda.Event()
            copied.record(self.stream)
        self.jobs.put((slot, copied, tuple(feat_view.shape), tuple(mask_view.shape)))

    def finish(self):
        self.jobs.put(None)
        self.worker.join()
        if self.error is not None:
            raise self.error
        return self.results


def _augmented_features_device(model, img_dev, mask_dev, mask_kind, grid):
    """The augmentation grid of one patient on the device, pipelined: ``grid`` = [(flip, angle), ...].
    1. every augmented MASK first (cheap) and all their union bounding boxes in ONE read-back -> the crop / ROI plans;
    2. per copy: image flip / rotation -> batched backbone -> ROI crops handed to the read-back ring (`_RoiReadback`).
    Per copy the same values as ``generate_features_device`` (same kernels, same plans); only the order of the waits differs.
    Returns [(features_list, mask_list)] per grid entry."""
    ex = getattr(model, "_extractor", None)
    if ex is None:
        ex = model._extractor = PointCloudExtractor(model)
    H, W, S = mask_dev.shape
    boxes = torch.empty((len(grid), 6), dtype=torch.int32, device=model.device)
    masks = []
    for k, (flip_type, angle) in enumerate(grid):
        m = ops.flip_rotate_volume(mask_dev, flip_type, angle, kind=mask_kind)
        ops.mask_bbox(m, out=boxes[k])
        masks.append(m)
    boxes_h = boxes.cpu().numpy()
    plans = []
    for k in range(len(grid)):
        cmin, cmax, rmin, rmax = (int(v) for v in boxes_h[k, :4])
        if cmax < cmin:
            raise ValueError("extract_coords: empty mask")
        plan = _plan_from_bbox(model, H, W, (rmin, rmax, cmin, cmax))
        plans.append(plan if plan is not None else _plan(model, masks[k].cpu().numpy()))
    if not hasattr(ex, "roi_slots"):
        ex.roi_slots = [dict() for _ in range(3)]
    ring = _RoiReadback(model.device, ex.roi_slots)
    d, off = model.feature_dim, model.token_offset
    gh, gw = model.grid
    image = None
    try:
        for k, (flip_type, angle) in enumerate(grid):
            plan = plans[k]
            image = ops.flip_rotate_volume(img_dev, flip_type, angle, kind="image", out=image)
            tok = _forward_volume(model, image, plan)
            fy0, fy1, fx0, fx1 = plan["feat_roi"]
            y0, y1, x0, x1 = plan["crop"]
            my0, my1, mx0, mx1 = plan["mask_roi"]
            ring.submit(tok.view(S, model.n_tokens, d)[:, off:, :].reshape(S, gh, gw, d)[:, fy0:fy1, fx0:fx1, :],
                        masks[k][y0 + my0:y0 + my1, x0 + mx0:x0 + mx1])
            masks[k] = None
    finally:
        results = ring.finish()
    out = []
    for feats, mask_c in results:
        mask_c = mask_c > 0
        out.append(([feats[s] for s in range(S)], [np.ascontiguousarray(mask_c[:, :, s]) for s in range(S)]))
    return out


def _as_pinned_pair(img_3d, mask_3d):
    img_t = img_3d if isinstance(img_3d, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(img_3d, dtype=np.float32))
    if isinstance(mask_3d, torch.Tensor):
        mask_t = mask_3d
    else:
        m = np.ascontiguousarray(mask_3d)
        mask_t = torch.as_tensor(m.view(np.uint8) if m.dtype == bool else (m > 0).view(np.uint8))
    return img_t, mask_t


class PointCloudExtractor:
    """Fused device path for a stream of patients: crop -> batched ViT forward -> ROI -> tumour-mask gather
    (+ positional encoding), i.e. what ``_get_features`` (train_models.py:143-182) returns for the features /
    masks ``generate_features`` produces, without the HDF5 hand-off and without leaving the GPU in between.

    Uploads are double-buffered on a copy stream: while patient i is in the backbone, patient i+1's volume and
    mask cross PCIe, and the union-mask bounding box that the crop geometry needs (tfds_dense_descriptor.py:257-263)
    is reduced on the device from the uploaded mask (24-byte read-back) instead of a host pass over 31 M voxels.
    """

    def __init__(self, model):
        self.model = model
        self.copy_stream = torch.cuda.Stream(device=model.device)
        self.slots = [dict(), dict()]

    def _upload(self, slot, img_t, mask_t):
        dev = self.model.device
        b = self.slots[slot]
        with torch.cuda.stream(self.copy_stream):
            if b.get("shape") != tuple(img_t.shape):
                # allocated in the COPY stream's pool: a block of the main stream's pool may have been freed by Python while kernels
                # that use it are still queued there (the previous patient's point cloud with to_host=False), and the copies below
                # would overwrite it without waiting for them.  Real datasets change shape with every patient.
                b.update(shape=tuple(img_t.shape), img=torch.empty(img_t.shape, dtype=torch.float32, device=dev),
                         mask=torch.empty(mask_t.shape, dtype=torch.uint8, device=dev),
                         bbox=torch.empty(6, dtype=torch.int32, device=dev),
                         bbox_host=torch.empty(6, dtype=torch.int32).pin_memory(), ev=torch.cuda.Event(), ev_box=torch.cuda.Event())
            if "free" in b:
                self.copy_stream.wait_event(b["free"])          # the previous user of this slot has been consumed
            b["mask"].copy_(mask_t, non_blocking=True)
            ops.mask_bbox(b["mask"], out=b["bbox"])
            b["bbox_host"].copy_(b["bbox"], non_blocking=True)
            b["ev_box"].record(self.copy_stream)
            b["img"].copy_(img_t, non_blocking=True)
            b["ev"].record(self.copy_stream)

    def _compute(self, slot, spatial_res, noise, add_pe):
        model, b = self.model, self.slots[slot]
        b["ev_box"].synchronize()                                # 24-byte bounding box is on the host
        cmin, cmax, rmin, rmax, _, _ = (int(v) for v in b["bbox_host"])
        if cmax < cmin:
            raise ValueError("extract_coords: empty mask")
        H, W, S = b["shape"]
        plan = _plan_from_bbox(model, H, W, (rmin, rmax, cmin, cmax))
        if plan is None:
            plan = _plan(model, b["mask"].cpu().numpy())
        main = torch.cuda.current_stream(model.device)
        main.wait_event(b["ev"])
        tok = _forward_volume(model, b["img"], plan)
        gh, gw = model.grid
        pe = dict(res=spatial_res, noise=noise, scale=0.25) if add_pe else None
        tokens, src, count = ops.mask_gather(tok, b["mask"], grid=(S, gh, gw, model.n_tokens, model.token_offset), feat_roi=plan["feat_roi"],
                                             mask_roi=_shift_roi(plan["mask_roi"], plan["crop"]), pe=pe, mask_layout="hws")
        b["free"] = torch.cuda.Event()
        b["free"].record(main)
        return plan, tokens, src, count

    def run(self, items, add_pe=True, to_host=True):
        """items: iterable of (img_3d, mask_3d, spatial_res[, noise]); yields one result dict per patient.

        ``to_host``: the point cloud of patient i is read back (row count first, then exactly that many rows, on a copy stream
        into pinned staging) AFTER the backbone of patient i + 1 has been enqueued, so the read-back and the host's wait for it
        never leave the GPU idle; results are yielded in order, one patient behind the launches."""
        it = iter(items)
        nxt = next(it, None)
        slot = 0
        if nxt is not None:
            self._upload(slot, *_as_pinned_pair(nxt[0], nxt[1]))
        pending = None
        while nxt is not None:
            cur, cur_slot = nxt, slot
            plan, tokens, src, count = self._compute(cur_slot, cur[2], cur[3] if len(cur) > 3 else (0.0, 0.0, 0.0), add_pe)
            nxt = next(it, None)
            slot ^= 1
            if nxt is not None:                                   # overlaps with the backbone of `cur`
                self._upload(slot, *_as_pinned_pair(nxt[0], nxt[1]))
            if not to_host:
                yield dict(plan=plan, tokens=tokens, src=src, count=count)
                continue
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(self.model.device))
            if pending is not None:
                yield self._read_back(*pending)
            pending = (plan, tokens, src, count, done)
        if pending is not None:
            yield self._read_back(*pending)

    def _read_back(self, plan, tokens, src, count, done):
        """D2H of one patient's point cloud on the read-back stream: 4-byte count, then count rows of tokens / source triplets."""
        if not hasattr(self, "d2h_stream"):
            self.d2h_stream = torch.cuda.Stream(device=self.model.device)
            self._n_host = torch.empty(1, dtype=torch.int32).pin_memory()
            self._stage = {}
        st = self.d2h_stream
        with torch.cuda.stream(st):
            st.wait_event(done)
            self._n_host.copy_(count, non_blocking=True)
            st.synchronize()
            n = int(self._n_host[0])
            key = (tokens.shape[1], src.shape[1])
            buf = self._stage.get(key)
            if buf is None or buf[0].shape[0] < n:
                rows = max(n, tokens.shape[0] // 2)
                buf = self._stage[key] = (torch.empty((rows, tokens.shape[1]), dtype=torch.float32).pin_memory(),
                                          torch.empty((rows, src.shape[1]), dtype=torch.int32).pin_memory())
            buf[0][:n].copy_(tokens[:n], non_blocking=True)
            buf[1][:n].copy_(src[:n], non_blocking=True)
            st.synchronize()
        # (tokens / src were allocated on the main stream and are referenced until here: the allocator cannot hand them out earlier)
        return dict(plan=plan, tokens=buf[0][:n].clone(), src=buf[1][:n].clone(), count=n)

    # ---- sharded extraction into ONE table (reference: the per-patient loop :421 + merge_dataframe_features.py)
    def plan_patient(self, mask_dev):
        """Crop / ROI plan of a device-resident (H, W, S) uint8 mask: the union-mask bounding box is reduced on the device
        (24-byte read-back), the rest is host integer geometry."""
        model = self.model
        H, W, S = mask_dev.shape
        box = ops.mask_bbox(mask_dev).cpu()
        cmin, cmax, rmin, rmax, _, _ = (int(v) for v in box)
        if cmax < cmin:
            raise ValueError("extract_coords: empty mask")
        plan = _plan_from_bbox(model, H, W, (rmin, rmax, cmin, cmax))
        return plan if plan is not None else _plan(model, mask_dev.cpu().numpy())

    def run_table(self, items, table, add_pe=True, max_rows_per_batch=131072):
        """This rank's patients of a sharded extraction, written straight into ``table`` (distributed.PointCloudTable).

        items: list of (patient_index, img (H, W, S) f32, mask (H, W, S) uint8, spatial_res[, noise]) for the patients of
        ``table.local_patients()``; tensors may be pinned host tensors (uploaded here) or already on the device.
        1. every local patient's row count from its mask alone (``ops.mask_count``), 2. counts all-gather + device scan
        (``table.exchange_counts``), 3. per patient backbone + gather at the patient's row offset of the table (the offset is read
        by the kernel from device memory), 4. one in-place all-gather of the rank's row range (``table.all_gather``).
        Returns the table's total row count; ``table.tokens[:total]`` / ``table.src[:total]`` then hold every rank's rows in
        (patient, candidate) order, identical on all ranks."""
        return self.emit_table(self.stage_table(items, table), table, add_pe=add_pe, max_rows_per_batch=max_rows_per_batch)

    def stage_table(self, items, table, count=True):
        """Step 1 of run_table: per patient the crop / ROI plan (24-byte read-back of the mask's bounding box) and, with
        ``count``, its row count into the table's count vector.  Returns the staged list ``emit_table`` takes."""
        model, dev = self.model, self.model.device
        gh, gw = model.grid
        staged = []
        items = list(items)
        # every mask's bounding box first (uploads + reductions enqueued back to back), then ONE read-back for all patients
        masks = [it[2] if it[2].is_cuda else it[2].to(dev, non_blocking=True) for it in items]
        boxes = torch.empty((max(len(items), 1), 6), dtype=torch.int32, device=dev)
        for i, m in enumerate(masks):
            ops.mask_bbox(m, out=boxes[i])
        boxes_host = boxes.cpu()
        for i, it in enumerate(items):
            pid, img_t, _, res = it[:4]
            noise = it[4] if len(it) > 4 else (0.0, 0.0, 0.0)
            mask_dev = masks[i]
            cmin, cmax, rmin, rmax, _, _ = (int(v) for v in boxes_host[i])
            if cmax < cmin:
                raise ValueError("extract_coords: empty mask")
            plan = _plan_from_bbox(model, mask_dev.shape[0], mask_dev.shape[1], (rmin, rmax, cmin, cmax))
            if plan is None:
                plan = _plan(model, mask_dev.cpu().numpy())
            S = mask_dev.shape[2]
            geo = dict(grid=(S, gh, gw, model.n_tokens, model.token_offset), feat_roi=plan["feat_roi"],
                       mask_roi=_shift_roi(plan["mask_roi"], plan["crop"]), mask_layout="hws")
            staged.append((pid, img_t, mask_dev, res, noise, plan, geo))
        if count:
            self.count_table(staged, table)
        return staged

    def count_table(self, staged, table):
        """Row counts of planned patients into the table's count vector (every step over the same patients starts here)."""
        for pid, _, mask_dev, _, _, _, geo in staged:
            ops.mask_count(mask_dev, out=table.count_out(pid), **geo)

    def emit_table(self, staged, table, add_pe=True, max_rows_per_batch=131072):
        """Steps 2-4 of run_table for patients staged by ``stage_table``."""
        model, dev = self.model, self.model.device
        table.exchange_counts()
        main = torch.cuda.current_stream(dev)
        # consecutive patients share one backbone batch while their token rows stay below what one 512 x 512 x 120 volume has
        # (small volumes do not fill the GPU on their own); the gather then runs per patient on its rows of the token matrix
        batchable = hasattr(model, "forward_volumes") and all(st[1].dim() == 3 for st in staged)
        groups, rows = [], 0
        for i, st in enumerate(staged):
            r = int(st[2].shape[2]) * model.n_tokens
            if groups and batchable and rows + r <= max_rows_per_batch:
                groups[-1].append(i)
                rows += r
            else:
                groups.append([i])
                rows = r

        def fetch(idx, stream_ctx):
            out = {}
            for i in idx:
                t = staged[i][1]
                if t.is_cuda:
                    out[i] = t
                else:
                    with stream_ctx():
                        out[i] = t.to(dev, non_blocking=True)
            return out

        import contextlib
        cur = fetch(groups[0], contextlib.nullcontext) if groups else {}
        host_side = any(not st[1].is_cuda for st in staged)
        fwd_prev = fwd_cur = None          # main-stream events behind the forward of the previous / this group (host-side volumes only)
        for gi, grp in enumerate(groups):
            vols = [(cur[i] if cur[i].is_contiguous() else cur[i].contiguous(), staged[i][5]["crop"]) for i in grp]
            if len(vols) > 1:
                tok = model.forward_volumes(vols)
            else:
                tok = _forward_volume(model, vols[0][0], staged[grp[0]][5])
            if host_side:
                fwd_prev, fwd_cur = fwd_cur, torch.cuda.Event()
                fwd_cur.record(main)
            nxt, ev = None, None
            if gi + 1 < len(groups) and any(not staged[i][1].is_cuda for i in groups[gi + 1]):
                # the next batch's volumes cross PCIe while this one is in the backbone.  Their buffers come from the copy stream's
                # pool, which hands out the previous group's volumes again as soon as Python dropped them -- while the host runs
                # several groups ahead of the GPU: the copies must not start before that group's forward has read them.
                if fwd_prev is not None:
                    self.copy_stream.wait_event(fwd_prev)
                nxt = fetch(groups[gi + 1], lambda: torch.cuda.stream(self.copy_stream))
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
            elif gi + 1 < len(groups):
                nxt = fetch(groups[gi + 1], contextlib.nullcontext)
            r0 = 0
            for i in grp:
                pid, _, mask_dev, res, noise, plan, geo = staged[i]
                r1 = r0 + int(mask_dev.shape[2]) * model.n_tokens
                pe = dict(res=res, noise=noise, scale=0.25) if add_pe else None
                ops.mask_gather(tok[r0:r1], mask_dev, pe=pe, table=table.slot(pid), **geo)
                r0 = r1
            if ev is not None:
                main.wait_event(ev)
            cur = nxt
        if fwd_cur is not None:
            self.copy_stream.wait_event(fwd_cur)    # the last volumes go back to the copy stream's pool when this returns
        return table.all_gather()


def _shift_roi(mask_roi, crop):
    """ROI of the cropped mask -> window of the full (H, W, S) mask."""
    my0, my1, mx0, mx1 = mask_roi
    y0, _, x0, _ = crop
    return (my0 + y0, my1 + y0, mx0 + x0, mx1 + x0)


def extract_point_cloud(model, img_3d, mask_3d, spatial_res, noise=(0.0, 0.0, 0.0), add_pe=True, to_host=True):
    """One patient through the fused device path (see PointCloudExtractor).

    img_3d (H, W, S) float32 in 0..1, mask_3d (H, W, S) bool/uint8 (numpy, or pinned torch tensors).
    Returns dict(tokens (n, D) f32, src (n, 3) int32 (slice, row, col) in ROI coords, count, plan).
    """
    ex = getattr(model, "_extractor", None)
    if ex is None:
        ex = model._extractor = PointCloudExtractor(model)
    return next(ex.run([(img_3d, mask_3d, spatial_res, noise)], add_pe=add_pe, to_host=to_host))


# --------------------------------------------------------------------------------------------- host helpers
def windowing_ct(width, level):
    """reference: tfds_dense_descriptor.py:204-239."""
    return level - width / 2, level + width / 2


def apply_window_ct(ct, width, level):
    """reference: tfds_dense_descriptor.py:287-303 -- HU window to 0..1."""
    lo, hi = windowing_ct(width, level)
    return np.clip((ct - lo) / (hi - lo), 0, 1)


def flip_image(image, mask, flip_type):
    """reference: tfds_dense_descriptor.py:306-325 (None | 'horizontal' | 'vertical')."""
    image, mask = image.copy(), mask.copy()
    if flip_type == "horizontal":
        return image[:, ::-1, ...], mask[:, ::-1, ...]
    if flip_type == "vertical":
        return image[::-1, ...], mask[::-1, ...]
    return image, mask


def rotate_image(image, mask, angle, axes=(0, 1)):
    """reference: tfds_dense_descriptor.py:328-350 -- offline augmentation on the host (scipy cubic
    spline, mode='nearest'); image clipped to 0..1, mask re-binarised."""
    from scipy.ndimage import rotate
    image, mask = image.copy(), mask.copy()
    if angle == 0:
        return image, mask
    image = np.clip(rotate(image, angle, axes=axes, reshape=False, mode="nearest"), 0, 1)
    mask = rotate(mask, angle, axes=axes, reshape=False, mode="nearest") > 0
    return image, mask


#: offline augmentation grid of the reference (tfds_dense_descriptor.py:463-465): 3 flips x 4 rotation angles
AUG_FLIPS = (None, "horizontal", "vertical")
AUG_ANGLES = tuple(range(0, 180, 45))


def normalize_volume(img_raw, modality, model_name):
    """Pixel normalisation of the extraction loop (:441-447): CT -> the lung window mapped to 0..1 for 'medsam', the tissue
    colour map / 255 (an RGB volume) for the other backbones; PET -> divided by its maximum."""
    from .visualization_utils import hu_to_rgb_vectorized
    if modality == "ct":
        if model_name == "medsam":
            return apply_window_ct(img_raw, width=800, level=40)
        return hu_to_rgb_vectorized(img_raw) / 255.0
    return img_raw / img_raw.max()


def extract_patient_features(model, img_raw, mask_raw, patient_id, label, dataset_name, modality, spatial_res, tqdm_text=None,
                             generate=None, device_augment=None):
    """One patient of the reference's extraction loop (tfds_dense_descriptor.py:452-488): every (flip, angle) of the offline
    augmentation grid goes through ``generate_features`` (all slices of a volume as one backbone batch here), the per-slice
    feature maps / masks are concatenated and described by the metadata table the trainer reads.

    Returns (df, all_features, all_masks); the caller writes ``df.to_parquet(df_path)`` and
    ``save_features(features_file, all_features, all_masks, patient_id)`` as the reference does (:489-490).
    Columns: feature_id (running index), slice, angle, flip, patient_id, label, dataset, modality, augmentation, spatial_res.
    ``augmentation`` is True on EVERY row, also for (flip None, angle 0): the reference evaluates ``df['flip'] is None`` on the
    Series object (:486), which is always False -- reproduced, because the trainer filters on this column."""
    import pandas as pd
    gen = generate or generate_features
    rows = {"slice": [], "angle": [], "flip": []}
    all_features, all_masks = [], []
    img_raw, mask_raw = np.asarray(img_raw), np.asarray(mask_raw)
    # The 12 whole-volume copies (3 flips x 4 angles, :463-466) are made ON THE DEVICE from one upload of the raw volume when the
    # inputs have the dtypes the kernels restate scipy for (float32 gray volume; bool or uint8 mask): csrc/augment.cu is
    # bit-identical to flip_image + rotate_image.  Other dtypes (a float64 volume keeps float64 through scipy) and custom
    # `generate` callbacks take the reference's own host functions.
    builtin = generate is None and generate_features is _BUILTIN_GENERATE_FEATURES      # (a patched module attribute is a custom callback too)
    on_device = device_augment if device_augment is not None else builtin
    on_device = on_device and builtin and img_raw.dtype == np.float32 and img_raw.ndim == 3 and mask_raw.dtype in (np.bool_, np.uint8)
    if device_augment and not on_device:
        raise ValueError("device_augment needs a float32 (H, W, S) volume, a bool / uint8 mask and the built-in generate_features")
    if on_device:
        img_dev = torch.as_tensor(np.ascontiguousarray(img_raw)).to(model.device)
        mask_kind = "mask_bool" if mask_raw.dtype == np.bool_ else "mask_u8"
        mask_dev = torch.as_tensor(np.ascontiguousarray(mask_raw).view(np.uint8)).to(model.device)
        # all 12 copies through the pipelined device loop (ROI read-backs overlap the next copy's backbone)
        aug_grid = [(f, a) for f in AUG_FLIPS for a in AUG_ANGLES]
        device_results = iter(_augmented_features_device(model, img_dev, mask_dev, mask_kind, aug_grid))
    for flip_type in AUG_FLIPS:
        if not on_device:
            image_flip, mask_flip = flip_image(img_raw, mask_raw, flip_type)
        for angle in AUG_ANGLES:
            if on_device:
                features, features_mask = next(device_results)
            else:
                image, mask = rotate_image(image_flip, mask_flip, angle)
                features, features_mask = gen(model=model, img_3d=image, mask_3d=mask,
                                              tqdm_text=tqdm_text or f"{modality} {patient_id}", display=False)
            all_masks += features_mask
            all_features += features
            rows["angle"] += [angle] * len(features)
            rows["flip"] += [flip_type] * len(features)
            rows["slice"] += list(range(0, len(features)))
    df = pd.DataFrame(rows)
    df.reset_index(drop=False, inplace=True)
    df = df.rename(columns={"index": "feature_id"})
    df["patient_id"] = patient_id
    df["label"] = label
    df["dataset"] = dataset_name.replace("_dataset", "")
    df["modality"] = modality
    df["augmentation"] = np.ones(df.shape[0], dtype=bool)          # see the docstring: the reference's expression is all-True
    df["spatial_res"] = [spatial_res] * df.shape[0]
    return df, all_features, all_masks


def _h5py():
    try:
        import h5py
        return h5py
    except ImportError as e:  # the reference's on-disk hand-off needs h5py; it is not in this image
        raise ImportError("h5py is required for the HDF5 hand-off (save_features / get_voxels)") from e


def save_features(filename, all_features, all_masks, patient_id):
    """reference: tfds_dense_descriptor.py:142-165 -- <pid>/features/<i>, <pid>/masks/<i>, lzf, one chunk."""
    h5py = _h5py()
    with h5py.File(filename, "a") as h5f:
        if patient_id in h5f:
            print(f"features for {patient_id} already exists")
            del h5f[patient_id]
        grp = h5f.create_group(patient_id)
        for i, (feature, mask) in enumerate(zip(all_features, all_masks)):
            grp.create_dataset(f"features/{i}", compression="lzf", data=feature, chunks=feature.shape)
            grp.create_dataset(f"masks/{i}", compression="lzf", data=mask, chunks=mask.shape)


def get_voxels(hdf5_path, patient_id, modality):
    """reference: tfds_dense_descriptor.py:353-362 -- (H, W, S) image + mask, isotropic 0.8 mm."""
    h5py = _h5py()
    spatial_res = np.array([0.8, 0.8, 0.8])
    with h5py.File(hdf5_path, "r") as h5f:
        idm = f"{patient_id}_{modality}"
        slices = sorted(int(k) for k in h5f[f"{idm}/img_exam"].keys())
        img = np.dstack([h5f[f"{idm}/img_exam/{k}"][()] for k in slices])
        mask = np.dstack([h5f[f"{idm}/mask_exam/{k}"][()] for k in slices])
    return img, mask, spatial_res


def build_arg_parser():
    """Same flags as the reference CLI (tfds_dense_descriptor.py:365-382)."""
    p = argparse.ArgumentParser(description="ViT patch embeddings of the lung_radiomics datasets (B200-native)")
    p.add_argument("-mn", "--model_name", type=str, default="medsam", help="medsam (the reference's default) | dinov2 | vit_s16 | vit_b16 | vit_l14")
    p.add_argument("-mp", "--model_path", type=str, default=os.path.join("models", "backbones", "medsam", "medsam_vit_b.pth"),
                   help="state-dict .pth (same default as the reference, :368-370)")
    p.add_argument("--random_init", action="store_true",
                   help="NOT in the reference: run with seeded random backbone weights instead of a checkpoint (tests / benchmarks only)")
    p.add_argument("-d", "--dataset_path", type=str, default=os.path.join("data", "lung_radiomics"))
    p.add_argument("-f", "--feature_folder", type=str, default=os.path.join("data", "features"))
    p.add_argument("-h5", "--hdf5_path", type=str, default=os.path.join("data", "lung_radiomics", "lung_radiomics_datasets_isotropic.hdf5"))
    p.add_argument("-df", "--df_path", type=str, default=os.path.join("data", "lung_radiomics", "lung_radiomics_datasets_isotropic.csv"))
    p.add_argument("-mod", "--modality", type=str, default="ct")
    return p


def process_patient(model, ds_path, patient_id, modality, label, dataset_name, df_path, features_file, voxels=None):
    """One (patient, modality) of the reference's HDF5 branch (:449-490): read the isotropic volume, run the augmentation grid
    through the backbone, write the metadata parquet and append the feature maps / masks to the modality's HDF5 file.
    ``voxels`` = (img, mask, spatial_res) replaces the HDF5 read (``get_voxels`` needs h5py)."""
    img_raw, mask_raw, spatial_res = voxels if voxels is not None else get_voxels(ds_path, patient_id, modality)
    df, all_features, all_masks = extract_patient_features(model, img_raw, mask_raw, patient_id, label, dataset_name, modality, spatial_res)
    df.to_parquet(df_path)
    save_features(features_file, all_features, all_masks, patient_id)
    return df


def main(argv=None):
    """The reference's extraction entry point for the HDF5 dataset (:364-490, ``use_tfds`` False; the tensorflow_datasets branch
    is not provided): same flags, same outputs -- ``<feature_folder>/<dataset>/<patient>_<modality>.parquet`` and
    ``<feature_folder>/features_masks_<modality>.hdf5``; patients whose parquet exists are skipped (:424)."""
    import pandas as pd
    args = build_arg_parser().parse_args(argv)
    if args.random_init:
        print("WARNING: --random_init: features come from an UNTRAINED backbone")
        model = load_model(args.model_name, None)
    else:
        model = load_model(args.model_name, args.model_path)       # a missing checkpoint is an error, never silent random weights
    modalities = ["pet", args.modality]
    meta = pd.read_csv(args.df_path)
    meta["label"] = (meta["egfr"] == "Mutant").astype(int)                               # :398
    patient2label = dict(zip(meta["patient_id"], meta["label"]))
    meta = meta[meta[f'has_{"".join(modalities)}']].reset_index(drop=True)                  # :400
    for dataset_name in ("santa_maria_dataset", "stanford_dataset"):
        features_dir = os.path.join(args.feature_folder, dataset_name)
        os.makedirs(features_dir, exist_ok=True)
        short = dataset_name.replace("_dataset", "")
        for patient_id in list(meta[meta["dataset"] == short]["patient_id"].unique()):
            for modality in modalities:
                df_path = os.path.join(features_dir, f"{patient_id}_{modality}.parquet")
                features_file = os.path.join(args.feature_folder, f"features_masks_{modality}.hdf5")
                if not os.path.exists(df_path):
                    process_patient(model, args.hdf5_path, patient_id, modality, patient2label[patient_id], dataset_name, df_path, features_file)


if __name__ == "__main__":
    main()
