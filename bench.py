#!/usr/bin/env python
"""Benchmark of the hot path: ViT dense-descriptor extraction + tumour-mask gather.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config C2]

A "step" is one pass of the path over one synthetic CT volume (BASELINE.json configs[1]:
ViT-B/16 over a 512x512x120 volume, mask-gathered point cloud): the batched backbone forward of
all slices + the stream-compaction gather.  Metric = CT slices per second (whole job, all GPUs).
  value : inputs already resident in HBM when the timed region starts
  e2e   : same metric through the public call `tfds_dense_descriptor.extract_point_cloud` with HOST
          (pinned) buffers: H2D of the volume + mask and D2H of the point cloud inside the timed region
One JSON line is printed by rank 0.  Under torchrun each rank processes its own volume (patients are
independent: weak scaling, no data-path collective); time = max over ranks of device time.
`--impl reference` times the CPU implementation of the same path on the host cores (the oracle port:
the reference's backbone lives in un-vendored third-party code and hard-codes .cuda()).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# ----------------------------------------------------------------------------------------- helpers
class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return dict(bf16=p.get("bf16_tflops_sustained", p.get("bf16_tflops")), burst=p.get("bf16_tflops"),
                    hbm=p.get("hbm_gbs"), source="MEASURED_PEAKS.json (sustained bf16: kernel timed inside a long step)")
    return dict(bf16=1400.0, burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


def cpu_baseline(case: str, sample_slices: int, threads: int | None = None):
    """Oracle port of the path on the host cores: fp32 ViT forward (oracle/vit_fp32.py) over a bounded
    sample of slices of the SAME synthetic volume + the NumPy gather on those slices."""
    from oracle import gather_np, vit_fp32
    from vit_deep_radiomics_b200 import synth
    from vit_deep_radiomics_b200.visualization_utils import crop_window, roi_window
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    img, mask, res, model_name = synth.make_case(case)
    cfg = vit_fp32.VIT_CONFIGS[model_name]
    H, W, S = img.shape
    s0 = max(0, S // 2 - sample_slices // 2)
    sl = slice(s0, s0 + sample_slices)
    w = vit_fp32.init_weights(cfg, (H, W), seed=1234)
    x = torch.from_numpy(np.ascontiguousarray(np.moveaxis(img[:, :, sl], -1, 0)))[:, None].expand(-1, 3, -1, -1).contiguous()
    t0 = time.perf_counter()
    with torch.no_grad():
        dense = vit_fp32.vit_forward(w, cfg, x).numpy()
    bigger = mask.sum(-1) > 0
    gh, gw = H // cfg["patch"], W // cfg["patch"]
    fx0, fy0, fx1, fy1 = roi_window((gh, gw), bigger, 1)
    mx0, my0, mx1, my1 = roi_window((H, W), bigger, 1)
    feats = [dense[i, fy0:fy1, fx0:fx1] for i in range(dense.shape[0])]
    masks = [mask[my0:my1, mx0:mx1, s0 + i] for i in range(dense.shape[0])]
    out = gather_np.token_gather(feats, masks, res)
    dt = time.perf_counter() - t0
    return dict(value=sample_slices / dt, unit="slices/s", cores=cores, kind="port",
                sample=f"{sample_slices} of {S} slices of the {case} volume ({model_name}, {H}x{W}): fp32 torch ViT forward "
                       f"+ NumPy mask gather ({out['flat'].size} tokens), {dt:.2f} s wall"), dt


# ----------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.sample_slices
    vals = []
    for i in range(args.warmup + args.steps):
        cb, dt = cpu_baseline(args.config, sample)
        if i >= args.warmup:
            vals.append(dt)
    v = sample * len(vals) / sum(vals)
    cb["value"] = v
    from vit_deep_radiomics_b200 import synth
    c = synth.CONFIGS[args.config]
    print(json.dumps({
        "impl": "reference", "metric": "CT slices/sec ViT dense-descriptor extraction + mask gather", "value": v,
        "unit": "slices/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(vals) / len(vals), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: {c['model']} dense descriptors over a synthetic "
                               f"{c['shape'][0]}x{c['shape'][1]}x{c['shape'][2]} CT volume with mask-gathered point cloud",
                   "step": f"bounded sample: {sample} slices per step on the host CPU"},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ----------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch.distributed as dist
    from vit_deep_radiomics_b200 import _C, ops, synth, tfds_dense_descriptor as tdd
    from vit_deep_radiomics_b200.distributed import init_distributed
    rank, world = init_distributed("nccl")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    img, mask, res, model_name = synth.make_case(args.config, seed=1235 + rank)
    H, W, S = img.shape
    model = tdd.load_model(model_name, img_hw=(H, W), device=dev, seed=1234)
    plan = tdd._plan(model, mask)
    gh, gw = model.grid

    # resident inputs for `value`
    img_dev = torch.as_tensor(img).to(dev)
    mask_s = torch.as_tensor(np.ascontiguousarray(np.moveaxis(mask, -1, 0)).view(np.uint8)).to(dev)
    pe = dict(res=res, noise=(0.0, 0.0, 0.0), scale=0.25)

    def step_resident():
        tok = tdd._forward_volume(model, img_dev, plan)
        return ops.mask_gather(tok, mask_s, grid=(S, gh, gw, model.n_tokens, 1), feat_roi=plan["feat_roi"],
                               mask_roi=plan["mask_roi"], pe=pe)

    # pinned host inputs for `e2e`
    img_pin = torch.as_tensor(img).pin_memory()
    mask_pin = torch.as_tensor(np.ascontiguousarray(mask).view(np.uint8)).pin_memory()

    extractor = tdd.PointCloudExtractor(model)

    def run_e2e(k):
        """k patients through the public streaming API: H2D of patient i+1 overlaps the backbone of patient i;
        every patient's point cloud is read back to the host."""
        last = None
        for out in extractor.run([(img_pin, mask_pin, res)] * k):
            last = out
        return last

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, profile=False):
        barrier()
        if profile:
            ops.PROFILE = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        prof, ops.PROFILE = ops.PROFILE, None
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out, prof

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _C.launch_count()
    ms, out, _ = timed(step_resident, args.steps)
    launches = _C.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    n_tokens = int(out[2].item())
    value = world * S * args.steps / (ms / 1e3)

    # per-kernel profile pass (CUDA events around every GEMM launch, same stream), separate from `value`; it runs the
    # op-by-op path (the native vdr_vit_forward call cannot be instrumented from here): one untimed step first, so that
    # its activation buffers exist
    ops.PROFILE = []
    step_resident()
    ops.PROFILE = None
    ms_p, _, prof = timed(step_resident, max(1, min(args.steps, 3)), profile=True)
    torch.cuda.synchronize()
    gemm_ms = sum(a.elapsed_time(b) for (kind, fl, a, b, _) in prof if kind == "gemm")
    gemm_fl = sum(fl for (kind, fl, a, b, _) in prof if kind == "gemm")
    attn_ms = sum(a.elapsed_time(b) for (kind, fl, a, b, _) in prof if kind == "attn")
    attn_fl = sum(fl for (kind, fl, a, b, _) in prof if kind == "attn")
    n_gemm = sum(1 for p in prof if p[0] == "gemm")
    by_label = {}
    for (kind, fl, a, b, label) in prof:
        t = by_label.setdefault(label, [0, 0.0, 0.0])
        t[0] += 1
        t[1] += a.elapsed_time(b)
        t[2] += fl
    kinds = {label: kind for (kind, fl, a, b, label) in prof}
    per_kernel = {}
    for k, v in sorted(by_label.items(), key=lambda kv: -kv[1][1]):
        e = {"launches": v[0], "ms_per_launch": v[1] / v[0]}
        if kinds[k] == "ln":      # HBM-bound: algorithmic bytes / time
            e["gbs"] = v[2] / (v[1] / 1e3) / 1e9
        else:
            e["tflops"] = v[2] / (v[1] / 1e3) / 1e12
        per_kernel[k] = e

    # The HBM-bound kernels of the path, timed alone (CUDA events, 10 launches each after 3 warm-ups):
    #   mask gather of this volume (C2: 5 k tokens out of a 14 k-candidate ROI -> launch / host-latency bound, reported as is)
    #   and of a dense mask over the whole token grid (every token selected: the bandwidth of the compaction + emit kernels);
    #   algorithmic bytes per SURVEY 8d: S*h*w mask bytes + n_sel * (D*4 read + D*4 written + 12).
    def time_gather(mask_dev, roi, reps=10):
        tok = model._workspace(S)["OUT"]
        froi, mroi = (plan["feat_roi"], plan["mask_roi"]) if roi else (None, None)
        call = lambda: ops.mask_gather(tok, mask_dev, grid=(S, gh, gw, model.n_tokens, 1), feat_roi=froi,  # noqa: E731
                                       mask_roi=mroi, pe=pe)
        for _ in range(3):
            out_g = call()
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(reps):
            out_g = call()
        g1.record()
        torch.cuda.synchronize()
        n_sel = int(out_g[2].item())
        r0, r1, c0, c1 = plan["feat_roi"] if roi else (0, gh, 0, gw)
        cand = S * (r1 - r0) * (c1 - c0)
        nbytes = cand + n_sel * (model.cfg["dim"] * 8 + 12)
        ms_g = g0.elapsed_time(g1) / reps
        return {"ms": ms_g, "selected": n_sel, "candidates": cand, "algorithmic_bytes": nbytes, "gbs": nbytes / (ms_g / 1e3) / 1e9}

    gather_c2 = time_gather(mask_s, roi=True)                        # the workload's own gather (ROI of the tumour, ~35 % selected)
    gather_dense = time_gather(torch.ones_like(mask_s), roi=False)   # whole 32x32xS grid, every token selected

    run_e2e(2)
    ms_e, out_e, _ = timed(lambda: run_e2e(args.steps), 1)
    e2e_value = world * S * args.steps / (ms_e / 1e3)
    h2d = img_pin.numel() * 4 + mask_pin.numel()
    d2h = out_e["count"] * (model.cfg["dim"] * 4 + 12) + 4

    if rank != 0:
        return
    medsam = medsam_side_measurement(dev) if args.medsam else None
    peaks = measured_peaks()
    achieved = gemm_fl / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.isfile(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    cb, _ = cpu_baseline(args.config, args.cpu_sample_slices)     # ~10 s of host work on rank 0
    c = synth.CONFIGS[args.config]
    flops_step = model.flops_per_slice() * S
    print(json.dumps({
        "metric": "CT slices/sec ViT dense-descriptor extraction + mask gather", "value": value, "unit": "slices/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{args.config}: {c['model']} dense descriptors over a synthetic {H}x{W}x{S} CT volume with "
                               f"mask-gathered point cloud ({n_tokens} tokens), one volume per GPU per step",
                   "l2": "inputs+activations per step (>1.5 GB) exceed the 126 MB L2; no explicit flush",
                   "weights": "seeded random init (no checkpoints offline)"},
        "model_tflops": flops_step * world * args.steps / (ms / 1e3) / 1e12,
        "e2e": {"value": e2e_value, "unit": "slices/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": ms_e / args.steps, "api": "tfds_dense_descriptor.PointCloudExtractor.run (pinned host buffers, uploads double-buffered on a copy stream)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05_kernel", "achieved": achieved, "peak": peaks["bf16"],
                     "unit": "TFLOP/s", "frac": (achieved / peaks["bf16"]) if achieved else None, "traffic": traffic,
                     "peak_source": peaks["source"], "launches_timed": n_gemm,
                     "share_of_step": gemm_ms / ms_p if ms_p else None,
                     "attention": {"achieved": attn_fl / (attn_ms / 1e3) / 1e12 if attn_ms else None,
                                   "share_of_step": attn_ms / ms_p if ms_p else None},
                     "per_kernel": per_kernel,
                     "hbm_kernels": {"peak_gbs": peaks["hbm"], "mask_gather_c2": gather_c2, "mask_gather_dense": gather_dense,
                                     "mask_gather_dense_frac": gather_dense["gbs"] / peaks["hbm"] if peaks["hbm"] else None}},
        "cpu_baseline": cb,
        "clocks": clocks,
        "n1_medsam": medsam}))


def medsam_side_measurement(dev, B=8, steps=3):
    """Not part of `value`: the reference's own default backbone (load_model('medsam'): SAM ViT-B image encoder, 1024 x 1024
    inputs, (64, 64, 256) descriptors) through the same libvdr kernels, B resident gray slices per pass."""
    from vit_deep_radiomics_b200 import _C, tfds_dense_descriptor as tdd
    model = tdd.load_model("medsam", device=dev, seed=1)
    x = torch.rand(B, 1024, 1024, device=dev)
    strides = (x.stride(0), 0, x.stride(1), x.stride(2))
    for _ in range(3):
        model.forward_tokens(x, strides, B)
    torch.cuda.synchronize()
    n0 = _C.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        model.forward_tokens(x, strides, B)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"metric": "MedSAM (SAM ViT-B) encoder slices/s, 1024x1024 gray slices resident, 1 GPU", "value": B / ms * 1e3, "batch": B,
            "ms_per_batch": ms, "model_tflops": model.flops_per_slice() * B / ms / 1e9, "gpu_launches_per_batch": (_C.launch_count() - n0) // steps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=str, default="C2")
    ap.add_argument("--sample-slices", type=int, default=8, dest="sample_slices")
    ap.add_argument("--cpu-sample-slices", type=int, default=32, dest="cpu_sample_slices",
                    help="slices of the workload the cpu_baseline leg of our arm runs on the host cores (~0.3 s per slice)")
    ap.add_argument("--no-medsam", action="store_false", dest="medsam", help="skip the MedSAM side measurement (key n1_medsam)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
