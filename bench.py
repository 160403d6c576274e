#!/usr/bin/env python
"""Benchmark of the hot path: ViT dense-descriptor extraction + tumour-mask gather (+ the table all-gather at N > 1).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config C2] [--no-sub]

A "step" is one pass of the path over one synthetic CT volume per GPU (BASELINE.json configs[1]: ViT-B/16 over a
512x512x120 volume, mask-gathered point cloud): the batched backbone forward of all slices + the stream-compaction gather.
Metric = CT slices per second (whole job, all GPUs).
  value : inputs already resident in HBM when the timed region starts
  e2e   : same metric through the public API with HOST (pinned) buffers: H2D of the volumes + masks and D2H of the
          point clouds inside the timed region
N > 1 (torchrun, one rank per GPU): the patients are sharded over the ranks and every step ALSO assembles the point-cloud
table on every rank INSIDE the timed region (distributed.PointCloudTable: row counts -> one small all-gather -> device scan ->
each rank's gather kernel writes at its row offset of the table -> one in-place NCCL all-gather of the row ranges), the
exchange step of SURVEY.md 8(e).  Time = max over ranks of the device time.
Sub-records (key "sub", each with its own roofline / cpu_baseline / e2e): C1 (ViT-S/16 224^2 x 8), C4 (ViT-L/14, patients
sharded + table all-gather), C5 (extraction -> gather -> classifier training step, patients/s), c3 (classifier data-parallel
training, samples/s, gradient all-reduce), medsam (the reference's default backbone).
`--impl reference` times the CPU implementation of the same path on the host cores (the oracle port: the reference's
backbone lives in un-vendored third-party code and hard-codes .cuda()).
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

NOMINAL_BF16_TFLOPS = 2250.0     # B200 dense bf16 (the north star's nominal denominator)
NOMINAL_HBM_GBS = 8000.0
NVLINK_PEER_GBS = 770.0          # measured peer copy per direction on this pool (B200_PROFILING.md); nominal 900


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


def tensor_fracs(tflops, peaks):
    """Fraction of the measured sustained / measured burst / nominal dense bf16 peak (BASELINE.md section 2 asks for all)."""
    if not tflops:
        return {"frac": None, "frac_burst": None, "frac_nominal": None}
    return {"frac": tflops / peaks["bf16"], "frac_burst": tflops / peaks["burst"], "frac_nominal": tflops / NOMINAL_BF16_TFLOPS}


class Dist:
    """rank / world + the barrier and max-over-ranks the timing contract asks for."""

    def __init__(self):
        from vit_deep_radiomics_b200.distributed import init_distributed
        self.rank, self.world = init_distributed("nccl")
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device(f"cuda:{self.local}")

    def barrier(self):
        if self.world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def max_ms(self, ms):
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            return float(t.item())
        return ms

    def timed(self, fn, steps):
        """barrier + synchronize, `steps` calls between two CUDA events on the launching stream, barrier + synchronize;
        returns (max over ranks of the device ms, last result)."""
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        self.barrier()
        return self.max_ms(e0.elapsed_time(e1)), out


# ----------------------------------------------------------------------------------------- CPU legs (oracle port)
def cpu_extraction(case: str, sample_slices: int, weights=None, threads: int | None = None, seed: int = 1235):
    """Oracle port of the path on the host cores: fp32 ViT forward (oracle/vit_fp32.py) over a bounded sample of slices of the
    SAME synthetic volume + the NumPy gather on those slices.  Returns (cpu_baseline dict, seconds, dense descriptors, slice range)."""
    from oracle import gather_np, vit_fp32
    from vit_deep_radiomics_b200 import synth
    from vit_deep_radiomics_b200.visualization_utils import roi_window
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    img, mask, res, model_name = synth.make_case(case, seed=seed)
    cfg = vit_fp32.VIT_CONFIGS[model_name]
    H, W, S = img.shape
    sample_slices = min(sample_slices, S)
    s0 = max(0, S // 2 - sample_slices // 2)
    sl = slice(s0, s0 + sample_slices)
    w = weights if weights is not None else vit_fp32.init_weights(cfg, (H, W), seed=1234)
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
    cb = dict(value=sample_slices / dt, unit="slices/s", cores=cores, kind="port",
              sample=f"{sample_slices} of {S} slices of the {case} volume ({model_name}, {H}x{W}): fp32 torch ViT forward (oracle/vit_fp32.py) "
                     f"+ NumPy mask gather (oracle/gather_np.py restatement of _get_features; the reference sources are not on the GPU box), "
                     f"{out['flat'].size} tokens, {dt:.2f} s wall")
    return cb, dt, dense, (s0, sample_slices)


# ----------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.sample_slices
    vals, cb = [], None
    for i in range(args.warmup + args.steps):
        cb, dt, _, _ = cpu_extraction(args.config, sample)
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


# ----------------------------------------------------------------------------------------- extraction (C1 / C2 / C4)
def bench_extraction(D: Dist, case: str, steps: int, warmup: int, *, profile: bool, cpu_slices: int, clocks: bool = False,
                     patients_per_step: int = 1):
    """One extraction config: `patients_per_step` synthetic volumes of `case` per rank per step.  Returns the record (rank 0
    keeps it; all ranks run it: at N > 1 it contains collectives)."""
    from vit_deep_radiomics_b200 import _C, ops, synth, tfds_dense_descriptor as tdd
    from vit_deep_radiomics_b200.distributed import PointCloudTable
    rank, world, dev = D.rank, D.world, D.dev
    P = patients_per_step
    n_pat = world * P                                    # patients per step over the whole job; rank r owns [r*P, (r+1)*P)
    vols = [synth.make_case(case, seed=1235 + rank * P + i) for i in range(P)]
    img, mask, res, model_name = vols[0]
    H, W, S = img.shape
    model = tdd.load_model(model_name, img_hw=(H, W), device=dev, seed=1234)
    ex = tdd.PointCloudExtractor(model)
    gh, gw = model.grid
    dim = model.cfg["dim"]
    peaks = measured_peaks()

    # resident inputs for `value`
    img_dev = [torch.as_tensor(v[0]).to(dev) for v in vols]
    mask_dev = [torch.as_tensor(np.ascontiguousarray(v[1]).view(np.uint8)).to(dev) for v in vols]
    plans = [ex.plan_patient(m) for m in mask_dev]
    geos = [dict(grid=(S, gh, gw, model.n_tokens, model.token_offset), feat_roi=p["feat_roi"],
                 mask_roi=tdd._shift_roi(p["mask_roi"], p["crop"]), mask_layout="hws") for p in plans]
    pe = dict(res=res, noise=(0.0, 0.0, 0.0), scale=0.25)
    cand = max(S * (p["feat_roi"][1] - p["feat_roi"][0]) * (p["feat_roi"][3] - p["feat_roi"][2]) for p in plans)
    table = PointCloudTable(n_pat, dim, cap_rows=n_pat * cand, device=dev, rank=rank, world=world)

    staged = ex.stage_table([(rank * P + i, img_dev[i], mask_dev[i], res) for i in range(P)], table, count=False)   # plans: once

    def step_resident():
        """counts -> (all-gather of counts, device scan) -> backbone (patients that fit share a batch) + per-patient gather at its
        table offset -> table all-gather: PointCloudExtractor.count_table + emit_table, the steps of run_table after planning"""
        ex.count_table(staged, table)
        return ex.emit_table(staged, table)

    # pinned host inputs for `e2e`
    img_pin = [torch.as_tensor(v[0]).pin_memory() for v in vols]
    mask_pin = [torch.as_tensor(np.ascontiguousarray(v[1]).view(np.uint8)).pin_memory() for v in vols]
    host_tok = host_src = None

    def run_e2e(k):
        """k steps through the public API from pinned host buffers.  N = 1: the streaming extractor (upload of patient i+1
        overlaps the backbone of patient i), every point cloud read back.  N > 1 (or several small patients per step): the extraction into one table
        (PointCloudExtractor.run_table): k*P patients per rank, one count exchange, one table all-gather, every rank reads its
        own row range back (the table itself stays on every GPU for the trainer)."""
        nonlocal host_tok, host_src
        if world == 1 and P == 1:
            last = None
            for out in ex.run([(img_pin[i % P], mask_pin[i % P], res) for i in range(k * P)]):
                last = out
            return last["count"] * P
        tb = PointCloudTable(world * k * P, dim, cap_rows=world * k * P * cand, device=dev, rank=rank, world=world)
        items = [(rank * k * P + j, img_pin[j % P], mask_pin[j % P], res) for j in range(k * P)]
        total = ex.run_table(items, tb)
        lo, hi = tb.rank_ranges()[rank]
        if host_tok is None or host_tok.shape[0] < hi - lo:
            host_tok = torch.empty((hi - lo, dim), dtype=torch.float32).pin_memory()
            host_src = torch.empty((hi - lo, 4), dtype=torch.int32).pin_memory()
        host_tok[:hi - lo].copy_(tb.tokens[lo:hi], non_blocking=True)
        host_src[:hi - lo].copy_(tb.src[lo:hi], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return (hi - lo) // k

    for _ in range(max(warmup, 3)):
        step_resident()
    sampler = ClockSampler(D.local) if clocks and rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = _C.launch_count()
    ms, total_rows = D.timed(step_resident, steps)
    launches = _C.launch_count() - launches0
    clk = sampler.stop() if sampler else None
    n_tokens = int(table.rank_ranges()[rank][1] - table.rank_ranges()[rank][0]) // P
    value = n_pat * S * steps / (ms / 1e3)
    flops_step = model.flops_per_slice() * S * P
    rec = {"workload": f"{case}: {model_name} dense descriptors over synthetic {H}x{W}x{S} CT volumes with mask-gathered point cloud "
                       f"({n_tokens} tokens per volume), {P} volume(s) per GPU per step" +
                       (f", patients sharded over {world} ranks + in-place NCCL all-gather of the point-cloud table every step" if world > 1 else ""),
           "value": value, "unit": "slices/s", "patients_per_s": n_pat * steps / (ms / 1e3), "ms_per_step": ms / steps,
           "model_tflops": flops_step * world * steps / (ms / 1e3) / 1e12, "gpu_launches": int(launches), "clocks": clk}
    rec.update({("model_" + k): v for k, v in tensor_fracs(rec["model_tflops"] / world, peaks).items()})

    # ---- the exchange step alone (N > 1): events around table.all_gather() after a barrier, max over ranks; bit-identity check
    if world > 1:
        ms_c, _ = D.timed(table.all_gather, 5)
        row_bytes = dim * 4 + 16
        tot_bytes = table.total * row_bytes
        rec["collective"] = {"kind": "in-place variable-length all-gather of the point-cloud table (ncclAllGather when the row ranges are equal, "
                                     "else grouped ncclBroadcast per rank) + all-gather of the int64 row counts",
                             "bytes_per_step": int(tot_bytes), "rows": int(table.total), "ms": ms_c / 5,
                             "busbw_gbs": tot_bytes * (world - 1) / world / (ms_c / 5 / 1e3) / 1e9,
                             "nvlink_peer_gbs_measured": NVLINK_PEER_GBS,
                             "share_of_step": (ms_c / 5) / (ms / steps),
                             "limiting": "launch/latency-bound at this size (a few MB per rank): the table rows are ~0.5 % of a step's HBM traffic"}
        # rank 0 recomputes patient P (rank 1's first volume) alone on its GPU through the 1-GPU path: the gathered rows must be bit-identical
        ok = None
        if rank == 0:
            v1 = synth.make_case(case, seed=1235 + 1 * P)
            o1 = tdd.extract_point_cloud(model, v1[0], v1[1], v1[2], to_host=False)
            n1 = int(o1["count"].item())
            lo = int(table.offsets[table._slot_index(P)].item())
            ok = bool(n1 == int(table.counts_host[table._slot_index(P)]) and torch.equal(table.tokens[lo:lo + n1], o1["tokens"][:n1])
                      and torch.equal(table.src[lo:lo + n1, 1:], o1["src"][:n1]) and bool((table.src[lo:lo + n1, 0] == P).all()))
        rec["collective"]["table_bit_identical_to_1gpu"] = ok

    # ---- per-kernel profile pass (CUDA events around every GEMM / attention launch, same stream), separate from `value`; it runs
    # the op-by-op path (the native vdr_vit_forward call cannot be instrumented from here)
    if profile:
        def prof_forward():      # the backbone batch exactly as a step issues it (several small patients share one batch)
            if P > 1:
                return model.forward_volumes([(img_dev[i], plans[i]["crop"]) for i in range(P)])
            return tdd._forward_volume(model, img_dev[0], plans[0])

        ops.PROFILE = []
        prof_forward()                                              # untimed: the op-by-op activation buffers get allocated
        ops.PROFILE = None
        torch.cuda.synchronize()
        ops.PROFILE = []
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        reps = max(1, min(steps, 3))
        for _ in range(reps):
            prof_forward()
        p1.record()
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
        ms_p = p0.elapsed_time(p1)
        tot = {}
        for kind, fl, a, b, label in prof:
            t = tot.setdefault(kind, [0.0, 0.0, 0])
            t[0] += a.elapsed_time(b)
            t[1] += fl
            t[2] += 1
        by_label = {}
        for kind, fl, a, b, label in prof:
            t = by_label.setdefault(label, [0, 0.0, 0.0, kind])
            t[0] += 1
            t[1] += a.elapsed_time(b)
            t[2] += fl
        per_kernel = {}
        for k, v in sorted(by_label.items(), key=lambda kv: -kv[1][1]):
            e = {"launches": v[0], "ms_per_launch": v[1] / v[0]}
            if v[3] == "ln":                                       # HBM-bound: algorithmic bytes / time
                e["gbs"] = v[2] / (v[1] / 1e3) / 1e9
            else:
                e["tflops"] = v[2] / (v[1] / 1e3) / 1e12
            per_kernel[k] = e
        g_ms, g_fl, g_n = tot.get("gemm", [0.0, 0.0, 0])
        a_ms, a_fl, _ = tot.get("attn", [0.0, 0.0, 0])
        achieved = g_fl / (g_ms / 1e3) / 1e12 if g_ms else None
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "gemm_traffic.json")
        if case == "C2" and os.path.isfile(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        attn = a_fl / (a_ms / 1e3) / 1e12 if a_ms else None
        rec["roofline"] = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel", "achieved": achieved, "peak": peaks["bf16"], "unit": "TFLOP/s",
                           **tensor_fracs(achieved, peaks), "peak_burst": peaks["burst"], "peak_nominal": NOMINAL_BF16_TFLOPS,
                           "traffic": traffic,
                           "traffic_source": "dram__bytes_read+write per launch from the committed ncu --set full capture (profiles/gemm_traffic.json); "
                                             "a constant of that capture, NOT measured in this run" if traffic else None,
                           "peak_source": peaks["source"], "launches_timed": g_n, "share_of_step": g_ms / ms_p if ms_p else None,
                           "attention": {"achieved": attn, **tensor_fracs(attn, peaks), "share_of_step": a_ms / ms_p if ms_p else None},
                           "per_kernel": per_kernel}

    # ---- the HBM-bound kernel of the path, the mask gather, timed alone: CUDA events around 10 launches after 3 warm-ups.
    # algorithmic bytes per SURVEY 8d: S*h*w mask bytes + n_sel * (D*4 read + D*4 written + 12).
    def time_gather(mask_t, geo, reps=10):
        """Device time of the gather alone: per repetition an L2 flush (1 GiB memset: also keeps the GPU busy long enough for the host
        to enqueue the next launch, so no launch gap is measured), then CUDA events right around the one cooperative launch."""
        tok = model._workspace(S)["OUT"]
        flush = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
        call = lambda: ops.mask_gather(tok, mask_t, pe=pe, **geo)          # noqa: E731
        for _ in range(3):
            out_g = call()
        torch.cuda.synchronize()
        evs = []
        for _ in range(reps):
            flush.zero_()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            out_g = call()
            g1.record()
            evs.append((g0, g1))
        torch.cuda.synchronize()
        n_sel = int(out_g[2].item())
        r0, r1, c0, c1 = geo["feat_roi"] if geo["feat_roi"] is not None else (0, gh, 0, gw)
        cand_g = S * (r1 - r0) * (c1 - c0)
        nbytes = cand_g + n_sel * (dim * 8 + 12)
        ms_g = sum(a.elapsed_time(b) for a, b in evs) / reps
        del flush
        return {"ms": ms_g, "selected": n_sel, "candidates": cand_g, "algorithmic_bytes": nbytes, "gbs": nbytes / (ms_g / 1e3) / 1e9,
                "frac_measured": nbytes / (ms_g / 1e3) / 1e9 / peaks["hbm"], "frac_nominal": nbytes / (ms_g / 1e3) / 1e9 / NOMINAL_HBM_GBS,
                "launches": 1, "timing": "CUDA events around each launch, L2 flushed before every repetition"}

    if profile:
        dense_geo = dict(grid=geos[0]["grid"], feat_roi=None, mask_roi=None, mask_layout="hws")
        rec["roofline"]["hbm_kernels"] = {
            "kernel": "g1_fused_kernel (one cooperative launch: PE table + predicate/ballots + scan + ranks + emit)", "peak_gbs": peaks["hbm"],
            "peak_nominal_gbs": NOMINAL_HBM_GBS,
            "mask_gather_own": time_gather(mask_dev[0], geos[0]),                      # the workload's own gather (ROI of the tumour)
            "mask_gather_dense": time_gather(torch.ones_like(mask_dev[0]), dense_geo)}  # whole token grid, every token selected

    # ---- e2e
    run_e2e(2)
    ms_e, per_step_rows = D.timed(lambda: run_e2e(steps), 1)
    rec["e2e"] = {"value": n_pat * S * steps / (ms_e / 1e3), "unit": "slices/s",
                  "h2d_bytes_per_step": int(P * (img_pin[0].numel() * 4 + mask_pin[0].numel())),
                  "d2h_bytes_per_step": int(per_step_rows * (dim * 4 + (12 if (world == 1 and P == 1) else 16)) + 4), "ms_per_step": ms_e / steps,
                  "api": "tfds_dense_descriptor.PointCloudExtractor.run (pinned host buffers, uploads double-buffered on a copy stream, each point cloud read back on a third stream while the next patient is in the backbone)" if (world == 1 and P == 1) else
                         "tfds_dense_descriptor.PointCloudExtractor.run_table into a distributed.PointCloudTable (pinned host buffers; counts all-gather, "
                         "table all-gather and the read-back of each rank's row range inside the timed region)"}

    # ---- CPU leg + parity of the device descriptors against the fp32 oracle on the CPU leg's slices (rank 0)
    if rank == 0:
        cb, _, dense, (s0, ns) = cpu_extraction(case, cpu_slices, weights=model.state_dict_f32)
        rec["cpu_baseline"] = cb
        # (through forward_volume = the timed path: slice staging -> TMA im2col patch embedding on channel-summed weights -> native forward)
        k = min(ns, 2)
        sub_vol = torch.from_numpy(np.ascontiguousarray(np.asarray(vols[0][0][:, :, s0:s0 + k], dtype=np.float32))).to(dev)
        gh_, gw_ = model.grid
        tok = model.forward_volume(sub_vol, (0, H, 0, W))
        got = tok.view(k, model.n_tokens, dim)[:, model.token_offset:, :].reshape(k, gh_, gw_, dim).cpu().numpy().astype(np.float64)
        want = dense[:got.shape[0]].astype(np.float64)
        g2, w2 = got.reshape(-1, dim), want.reshape(-1, dim)
        cos = (g2 * w2).sum(1) / (np.linalg.norm(g2, axis=1) * np.linalg.norm(w2, axis=1))
        rec["parity"] = {"what": f"bf16 device descriptors (forward_volume, the timed path) vs the fp32 oracle, {got.shape[0]} slices of this workload ({model_name}@{H}x{W})",
                         "max_abs": float(np.abs(got - want).max()), "rms_rel": float(np.sqrt(((got - want) ** 2).sum() / (want ** 2).sum())),
                         "min_cosine": float(cos.min()), "gather_indices": "bit-exact (tests/test_gpu_gather.py, test_gpu_pipeline.py)"}
    return rec


# ----------------------------------------------------------------------------------------- classifier (c3) and pipeline (C5)
def _classifier_flops(n, d, ff, layers):
    """F_fwd = L*(8 n d^2 + 4 n d ff + 4 n^2 d) + 4 d^2 + 8 d with n including the CLS token."""
    return layers * (8.0 * n * d * d + 4.0 * n * d * ff + 4.0 * n * n * d) + 4.0 * d * d + 8.0 * d


def bench_classifier(D: Dist, samples: int = 64, epochs: int = 2, cpu_samples: int = 2):
    """C3: point-cloud transformer training (fwd + focal loss + bwd, AdamW every 32-sample virtual batch split over the ranks,
    ONE flat-bucket gradient all-reduce per optimizer step) on synthetic clouds; samples/s over all ranks."""
    from vit_deep_radiomics_b200 import _C, synth
    from vit_deep_radiomics_b200.distributed import allreduce_grads, grad_bucket, zero_grads
    from vit_deep_radiomics_b200.models_archs import TransformerNoduleClassifier
    from vit_deep_radiomics_b200.train_models import FocalLoss
    rank, world, dev = D.rank, D.world, D.dev
    ids, labels, sizes, cloud = synth.point_cloud_patients(samples, d=256, n_range=(512, 4096), seed=1236)
    torch.manual_seed(0)
    model = TransformerNoduleClassifier(256, 1024, 4, 2, 2).to(dev)
    if world > 1:
        with torch.no_grad():
            for p in model.parameters():
                torch.distributed.broadcast(p, 0)
    opt = torch.optim.AdamW(model.parameters(), lr=5e-4, weight_decay=0.01)
    crit = FocalLoss(alpha=torch.tensor([0.25, 0.75], device=dev), gamma=2)
    mine = list(range(rank, samples, world))
    data = [(torch.from_numpy(cloud(i)).to(dev), torch.eye(2, device=dev)[int(labels[i])]) for i in mine]
    window = max(1, 32 // world)
    bucket = grad_bucket(model)
    from vit_deep_radiomics_b200.graph_step import graphed_step
    step = graphed_step(model, crit)          # forward + loss + backward of a sample as one CUDA graph per cloud length

    def epoch():
        zero_grads(model, opt)
        tot = torch.zeros((), device=dev)
        for k, (x, y) in enumerate(data):
            loss, _ = step(x, y, 1.0 / 32)                                   # loss / iters_to_accumulate, train_models.py:674
            tot += loss
            if (k + 1) % window == 0 or k + 1 == len(data):                  # :685
                allreduce_grads(model)
                opt.step()
                zero_grads(model, opt)
        return tot

    l0 = float(epoch()) * 32 / len(data)      # first visit of every length: eager
    epoch()                                    # second visit: captured; from here on every sample is one graph launch
    n0 = _C.launch_count()
    ms, tot = D.timed(epoch, epochs)
    launches = _C.launch_count() - n0
    l1 = float(tot) * 32 / len(data)
    rec = {"workload": f"C3: TransformerNoduleClassifier(256, ff 1024, 4 heads, 2 layers) training on {samples} synthetic point clouds of 512..4096 tokens, "
                       f"virtual batch 32 split over {world} rank(s), AdamW, focal loss",
           "value": samples * epochs / (ms / 1e3), "unit": "samples/s", "ms_per_sample_per_gpu": ms / (epochs * len(data)),
           "gpu_launches": int(launches), "loss_first_epoch": l0, "loss_last_epoch": l1,
           "cuda_graphs": {"lengths_captured": len(step.graphs), "replays": step.replays, "eager_steps": step.eager, "pool_bytes": int(step.bytes),
                           "note": "gpu_launches counts the library's launch calls; replayed graphs run the captured kernels without them"}}
    flops = sum(3.0 * _classifier_flops(int(sizes[i]) + 1, 256, 1024, 2) for i in range(samples)) * epochs
    peaks = measured_peaks()
    tfl = flops / (ms / 1e3) / 1e12
    rec["roofline"] = {"bound": "tensor", "achieved": tfl, "unit": "TFLOP/s", "peak": peaks["bf16"], **tensor_fracs(tfl / world, peaks),
                       "note": "fwd+bwd algorithmic flops (3 x forward) over the whole step incl. optimizer; batch-1 sequences of <= 4k tokens are launch- "
                               "and latency-bound, not tensor-bound"}
    if world > 1:
        ms_c, _ = D.timed(bucket.allreduce, 10)
        nbytes = bucket.flat.numel() * 4
        rec["collective"] = {"kind": "ncclAllReduce(sum) of the persistent flat fp32 gradient bucket (param.grad tensors are views of it)",
                             "bytes_per_step": int(nbytes), "ms": ms_c / 10, "busbw_gbs": nbytes * 2 * (world - 1) / world / (ms_c / 10 / 1e3) / 1e9,
                             "per_optimizer_step": True, "limiting": "latency-bound (6.85 MB)"}
        bucket.zero()
    if rank == 0:
        from oracle import classifier_fp32 as C
        torch.set_num_threads(os.cpu_count() or 1)
        sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
        t1 = time.perf_counter()
        for i in range(cpu_samples):
            lg, _ = C.classifier_forward(sd, torch.from_numpy(cloud(i))[None], 4, 2)
            C.focal_loss(lg[0], torch.eye(2)[int(labels[i])], 2.0, torch.tensor([0.25, 0.75])).backward()
        cpu_dt = time.perf_counter() - t1
        rec["cpu_baseline"] = {"value": cpu_samples / cpu_dt, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": f"first {cpu_samples} clouds ({int(sizes[0])}, {int(sizes[1])} tokens), fp32 oracle fwd+bwd (oracle/classifier_fp32.py, "
                                         f"pinned to the unmodified models_archs module), {cpu_dt:.2f} s wall"}
    # e2e: clouds start in pinned host memory, the loss comes back per sample
    pin = [(torch.from_numpy(cloud(i)).pin_memory(), torch.eye(2)[int(labels[i])].pin_memory()) for i in mine[:8]]

    def e2e_pass():
        zero_grads(model, opt)
        for k, (x, y) in enumerate(pin):
            loss, _ = step(x.to(dev, non_blocking=True), y.to(dev, non_blocking=True), 1.0 / 32)
            float(loss.item())                                               # :681 the reference reads the loss back every sample
        allreduce_grads(model)
        opt.step()

    e2e_pass()
    ms_e, _ = D.timed(e2e_pass, 1)
    rec["e2e"] = {"value": world * len(pin) / (ms_e / 1e3), "unit": "samples/s",
                  "h2d_bytes_per_step": int(sum(x.numel() * 4 + 8 for x, _ in pin) / len(pin)), "d2h_bytes_per_step": 4,
                  "api": "graph_step.GraphedTrainStep (= train_models.train_epoch(cuda_graphs=True): TransformerNoduleClassifier.forward + FocalLoss + backward) per sample from pinned host clouds, loss.item() per sample "
                         "(as train_models.py:681), one optimizer step per 8 samples per rank"}
    return rec


def bench_classifier_bimodal(D: Dist, samples: int = 32, epochs: int = 2):
    """N4: the PET/CT bimodal classifier exactly as conf/parameters_models.yaml builds it for 'petct' (two encoders,
    CLS-query cross attention both ways, three heads; CrossModalFocalLoss), one training sample = one CUDA graph per (CT, PET) pair
    of token counts.  Single GPU; samples/s with the optimizer step of a 32-sample virtual batch inside the timed region."""
    from oracle import classifier_fp32 as C
    from vit_deep_radiomics_b200 import _C, config_manager, train_models as tm
    from vit_deep_radiomics_b200.distributed import zero_grads
    from vit_deep_radiomics_b200.graph_step import graphed_step
    dev = D.dev
    cfg = config_manager.load_conf(project_dir=os.path.dirname(os.path.abspath(__file__)))
    cm = cfg["models"]["transformer"]
    d = cm["feature_dim"]
    gen = torch.Generator().manual_seed(1239)
    n_ct = torch.randint(512, 4096, (samples,), generator=gen)
    n_pet = torch.randint(64, 1024, (samples,), generator=gen)
    host = [(torch.randn(int(a), d, generator=gen), torch.randn(int(b), d, generator=gen), torch.eye(2)[i % 2]) for i, (a, b) in enumerate(zip(n_ct, n_pet))]
    data = [tuple(t.to(dev) for t in smp) for smp in host]
    torch.manual_seed(0)
    model = tm.build_model(cfg, "transformer", "petct").to(dev)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.01)
    crit = tm.make_criterion("crossmodal", dev)
    step = graphed_step(model, crit)
    model.train()

    def epoch():
        zero_grads(model, opt)
        tot = torch.zeros((), device=dev)
        for k, (xc, xp, y) in enumerate(data):
            loss, _ = step((xc, xp), y, 1.0 / 32)
            tot += loss
            if (k + 1) % 32 == 0 or k + 1 == len(data):
                opt.step()
                zero_grads(model, opt)
        return tot

    epoch()                                     # first visit of every pair of lengths: eager
    epoch()                                     # second visit: captured
    n0 = _C.launch_count()
    ms, _ = D.timed(epoch, epochs)
    launches = _C.launch_count() - n0
    tm.train_epoch(model, data[:8], crit, opt, virtual_batch_size=32, cuda_graphs=False)      # (warm: per-shape kernel attributes of the eager path)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tm.train_epoch(model, data[:8], crit, opt, virtual_batch_size=32, cuda_graphs=False)
    torch.cuda.synchronize()
    eager_ms = (time.perf_counter() - t0) / 8 * 1e3
    rec = {"workload": f"N4: TransformerNoduleBimodalClassifier ('petct' of conf/parameters_models.yaml, feature_dim {d}) training on {samples} synthetic "
                       "(CT 512..4096, PET 64..1024 token) cloud pairs, virtual batch 32, AdamW, CrossModalFocalLoss",
           "value": samples * epochs / (ms / 1e3), "unit": "samples/s", "ms_per_sample": ms / (epochs * samples), "gpu_launches": int(launches),
           "op_by_op_ms_per_sample": eager_ms,
           "cuda_graphs": {"pairs_captured": len(step.graphs), "replays": step.replays, "eager_steps": step.eager, "pool_bytes": int(step.bytes)}}
    ct_c, pet_c = cm["ct"], cm["pet"]
    ff_ct, ff_pet = int(d * ct_c["mlp_ratio"]), int(d * pet_c["mlp_ratio"])
    flops = sum(3.0 * (_classifier_flops(int(a) + 1, d, ff_ct, ct_c["num_layers"]) + _classifier_flops(int(b) + 1, d, ff_pet, pet_c["num_layers"]))
                for a, b in zip(n_ct, n_pet)) * epochs
    peaks = measured_peaks()
    tfl = flops / (ms / 1e3) / 1e12
    rec["roofline"] = {"bound": "tensor", "achieved": tfl, "unit": "TFLOP/s", "peak": peaks["bf16"], **tensor_fracs(tfl, peaks),
                       "note": "fwd+bwd algorithmic flops of the two encoders (3 x forward); batch-1 sequences: launch- and latency-bound"}
    torch.set_num_threads(os.cpu_count() or 1)
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    crit_c = tm.CrossModalFocalLoss(alpha=torch.tensor([0.25, 0.75]), gamma_unimodal=2.0, gamma_bimodal=1.0, beta=0.6)
    t1 = time.perf_counter()
    xc, xp, y = host[0]
    lg, _, lc, lp = C.bimodal_forward(sd, xc[None], xp[None], ct_c["num_heads"], pet_c["num_heads"], ct_c["num_layers"], pet_c["num_layers"])
    crit_c(lg[0], lc[0], lp[0], y).backward()
    cpu_dt = time.perf_counter() - t1
    rec["cpu_baseline"] = {"value": 1.0 / cpu_dt, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                           "sample": f"first cloud pair ({int(n_ct[0])}, {int(n_pet[0])} tokens), fp32 oracle fwd+bwd (oracle/classifier_fp32.bimodal_forward, "
                                     f"pinned to the unmodified models_archs module), {cpu_dt:.2f} s wall"}
    # e2e: cloud pairs start in pinned host memory, the loss comes back per sample
    pin = [tuple(t.pin_memory() for t in smp) for smp in host[:8]]

    def e2e_pass():
        zero_grads(model, opt)
        for xc_, xp_, y_ in pin:
            loss, _ = step((xc_.to(dev, non_blocking=True), xp_.to(dev, non_blocking=True)), y_.to(dev, non_blocking=True), 1.0 / 32)
            float(loss.item())
        opt.step()

    e2e_pass()
    ms_e, _ = D.timed(e2e_pass, 2)
    rec["e2e"] = {"value": len(pin) * 2 / (ms_e / 1e3), "unit": "samples/s",
                  "h2d_bytes_per_step": int(sum(a.numel() + b.numel() for a, b, _ in pin) * 4 // len(pin)), "d2h_bytes_per_step": 4,
                  "api": "graph_step.GraphedTrainStep((x_ct, x_pet), label) = train_models.train_epoch(cuda_graphs=True) for the bimodal model, pinned host clouds"}
    return rec


def bench_pipeline(D: Dist, patients: int = 4, cpu_slices: int = 4, model: str | None = None):
    """C5: per patient a 512x512x120 volume goes through ViT-B/16 extraction, the mask gather (+ PE) and ONE training step of the
    point-cloud classifier on the 768-wide descriptors (feature_dim 768 / 12 heads in the YAML schema; the reference's 256 comes
    from MedSAM's neck); volumes uploaded inside the timed region; patients/s over all ranks."""
    from vit_deep_radiomics_b200 import _C, synth, tfds_dense_descriptor as tdd
    from vit_deep_radiomics_b200.distributed import allreduce_grads, zero_grads
    from vit_deep_radiomics_b200.models_archs import TransformerNoduleClassifier
    from vit_deep_radiomics_b200.train_models import FocalLoss
    rank, world, dev = D.rank, D.world, D.dev
    img, mask, res, name = synth.make_case("C2", seed=1238 + rank)
    H, W, S = img.shape
    if model is not None:     # e.g. "medsam": the reference's DEFAULT pipeline -- SAM ViT-B encoder on the crop window resized to 1024^2
        name = model          # (on the device) -> 256-wide descriptors -> the classifier exactly as conf/parameters_models.yaml configures it
    backbone = tdd.load_model(name, img_hw=None if model is not None else (H, W), device=dev, seed=1234)
    dim = backbone.feature_dim
    torch.manual_seed(0)
    clf = TransformerNoduleClassifier(dim, 4 * dim, dim // 64, 2, 2).to(dev)
    if world > 1:
        with torch.no_grad():
            for p in clf.parameters():
                torch.distributed.broadcast(p, 0)
    opt = torch.optim.AdamW(clf.parameters(), lr=5e-4, weight_decay=0.01)
    crit = FocalLoss(alpha=torch.tensor([0.25, 0.75], device=dev), gamma=2)
    img_pin = torch.as_tensor(img).pin_memory()
    mask_pin = torch.as_tensor(np.ascontiguousarray(mask).view(np.uint8)).pin_memory()
    ex = tdd.PointCloudExtractor(backbone)
    window = max(1, 32 // world)
    labels = [torch.eye(2, device=dev)[i % 2] for i in range(2)]
    n_tok = 0

    from vit_deep_radiomics_b200.graph_step import graphed_step
    step = graphed_step(clf, crit)

    def run(k):
        nonlocal n_tok
        zero_grads(clf, opt)
        last = None
        for i, out in enumerate(ex.run([(img_pin, mask_pin, res)] * k, to_host=False)):
            n_tok = int(out["count"].item())                # the point cloud's size is data-dependent: one 4-byte read-back
            last, _ = step(out["tokens"][:n_tok], labels[i % 2], 1.0 / 32)
            if (i + 1) % window == 0 or i + 1 == k:
                allreduce_grads(clf)
                opt.step()
                zero_grads(clf, opt)
        return float(last.detach())

    for _ in range(3):                                      # the first passes pay the caching allocator's growth (every workspace size is new) and
        run(patients)                                       # the clock ramp: measured 57 -> 32 ms per patient between the first and the third pass
    n0 = _C.launch_count()
    ms, loss = D.timed(lambda: run(patients), 1)
    launches = _C.launch_count() - n0
    peaks = measured_peaks()
    flops = patients * world * (backbone.flops_per_slice() * S + 3.0 * _classifier_flops(n_tok + 1, dim, 4 * dim, 2))
    tfl = flops / (ms / 1e3) / 1e12
    v = world * patients / (ms / 1e3)
    rec = {"workload": f"C5: {name} extraction over a {H}x{W}x{S} volume{' (crop window resized to 1024x1024 on the device)' if model is not None else ''} -> mask gather + PE ({n_tok} tokens) -> 2-layer transformer classifier "
                       f"(d {dim}, {dim // 64} heads) training step per patient, {patients} patients per rank, gradient all-reduce per {window} patients per rank",
           "value": v, "unit": "patients/s", "slices_per_s": v * S, "ms_per_patient": ms / patients, "gpu_launches": int(launches), "loss": loss,
           "roofline": {"bound": "tensor", "achieved": tfl, "unit": "TFLOP/s", "peak": peaks["bf16"], **tensor_fracs(tfl / world, peaks),
                        "note": "algorithmic flops of the backbone + 3 x classifier forward over the whole per-patient time (uploads, gather, optimizer included)"},
           "e2e": {"value": v, "unit": "patients/s", "h2d_bytes_per_step": int(img_pin.numel() * 4 + mask_pin.numel()), "d2h_bytes_per_step": 8,
                   "api": "PointCloudExtractor.run(to_host=False) -> TransformerNoduleClassifier -> FocalLoss.backward; the timed region IS end to end "
                          "(pinned host volumes in, the token count and the loss out)"}}
    if rank == 0 and model is not None:
        from oracle import sam_fp32
        torch.set_num_threads(os.cpu_count() or 1)
        t0 = time.perf_counter()
        with torch.no_grad():
            sam_fp32.sam_dense_descriptor(backbone.state_dict_f32, backbone.cfg, torch.rand(1, 3, 1024, 1024))
        cpu_dt = time.perf_counter() - t0
        rec["cpu_baseline"] = {"value": (1.0 / S) / cpu_dt, "unit": "patients/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": f"1 of {S} slices through oracle/sam_fp32.py at 1024x1024 (the encoder alone: resize, gather and classifier step are "
                                         f"< 2 % of the host time), {cpu_dt:.2f} s wall, scaled by slices"}
    elif rank == 0:
        from oracle import classifier_fp32 as C, gather_np, vit_fp32
        torch.set_num_threads(os.cpu_count() or 1)
        cfg = vit_fp32.VIT_CONFIGS[name]
        s0 = S // 2 - cpu_slices // 2
        x = torch.from_numpy(np.ascontiguousarray(np.moveaxis(img[:, :, s0:s0 + cpu_slices], -1, 0)))[:, None].expand(-1, 3, -1, -1).contiguous()
        t0 = time.perf_counter()
        with torch.no_grad():
            dense = vit_fp32.vit_forward(backbone.state_dict_f32, cfg, x).numpy()
        o = gather_np.token_gather([dense[i] for i in range(dense.shape[0])], [mask[:, :, s0 + i] for i in range(dense.shape[0])], res)
        sd = {k: v_.detach().cpu().clone().requires_grad_(True) for k, v_ in clf.state_dict().items()}
        lg, _ = C.classifier_forward(sd, torch.from_numpy(o["tokens"].astype(np.float32))[None], dim // 64, 2)
        C.focal_loss(lg[0], torch.eye(2)[0], 2.0, torch.tensor([0.25, 0.75])).backward()
        cpu_dt = time.perf_counter() - t0
        rec["cpu_baseline"] = {"value": (cpu_slices / S) / cpu_dt, "unit": "patients/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": f"{cpu_slices} of {S} slices through the fp32 oracle ViT + NumPy gather + oracle classifier fwd/bwd "
                                         f"({o['flat'].size} tokens), {cpu_dt:.2f} s wall, scaled by slices"}
    return rec


def bench_medsam(D: Dist, B: int = 8, steps: int = 3, cpu: bool = True):
    """The reference's own default backbone (load_model('medsam'): SAM ViT-B image encoder, 1024 x 1024 inputs, (64, 64, 256)
    descriptors) through the same libvdr kernels, B gray slices per pass on rank 0's GPU."""
    from vit_deep_radiomics_b200 import _C, tfds_dense_descriptor as tdd
    dev = D.dev
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
    launches = (_C.launch_count() - n0) // steps
    peaks = measured_peaks()
    tfl = model.flops_per_slice() * B / ms / 1e9
    rec = {"workload": f"MedSAM (SAM ViT-B image encoder, the reference's default backbone): {B} gray 1024x1024 slices per pass -> (64, 64, 256) descriptors, 1 GPU",
           "value": B / ms * 1e3, "unit": "slices/s", "ms_per_batch": ms, "gpu_launches_per_batch": int(launches),
           "roofline": {"bound": "tensor", "achieved": tfl, "unit": "TFLOP/s", "peak": peaks["bf16"], **tensor_fracs(tfl, peaks),
                        "note": "942 GFLOP per slice (patch embed + block GEMMs 701, global attention incl. bias 209, windowed attention 25, neck 6) over the whole pass"}}
    # e2e: pinned host slices in, descriptors back to the host
    xp = torch.rand(B, 1024, 1024).pin_memory()
    out_host = torch.empty((B, 64, 64, 256), dtype=torch.float32).pin_memory()

    def e2e():
        xd = xp.to(dev, non_blocking=True)
        d = model.dense_descriptors(xd)
        out_host.copy_(d, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e()
    ms_e, _ = D.timed(e2e, 2) if D.world == 1 else (None, None)
    if ms_e:
        rec["e2e"] = {"value": 2 * B / (ms_e / 1e3), "unit": "slices/s", "h2d_bytes_per_step": int(xp.numel() * 4), "d2h_bytes_per_step": int(out_host.numel() * 4),
                      "api": "SamImageEncoder.dense_descriptors on pinned host slices, descriptors copied back (what get_dense_descriptor returns, batched)"}
    if cpu:
        from oracle import sam_fp32
        torch.set_num_threads(os.cpu_count() or 1)
        t0 = time.perf_counter()
        with torch.no_grad():
            sam_fp32.sam_dense_descriptor(model.state_dict_f32, model.cfg, xp[:1, None].expand(-1, 3, -1, -1))
        dt = time.perf_counter() - t0
        rec["cpu_baseline"] = {"value": 1.0 / dt, "unit": "slices/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": f"1 slice through oracle/sam_fp32.py (fp32 restatement of segment_anything's ImageEncoderViT, pinned to HF SamVisionEncoder), {dt:.2f} s wall"}
    return rec


def bench_augment(D: Dist, host_slices: int = 6):
    """V4 / N2: one patient of the reference's offline extraction loop (tfds_dense_descriptor.py:452-488: 3 flips x 4 angles of the
    whole volume, each through generate_features) -- the 12 volume copies made on the device (csrc/augment.cu, bit-identical to
    scipy.ndimage.rotate) against the reference's host flip_image + rotate_image, timed on a slab of the same volume."""
    from vit_deep_radiomics_b200 import _C, ops, synth, tfds_dense_descriptor as tdd
    dev = D.dev
    img, mask, res, name = synth.make_case("C2", seed=1240)
    H, W, S = img.shape
    model = tdd.load_model(name, img_hw=(H, W), device=dev, seed=1234)

    def patient():
        df, feats, masks = tdd.extract_patient_features(model, img, mask, "p0", 1, "synthetic_dataset", "CT", res, device_augment=True)
        return len(feats)

    patient()
    torch.cuda.synchronize()
    n0 = _C.launch_count()
    t0 = time.perf_counter()
    n_maps = patient()
    torch.cuda.synchronize()
    dt_dev = time.perf_counter() - t0
    launches = _C.launch_count() - n0
    # the 24 whole-volume transforms alone (image + mask per (flip, angle)), device time
    img_d = torch.as_tensor(img).to(dev)
    mask_d = torch.as_tensor(np.ascontiguousarray(mask).view(np.uint8)).to(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for flip in tdd.AUG_FLIPS:
        for angle in tdd.AUG_ANGLES:
            ops.flip_rotate_volume(img_d, flip, angle, kind="image")
            ops.flip_rotate_volume(mask_d, flip, angle, kind="mask_bool")
    e1.record()
    torch.cuda.synchronize()
    ms_aug = e0.elapsed_time(e1)
    voxels = H * W * S
    rec = {"workload": f"offline augmentation loop of one patient: {len(tdd.AUG_FLIPS)} flips x {len(tdd.AUG_ANGLES)} angles of a {H}x{W}x{S} volume + mask on the device "
                       f"(cubic-spline rotation bit-identical to scipy), each copy through {name} + ROI read-back ({n_maps} feature maps)",
           "value": 1.0 / dt_dev, "unit": "patients/s", "s_per_patient": dt_dev, "gpu_launches": int(launches),
           "augmentation_ms": ms_aug, "augmentation_share": ms_aug / 1e3 / dt_dev,
           "roofline": {"bound": "hbm", "kernel": "rot_prefilter_kernel + rot_interp_kernel (+ flip_copy_kernel)",
                        "achieved": 12 * voxels * (4 + 4 + 1 + 1) / (ms_aug / 1e3) / 1e9, "unit": "GB/s", "peak": measured_peaks()["hbm"],
                        "frac": 12 * voxels * 10 / (ms_aug / 1e3) / 1e9 / measured_peaks()["hbm"],
                        "note": "algorithmic bytes = volume in + out (f32) and mask in + out (u8) per (flip, angle); the rotation itself is f64 arithmetic "
                                "(prefilter planes in a padded f64 workspace), so the fraction is an upper bound on what the memory system sees"},
           "e2e": {"value": 1.0 / dt_dev, "unit": "patients/s", "h2d_bytes_per_step": int(voxels * 5), "d2h_bytes_per_step": None,
                   "api": "tfds_dense_descriptor.extract_patient_features(..., device_augment=True): NumPy volume in, metadata table + per-slice feature maps / masks out (wall clock)"}}
    if D.rank == 0:
        sub_i, sub_m = np.ascontiguousarray(img[:, :, :host_slices]), np.ascontiguousarray(mask[:, :, :host_slices])
        t0 = time.perf_counter()
        fi, fm = tdd.flip_image(sub_i, sub_m, "horizontal")
        tdd.rotate_image(fi, fm, 45)
        dt = time.perf_counter() - t0
        host_aug = dt * (S / host_slices) * 9 + 0.0            # 9 of the 12 copies are rotated (angle != 0)
        host_total = host_aug + (dt_dev - ms_aug / 1e3)         # the backbone / read-back part is the same on both paths
        rec["cpu_baseline"] = {"value": 1.0 / host_total, "unit": "patients/s", "cores": 1, "kind": "port",
                               "sample": f"flip_image + rotate_image (the reference's scipy.ndimage.rotate calls, one thread as scipy runs them) on a {host_slices}-slice slab, "
                                         f"{dt:.2f} s, scaled to {S} slices x 9 rotated copies = {host_aug:.1f} s per patient, plus the device path's backbone time"}
    return rec


# ----------------------------------------------------------------------------------------- our arm
def run_ours(args):
    D = Dist()
    rank, world = D.rank, D.world
    from vit_deep_radiomics_b200 import synth
    main = bench_extraction(D, args.config, args.steps, args.warmup, profile=True, cpu_slices=args.cpu_sample_slices, clocks=True)
    sub = {}
    if args.sub:
        torch.cuda.empty_cache()
        if world == 1:
            sub["C1"] = bench_extraction(D, "C1", 20, 3, profile=True, cpu_slices=8, patients_per_step=16)   # 16 x 8 slices share a backbone batch
        torch.cuda.empty_cache()
        sub["C4"] = bench_extraction(D, "C4", 5, 3, profile=(world == 1), cpu_slices=8, patients_per_step=8)
        torch.cuda.empty_cache()
        sub["c3"] = bench_classifier(D)
        torch.cuda.empty_cache()
        if world == 1:
            sub["c3_bimodal"] = bench_classifier_bimodal(D)
        torch.cuda.empty_cache()
        sub["C5"] = bench_pipeline(D)
        torch.cuda.empty_cache()
        if args.medsam:
            sub["C5_medsam"] = bench_pipeline(D, patients=2, model="medsam")      # the reference's default backbone + shipped classifier config
        torch.cuda.empty_cache()
        if world == 1 and args.medsam:
            sub["medsam"] = bench_medsam(D)
        torch.cuda.empty_cache()
        if world == 1:
            sub["augment"] = bench_augment(D)
    if rank != 0:
        return
    c = synth.CONFIGS[args.config]
    line = {
        "metric": "CT slices/sec ViT dense-descriptor extraction + mask gather", "value": main["value"], "unit": "slices/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": main["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": main["workload"],
                   "l2": "inputs+activations per step (>1.5 GB) exceed the 126 MB L2; no explicit flush",
                   "weights": "seeded random init (no checkpoints offline)", "model": c["model"]},
        "model_tflops": main["model_tflops"], "model_frac": main["model_frac"], "model_frac_burst": main["model_frac_burst"],
        "model_frac_nominal": main["model_frac_nominal"],
        "e2e": main["e2e"], "gpu_launches": main["gpu_launches"], "roofline": main.get("roofline"), "cpu_baseline": main.get("cpu_baseline"),
        "parity": main.get("parity"), "clocks": main["clocks"], "collective": main.get("collective"), "sub": sub}
    print(json.dumps(line))


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
    ap.add_argument("--no-sub", action="store_false", dest="sub", help="skip the sub-records (C1, C4, c3, C5, medsam, augment)")
    ap.add_argument("--no-medsam", action="store_false", dest="medsam", help="skip the MedSAM sub-record")
    args = ap.parse_args()
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
    finally:
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
