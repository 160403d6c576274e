"""GPU diagnostics: run one kernel family per process and print error statistics (no asserts).
Usage: python tools/gpu_diag.py {gemm|attn|ln|gather|im2col|voxel} ; used during bring-up via gpurun."""
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_deep_radiomics_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def stats(name, got, ref):
    got, ref = got.double().cpu(), ref.double().cpu()
    diff = (got - ref).abs()
    rel = diff.max().item() / max(ref.abs().max().item(), 1e-30)
    bad = int((~torch.isfinite(got)).sum())
    print(f"  {name}: max_abs={diff.max().item():.4e} max_rel_to_peak={rel:.3e} mean_abs={diff.mean().item():.3e} "
          f"ref_absmax={ref.abs().max().item():.3e} nonfinite={bad}", flush=True)
    return rel


def diag_gemm():
    torch.manual_seed(0)
    for (M, N, K, epi, cdt) in [(128, 256, 64, "bias", torch.float32), (128, 256, 128, "bias", torch.float32),
                                (256, 512, 768, "bias", torch.bfloat16), (300, 384, 200, "gelu", torch.bfloat16),
                                (1025, 768, 768, "residual", torch.bfloat16), (4100, 2304, 768, "bias", torch.bfloat16),
                                (64, 64, 64, "bias", torch.float32), (2050, 3072, 768, "gelu", torch.bfloat16),
                                (2050, 768, 3072, "residual", torch.bfloat16), (130, 8, 72, "bias", torch.float32)]:
        a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
        b = torch.randn(N, device=dev) * 0.1
        r = torch.randn(M, N, device=dev).bfloat16() if epi == "residual" else None
        ref = a.float() @ w.float().t() + b
        if epi == "gelu":
            ref = torch.nn.functional.gelu(ref)
        if epi == "residual":
            ref = ref + r.float()
        print(f"gemm M={M} N={N} K={K} epi={epi} out={cdt}", flush=True)
        out = ops.gemm(a, w, b, epilogue=epi, residual=r, out_dtype=cdt)
        torch.cuda.synchronize()
        stats("out", out, ref)
    # row remap (patch-embed style): M = B*Np rows written behind a CLS row, residual = pos-embed
    B, Np, d, K = 3, 196, 384, 768
    a = (torch.randn(B * Np, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(d, K, device=dev) * 0.05).bfloat16()
    b = torch.randn(d, device=dev) * 0.1
    pos = torch.randn(Np + 1, d, device=dev)
    out = torch.zeros(B * (Np + 1), d, device=dev, dtype=torch.bfloat16)
    ops.gemm(a, w, b, epilogue="residual", residual=pos, out=out, out_group=(Np, Np + 1, 1), res_mod=(Np, 1))
    torch.cuda.synchronize()
    ref = (a.float() @ w.float().t() + b).reshape(B, Np, d) + pos[1:]
    print("gemm remap", flush=True)
    stats("patch rows", out.reshape(B, Np + 1, d)[:, 1:], ref)
    print("  cls rows untouched:", bool((out.reshape(B, Np + 1, d)[:, 0] == 0).all()))
    # timing
    for (M, N, K) in [(123000, 2304, 768), (123000, 768, 768), (123000, 3072, 768), (123000, 768, 3072)]:
        a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
        b = torch.randn(N, device=dev)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            ops.gemm(a, w, b, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.gemm(a, w, b, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"gemm timing M={M} N={N} K={K}: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
        e0.record()
        for _ in range(10):
            torch.nn.functional.linear(a, w)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"   cublas (no bias)      : {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)


def diag_attn():
    torch.manual_seed(1)
    for (B, N, heads) in [(1, 128, 1), (1, 256, 1), (2, 197, 2), (1, 1025, 2), (3, 300, 4)]:
        d = heads * 64
        qkv = (torch.randn(B * N, 3 * d, device=dev)).bfloat16()
        q, k, v = qkv.float().reshape(B, N, 3, heads, 64).permute(2, 0, 3, 1, 4)
        a = torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1)
        ref = (a @ v).transpose(1, 2).reshape(B * N, d)
        ref_lse = torch.logsumexp(q @ k.transpose(-1, -2) / 8.0, dim=-1)
        print(f"attn B={B} N={N} heads={heads}", flush=True)
        out, lse = ops.flash_attn(qkv, B, N, heads, return_lse=True)
        torch.cuda.synchronize()
        stats("out", out, ref)
        stats("lse", lse, ref_lse)
    B, N, heads = 120, 1025, 12
    qkv = torch.randn(B * N, 3 * heads * 64, device=dev).bfloat16()
    out = torch.empty(B * N, heads * 64, device=dev, dtype=torch.bfloat16)
    for _ in range(2):
        ops.flash_attn(qkv, B, N, heads, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.flash_attn(qkv, B, N, heads, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fl = 4.0 * B * heads * N * N * 64
    print(f"attn timing B={B} N={N} h={heads}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)


def diag_ln():
    torch.manual_seed(2)
    for (rows, d) in [(1000, 256), (1025, 768), (333, 384), (77, 1024), (50, 2048)]:
        x = (torch.randn(rows, d, device=dev) * 2 + 0.5).bfloat16()
        g = torch.randn(d, device=dev)
        b = torch.randn(d, device=dev)
        ref = torch.nn.functional.layer_norm(x.float(), (d,), g, b, 1e-6)
        print(f"ln rows={rows} d={d}", flush=True)
        y32 = ops.layernorm(x, g, b, 1e-6, out_dtype=torch.float32)
        y16, mu, rs = ops.layernorm(x, g, b, 1e-6, save_stats=True)
        torch.cuda.synchronize()
        stats("f32", y32, ref)
        stats("bf16", y16, ref)
        stats("mean", mu, x.float().mean(1))
        if d <= 1024:
            dy = torch.randn(rows, d, device=dev).bfloat16()
            xr = x.float().requires_grad_(True)
            gr = g.clone().requires_grad_(True)
            br = b.clone().requires_grad_(True)
            torch.nn.functional.layer_norm(xr, (d,), gr, br, 1e-6).backward(dy.float())
            dg = torch.zeros(d, device=dev)
            db = torch.zeros(d, device=dev)
            dx = ops.layernorm_bwd(dy, x, g, mu, rs, dg, db)
            torch.cuda.synchronize()
            stats("dx", dx, xr.grad)
            stats("dgamma", dg, gr.grad)
            stats("dbeta", db, br.grad)
    n, d = 777, 256
    x = torch.randn(n, d, device=dev)
    cls = torch.randn(d, device=dev)
    g, b = torch.randn(d, device=dev), torch.randn(d, device=dev)
    y = ops.cls_concat_layernorm(x, cls, g, b, 1e-5)
    ref = torch.nn.functional.layer_norm(torch.cat([cls[None], x]), (d,), g, b, 1e-5)
    torch.cuda.synchronize()
    print("cls_concat_ln", flush=True)
    stats("y", y, ref)
    rows, d = 123000, 768
    x = torch.randn(rows, d, device=dev).bfloat16()
    g, b = torch.ones(d, device=dev), torch.zeros(d, device=dev)
    out = torch.empty_like(x)
    for _ in range(3):
        ops.layernorm(x, g, b, 1e-6, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.layernorm(x, g, b, 1e-6, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"ln timing rows={rows} d={d}: {ms:.3f} ms  {2 * rows * d * 2 / ms / 1e6:.0f} GB/s", flush=True)


def diag_gather():
    from oracle import gather_np as G
    rng = np.random.default_rng(3)
    for (S, h, w, hm, wm, D, dens) in [(5, 6, 6, 20, 20, 16, 0.3), (7, 18, 15, 61, 47, 48, 0.25), (14, 17, 13, 55, 40, 256, 0.2),
                                       (120, 32, 32, 512, 512, 768, 0.04), (4, 5, 5, 16, 16, 8, 0.0), (2, 3, 4, 9, 12, 8, 1.1)]:
        feats = rng.standard_normal((S, h, w, D)).astype(np.float32)
        masks = rng.random((S, hm, wm)) < dens
        res, noise = rng.uniform(0.5, 1.5, 3), rng.uniform(-5, 5, 3)
        o = G.token_gather(list(feats), list(masks), res, noise, D)
        f_t = torch.from_numpy(feats).to(dev)
        m_t = torch.from_numpy(masks.astype(np.uint8)).to(dev)
        tok, src, cnt = ops.mask_gather(f_t, m_t, pe=dict(res=res, noise=noise))
        raw, src2, cnt2 = ops.mask_gather(f_t, m_t)
        torch.cuda.synchronize()
        n = int(cnt.item())
        print(f"gather S={S} h={h} w={w} D={D}: count={n} ref={o['flat'].size}", flush=True)
        if n == o["flat"].size and n > 0:
            print("  src bit-exact:", bool(np.array_equal(src[:n].cpu().numpy(), o["src"])),
                  " raw bit-exact:", bool(np.array_equal(raw[:n].cpu().numpy(), o["raw"])))
            ref32 = o["tokens"].astype(np.float32)
            got = tok[:n].cpu().numpy()
            print("  tokens+PE: max_abs", float(np.abs(got - ref32).max()), "bit-identical frac", float((got == ref32).mean()))
        tb, _, _ = ops.mask_gather(f_t.bfloat16(), m_t)
        torch.cuda.synchronize()
        if n:
            print("  bf16 feat path exact:", bool(torch.equal(tb[:n], f_t.bfloat16().float().reshape(-1, D)[torch.from_numpy(
                (o['src'][:, 0].astype(np.int64) * h + o['src'][:, 1]) * w + o['src'][:, 2]).to(dev)])))


def diag_voxel():
    from oracle import gather_np as G
    rng = np.random.default_rng(4)
    for shape in [(16, 16, 6), (12, 20, 5), (64, 64, 10), (512, 512, 120)]:
        H, W, S = shape
        img = rng.normal(-300, 350, shape).astype(np.float32)
        zz = np.indices(shape)
        mask = (((zz[0] - H * 0.5) / (H * 0.2)) ** 2 + ((zz[1] - W * 0.45) / (W * 0.15)) ** 2 + ((zz[2] - S * 0.5) / (S * 0.3)) ** 2) <= 1
        o = G.voxel_pointcloud_box(img, mask, (0.8, 0.8, 0.8))
        i_t, m_t = torch.from_numpy(img).to(dev), torch.from_numpy(mask.astype(np.uint8)).to(dev)
        bbox = ops.voxel_bbox(m_t)
        bb = bbox.cpu().numpy()
        cap = int(max(0, bb[1] - bb[0] + 1) * max(0, bb[3] - bb[2] + 1) * max(0, bb[5] - bb[4] + 1))
        flat, raw, mk, cnt = ops.voxel_gather(i_t, m_t, bbox, max(cap, 1))
        torch.cuda.synchronize()
        n = int(cnt.item())
        print(f"voxel {shape}: bbox={bb.tolist()} count={n} ref={o['flat'].size}", flush=True)
        if n == o["flat"].size:
            print("  flat exact:", bool(np.array_equal(flat[:n].cpu().numpy(), o["flat"])), " raw exact:",
                  bool(np.array_equal(raw[:n].cpu().numpy(), o["raw"])), " mask exact:",
                  bool(np.array_equal(mk[:n].cpu().numpy().astype(bool), o["mask"].astype(bool))))


def diag_im2col():
    torch.manual_seed(5)
    for (B, H, W, p, gray) in [(2, 64, 64, 16, True), (2, 56, 42, 14, False), (3, 224, 224, 16, True)]:
        if gray:
            vol = torch.randn(H, W, B, device=dev)  # (H, W, S) layout as np.dstack gives
            img = vol.permute(2, 0, 1)[:, None].expand(B, 3, H, W)
            A = ops.im2col_patches(vol, (1, 0, W * B, B), B, H, W, p)
        else:
            img = torch.randn(B, 3, H, W, device=dev)
            A = ops.im2col_patches(img, img.stride(), B, H, W, p)
        torch.cuda.synchronize()
        ref = torch.nn.functional.unfold(img.contiguous(), kernel_size=p, stride=p).transpose(1, 2).reshape(-1, 3 * p * p)
        K = 3 * p * p
        print(f"im2col B={B} {H}x{W} p={p} gray={gray}: exact={bool(torch.equal(A[:, :K].float(), ref.bfloat16().float()))} "
              f"pad_zero={bool((A[:, K:] == 0).all())}", flush=True)


if __name__ == "__main__":
    which = sys.argv[1]
    print(f"== {which} on {torch.cuda.get_device_name(0)}", flush=True)
    {"gemm": diag_gemm, "attn": diag_attn, "ln": diag_ln, "gather": diag_gather, "voxel": diag_voxel,
     "im2col": diag_im2col}[which]()
    torch.cuda.synchronize()
    print(f"== {which} done", flush=True)
