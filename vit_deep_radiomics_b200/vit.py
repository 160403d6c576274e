"""Plain ViT backbone (ViT-S/16, ViT-B/16, ViT-L/14 of BASELINE.json) whose forward pass runs
entirely in libvdr.so kernels.

Replaces the backbone call of the reference, ``model.image_encoder(img_tensor)``
(src/tfds_dense_descriptor.py:123), and produces what ``get_dense_descriptor`` returns
(:124-133): per-slice dense descriptor maps (H/p, W/p, D), patch tokens only.

Weights live in an ordinary state-dict (timm / DINOv2 key names, fp32) so checkpoints stay
interchangeable; ``prepare()`` makes the bf16 operand copies the tcgen05 GEMMs read.  Activations
are bf16 with fp32 accumulation; the final LayerNorm writes fp32 descriptors.
"""
from __future__ import annotations

import math
import os

import torch

from . import _C, ops

VIT_CONFIGS = {
    "vit_s16": dict(dim=384, depth=12, heads=6, patch=16),
    "vit_b16": dict(dim=768, depth=12, heads=12, patch=16),
    "vit_l14": dict(dim=1024, depth=24, heads=16, patch=14),
    "vit_t16": dict(dim=128, depth=2, heads=2, patch=16),   # tiny, for tests
}


def init_vit_state_dict(cfg: dict, img_hw, seed: int = 1234) -> dict:
    """Seeded random init (no checkpoints are available offline): trunc_normal(0.02) weights,
    LayerNorm gamma ~ 1, small biases.  Same law as the oracle's init so parity tests can share it."""
    g = torch.Generator().manual_seed(seed)
    d, L, p = cfg["dim"], cfg["depth"], cfg["patch"]
    n_tok = (img_hw[0] // p) * (img_hw[1] // p) + 1

    def tn(*shape, std=0.02):
        t = torch.empty(*shape, dtype=torch.float32)
        torch.nn.init.trunc_normal_(t, std=std, a=-2 * std, b=2 * std, generator=g)
        return t

    w = {"patch_embed.weight": tn(d, 3, p, p), "patch_embed.bias": tn(d, std=0.01),
         "cls_token": tn(1, 1, d), "pos_embed": tn(1, n_tok, d),
         "norm.weight": 1.0 + tn(d, std=0.05), "norm.bias": tn(d, std=0.01)}
    for i in range(L):
        b = f"blocks.{i}."
        w[b + "norm1.weight"] = 1.0 + tn(d, std=0.05)
        w[b + "norm1.bias"] = tn(d, std=0.01)
        w[b + "attn.qkv.weight"] = tn(3 * d, d)
        w[b + "attn.qkv.bias"] = tn(3 * d, std=0.01)
        w[b + "attn.proj.weight"] = tn(d, d)
        w[b + "attn.proj.bias"] = tn(d, std=0.01)
        w[b + "norm2.weight"] = 1.0 + tn(d, std=0.05)
        w[b + "norm2.bias"] = tn(d, std=0.01)
        w[b + "mlp.fc1.weight"] = tn(4 * d, d)
        w[b + "mlp.fc1.bias"] = tn(4 * d, std=0.01)
        w[b + "mlp.fc2.weight"] = tn(d, 4 * d)
        w[b + "mlp.fc2.bias"] = tn(d, std=0.01)
    return w


class ViTBackbone:
    """Holds device weights + workspaces and runs the forward pass through the C ABI."""

    def __init__(self, name: str, img_hw=(512, 512), state_dict: dict | None = None, device="cuda:0",
                 seed: int = 1234):
        if name not in VIT_CONFIGS:
            raise ValueError(f"unknown backbone {name!r}; choose from {sorted(VIT_CONFIGS)}")
        self.model_name = name
        self.cfg = dict(VIT_CONFIGS[name])
        self.img_hw = (int(img_hw[0]), int(img_hw[1]))
        p = self.cfg["patch"]
        if self.img_hw[0] % p or self.img_hw[1] % p:
            raise ValueError(f"image size {self.img_hw} is not a multiple of the patch size {p}")
        self.grid = (self.img_hw[0] // p, self.img_hw[1] // p)
        self.n_patches = self.grid[0] * self.grid[1]
        self.n_tokens = self.n_patches + 1
        self.token_offset = 1                 # the CLS row in front of every image's patch tokens
        self.feature_dim = self.cfg["dim"]
        self.device = torch.device(device)
        self.state_dict_f32 = state_dict if state_dict is not None else init_vit_state_dict(self.cfg, self.img_hw, seed)
        self._ws: dict = {}
        # Patch sizes below 16 (ViT-L/14, DINOv2): the slices are staged CELL-PADDED -- every p x p patch in the corner of a 16 x 16 cell
        # of a zeroed (gh*16, gw*16) image (vdr_volume_to_slices_cells) -- and the patch weights get zero columns at the pad positions, so
        # the patch embedding is the same TMA im2col GEMM as for 16-pixel patches (a 14-pixel box row is 28 bytes, not a TMA box).
        self.cell = None
        if (p < 16 and not ops.patch_embed_supported(self.img_hw[0], self.img_hw[1], p) and os.environ.get("VDR_NO_CELL_PAD") is None
                and ops.patch_embed_supported(self.grid[0] * 16, self.grid[1] * 16, 16)):
            self.cell = (p, 16)
        self.pe_patch = 16 if self.cell else p
        self.stage_hw = (self.grid[0] * 16, self.grid[1] * 16) if self.cell else self.img_hw
        self.prepare()

    # -- weights -----------------------------------------------------------------------------
    def prepare(self):
        sd, dev, d, p = self.state_dict_f32, self.device, self.cfg["dim"], self.cfg["patch"]
        if sd["pos_embed"].shape[1] != self.n_tokens:
            raise ValueError(f"pos_embed has {sd['pos_embed'].shape[1]} tokens, image needs {self.n_tokens}")
        K = 3 * p * p
        ldk = (K + 7) // 8 * 8
        w_pe = torch.zeros(d, ldk, dtype=torch.bfloat16, device=dev)
        w_pe[:, :K] = sd["patch_embed.weight"].reshape(d, K).to(dev).bfloat16()
        f32 = lambda k: sd[k].to(dev, torch.float32).contiguous()   # noqa: E731
        bf = lambda k: sd[k].to(dev).bfloat16().contiguous()        # noqa: E731
        w_pe_tma = w_pe
        if self.cell:   # [d, 3, p, p] -> [d, 3, 16, 16] with zeros behind every patch row / below every patch
            w4 = torch.zeros(d, 3, 16, 16, dtype=torch.bfloat16, device=dev)
            w4[:, :, :p, :p] = sd["patch_embed.weight"].to(dev).bfloat16()
            w_pe_tma = w4.reshape(d, 3 * 256).contiguous()
        # gray slices (gray2rgb feeds one picture to the three input channels): sum_c x . W_c = x . (W_r + W_g + W_b), summed in f32 and
        # rounded to bf16 once -> K = p*p for the TMA im2col GEMM (vdr_patch_embed_gemm_gray), in the cell layout when cell-padded
        w_sum = sd["patch_embed.weight"].to(dev, torch.float32).sum(dim=1)                     # (d, p, p)
        if self.cell:
            w_g = torch.zeros(d, 16, 16, dtype=torch.float32, device=dev)
            w_g[:, :p, :p] = w_sum
            w_sum = w_g
        w_pe_gray = w_sum.reshape(d, -1).bfloat16().contiguous()
        self.w = dict(pe_w=w_pe, pe_w_tma=w_pe_tma, pe_w_gray=w_pe_gray, pe_b=f32("patch_embed.bias"), cls=f32("cls_token").reshape(d),
                      pos=f32("pos_embed").reshape(self.n_tokens, d),
                      norm_w=f32("norm.weight"), norm_b=f32("norm.bias"), blocks=[])
        for i in range(self.cfg["depth"]):
            b = f"blocks.{i}."
            self.w["blocks"].append(dict(
                n1w=f32(b + "norm1.weight"), n1b=f32(b + "norm1.bias"),
                qkv_w=bf(b + "attn.qkv.weight"), qkv_b=f32(b + "attn.qkv.bias"),
                proj_w=bf(b + "attn.proj.weight"), proj_b=f32(b + "attn.proj.bias"),
                n2w=f32(b + "norm2.weight"), n2b=f32(b + "norm2.bias"),
                fc1_w=bf(b + "mlp.fc1.weight"), fc1_b=f32(b + "mlp.fc1.bias"),
                fc2_w=bf(b + "mlp.fc2.weight"), fc2_b=f32(b + "mlp.fc2.bias")))
        self._folded = bool(self.fold_layernorm)
        if self._folded:
            # norm1 -> qkv and norm2 -> fc1 folded (vdr_fold_layernorm): the blocks then run no LayerNorm kernel
            for blk in self.w["blocks"]:
                blk["qkv_wf"], blk["qkv_bf"], blk["qkv_cs"] = ops.fold_layernorm(blk["qkv_w"], blk["qkv_b"], blk["n1w"], blk["n1b"])
                blk["fc1_wf"], blk["fc1_bf"], blk["fc1_cs"] = ops.fold_layernorm(blk["fc1_w"], blk["fc1_b"], blk["n2w"], blk["n2b"])
        self.K, self.ldk = K, ldk
        # the same pointers as a vdr_vit_weights struct: the whole forward is then one C call (vdr_vit_forward)
        import ctypes as C
        blocks = (_C.VitBlock * self.cfg["depth"])()
        for i, blk in enumerate(self.w["blocks"]):
            for name in ("n1w", "n1b", "qkv_w", "qkv_b", "proj_w", "proj_b", "n2w", "n2b", "fc1_w", "fc1_b", "fc2_w", "fc2_b",
                         "qkv_wf", "qkv_bf", "qkv_cs", "fc1_wf", "fc1_bf", "fc1_cs"):
                if name in blk:
                    setattr(blocks[i], name, blk[name].data_ptr())
        w = self.w
        self._native_blocks = blocks          # keeps the array alive
        self._native = _C.VitWeights(d, self.cfg["depth"], self.cfg["heads"], self.pe_patch, self.stage_hw[0], self.stage_hw[1], 1e-6,
                                     w["pe_w_tma"].data_ptr(), w["pe_w_tma"].stride(0), w["pe_b"].data_ptr(), w["cls"].data_ptr(),
                                     w["pos"].data_ptr(), w["norm_w"].data_ptr(), w["norm_b"].data_ptr(),
                                     C.cast(blocks, C.POINTER(_C.VitBlock)),
                                     w["pe_w_gray"].data_ptr() if self.gray_fold else None, w["pe_w_gray"].stride(0))

    def _workspace(self, B: int) -> dict:
        ws = self._ws.get(B)
        if ws is None:
            d, N, dev = self.cfg["dim"], self.n_tokens, self.device
            ws = dict(OUT=torch.empty(B * N, d, dtype=torch.float32, device=dev))
            self._ws = {B: ws}   # keep one batch size resident
        return ws

    def _op_buffers(self, B: int) -> dict:
        """Activation buffers of the op-by-op path (the native path, vdr_vit_forward, carves its own out of one workspace)."""
        ws = self._workspace(B)
        if "X" not in ws:
            d, N, dev, bf = self.cfg["dim"], self.n_tokens, self.device, torch.bfloat16
            ws.update(A=torch.empty(B * self.n_patches, self.ldk, dtype=bf, device=dev),
                      X=torch.empty(B * N, d, dtype=bf, device=dev),
                      Y=torch.empty(B * N, d, dtype=bf, device=dev),
                      QKV=torch.empty(B * N, 3 * d, dtype=bf, device=dev),
                      H=torch.empty(B * N, 4 * d, dtype=bf, device=dev))
        return ws

    # -- forward -----------------------------------------------------------------------------
    def forward_volume(self, vol: torch.Tensor, crop) -> torch.Tensor:
        """vol: (H, W, S) f32 CUDA volume (np.dstack layout), crop = (y0, y1, x0, x1); a window whose size differs from
        self.img_hw is resized to it as prepare_image does.  All S slices go through the backbone as one batch; returns the
        (S*N, d) f32 token matrix."""
        S = vol.shape[2]
        ws = self._workspace(S)
        if "SL" not in ws:   # (zeroed once: the pad pixels of the cell-padded layout are never written)
            ws["SL"] = torch.zeros((S,) + self.stage_hw, dtype=torch.bfloat16, device=self.device)
        ops.volume_to_slices(vol, crop, out=ws["SL"], out_hw=self.img_hw, cells=self.cell)
        return self._encode_slices(S, ws["SL"])

    def forward_volumes(self, vols) -> torch.Tensor:
        """Several patients in ONE backbone batch: vols = [(vol (H, W, S_i) f32 CUDA, crop), ...]; their slices are staged one
        after the other and go through the encoder together (small volumes -- 16 slices of 224 x 224 -- do not fill the GPU on their
        own).  Returns the (sum(S_i) * N, d) f32 token matrix, patient i's rows behind patient i - 1's."""
        if len(vols) == 1:
            return self.forward_volume(*vols[0])
        S = sum(int(v.shape[2]) for v, _ in vols)
        ws = self._workspace(S)
        if "SL" not in ws:
            ws["SL"] = torch.zeros((S,) + self.stage_hw, dtype=torch.bfloat16, device=self.device)
        s0 = 0
        for vol, crop in vols:
            ops.volume_to_slices(vol, crop, out=ws["SL"][s0:s0 + vol.shape[2]], out_hw=self.img_hw, cells=self.cell)
            s0 += int(vol.shape[2])
        return self._encode_slices(S, ws["SL"])

    def _encode_slices(self, S: int, slices: torch.Tensor) -> torch.Tensor:
        """The staged (S,) + stage_hw bf16 slices through the encoder: the native forward, else the op-by-op path with the TMA im2col view,
        else (geometries the view cannot address) the materialised im2col matrix."""
        if ops.PROFILE is None and self.use_native_forward:
            return self._encode_native(S, slices)            # one C call enqueues the whole forward (same kernels, same order)
        if ops.patch_embed_supported(self.stage_hw[0], self.stage_hw[1], self.pe_patch):
            return self._encode(S, images=slices)            # patch embedding reads the slices through a TMA im2col view
        ops.im2col_gray_bf16(slices, self.cfg["patch"], out=self._op_buffers(S)["A"])
        return self._encode(S)

    def forward_tokens(self, src: torch.Tensor, strides, B: int) -> torch.Tensor:
        """src: f32 CUDA storage holding B images of self.img_hw addressed by element `strides`
        (batch, channel, row, col).  Returns the final-LayerNorm token matrix (B*N, d) f32
        (row b*N is the CLS token, rows b*N+1.. the patch tokens in (py, px) order)."""
        H, W = self.img_hw
        ops.im2col_patches(src, strides, B, H, W, self.cfg["patch"], out=self._op_buffers(B)["A"])
        return self._encode(B)

    def _encode(self, B: int, images: torch.Tensor | None = None) -> torch.Tensor:
        """Patch-embedding GEMM + transformer blocks + final LayerNorm.  The patch embedding reads either `images`
        ((B, H, W) bf16 slices, through the TMA im2col view) or the materialised im2col matrix in ws['A']."""
        cfg, w, ws = self.cfg, self.w, self._workspace(B)
        d, heads, N, Np = cfg["dim"], cfg["heads"], self.n_tokens, self.n_patches
        ws = self._op_buffers(B)
        # patch embedding GEMM: bias + pos-embed fused, rows written behind each image's CLS row
        if images is not None and images.dim() == 3 and self.gray_fold:
            ops.patch_embed(images, w["pe_w_gray"], w["pe_b"], w["pos"], self.pe_patch, out=ws["X"], channel_summed=True)
        elif images is not None:
            ops.patch_embed(images, w["pe_w_tma"], w["pe_b"], w["pos"], self.pe_patch, out=ws["X"])
        else:
            ops.gemm(ws["A"], w["pe_w"], w["pe_b"], epilogue="residual", residual=w["pos"], out=ws["X"], k=self.K,
                     out_group=(Np, N, 1), res_mod=(Np, 1))
        ops.write_cls_rows(w["cls"], w["pos"], ws["X"], B, N, d)
        X, Y, QKV, Hb = ws["X"], ws["Y"], ws["QKV"], ws["H"]
        scale = 1.0 / math.sqrt(64)
        if self._folded:
            # same kernels in the same order as vdr_vit_forward's folded path: the residual GEMMs leave the row statistics in ST,
            # the qkv / fc1 GEMMs read the raw residual stream X and normalise in their epilogue
            if "ST" not in ws:
                ws["ST"] = torch.empty(d // 64, B * N, 2, dtype=torch.float32, device=self.device)
            ST = ws["ST"]
            ops.row_stats(X, out=ST[:1])
            depth = len(w["blocks"])
            for i, blk in enumerate(w["blocks"]):
                ops.gemm(X, blk["qkv_wf"], blk["qkv_bf"], out=QKV, ln_stats=ST[:1] if i == 0 else ST, ln_colsum=blk["qkv_cs"])
                ops.flash_attn(QKV, B, N, heads, scale, out=Y)
                ops.gemm(Y, blk["proj_w"], blk["proj_b"], epilogue="residual", residual=X, out=X, stats_out=ST)
                ops.gemm(X, blk["fc1_wf"], blk["fc1_bf"], epilogue="gelu", out=Hb, ln_stats=ST, ln_colsum=blk["fc1_cs"])
                ops.gemm(Hb, blk["fc2_w"], blk["fc2_b"], epilogue="residual", residual=X, out=X, stats_out=ST if i + 1 < depth else None)
            ops.layernorm(X, w["norm_w"], w["norm_b"], 1e-6, out=ws["OUT"])
            return ws["OUT"]
        for blk in w["blocks"]:
            ops.layernorm(X, blk["n1w"], blk["n1b"], 1e-6, out=Y)
            ops.gemm(Y, blk["qkv_w"], blk["qkv_b"], out=QKV)
            ops.flash_attn(QKV, B, N, heads, scale, out=Y)
            ops.gemm(Y, blk["proj_w"], blk["proj_b"], epilogue="residual", residual=X, out=X)
            ops.layernorm(X, blk["n2w"], blk["n2b"], 1e-6, out=Y)
            ops.gemm(Y, blk["fc1_w"], blk["fc1_b"], epilogue="gelu", out=Hb)
            ops.gemm(Hb, blk["fc2_w"], blk["fc2_b"], epilogue="residual", residual=X, out=X)
        ops.layernorm(X, w["norm_w"], w["norm_b"], 1e-6, out=ws["OUT"])
        return ws["OUT"]

    #: fold norm1 / norm2 into the qkv / fc1 GEMMs (set False before prepare() for the LayerNorm-kernel path)
    fold_layernorm = os.environ.get("VDR_NO_LN_FOLD") is None     # the env switch exists for A/B timing only
    gray_fold = os.environ.get("VDR_NO_GRAY_FOLD") is None        # channel-summed patch weights for gray slices (A/B switch)

    #: route forward_volume through vdr_vit_forward (per-kernel profiling, ops.PROFILE, always uses the op-by-op path)
    use_native_forward = True

    def _encode_native(self, B: int, images: torch.Tensor) -> torch.Tensor:
        import ctypes as C
        ws = self._workspace(B)
        need = _C.lib().vdr_vit_forward_workspace_bytes(C.byref(self._native), B)
        buf = ws.get("NATIVE")
        if buf is None or buf.numel() < need:
            # the per-op buffers are not needed on this path: reuse their storage for the native workspace when it fits
            buf = ws["NATIVE"] = torch.empty(need, dtype=torch.uint8, device=self.device)
        out = ws["OUT"]
        Cimg = 1 if images.dim() == 3 else 3
        _C.check(_C.lib().vdr_vit_forward(C.byref(self._native), images.data_ptr(), B, Cimg, out.data_ptr(), out.stride(0),
                                          buf.data_ptr(), buf.numel(), ops._stream()), "vdr_vit_forward")
        return out

    def dense_descriptors(self, images: torch.Tensor) -> torch.Tensor:
        """images (B, 3, H, W) or (B, H, W) f32 CUDA -> (B, H/p, W/p, d) f32 (a copy)."""
        if images.dim() == 3:
            B = images.shape[0]
            strides = (images.stride(0), 0, images.stride(1), images.stride(2))
        else:
            B = images.shape[0]
            strides = images.stride()
        if tuple(images.shape[-2:]) != self.img_hw:
            raise ValueError(f"expected images of {self.img_hw}, got {tuple(images.shape[-2:])}")
        tok = self.forward_tokens(images, strides, B)
        d = self.cfg["dim"]
        return tok.view(B, self.n_tokens, d)[:, 1:, :].reshape(B, self.grid[0], self.grid[1], d)

    def flops_per_slice(self) -> float:
        """F = 2*Np*3p^2*d + L*(24*N*d^2 + 4*N^2*d)."""
        d, L, p = self.cfg["dim"], self.cfg["depth"], self.cfg["patch"]
        N, Np = self.n_tokens, self.n_patches
        return 2.0 * Np * 3 * p * p * d + L * (24.0 * N * d * d + 4.0 * N * N * d)
