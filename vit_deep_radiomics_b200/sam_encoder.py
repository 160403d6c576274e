"""The two backbones the reference's ``load_model`` actually knows (src/tfds_dense_descriptor.py:51-67), on libvdr kernels:

* ``SamImageEncoder`` -- ``model.image_encoder`` of ``sam_model_registry['vit_b']`` (MedSAM, :104,123): patch embedding +
  absolute position embedding, 12 pre-norm blocks with 14x14 windowed attention (global at blocks 2, 5, 8, 11) and the
  decomposed relative-position bias, neck (1x1 conv, LayerNorm2d, 3x3 conv, LayerNorm2d) -> (H/16, W/16, 256) descriptors,
  which is what gives the classifiers their ``feature_dim: 256`` (conf/parameters_models.yaml:4).
* ``DinoV2PatchEmbed`` -- ``model.patch_embed`` of the torch.hub DINOv2 ViT-S/14 (:87,128): the reference's 'dinov2' mode
  only runs the 14x14 strided convolution.

State-dict keys are segment_anything's / DINOv2's, with or without the ``image_encoder.`` prefix of a full SAM checkpoint,
so ``medsam_vit_b.pth`` loads unchanged.  The GEMMs are the tcgen05 kernel (vdr_gemm, LayerNorms folded into their epilogues);
the global blocks' attention is the tcgen05 flash kernel instantiated with the bias (vdr_relpos_tables +
vdr_flash_attn_relpos_fwd), the windowed blocks read their 14x14 windows in place (vdr_attn_relpos_windows_fwd, mma.sync:
partition, attention and unpartition in one launch).  There is no CPU path: every op is a libvdr call on CUDA tensors.
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn.functional as F

from . import _C, ops

SAM_CONFIGS = {
    "medsam": dict(dim=768, depth=12, heads=12, global_attn=(2, 5, 8, 11), window=14, out_chans=256, patch=16),
    "sam_tiny": dict(dim=128, depth=4, heads=2, global_attn=(1, 3), window=14, out_chans=64, patch=16),   # tests
    "sam_small": dict(dim=128, depth=2, heads=2, global_attn=(1,), window=14, out_chans=256, patch=16),   # tests: MedSAM's 256-wide neck
}


def init_sam_state_dict(cfg: dict, img_hw, seed: int = 1234) -> dict:
    """Seeded random weights with segment_anything key names (no checkpoints are available offline)."""
    g = torch.Generator().manual_seed(seed)
    d, p, oc, ws = cfg["dim"], cfg["patch"], cfg["out_chans"], cfg["window"]
    gh, gw = img_hw[0] // p, img_hw[1] // p
    hd = d // cfg["heads"]

    def tn(*shape, std=0.02):
        t = torch.empty(*shape, dtype=torch.float32)
        torch.nn.init.trunc_normal_(t, std=std, a=-2 * std, b=2 * std, generator=g)
        return t

    w = {"pos_embed": tn(1, gh, gw, d), "patch_embed.proj.weight": tn(d, 3, p, p), "patch_embed.proj.bias": tn(d, std=0.01)}
    for i in range(cfg["depth"]):
        b = f"blocks.{i}."
        sh, sw = (gh, gw) if i in cfg["global_attn"] else (ws, ws)
        w[b + "norm1.weight"] = 1.0 + tn(d, std=0.05)
        w[b + "norm1.bias"] = tn(d, std=0.01)
        w[b + "attn.qkv.weight"] = tn(3 * d, d, std=0.04)
        w[b + "attn.qkv.bias"] = tn(3 * d, std=0.01)
        w[b + "attn.proj.weight"] = tn(d, d)
        w[b + "attn.proj.bias"] = tn(d, std=0.01)
        w[b + "attn.rel_pos_h"] = tn(2 * sh - 1, hd, std=0.1)
        w[b + "attn.rel_pos_w"] = tn(2 * sw - 1, hd, std=0.1)
        w[b + "norm2.weight"] = 1.0 + tn(d, std=0.05)
        w[b + "norm2.bias"] = tn(d, std=0.01)
        w[b + "mlp.lin1.weight"] = tn(4 * d, d)
        w[b + "mlp.lin1.bias"] = tn(4 * d, std=0.01)
        w[b + "mlp.lin2.weight"] = tn(d, 4 * d)
        w[b + "mlp.lin2.bias"] = tn(d, std=0.01)
    w["neck.0.weight"] = tn(oc, d, 1, 1, std=0.05)
    w["neck.1.weight"] = 1.0 + tn(oc, std=0.05)
    w["neck.1.bias"] = tn(oc, std=0.01)
    w["neck.2.weight"] = tn(oc, oc, 3, 3, std=0.05)
    w["neck.3.weight"] = 1.0 + tn(oc, std=0.05)
    w["neck.3.bias"] = tn(oc, std=0.01)
    return w


def sam_flops_per_slice(cfg: dict, grid) -> float:
    """Algorithmic flops of one image through the SAM encoder (mul-add = 2; softmax / LayerNorm / GELU excluded, SURVEY.md
    section 8d convention): patch embedding, per block 24*N*d^2 for qkv / proj / MLP (the padding rows of partitioned windows are
    NOT counted), 4*n^2*d per attention extent plus the decomposed rel-pos terms 2*n*d*(Sh+Sw), neck 1x1 and 3x3 convolutions.
    medsam at 64 x 64 tokens: 941.7 GFLOP (block GEMMs + patch embedding 700.6, global attention 209.4, windows 25.3, neck 6.4)."""
    d, oc, win, p = cfg["dim"], cfg["out_chans"], cfg["window"], cfg["patch"]
    gh, gw = grid
    N = gh * gw
    f = 2.0 * N * (3 * p * p) * d
    for i in range(cfg["depth"]):
        f += 24.0 * N * d * d
        if i in cfg["global_attn"]:
            f += 4.0 * N * N * d + 2.0 * N * d * (gh + gw)
        else:
            nwin = (-(-gh // win)) * (-(-gw // win))
            f += nwin * (4.0 * (win * win) ** 2 * d + 2.0 * win * win * d * 2 * win)
    return f + 2.0 * N * d * oc + 2.0 * N * 9 * oc * oc


def _strip_prefix(sd: dict, prefix: str) -> dict:
    if any(k.startswith(prefix) for k in sd):
        return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    return sd


def _fit_rel_pos(rel_pos: torch.Tensor, size: int) -> torch.Tensor:
    """segment_anything get_rel_pos: a table whose length is not 2*size-1 is linearly interpolated to it."""
    L = 2 * size - 1
    if rel_pos.shape[0] == L:
        return rel_pos
    r = F.interpolate(rel_pos.reshape(1, rel_pos.shape[0], -1).permute(0, 2, 1), size=L, mode="linear")
    return r.reshape(-1, L).permute(1, 0)


class SamImageEncoder:
    """MedSAM / SAM ViT-B image encoder; same call surface as ``vit.ViTBackbone`` for the extraction code."""

    has_cls = False

    def __init__(self, name: str = "medsam", img_hw=(1024, 1024), state_dict: dict | None = None, device="cuda:0", seed: int = 1234):
        if name not in SAM_CONFIGS:
            raise ValueError(f"unknown SAM encoder {name!r}; choose from {sorted(SAM_CONFIGS)}")
        self.model_name = name
        self.cfg = dict(SAM_CONFIGS[name])
        self.img_hw = (int(img_hw[0]), int(img_hw[1]))
        p = self.cfg["patch"]
        if self.img_hw[0] % p or self.img_hw[1] % p:
            raise ValueError(f"image size {self.img_hw} is not a multiple of the patch size {p}")
        if self.cfg["dim"] != 64 * self.cfg["heads"]:
            raise ValueError("libvdr attention kernels need head_dim 64")
        self.grid = (self.img_hw[0] // p, self.img_hw[1] // p)
        self.n_patches = self.n_tokens = self.grid[0] * self.grid[1]
        self.token_offset = 0                                   # no CLS row in front of an image's tokens
        self.feature_dim = self.cfg["out_chans"]
        self.device = torch.device(device)
        if state_dict is not None:
            state_dict = _strip_prefix(state_dict, "image_encoder.")
        self.state_dict_f32 = state_dict if state_dict is not None else init_sam_state_dict(self.cfg, self.img_hw, seed)
        self._ws: dict = {}
        self.prepare()

    def prepare(self):
        sd, dev, cfg = self.state_dict_f32, self.device, self.cfg
        d, p, oc, ws = cfg["dim"], cfg["patch"], cfg["out_chans"], cfg["window"]
        gh, gw = self.grid
        if tuple(sd["pos_embed"].shape) != (1, gh, gw, d):
            raise ValueError(f"pos_embed is {tuple(sd['pos_embed'].shape)}, the image needs (1, {gh}, {gw}, {d})")
        f32 = lambda k: sd[k].to(dev, torch.float32).contiguous()   # noqa: E731
        bf = lambda k: sd[k].to(dev).bfloat16().contiguous()        # noqa: E731
        self.K = 3 * p * p
        self.__dict__.pop("_native", None)
        self.w = dict(pe_w=sd["patch_embed.proj.weight"].reshape(d, self.K).to(dev).bfloat16().contiguous(),
                      # gray slices (gray2rgb, :41): sum_c x . W_c = x . (W_r + W_g + W_b) -- K = p*p for the TMA im2col patch embedding
                      pe_w_gray=sd["patch_embed.proj.weight"].to(dev, torch.float32).sum(dim=1).reshape(d, p * p).bfloat16().contiguous(),
                      pe_b=f32("patch_embed.proj.bias"), pos=f32("pos_embed").reshape(gh * gw, d), blocks=[],
                      neck0=sd["neck.0.weight"].reshape(oc, d).to(dev).bfloat16().contiguous(),
                      neck1_w=f32("neck.1.weight"), neck1_b=f32("neck.1.bias"),
                      # 3x3 conv weight (out, in, ky, kx) -> (out, ky, kx, in): the K order vdr_im2col3x3_tokens writes
                      neck2=sd["neck.2.weight"].permute(0, 2, 3, 1).reshape(oc, 9 * oc).to(dev).bfloat16().contiguous(),
                      neck3_w=f32("neck.3.weight"), neck3_b=f32("neck.3.bias"))
        for i in range(cfg["depth"]):
            b = f"blocks.{i}."
            sh, sw = (gh, gw) if i in cfg["global_attn"] else (ws, ws)
            rel_hi, rel_lo = ops.relpos_split(_fit_rel_pos(sd[b + "attn.rel_pos_h"].float(), sh).to(dev),
                                              _fit_rel_pos(sd[b + "attn.rel_pos_w"].float(), sw).to(dev))
            self.w["blocks"].append(dict(
                n1w=f32(b + "norm1.weight"), n1b=f32(b + "norm1.bias"),
                qkv_w=bf(b + "attn.qkv.weight"), qkv_b=f32(b + "attn.qkv.bias"),
                proj_w=bf(b + "attn.proj.weight"), proj_b=f32(b + "attn.proj.bias"),
                rel_hi=rel_hi, rel_lo=rel_lo,
                n2w=f32(b + "norm2.weight"), n2b=f32(b + "norm2.bias"),
                fc1_w=bf(b + "mlp.lin1.weight"), fc1_b=f32(b + "mlp.lin1.bias"),
                fc2_w=bf(b + "mlp.lin2.weight"), fc2_b=f32(b + "mlp.lin2.bias"),
                window=0 if i in cfg["global_attn"] else ws))
        if self.fold_layernorm:
            # norm1 -> qkv and norm2 -> lin1 are folded into the GEMMs (vdr_fold_layernorm): the producing residual GEMM leaves the
            # row statistics, the consumer normalises in its epilogue.  The windowed blocks can do that
            # because their windows are read in place (vdr_attn_relpos_windows_fwd): a pad token is "the qkv bias", not a row.
            for blk in self.w["blocks"]:
                blk["fc1_wf"], blk["fc1_bf"], blk["fc1_cs"] = ops.fold_layernorm(blk["fc1_w"], blk["fc1_b"], blk["n2w"], blk["n2b"])
                blk["qkv_wf"], blk["qkv_bf"], blk["qkv_cs"] = ops.fold_layernorm(blk["qkv_w"], blk["qkv_b"], blk["n1w"], blk["n1b"])

    def _native_struct(self):
        """The weights as a vdr_sam_weights struct (one C call per forward, vdr_sam_forward); None for configurations the native
        forward does not cover (unfolded LayerNorms, explicit window copies, the table-reading global kernel)."""
        import ctypes as C
        cfg, w = self.cfg, self.w
        win = cfg["window"]
        # (the switches below are class attributes that tests / A/B runs flip on an instance: evaluated on every call)
        if not (self.use_native_forward and "fc1_wf" in w["blocks"][0] and self.windows_in_place and win * win <= 208
                and self.global_attn_kernel == "auto"):
            return None
        gh, gw = self.grid
        if not (gw == 64 and gh % 4 == 0 and gh <= 64) and (gw == 64 and gh % 4 == 0):
            return None                                       # "auto" would pick the bias-table kernel here (needs the REL scratch)
        nat = self.__dict__.get("_native")
        if nat is not None:
            return nat
        blocks = (_C.SamBlock * cfg["depth"])()
        for i, blk in enumerate(w["blocks"]):
            for name in ("qkv_wf", "qkv_bf", "qkv_cs", "qkv_b", "proj_w", "proj_b", "fc1_wf", "fc1_bf", "fc1_cs", "fc2_w", "fc2_b", "rel_hi", "rel_lo"):
                setattr(blocks[i], name, blk[name].data_ptr())
            blocks[i].window = int(blk["window"])
        self._native_blocks = blocks                          # keeps the array alive
        self._native = _C.SamWeights(cfg["dim"], cfg["depth"], cfg["heads"], cfg["patch"], self.img_hw[0], self.img_hw[1], cfg["out_chans"], 1e-6,
                                     w["pe_w"].data_ptr(), w["pe_w"].stride(0), w["pe_w_gray"].data_ptr(), w["pe_w_gray"].stride(0),
                                     w["pe_b"].data_ptr(), w["pos"].data_ptr(), w["neck0"].data_ptr(), w["neck1_w"].data_ptr(),
                                     w["neck1_b"].data_ptr(), w["neck2"].data_ptr(), w["neck3_w"].data_ptr(), w["neck3_b"].data_ptr(),
                                     C.cast(blocks, C.POINTER(_C.SamBlock)))
        return self._native

    #: the whole encoder as one C call (vdr_sam_forward) where it applies; False = the op-by-op path (profiling, A/B)
    use_native_forward = os.environ.get("VDR_SAM_OP_BY_OP") is None

    def _encode_native(self, B: int, images: torch.Tensor | None, im2col: torch.Tensor | None, out: torch.Tensor | None) -> torch.Tensor:
        import ctypes as C
        nat = self._native_struct()
        nb = self.__dict__.setdefault("_native_ws", {})
        need = _C.lib().vdr_sam_forward_workspace_bytes(C.byref(nat), B)
        if need == 0:
            _C.check(1, "vdr_sam_forward_workspace_bytes")
        buf = nb.get("buf")
        if buf is None or buf.numel() < need:
            buf = nb["buf"] = torch.empty(need, dtype=torch.uint8, device=self.device)
        if out is None:
            out = nb.get(("out", B))
            if out is None:
                out = nb[("out", B)] = torch.empty(B * self.n_tokens, self.feature_dim, dtype=torch.float32, device=self.device)
        _C.check(_C.lib().vdr_sam_forward(C.byref(nat), images.data_ptr() if images is not None else None,
                                          im2col.data_ptr() if im2col is not None else None, B, out.data_ptr(), out.stride(0),
                                          buf.data_ptr(), buf.numel(), ops._stream()), "vdr_sam_forward")
        return out

    def _buffers(self, B: int) -> dict:
        ws = self._ws.get(B)
        if ws is None:
            cfg, dev, bf = self.cfg, self.device, torch.bfloat16
            d, oc, win = cfg["dim"], cfg["out_chans"], cfg["window"]
            gh, gw = self.grid
            N = gh * gw
            nwin = (-(-gh // win)) * (-(-gw // win))
            NW = nwin * win * win                                    # rows per image in the windowed layout (>= N)
            in_place = self.windows_in_place and win * win <= 208
            ws = dict(A=torch.empty(B * N, self.K, dtype=bf, device=dev),
                      X=torch.empty(B * N, d, dtype=bf, device=dev), Y=torch.empty(B * N, d, dtype=bf, device=dev),
                      QKV=torch.empty(B * (N if in_place else max(N, NW)), 3 * d, dtype=bf, device=dev),
                      H=torch.empty(B * N, 4 * d, dtype=bf, device=dev),
                      N0=torch.empty(B * N, oc, dtype=bf, device=dev), N1=torch.empty(B * N, oc, dtype=bf, device=dev),
                      NA=torch.empty(B * N, 9 * oc, dtype=bf, device=dev),
                      OUT=torch.empty(B * N, oc, dtype=torch.float32, device=dev))
            if not in_place:                                         # window_partition / unpartition copies of the explicit path
                ws.update(YW=torch.empty(B * NW, d, dtype=bf, device=dev), OW=torch.empty(B * NW, d, dtype=bf, device=dev))
            if gw == 64 and gh % 4 == 0 and self.global_attn_kernel == "tcgen05":   # bias table of the table-reading tcgen05 kernel (A/B)
                ws["REL"] = torch.empty(B * cfg["heads"] * N * (gh + gw), dtype=torch.float32, device=dev)
            if len(self._ws) >= 2:                                   # a volume's full batches + its last partial one stay resident
                self._ws.pop(next(iter(self._ws)))
            self._ws[B] = ws
        return ws

    #: fold norm2 (all blocks) / norm1 (global blocks) into the GEMMs that consume them (set False before prepare() for A/B)
    fold_layernorm = True

    #: windowed blocks read their windows in place (vdr_attn_relpos_windows_fwd); False = window_partition / unpartition copies
    windows_in_place = True

    #: "auto" (tcgen05 flash kernel computing its own bias terms where the token grid is Sh x 64, else mma.sync) | "tcgen05" (the same
    #: kernel reading a bias table from vdr_relpos_tables) | "mma" (A/B timing, parity tests)
    global_attn_kernel = "auto"

    # -- forward -----------------------------------------------------------------------------
    def forward_tokens(self, src: torch.Tensor, strides, B: int) -> torch.Tensor:
        """src: f32 CUDA storage of B images addressed by element strides (batch, channel, row, col); channel stride 0 =
        gray2rgb.  Returns the neck output as a token matrix (B*gh*gw, out_chans) f32, tokens in (row, col) order."""
        H, W = self.img_hw
        if ops.PROFILE is None and self._native_struct() is not None:
            nb = self.__dict__.setdefault("_native_ws", {})
            if (src.dim() == 3 and strides[1] == 0 and src.is_contiguous() and tuple(src.shape) == (B, H, W)
                    and ops.patch_embed_supported(H, W, self.cfg["patch"])):
                # gray pictures: bf16 slices straight into the TMA im2col patch embedding (channel-summed weights), no patch matrix
                sl = nb.get(("SL", B))
                if sl is None:
                    for k in [k for k in nb if isinstance(k, tuple) and k[0] in ("SL", "A")]:
                        del nb[k]
                    sl = nb[("SL", B)] = torch.empty((B, H, W), dtype=torch.bfloat16, device=self.device)
                sl.copy_(src)
                return self._encode_native(B, sl, None, None)
            A = nb.get(("A", B))
            if A is None:
                for k in [k for k in nb if isinstance(k, tuple) and k[0] == "A"]:
                    del nb[k]
                A = nb[("A", B)] = torch.empty(B * self.n_tokens, self.K, dtype=torch.bfloat16, device=self.device)
            ops.im2col_patches(src, strides, B, H, W, self.cfg["patch"], out=A)
            return self._encode_native(B, None, A, None)
        ws = self._buffers(B)
        ops.im2col_patches(src, strides, B, H, W, self.cfg["patch"], out=ws["A"])
        return self._encode(B)

    def forward_volume(self, vol: torch.Tensor, crop) -> torch.Tensor:
        """(H, W, S) f32 CUDA volume + crop window -> (S*gh*gw, out_chans) f32; the window is resized to the encoder input
        as prepare_image does (tfds_dense_descriptor.py:40-44)."""
        S = vol.shape[2]
        Bc = min(S, max(1, int(self.volume_batch)))
        vb = self.__dict__.get("_vol")
        if vb is None or vb["S"] != S:
            vb = self._vol = dict(S=S, SL=torch.empty((S,) + self.img_hw, dtype=torch.bfloat16, device=self.device),
                                  OUT=torch.empty(S * self.n_tokens, self.feature_dim, dtype=torch.float32, device=self.device))
        ops.volume_to_slices(vol, crop, out=vb["SL"], out_hw=self.img_hw)      # all slices staged (and resized) by one kernel
        N = self.n_tokens
        native = ops.PROFILE is None and self._native_struct() is not None
        for s0 in range(0, S, Bc):                                             # encoder batches of `volume_batch` slices
            b = min(Bc, S - s0)
            if native:                                                         # one C call per batch: gray slices -> descriptors
                self._encode_native(b, vb["SL"][s0:s0 + b], None, vb["OUT"][s0 * N:(s0 + b) * N])
                continue
            ws = self._buffers(b)
            ops.im2col_gray_bf16(vb["SL"][s0:s0 + b], self.cfg["patch"], out=ws["A"])
            self._encode(b, out=vb["OUT"][s0 * N:(s0 + b) * N])
        return vb["OUT"]

    #: slices per encoder batch of forward_volume.  Throughput is flat in the batch size (one 120-slice volume on one box: 139.5 ms at 8,
    #: 136.0 at 16, 132.0 at 40, 132.4 at 60, 133.4-134.3 with all 120 in one batch); 40 keeps the activation workspace at 4.7 GB
    #: instead of 14 GB per resident volume.
    volume_batch = 40

    def _encode(self, B: int, out: torch.Tensor | None = None) -> torch.Tensor:
        cfg, w, ws = self.cfg, self.w, self._buffers(B)
        d, heads, win = cfg["dim"], cfg["heads"], cfg["window"]
        gh, gw = self.grid
        N = gh * gw
        nwh, nww = -(-gh // win), -(-gw // win)
        NW = nwh * nww * win * win
        X, Y, Hb = ws["X"], ws["Y"], ws["H"]
        scale = 1.0 / math.sqrt(64)
        # patch embedding (Conv2d 16x16 stride 16) + absolute position embedding in the GEMM epilogue
        ops.gemm(ws["A"], w["pe_w"], w["pe_b"], epilogue="residual", residual=w["pos"], out=X, res_mod=(N, 0))
        fold = "fc1_wf" in w["blocks"][0]
        if fold and "ST" not in ws:
            ws["ST"] = torch.empty(d // 64, B * N, 2, dtype=torch.float32, device=self.device)
        ST = ws.get("ST")
        in_place = self.windows_in_place and win * win <= 208          # windows read in place (no partition / unpartition copies)
        if fold:
            ops.row_stats(X, out=ST[:1])
        for i, blk in enumerate(w["blocks"]):
            qkv = ws["QKV"][:B * N]
            if blk["window"] and not in_place:
                ops.layernorm(X, blk["n1w"], blk["n1b"], 1e-6, out=Y)
                ops.window_rows(Y, B, gh, gw, win, True, out=ws["YW"])
                qkv = ws["QKV"][:B * NW]
                ops.gemm(ws["YW"], blk["qkv_w"], blk["qkv_b"], out=qkv)
                ops.attn_relpos(qkv, B * nwh * nww, win, win, heads, blk["rel_hi"], blk["rel_lo"], scale, out=ws["OW"])
                ops.window_rows(ws["OW"], B, gh, gw, win, False, out=Y)
            else:
                if fold:
                    ops.gemm(X, blk["qkv_wf"], blk["qkv_bf"], out=qkv, ln_stats=ST[:1] if i == 0 else ST, ln_colsum=blk["qkv_cs"])
                else:
                    ops.layernorm(X, blk["n1w"], blk["n1b"], 1e-6, out=Y)
                    ops.gemm(Y, blk["qkv_w"], blk["qkv_b"], out=qkv)
                if blk["window"]:
                    ops.attn_relpos_windows(qkv, blk["qkv_b"], B, gh, gw, win, heads, blk["rel_hi"], blk["rel_lo"], scale, out=Y)
                else:
                    ops.attn_relpos(qkv, B, gh, gw, heads, blk["rel_hi"], blk["rel_lo"], scale, out=Y, rel=ws.get("REL"), kernel=self.global_attn_kernel)
            if fold:
                ops.gemm(Y, blk["proj_w"], blk["proj_b"], epilogue="residual", residual=X, out=X, stats_out=ST)
                ops.gemm(X, blk["fc1_wf"], blk["fc1_bf"], epilogue="gelu", out=Hb, ln_stats=ST, ln_colsum=blk["fc1_cs"])
                ops.gemm(Hb, blk["fc2_w"], blk["fc2_b"], epilogue="residual", residual=X, out=X, stats_out=ST)
            else:
                ops.gemm(Y, blk["proj_w"], blk["proj_b"], epilogue="residual", residual=X, out=X)
                ops.layernorm(X, blk["n2w"], blk["n2b"], 1e-6, out=Y)
                ops.gemm(Y, blk["fc1_w"], blk["fc1_b"], epilogue="gelu", out=Hb)
                ops.gemm(Hb, blk["fc2_w"], blk["fc2_b"], epilogue="residual", residual=X, out=X)
        # neck: 1x1 conv (a GEMM), LayerNorm2d = LayerNorm over the channels of each token, 3x3 conv as im2col + GEMM, LayerNorm2d
        ops.gemm(X, w["neck0"], None, out=ws["N0"])
        ops.layernorm(ws["N0"], w["neck1_w"], w["neck1_b"], 1e-6, out=ws["N1"])
        ops.im2col3x3_tokens(ws["N1"], B, gh, gw, out=ws["NA"])
        ops.gemm(ws["NA"], w["neck2"], None, out=ws["N0"])
        res = ws["OUT"] if out is None else out
        ops.layernorm(ws["N0"], w["neck3_w"], w["neck3_b"], 1e-6, out=res)
        return res

    def dense_descriptors(self, images: torch.Tensor) -> torch.Tensor:
        """images (B, 3, H, W) or (B, H, W) f32 CUDA -> (B, H/16, W/16, out_chans) f32 (a copy): get_dense_descriptor's result
        for 'medsam' (:123-126), batched."""
        B = images.shape[0]
        strides = (images.stride(0), 0, images.stride(1), images.stride(2)) if images.dim() == 3 else images.stride()
        if tuple(images.shape[-2:]) != self.img_hw:
            raise ValueError(f"expected images of {self.img_hw}, got {tuple(images.shape[-2:])}")
        tok = self.forward_tokens(images, strides, B)
        return tok.view(B, self.grid[0], self.grid[1], self.feature_dim).clone()     # a copy: the workspace is reused by the next call

    def flops_per_slice(self) -> float:
        """Algorithmic flops of one image (``sam_flops_per_slice``)."""
        return sam_flops_per_slice(self.cfg, self.grid)


class DinoV2PatchEmbed:
    """The reference's 'dinov2' mode (:128-133): only ``model.patch_embed`` -- Conv2d(3, 384, 14, stride 14) + bias."""

    has_cls = False

    def __init__(self, img_hw=(896, 896), state_dict: dict | None = None, device="cuda:0", seed: int = 1234, dim: int = 384, patch: int = 14):
        self.model_name = "dinov2"
        self.img_hw = (int(img_hw[0]), int(img_hw[1]))
        if self.img_hw[0] % patch or self.img_hw[1] % patch:
            raise ValueError(f"image size {self.img_hw} is not a multiple of the patch size {patch}")     # DINOv2 PatchEmbed asserts this
        self.device = torch.device(device)
        if state_dict is None:
            g = torch.Generator().manual_seed(seed)
            wt = torch.empty(dim, 3, patch, patch)
            torch.nn.init.trunc_normal_(wt, std=0.02, a=-0.04, b=0.04, generator=g)
            bs = torch.empty(dim)
            torch.nn.init.trunc_normal_(bs, std=0.01, a=-0.02, b=0.02, generator=g)
            state_dict = {"patch_embed.proj.weight": wt, "patch_embed.proj.bias": bs}
        self.state_dict_f32 = state_dict
        wt = state_dict["patch_embed.proj.weight"]
        dim, patch = wt.shape[0], wt.shape[-1]
        self.cfg = dict(dim=dim, patch=patch, depth=0, heads=0)
        self.grid = (self.img_hw[0] // patch, self.img_hw[1] // patch)
        self.n_patches = self.n_tokens = self.grid[0] * self.grid[1]
        self.token_offset = 0
        self.feature_dim = dim
        self.K = 3 * patch * patch
        ldk = (self.K + 7) // 8 * 8
        w = torch.zeros(dim, ldk, dtype=torch.bfloat16, device=self.device)
        w[:, :self.K] = wt.reshape(dim, self.K).to(self.device).bfloat16()
        self.w = dict(pe_w=w, pe_b=state_dict["patch_embed.proj.bias"].to(self.device, torch.float32).contiguous())

    def forward_tokens(self, src: torch.Tensor, strides, B: int) -> torch.Tensor:
        H, W = self.img_hw
        A = ops.im2col_patches(src, strides, B, H, W, self.cfg["patch"])
        return ops.gemm(A, self.w["pe_w"], self.w["pe_b"], out_dtype=torch.float32, k=self.K)

    def forward_volume(self, vol: torch.Tensor, crop) -> torch.Tensor:
        sl = ops.volume_to_slices(vol, crop, out_hw=self.img_hw)
        A = ops.im2col_gray_bf16(sl, self.cfg["patch"])
        return ops.gemm(A, self.w["pe_w"], self.w["pe_b"], out_dtype=torch.float32, k=self.K)

    def dense_descriptors(self, images: torch.Tensor) -> torch.Tensor:
        B = images.shape[0]
        strides = (images.stride(0), 0, images.stride(1), images.stride(2)) if images.dim() == 3 else images.stride()
        if tuple(images.shape[-2:]) != self.img_hw:
            raise ValueError(f"expected images of {self.img_hw}, got {tuple(images.shape[-2:])}")
        return self.forward_tokens(images, strides, B).view(B, self.grid[0], self.grid[1], self.feature_dim)

    def flops_per_slice(self) -> float:
        return 2.0 * self.n_patches * self.K * self.feature_dim
