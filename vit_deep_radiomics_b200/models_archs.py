"""Point-cloud transformer classifier -- drop-in for the reference's src/models_archs.py.

``TransformerNoduleClassifier(input_dim, dim_feedforward, num_heads, num_classes, num_layers)``
has the reference's constructor, ``forward(x) -> (logits, cls)`` contract and state-dict keys
(``cls_token``, ``norm.*``, ``transformer_encoder.layers.{i}.self_attn.in_proj_weight`` ...,
``classifier.dense{1,2}.*``) so ``.pth`` files interchange with the
reference (``save_checkpoint`` / ``load_checkpoint`` keep its file naming, :14-35).

The arithmetic does not go through torch.nn: parameters are only *stored* in torch modules.
Forward and backward run in libvdr.so (CLS-concat+LayerNorm kernel, tcgen05 GEMMs with fused
bias/GELU/residual epilogues, fused attention, LayerNorm) -- see ``classifier_kernels.py``.
Train mode applies the reference's dropouts (encoder layers 0.1 / 0.5, MLPLayer 0.1, :51,58,135,187-199) inside the kernels from a
counter-based generator seeded per forward pass from torch's default generator; RNG streams differ from PyTorch's by
construction, so parity with the reference is exact in ``eval()`` and statistical in ``train()``.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import bimodal_kernels as bk
from . import classifier_kernels as ck


def save(model, model_path):
    torch.save(model.state_dict(), model_path)


def load(model, model_path):
    device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    model.load_state_dict(torch.load(model_path, map_location=device))
    return model.to(device)


def save_checkpoint(model, save_dir, epoch):
    """reference: models_archs.py:14-19 -- <save_dir>/model_epoch_<epoch:04d>.pth (state-dict only)."""
    os.makedirs(save_dir, exist_ok=True)
    save(model, os.path.join(save_dir, f"model_epoch_{str(epoch).zfill(4)}.pth"))


def load_checkpoint(model, save_dir, epoch):
    """reference: models_archs.py:22-25."""
    return load(model, os.path.join(save_dir, f"model_epoch_{str(epoch).zfill(4)}.pth"))


def set_dropout(model: nn.Module, p_encoder: float | None = None, p_head: float | None = None) -> nn.Module:
    """Change the dropout rates of a classifier built by this module (the reference hard-codes 0.1 / 0.5 in its constructors):
    ``p_encoder`` for every nn.TransformerEncoderLayer (its sub-layer dropouts and the attention probabilities), ``p_head`` for every
    MLPLayer.  ``set_dropout(model, 0.0, 0.0)`` makes ``train()`` deterministic -- what the parity tests against the fp32 oracle and
    the reference's gradients use ("dropout 0 for parity runs")."""
    for m in model.modules():
        if p_encoder is not None and isinstance(m, nn.TransformerEncoderLayer):
            for name in ("dropout", "dropout1", "dropout2"):
                getattr(m, name).p = float(p_encoder)
            m.self_attn.dropout = float(p_encoder)
        if p_head is not None and isinstance(m, MLPLayer):
            m.dropout_rate = float(p_head)
    return model


class MLPLayer(nn.Module):
    """Parameter container for the reference's MLPLayer (:186-200): dense1 -> GELU -> dense2."""

    def __init__(self, input_dim, hidden_features, out_features, dropout_rate=0.1):
        super().__init__()
        self.dense1 = nn.Linear(input_dim, hidden_features, bias=True)
        self.dense2 = nn.Linear(hidden_features, out_features, bias=True)
        self.dropout_rate = dropout_rate


class TransformerNoduleClassifier(nn.Module):
    """reference: models_archs.py:127-147."""

    def __init__(self, input_dim, dim_feedforward, num_heads, num_classes, num_layers):
        super().__init__()
        if input_dim % num_heads or input_dim // num_heads != 64:
            raise ValueError("the fused attention kernel needs head_dim == 64 "
                             f"(input_dim {input_dim} / num_heads {num_heads})")
        # torch modules are used as parameter containers only: identical names, shapes and init laws
        # to the reference, never called.
        layer = nn.TransformerEncoderLayer(d_model=input_dim, dim_feedforward=dim_feedforward, nhead=num_heads,
                                           activation="gelu", batch_first=True, dropout=0.1)
        self.norm = nn.LayerNorm(input_dim)
        self.transformer_encoder = nn.TransformerEncoder(layer, num_layers=num_layers, enable_nested_tensor=False)
        self.cls_token = nn.Parameter(torch.randn(1, 1, input_dim))
        self.classifier = MLPLayer(input_dim, input_dim * 2, num_classes)
        self.input_dim, self.dim_feedforward = input_dim, dim_feedforward
        self.num_heads, self.num_layers, self.num_classes = num_heads, num_layers, num_classes

    def _drop_cfg(self):
        """Dropout of this forward pass: active in train mode (model.train(), train_models.py:652) with the rates the modules were
        built with (0.1 in every encoder sub-layer and on the attention probabilities, :135; 0.1 in the head, :139,187)."""
        if not self.training:
            return None
        p_enc, p_head = self._drop_rates()
        if p_enc <= 0 and p_head <= 0:
            return None
        override = self.__dict__.get("_drop_override")        # graph_step: seed taken from a device counter during capture
        if override is not None:
            return override()
        return ck.DropCfg(ck.new_seed(), p_enc, p_head)

    def _drop_rates(self):
        return float(self.transformer_encoder.layers[0].dropout.p), float(self.classifier.dropout_rate)

    def param_list(self):
        """Parameters in the fixed order the kernels expect (see classifier_kernels.PARAM_ORDER)."""
        ps = [self.cls_token, self.norm.weight, self.norm.bias]
        for lyr in self.transformer_encoder.layers:
            ps += [lyr.self_attn.in_proj_weight, lyr.self_attn.in_proj_bias,
                   lyr.self_attn.out_proj.weight, lyr.self_attn.out_proj.bias,
                   lyr.norm1.weight, lyr.norm1.bias,
                   lyr.linear1.weight, lyr.linear1.bias, lyr.linear2.weight, lyr.linear2.bias,
                   lyr.norm2.weight, lyr.norm2.bias]
        ps += [self.classifier.dense1.weight, self.classifier.dense1.bias,
               self.classifier.dense2.weight, self.classifier.dense2.bias]
        return ps

    def forward(self, x):
        """x (batch, seq_len, feature_dim) f32 CUDA -> (logits (batch, classes), cls (batch, feature_dim)).
        The reference runs batch 1 (variable-length clouds, no padding); batches are looped."""
        if x.dim() != 3 or x.shape[2] != self.input_dim:
            raise ValueError(f"expected (batch, seq_len, {self.input_dim}), got {tuple(x.shape)}")
        params = self.param_list()
        train = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        logits, cls = [], []
        for b in range(x.shape[0]):
            drop = self._drop_cfg()
            if train:
                lg, c = ck.ClassifierFunction.apply(x[b], self.num_heads, self.num_layers, drop, *params)
            else:
                lg, c = ck.classifier_forward(x[b], self.num_heads, self.num_layers, params, drop=drop)
            logits.append(lg)
            cls.append(c)
        return torch.stack(logits, 0), torch.stack(cls, 0)


class CrossAttentionLayer(nn.Module):
    """Parameter container for the reference's CrossAttentionLayer (:174-183)."""

    def __init__(self, input_dim, num_heads):
        super().__init__()
        self.multihead_attn = nn.MultiheadAttention(embed_dim=input_dim, num_heads=num_heads, batch_first=True)


def _encoder_params(cls_token, norm, encoder):
    ps = [cls_token, norm.weight, norm.bias]
    for lyr in encoder.layers:
        ps += [lyr.self_attn.in_proj_weight, lyr.self_attn.in_proj_bias, lyr.self_attn.out_proj.weight, lyr.self_attn.out_proj.bias,
               lyr.norm1.weight, lyr.norm1.bias, lyr.linear1.weight, lyr.linear1.bias, lyr.linear2.weight, lyr.linear2.bias,
               lyr.norm2.weight, lyr.norm2.bias]
    return ps


def _mlp_params(m):
    return [m.dense1.weight, m.dense1.bias, m.dense2.weight, m.dense2.bias]


class TransformerNoduleBimodalClassifier(nn.Module):
    """reference: models_archs.py:38-124 -- same constructor, ``forward(x_ct=None, x_pet=None) ->
    (logits_petct, petct_cls_token, logits_ct, logits_pet)`` and state-dict keys; arithmetic in libvdr (bimodal_kernels.py)."""

    def __init__(self, input_dim, mlp_ratio_ct, mlp_ratio_pet, num_heads_ct, num_heads_pet, num_layers_ct, num_layers_pet,
                 num_classes):
        super().__init__()
        for h in (num_heads_ct, num_heads_pet):
            if input_dim % h or input_dim // h != 64:
                raise ValueError(f"the fused attention kernels need head_dim == 64 (input_dim {input_dim} / num_heads {h})")

        def enc(ratio, heads, layers):
            layer = nn.TransformerEncoderLayer(d_model=input_dim, dim_feedforward=int(ratio * input_dim), nhead=heads,
                                               activation="gelu", batch_first=True, dropout=0.5)
            return nn.TransformerEncoder(layer, num_layers=layers, enable_nested_tensor=False)

        self.transformer_encoder_ct = enc(mlp_ratio_ct, num_heads_ct, num_layers_ct)
        self.transformer_encoder_pet = enc(mlp_ratio_pet, num_heads_pet, num_layers_pet)
        self.norm_ct = nn.LayerNorm(input_dim)
        self.norm_pet = nn.LayerNorm(input_dim)
        self.cls_token_ct = nn.Parameter(torch.randn(1, 1, input_dim))
        self.cls_token_pet = nn.Parameter(torch.randn(1, 1, input_dim))
        self.classifier_ct = MLPLayer(input_dim, input_dim * 2, num_classes, dropout_rate=0.1)
        self.classifier_pet = MLPLayer(input_dim, input_dim * 2, num_classes, dropout_rate=0.1)
        self.projection_petct = MLPLayer(input_dim * 2, input_dim, input_dim, dropout_rate=0.1)
        self.cross_attention_ct = CrossAttentionLayer(input_dim, num_heads_ct)
        self.cross_attention_pet = CrossAttentionLayer(input_dim, num_heads_ct)          # (sic) the reference uses num_heads_ct for both (:70-71)
        self.classifier_petct = MLPLayer(input_dim, input_dim * 2, num_classes, dropout_rate=0.1)
        self.input_dim = input_dim
        self.cfg = dict(heads_ct=num_heads_ct, heads_pet=num_heads_pet, layers_ct=num_layers_ct, layers_pet=num_layers_pet)

    def param_groups(self):
        ca = lambda m: [m.multihead_attn.in_proj_weight, m.multihead_attn.in_proj_bias,          # noqa: E731
                        m.multihead_attn.out_proj.weight, m.multihead_attn.out_proj.bias]
        return dict(enc_ct=_encoder_params(self.cls_token_ct, self.norm_ct, self.transformer_encoder_ct),
                    enc_pet=_encoder_params(self.cls_token_pet, self.norm_pet, self.transformer_encoder_pet),
                    cross_ct=ca(self.cross_attention_ct), cross_pet=ca(self.cross_attention_pet),
                    head_ct=_mlp_params(self.classifier_ct), head_pet=_mlp_params(self.classifier_pet),
                    proj=_mlp_params(self.projection_petct), head_petct=_mlp_params(self.classifier_petct))

    def _drop_rates(self):
        return float(self.transformer_encoder_ct.layers[0].dropout.p), float(self.classifier_ct.dropout_rate)

    def forward(self, x_ct=None, x_pet=None):
        """x_* (batch, seq_len, feature_dim) f32 CUDA or None (at least one).  Batches are looped (the reference runs batch 1)."""
        if x_ct is None and x_pet is None:
            raise AssertionError("At least one modality should be used")
        groups = self.param_groups()
        flat = [p for k in bk.GROUPS for p in groups[k]]
        sizes = tuple(len(groups[k]) for k in bk.GROUPS)
        train = torch.is_grad_enabled() and any(p.requires_grad for p in flat)
        B = (x_ct if x_ct is not None else x_pet).shape[0]
        outs = [[], [], [], []]
        for b in range(B):
            xc = x_ct[b] if x_ct is not None else None
            xp = x_pet[b] if x_pet is not None else None
            drop = None
            if self.training:   # 0.5 in both encoders (:51,58), 0.1 in the four MLPLayers (:66-75)
                p_enc, p_head = self._drop_rates()
                if p_enc > 0 or p_head > 0:
                    override = self.__dict__.get("_drop_override")    # graph_step: seed taken from a device counter during capture
                    drop = override() if override is not None else ck.DropCfg(ck.new_seed(), p_enc, p_head)
            if train:
                r = bk.BimodalFunction.apply(xc, xp, self.cfg, sizes, drop, *flat)
            else:
                r = bk.bimodal_forward(xc, xp, self.cfg, groups, drop=drop)
            for o, v in zip(outs, r):
                o.append(v)
        return tuple(torch.stack(o, 0) for o in outs)
