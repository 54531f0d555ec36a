"""Frozen target encoders of the REED train step: the DINOv2 ViT forward on the package's kernels.

Replaces, for the ``dinov2-vit-{s,b,l,g}`` encoder types, what the reference does at /root/reference/image/utils.py:92-105
(``torch.hub.load('facebookresearch/dinov2', 'dinov2_vit{b}14[_reg]')``, head removed, ``pos_embed`` resampled to the
16 x 16 grid of a 224-pixel input) and /root/reference/image/train.py:348-360 (``encoder.forward_features(x)`` under bf16
autocast, ``z['x_norm_patchtokens']``).  The hub code itself is not vendored in the reference; the architecture restated
here is the published ``DinoVisionTransformer``: patch-14 conv embedding, cls (+ optional register) tokens, learned
absolute position embedding, pre-norm blocks ``x += ls1 * attn(norm1(x)); x += ls2 * mlp(norm2(x))`` with affine LayerNorm
(eps 1e-6), fused-QKV attention (scale head_dim^-0.5), exact-GELU MLP, LayerScale, final LayerNorm.  Parameter names follow
that model's ``state_dict`` so published checkpoints load unchanged.

Kernels: every Linear is the tcgen05 GEMM (bias in the epilogue; LayerScale + residual = the gate+residual epilogue with
one gate row), LayerNorm = ``reed_ln_modulate_fwd`` with (shift, scale) = (bias, weight - 1), attention = ``reed_attn_fwd``
(257 tokens: the ragged-sequence kernel), GELU(erf) = ``reed_unary`` op 2.  Inference only: the encoders are frozen.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops

# model_config letter -> (embed_dim, depth, heads); patch 14, mlp_ratio 4 (ViT-g uses a SwiGLU MLP: not covered)
DINOV2_CONFIGS = {"s": (384, 12, 6), "b": (768, 12, 12), "l": (1024, 24, 16)}


class _Block(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = nn.Module()
        self.attn.qkv = nn.Linear(dim, 3 * dim)
        self.attn.proj = nn.Linear(dim, dim)
        self.ls1 = nn.Module()
        self.ls1.gamma = nn.Parameter(torch.full((dim,), 1e-5))
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = nn.Module()
        self.mlp.fc1 = nn.Linear(dim, 4 * dim)
        self.mlp.fc2 = nn.Linear(4 * dim, dim)
        self.ls2 = nn.Module()
        self.ls2.gamma = nn.Parameter(torch.full((dim,), 1e-5))


class DinoV2(nn.Module):
    """``DinoVisionTransformer`` (patch 14) with the head removed; ``forward_features`` returns the hub model's dict."""

    def __init__(self, embed_dim=768, depth=12, num_heads=12, img_size=224, patch_size=14, num_register_tokens=0,
                 precision: str = "bf16"):
        super().__init__()
        self.embed_dim, self.num_heads, self.patch_size = embed_dim, num_heads, patch_size
        self.num_register_tokens = num_register_tokens
        self.reed_precision = precision
        grid = img_size // patch_size
        self.patch_embed = nn.Module()
        self.patch_embed.proj = nn.Conv2d(3, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, grid * grid + 1, embed_dim))
        self.register_tokens = nn.Parameter(torch.zeros(1, num_register_tokens, embed_dim)) if num_register_tokens else None
        self.mask_token = nn.Parameter(torch.zeros(1, embed_dim))          # present in the checkpoints, unused at inference
        self.blocks = nn.ModuleList([_Block(embed_dim) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        self.head = nn.Identity()
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.normal_(self.cls_token, std=1e-6)
        for p in self.parameters():
            p.requires_grad_(False)

    # -- pieces ------------------------------------------------------------------------------------------------
    def _act_dtype(self):
        return torch.bfloat16 if self.reed_precision == "bf16" else torch.float32

    def _patches(self, x):
        """im2col of the stride-14 conv: [B,3,H,W] -> [B*N, 3*14*14 padded to a multiple of 8] in the act dtype."""
        B, C, H, W = x.shape
        p = self.patch_size
        cols = x.reshape(B, C, H // p, p, W // p, p).permute(0, 2, 4, 1, 3, 5).reshape(B * (H // p) * (W // p), C * p * p)
        k = cols.shape[1]
        pad = (-k) % 8
        cols = F.pad(cols.to(self._act_dtype()), (0, pad))
        w = getattr(self, "_reed_patch_w", None)
        wsrc = self.patch_embed.proj.weight
        if w is None or w.device != wsrc.device or self._reed_patch_w_version != wsrc._version or w.dtype != cols.dtype:
            w = F.pad(wsrc.detach().reshape(wsrc.shape[0], -1), (0, pad)).to(cols.dtype).contiguous()
            self._reed_patch_w, self._reed_patch_w_version = w, wsrc._version
        return cols.contiguous(), w

    def _w(self, p):
        return ops.weight_for(p, self._act_dtype())

    def _layer_norm(self, x2, norm: nn.LayerNorm):
        """Affine LayerNorm on the fused LN+modulate kernel: LN(x) * w + b = LN(x) * (1 + (w - 1)) + b, one group."""
        M, _ = x2.shape
        shift = norm.bias.detach().float().view(1, -1)
        scale = (norm.weight.detach().float() - 1.0).view(1, -1)
        return ops.ln_modulate_fwd(x2, shift, scale, M, self._act_dtype(), eps=norm.eps)[0]

    @torch.no_grad()
    def forward_features(self, x: torch.Tensor, masks=None) -> Dict[str, Optional[torch.Tensor]]:
        if not x.is_cuda:
            raise RuntimeError("reed_b200 encoders run on CUDA (sm_100a) tensors only; there is no CPU fallback")
        if masks is not None:
            raise NotImplementedError("masked forward is a pre-training feature of DINOv2, not used by REED")
        B = x.shape[0]
        D, Hh = self.embed_dim, self.num_heads
        act = self._act_dtype()
        cols, w = self._patches(x)
        n = cols.shape[0] // B
        if n + 1 != self.pos_embed.shape[1]:
            raise ValueError(f"{n} patches but pos_embed holds {self.pos_embed.shape[1] - 1}: resample it to the input grid "
                             "(load_encoders does, like utils.py:98-101)")
        tok = ops.gemm(cols, w, out_dtype=torch.float32, bias=self.patch_embed.proj.bias.detach().float()).view(B, n, D)
        seq = torch.cat([self.cls_token.float().expand(B, -1, -1), tok], dim=1) + self.pos_embed.float()
        if self.register_tokens is not None:
            seq = torch.cat([seq[:, :1], self.register_tokens.float().expand(B, -1, -1), seq[:, 1:]], dim=1)
        T = seq.shape[1]
        M = B * T
        xr = seq.reshape(M, D).contiguous()                                   # fp32 residual stream
        for blk in self.blocks:
            h = self._layer_norm(xr, blk.norm1)
            qkv = ops.gemm(h, self._w(blk.attn.qkv.weight), out_dtype=act, bias=blk.attn.qkv.bias.detach().float())
            o, _ = ops.attention_fwd(qkv, B, T, Hh, D // Hh)
            xr = ops.gemm(o, self._w(blk.attn.proj.weight), out_dtype=torch.float32, bias=blk.attn.proj.bias.detach().float(),
                          epilogue=ops.EPI_GATE_RES, aux=xr, gate=blk.ls1.gamma.detach().float().view(1, D), rows_per_group=M)
            h = self._layer_norm(xr, blk.norm2)
            a = ops.gemm(h, self._w(blk.mlp.fc1.weight), out_dtype=act, bias=blk.mlp.fc1.bias.detach().float())
            a = ops.cast(a, act, op=2)                                        # exact (erf) GELU
            xr = ops.gemm(a, self._w(blk.mlp.fc2.weight), out_dtype=torch.float32, bias=blk.mlp.fc2.bias.detach().float(),
                          epilogue=ops.EPI_GATE_RES, aux=xr, gate=blk.ls2.gamma.detach().float().view(1, D), rows_per_group=M)
        x_norm = self._layer_norm(xr, self.norm).view(B, T, D)
        R = self.num_register_tokens
        return {"x_norm_clstoken": x_norm[:, 0], "x_norm_regtokens": x_norm[:, 1:R + 1],
                "x_norm_patchtokens": x_norm[:, R + 1:], "x_prenorm": xr.view(B, T, D), "masks": None}

    def forward(self, x):
        return self.head(self.forward_features(x)["x_norm_clstoken"])


def resample_abs_pos_embed(posemb: torch.Tensor, new_size, num_prefix_tokens: int = 1) -> torch.Tensor:
    """timm.layers.pos_embed.resample_abs_pos_embed as the reference calls it (utils.py:98-101): bicubic, antialiased
    interpolation of the grid part of a [1, prefix + h*w, D] table; prefix tokens are kept."""
    num_new = new_size[0] * new_size[1] + num_prefix_tokens
    if num_new == posemb.shape[1] and new_size[0] == new_size[1]:
        return posemb
    hw = int(math.sqrt(posemb.shape[1] - num_prefix_tokens))
    prefix, grid = posemb[:, :num_prefix_tokens], posemb[:, num_prefix_tokens:]
    dim, dtype = grid.shape[-1], grid.dtype
    grid = grid.float().reshape(1, hw, hw, dim).permute(0, 3, 1, 2)
    grid = F.interpolate(grid, size=tuple(new_size), mode="bicubic", antialias=True)
    grid = grid.permute(0, 2, 3, 1).reshape(1, -1, dim).to(dtype)
    return torch.cat([prefix, grid], dim=1)


def build_dinov2(model_config: str, resolution: int = 256, registers: bool = False, state_dict=None, precision="bf16") -> DinoV2:
    """A DINOv2 ViT-{s,b,l}/14 for ``resolution``-pixel training images (fed at 224 * (resolution // 256), see
    preprocess_raw_image); ``state_dict``: a published checkpoint (37 x 37 position grid) or None for random weights."""
    if model_config not in DINOV2_CONFIGS:
        raise NotImplementedError(f"dinov2 vit-{model_config} (SwiGLU MLP) is not covered; available: {sorted(DINOV2_CONFIGS)}")
    dim, depth, heads = DINOV2_CONFIGS[model_config]
    grid = 16 * (resolution // 256)
    model = DinoV2(dim, depth, heads, img_size=grid * 14, num_register_tokens=4 if registers else 0, precision=precision)
    if state_dict is not None:
        sd = dict(state_dict)
        sd["pos_embed"] = resample_abs_pos_embed(sd["pos_embed"], [grid, grid])
        model.load_state_dict(sd, strict=True)
    return model


def load_encoders(enc_type: str, device, resolution: int = 256, ckpt_dir: Optional[str] = None):
    """Drop-in for utils.py:55-164 restricted to the DINOv2 family.  Checkpoints are read from ``ckpt_dir`` (default
    ``$REED_CKPT_DIR`` or ./ckpts) as ``dinov2_vit{b}14[_reg4]_pretrain.pth`` - the files torch.hub would download; there is
    no network on the training box, so a missing file is an error, not a download."""
    assert resolution in (256, 512), "the reference feeds 224 * (resolution // 256) pixels to the encoders"
    ckpt_dir = ckpt_dir or os.environ.get("REED_CKPT_DIR", "./ckpts")
    encoders, encoder_types, architectures = [], [], []
    for enc_name in enc_type.split(","):
        encoder_type, architecture, model_config = enc_name.split("-")
        if "dinov2" not in encoder_type:
            raise NotImplementedError(f"encoder type {encoder_type!r}: only the DINOv2 family runs on the reed_b200 kernels")
        reg = "reg" in encoder_type
        fname = os.path.join(ckpt_dir, f"dinov2_vit{model_config}14{'_reg4' if reg else ''}_pretrain.pth")
        if not os.path.exists(fname):
            raise FileNotFoundError(f"{fname} not found: place the published DINOv2 checkpoint there (no network access)")
        sd = torch.load(fname, map_location="cpu")
        encoders.append(build_dinov2(model_config, resolution, reg, sd).to(device).eval())
        encoder_types.append(encoder_type)
        architectures.append(architecture)
    return encoders, encoder_types, architectures
