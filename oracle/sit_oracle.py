"""Functional CPU restatement of the reference SiT forward.  TEST INFRASTRUCTURE ONLY.

Everything is a pure function of a ``state_dict``-shaped mapping ``sd`` (name -> tensor) and an
``ArchSpec``; there are no modules, so autograd on the leaves of ``sd`` gives per-parameter gradients.

Reference lines followed (all under /root/reference/image/models/sit.py unless noted):
  timestep features 46-64, t-MLP 66-70, label dropout + table gather 84-100, adaLN block 130-137,
  final layer 153-158, unpatchify 256-269, forward 271-311, projector MLP 17-24,
  2-D sin-cos table 319-366, zoo 373-415.
timm semantics restated (timm is un-vendored, see oracle/timm_shim): PatchEmbed = strided conv then
flatten(2).transpose(1,2); Attention = fused-QKV MHA with columns ordered (3, heads, head_dim), optional
LayerNorm(head_dim, eps 1e-5, affine) on q and k, softmax(q k^T / sqrt(d)) v; Mlp = fc2(act(fc1(x))).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Mapping, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

# name -> (depth, hidden, heads); decoder width quirk handled in zoo_spec (sit.py:373-407)
_FAMILIES = {"XL": (28, 1152, 16), "L": (24, 1024, 16), "B": (12, 768, 12), "S": (12, 384, 6)}


@dataclass
class ArchSpec:
    input_size: int = 32
    patch_size: int = 2
    in_channels: int = 4
    hidden_size: int = 1152
    decoder_hidden_size: int = 768
    depth: int = 28
    num_heads: int = 16
    mlp_ratio: float = 4.0
    class_dropout_prob: float = 0.1
    num_classes: int = 1000
    encoder_depth: int = 8
    encoder_depth_text: Optional[int] = None
    z_dims: Sequence[int] = field(default_factory=lambda: [768])
    z_types: Sequence[str] = field(default_factory=lambda: ["i"])
    projector_dim: int = 2048
    qk_norm: bool = False

    @property
    def tokens(self) -> int:
        return (self.input_size // self.patch_size) ** 2


def zoo_spec(name: str, **overrides) -> ArchSpec:
    """'SiT-XL/2' etc.  S models inherit decoder_hidden_size=768 unless overridden (reference quirk)."""
    fam, patch = name.split("-")[1].split("/")
    depth, hidden, heads = _FAMILIES[fam]
    kw = dict(depth=depth, hidden_size=hidden, num_heads=heads, patch_size=int(patch))
    if fam != "S":
        kw["decoder_hidden_size"] = hidden
    kw.update(overrides)
    return ArchSpec(**kw)


# --------------------------------------------------------------------------------------------------
# small pieces
# --------------------------------------------------------------------------------------------------

def sincos_table_2d(width: int, grid: int) -> torch.Tensor:
    """(grid*grid, width) fp32 table; float64 math; first half of the channels encodes the column index."""
    quarter = width // 4
    omega = 1.0 / (10000.0 ** (np.arange(quarter, dtype=np.float64) / quarter))
    col = np.tile(np.arange(grid, dtype=np.float64), grid)        # w index varies fastest
    row = np.repeat(np.arange(grid, dtype=np.float64), grid)
    parts = []
    for pos in (col, row):
        ang = pos[:, None] * omega[None, :]
        parts += [np.sin(ang), np.cos(ang)]
    return torch.from_numpy(np.concatenate(parts, axis=1)).float()


def timestep_features(t: torch.Tensor, width: int = 256) -> torch.Tensor:
    half = width // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half).to(t.device)
    ang = t.float()[:, None] * freqs[None, :]
    return torch.cat([ang.cos(), ang.sin()], dim=1).to(t.dtype)


def _affine(x, sd, prefix):
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


def _ln(x, eps):
    return F.layer_norm(x, (x.shape[-1],), None, None, eps)


def _modulated_norm(x, shift, scale):
    return _ln(x, 1e-6) * (1.0 + scale[:, None, :]) + shift[:, None, :]


def attention(x, sd, prefix, heads, qk_norm):
    n, t, d = x.shape
    hd = d // heads
    qkv = _affine(x, sd, prefix + ".qkv").view(n, t, 3, heads, hd)
    q, k, v = (qkv[:, :, i].transpose(1, 2) for i in range(3))          # (n, heads, t, hd)
    if qk_norm:
        q = F.layer_norm(q, (hd,), sd[prefix + ".q_norm.weight"], sd[prefix + ".q_norm.bias"], 1e-5)
        k = F.layer_norm(k, (hd,), sd[prefix + ".k_norm.weight"], sd[prefix + ".k_norm.bias"], 1e-5)
    prob = torch.softmax((q @ k.transpose(-1, -2)) * hd ** -0.5, dim=-1)
    ctx = (prob @ v).transpose(1, 2).reshape(n, t, d)
    return _affine(ctx, sd, prefix + ".proj")


def block(x, c, sd, prefix, heads, qk_norm):
    mod = _affine(F.silu(c), sd, prefix + ".adaLN_modulation.1")
    sh_a, sc_a, g_a, sh_m, sc_m, g_m = mod.chunk(6, dim=-1)
    x = x + g_a[:, None, :] * attention(_modulated_norm(x, sh_a, sc_a), sd, prefix + ".attn", heads, qk_norm)
    h = F.gelu(_affine(_modulated_norm(x, sh_m, sc_m), sd, prefix + ".mlp.fc1"), approximate="tanh")
    return x + g_m[:, None, :] * _affine(h, sd, prefix + ".mlp.fc2")


def projector(x, sd, prefix):
    x = F.silu(_affine(x, sd, prefix + ".0"))
    x = F.silu(_affine(x, sd, prefix + ".2"))
    return _affine(x, sd, prefix + ".4")


def patchify_embed(x, sd, patch):
    w, b = sd["x_embedder.proj.weight"], sd["x_embedder.proj.bias"]
    y = F.conv2d(x, w, b, stride=patch)
    return y.flatten(2).transpose(1, 2) + sd["pos_embed"]


def unpatchify(tok, channels, patch):
    n, t, _ = tok.shape
    g = int(round(t ** 0.5))
    assert g * g == t
    return tok.view(n, g, g, patch, patch, channels).permute(0, 5, 1, 3, 2, 4).reshape(n, channels, g * patch, g * patch)


# --------------------------------------------------------------------------------------------------
# full forward
# --------------------------------------------------------------------------------------------------

def sit_forward(sd: Mapping[str, torch.Tensor], spec: ArchSpec, x, t, y, *, inference: bool = True,
                training: bool = False, drop_mask: Optional[torch.Tensor] = None):
    """Returns (prediction (N,C,H,W), zs or None).

    ``drop_mask``: explicit boolean label-dropout mask.  When None and dropout applies
    (training and class_dropout_prob>0) it is drawn as ``torch.rand(N, device) < p`` exactly like
    sit.py:89, consuming the device generator at the same point of the call sequence.
    """
    tok = patchify_embed(x, sd, spec.patch_size)
    n, tcount, width = tok.shape

    feats = timestep_features(t)
    t_emb = _affine(F.silu(_affine(feats, sd, "t_embedder.mlp.0")), sd, "t_embedder.mlp.2")
    labels = y
    if training and spec.class_dropout_prob > 0:
        if drop_mask is None:
            drop_mask = torch.rand(labels.shape[0], device=labels.device) < spec.class_dropout_prob
        labels = torch.where(drop_mask, torch.full_like(labels, spec.num_classes), labels)
    c = t_emb + sd["y_embedder.embedding_table.weight"][labels]

    split_taps = spec.encoder_depth_text is not None and spec.encoder_depth_text != spec.encoder_depth
    zs = None
    z_img = z_txt = None
    for i in range(spec.depth):
        tok = block(tok, c, sd, f"blocks.{i}", spec.num_heads, spec.qk_norm)
        if inference:
            continue
        layer = i + 1
        if layer == spec.encoder_depth:
            if not split_taps:
                zs = [projector(tok.reshape(-1, width), sd, f"projectors.{k}").view(n, tcount, -1) if kind == "i"
                      else projector(tok.mean(dim=1), sd, f"projectors.{k}")
                      for k, kind in enumerate(spec.z_types)]
            else:
                for k, kind in enumerate(spec.z_types):
                    if kind == "i":
                        z_img = projector(tok.reshape(-1, width), sd, f"projectors.{k}").view(n, tcount, -1)
        if split_taps and layer == spec.encoder_depth_text:
            for k, kind in enumerate(spec.z_types):
                if kind == "t":
                    z_txt = projector(tok.mean(dim=1), sd, f"projectors.{k}")
    if not inference and split_taps:
        zs = [z_img, z_txt]

    shift, scale = _affine(F.silu(c), sd, "final_layer.adaLN_modulation.1").chunk(2, dim=-1)
    out = _affine(_modulated_norm(tok, shift, scale), sd, "final_layer.linear")
    return unpatchify(out, spec.in_channels, spec.patch_size), zs


def as_model(sd, spec: ArchSpec, training: bool = False, drop_mask=None):
    """Callable with the reference call convention ``model(x, t, y=..., inference=True) -> (pred, zs)``."""
    def call(x, t, y=None, inference=True, **_):
        return sit_forward(sd, spec, x, t, y, inference=inference, training=training, drop_mask=drop_mask)
    return call


def flops_per_image(spec: ArchSpec, train: bool = True) -> float:
    """Algorithmic FLOPs (2*MAC) per image, formula of BASELINE.md section 3 / SURVEY.md section 8(d)."""
    L, T, D = spec.depth, spec.tokens, spec.hidden_size
    f = L * T * 24 * D * D + L * 4 * T * T * D + L * 12 * D * D
    for z, kind in zip(spec.z_dims, spec.z_types):
        f += 2 * (D * spec.projector_dim + spec.projector_dim ** 2 + spec.projector_dim * z) * (T if kind == "i" else 1)
    f += 64 * T * D + 4 * D * D + 2 * (256 * D + D * D)
    return float(f) * (3.0 if train else 1.0)
