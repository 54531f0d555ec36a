"""Functional CPU restatement of the DINOv2 ViT forward (torch fp32).  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference obtains this model from ``torch.hub.load('facebookresearch/dinov2', ...)``
(/root/reference/image/utils.py:92-105) - third-party code that is neither vendored under /root/reference nor reachable
from this container, and the reference holds no golden outputs for it.  What is restated is the published architecture
(``DinoVisionTransformer``, patch 14): conv patch embedding, [cls | registers | patches] + learned position table, pre-norm
blocks with affine LayerNorm(eps 1e-6), fused-QKV softmax attention (scale head_dim^-0.5), exact GELU MLP, LayerScale,
final LayerNorm; ``forward_features`` keys as consumed at /root/reference/image/train.py:354-358.
"""
from __future__ import annotations

from typing import Dict, Mapping

import torch
import torch.nn.functional as F


def forward_features(sd: Mapping[str, torch.Tensor], x: torch.Tensor, num_heads: int, patch: int = 14) -> Dict[str, torch.Tensor]:
    B = x.shape[0]
    D = sd["cls_token"].shape[-1]
    depth = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))
    tok = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=patch).flatten(2).transpose(1, 2)
    seq = torch.cat([sd["cls_token"].expand(B, -1, -1), tok], dim=1) + sd["pos_embed"]
    R = 0
    if "register_tokens" in sd and sd["register_tokens"] is not None:
        R = sd["register_tokens"].shape[1]
        seq = torch.cat([seq[:, :1], sd["register_tokens"].expand(B, -1, -1), seq[:, 1:]], dim=1)
    T = seq.shape[1]
    hd = D // num_heads
    for i in range(depth):
        p = f"blocks.{i}."
        h = F.layer_norm(seq, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps=1e-6)
        qkv = F.linear(h, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]).view(B, T, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
        att = torch.softmax(qkv[0] @ qkv[1].transpose(-1, -2) * hd ** -0.5, dim=-1) @ qkv[2]
        att = att.transpose(1, 2).reshape(B, T, D)
        seq = seq + sd[p + "ls1.gamma"] * F.linear(att, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
        h = F.layer_norm(seq, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps=1e-6)
        h = F.linear(F.gelu(F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])), sd[p + "mlp.fc2.weight"],
                     sd[p + "mlp.fc2.bias"])
        seq = seq + sd[p + "ls2.gamma"] * h
    xn = F.layer_norm(seq, (D,), sd["norm.weight"], sd["norm.bias"], eps=1e-6)
    return {"x_norm_clstoken": xn[:, 0], "x_norm_regtokens": xn[:, 1:R + 1], "x_norm_patchtokens": xn[:, R + 1:],
            "x_prenorm": seq}
