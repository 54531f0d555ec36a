"""Deterministic synthetic states/inputs shared by make_golden.py and the tests.  TEST INFRASTRUCTURE ONLY.

The reference zero-initialises every adaLN linear and the output linear (sit.py:245-254), which makes the
network the identity and most gradients exactly zero; parity cases therefore use a fully random state built
here from a seeded CPU generator instead of the reference initialiser.
"""
from __future__ import annotations

from collections import OrderedDict

import torch

from .sit_oracle import ArchSpec, sincos_table_2d


def parameter_shapes(spec: ArchSpec) -> "OrderedDict[str, tuple]":
    """state_dict layout of the reference SiT; order = named_parameters order."""
    D, Dd, p, C = spec.hidden_size, spec.decoder_hidden_size, spec.patch_size, spec.in_channels
    hd = D // spec.num_heads
    hidden = int(D * spec.mlp_ratio)
    table_rows = spec.num_classes + (1 if spec.class_dropout_prob > 0 else 0)
    out = OrderedDict()
    out["pos_embed"] = (1, spec.tokens, D)
    out["x_embedder.proj.weight"] = (D, C, p, p)
    out["x_embedder.proj.bias"] = (D,)
    out["t_embedder.mlp.0.weight"] = (D, 256)
    out["t_embedder.mlp.0.bias"] = (D,)
    out["t_embedder.mlp.2.weight"] = (D, D)
    out["t_embedder.mlp.2.bias"] = (D,)
    out["y_embedder.embedding_table.weight"] = (table_rows, D)
    for i in range(spec.depth):
        b = f"blocks.{i}."
        out[b + "attn.qkv.weight"] = (3 * D, D)
        out[b + "attn.qkv.bias"] = (3 * D,)
        if spec.qk_norm:
            for nm in ("q_norm", "k_norm"):
                out[b + f"attn.{nm}.weight"] = (hd,)
                out[b + f"attn.{nm}.bias"] = (hd,)
        out[b + "attn.proj.weight"] = (D, D)
        out[b + "attn.proj.bias"] = (D,)
        out[b + "mlp.fc1.weight"] = (hidden, D)
        out[b + "mlp.fc1.bias"] = (hidden,)
        out[b + "mlp.fc2.weight"] = (D, hidden)
        out[b + "mlp.fc2.bias"] = (D,)
        out[b + "adaLN_modulation.1.weight"] = (6 * D, D)
        out[b + "adaLN_modulation.1.bias"] = (6 * D,)
    for k, z in enumerate(spec.z_dims):
        P = spec.projector_dim
        for idx, (o, i_) in zip((0, 2, 4), ((P, D), (P, P), (z, P))):
            out[f"projectors.{k}.{idx}.weight"] = (o, i_)
            out[f"projectors.{k}.{idx}.bias"] = (o,)
    out["final_layer.linear.weight"] = (p * p * C, Dd)
    out["final_layer.linear.bias"] = (p * p * C,)
    out["final_layer.adaLN_modulation.1.weight"] = (2 * Dd, Dd)
    out["final_layer.adaLN_modulation.1.bias"] = (2 * Dd,)
    return out


def random_state(spec: ArchSpec, seed: int) -> "OrderedDict[str, torch.Tensor]":
    """Weights ~ N(0, 1/fan_in) (adaLN/final smaller), biases ~ N(0, 0.05^2), norm scales ~ 1 + N(0, 0.1^2)."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for name, shape in parameter_shapes(spec).items():
        if name == "pos_embed":
            sd[name] = sincos_table_2d(spec.hidden_size, int(spec.tokens ** 0.5))[None]
            continue
        w = torch.randn(shape, generator=g)
        if name.endswith("_norm.weight"):
            w = 1 + 0.1 * w
        elif name.endswith(".bias"):
            w = 0.05 * w
        elif name.endswith("embedding_table.weight"):
            w = 0.5 * w
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            scale = fan_in ** -0.5
            if "adaLN_modulation" in name:
                scale *= 0.5
            w = scale * w
        sd[name] = w
    return sd


def random_batch(spec: ArchSpec, batch: int, seed: int, text_dims=()):
    """latents (B,C,S,S), labels (B,), target features list, t (B,1,1,1), noise, label-drop mask."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, spec.in_channels, spec.input_size, spec.input_size, generator=g)
    y = torch.randint(0, spec.num_classes, (batch,), generator=g)
    zs = []
    for z, kind in zip(spec.z_dims, spec.z_types):
        zs.append(torch.randn(batch, spec.tokens, z, generator=g) if kind == "i" else torch.randn(batch, z, generator=g))
    t = torch.rand(batch, 1, 1, 1, generator=g)
    noise = torch.randn(x.shape, generator=g)
    drop = torch.rand(batch, generator=g) < spec.class_dropout_prob
    return dict(x=x, y=y, zs=zs, t=t, noise=noise, drop=drop)


def checksum(t: torch.Tensor) -> float:
    return float(t.double().abs().sum())
