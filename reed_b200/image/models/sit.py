"""Drop-in replacement for the reference ``image/models/sit.py`` running on hand-written sm_100a kernels.

API surface kept from /root/reference/image/models/sit.py: ``SiT`` constructor kwargs (165-185), the module tree and
therefore the ``state_dict`` keys/shapes (198-215), ``initialize_weights`` RNG consumption order (218-254),
``SiT.forward(x, t, y, inference=True) -> (pred, zs)`` (271-311), ``unpatchify`` (256-269), ``build_mlp`` (17-24),
``modulate`` (26-27), the twelve ``SiT_models`` entries (373-415) including the SiT-S ``decoder_hidden_size=768``
quirk, and the requirement that ``qk_norm`` is passed (115).

The modules below only OWN parameters (in the reference's creation order).  The arithmetic of ``forward`` is done by
``reed_b200.ops``: fused LayerNorm+modulate, tcgen05 GEMMs with bias/GELU/SiLU/gate+residual epilogues, fused
attention, all through the C-ABI library.  CPU tensors raise - there is no fallback path.

Precision: inside ``torch.autocast(device_type='cuda', dtype=torch.bfloat16)`` (what ``accelerate`` sets up for
``--mixed-precision bf16``) the bf16 tensor-core path runs; otherwise the fp32 path.  ``model.reed_precision`` may be
set to ``'bf16'`` / ``'fp32'`` to override.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from ... import ops


def build_mlp(hidden_size, projector_dim, z_dim):
    layers = [nn.Linear(hidden_size, projector_dim), nn.SiLU(),
              nn.Linear(projector_dim, projector_dim), nn.SiLU(),
              nn.Linear(projector_dim, z_dim)]
    return nn.Sequential(*layers)


def modulate(x, shift, scale):
    """API-compatibility helper (sit.py:26-27); the model itself uses the fused LayerNorm+modulate kernel."""
    return shift.unsqueeze(1) + x * (scale.unsqueeze(1) + 1)


# ----------------------------------------------------------------------------------------------------
# parameter containers (names and creation order follow the reference / timm)
# ----------------------------------------------------------------------------------------------------

class PatchEmbed(nn.Module):
    """Parameters of timm's PatchEmbed: a strided conv used as a (C*p*p -> D) linear map per patch."""

    def __init__(self, img_size, patch_size, in_chans, embed_dim, bias=True):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.grid_size = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, bias=bias)

    def patches(self, x):
        """(N,C,H,W) -> (N*T, C*p*p) rows in the conv-weight order (c, p, q)."""
        n, c, hgt, wid = x.shape
        p = self.patch_size[0]
        gh, gw = hgt // p, wid // p
        return x.reshape(n, c, gh, p, gw, p).permute(0, 2, 4, 1, 3, 5).reshape(n * gh * gw, c * p * p)


class TimestepEmbedder(nn.Module):
    def __init__(self, hidden_size, frequency_embedding_size=256):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(frequency_embedding_size, hidden_size, bias=True), nn.SiLU(),
                                 nn.Linear(hidden_size, hidden_size, bias=True))
        self.frequency_embedding_size = frequency_embedding_size

    _freq_cache = {}

    @staticmethod
    def positional_embedding(t, dim, max_period=10000):
        half = dim // 2
        key = (half, max_period, t.device)
        freqs = TimestepEmbedder._freq_cache.get(key)
        if freqs is None:      # CPU exp like the reference (sit.py:57-59), moved once: no H2D copy inside a captured graph
            freqs = torch.exp(torch.arange(half, dtype=torch.float32) * (-math.log(max_period) / half)).to(t.device)
            TimestepEmbedder._freq_cache[key] = freqs
        ang = t.float().unsqueeze(1) * freqs.unsqueeze(0)
        emb = torch.cat((ang.cos(), ang.sin()), dim=1)
        if dim % 2:
            emb = torch.cat((emb, emb.new_zeros(emb.shape[0], 1)), dim=1)
        return emb

    def forward(self, t, act_dtype=torch.float32):
        feats = self.positional_embedding(t, self.frequency_embedding_size).to(t.dtype).float()
        hid = ops.linear(feats, self.mlp[0].weight, self.mlp[0].bias, act=ops.ACT_SILU, act_dtype=act_dtype)
        return ops.linear(hid, self.mlp[2].weight, self.mlp[2].bias, act_dtype=act_dtype, out_dtype=torch.float32)


class LabelEmbedder(nn.Module):
    def __init__(self, num_classes, hidden_size, dropout_prob):
        super().__init__()
        self.embedding_table = nn.Embedding(num_classes + int(dropout_prob > 0), hidden_size)
        self.num_classes = num_classes
        self.dropout_prob = dropout_prob

    def token_drop(self, labels, force_drop_ids=None):
        if force_drop_ids is None:
            dropped = torch.rand(labels.shape[0], device=labels.device) < self.dropout_prob   # same RNG draw as sit.py:89
        else:
            dropped = force_drop_ids == 1
        return torch.where(dropped, self.num_classes, labels)

    def forward(self, labels, train, force_drop_ids=None):
        if (train and self.dropout_prob > 0) or force_drop_ids is not None:
            labels = self.token_drop(labels, force_drop_ids)
        return self.embedding_table(labels)


class _AttentionParams(nn.Module):
    def __init__(self, dim, num_heads, qk_norm):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.fused_attn = True
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.q_norm = nn.LayerNorm(self.head_dim) if qk_norm else nn.Identity()
        self.k_norm = nn.LayerNorm(self.head_dim) if qk_norm else nn.Identity()
        self.proj = nn.Linear(dim, dim)


class _MlpParams(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden, bias=True)
        self.fc2 = nn.Linear(hidden, dim, bias=True)


class SiTBlock(nn.Module):
    def __init__(self, hidden_size, num_heads, mlp_ratio=4.0, **block_kwargs):
        super().__init__()
        self.qk_norm = block_kwargs["qk_norm"]            # KeyError when absent, like sit.py:115
        self.num_heads = num_heads
        self.attn = _AttentionParams(hidden_size, num_heads, self.qk_norm)
        if "fused_attn" in block_kwargs:
            self.attn.fused_attn = block_kwargs["fused_attn"]
        self.mlp = _MlpParams(hidden_size, int(hidden_size * mlp_ratio))
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size, 6 * hidden_size, bias=True))

    def forward(self, x, c_act, act_dtype, c_acc=None, ada=None, link_in=None, link_out=None):
        """x: (N,T,D) fp32; c_act = silu(c) already in the act dtype (shared by every block); c_acc: the side
        accumulator the block adds its gradient w.r.t. c_act into (ops.silu_cast); ada: (ops.AdaLNAll, block index) when
        the modulation vectors of all blocks were computed by one grouped GEMM."""
        lin = self.adaLN_modulation[1]
        a, m = self.attn, self.mlp
        return ops.SiTBlockFn.apply(x, c_act, lin.weight, lin.bias, a.qkv.weight, a.qkv.bias, a.proj.weight, a.proj.bias,
                                    m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias, self.num_heads, act_dtype,
                                    getattr(self, "_reed_after_backward", None), c_acc,
                                    *((a.q_norm.weight, a.q_norm.bias, a.k_norm.weight, a.k_norm.bias) if self.qk_norm
                                      else (None, None, None, None)), ada, link_in, link_out)


class FinalLayer(nn.Module):
    def __init__(self, hidden_size, patch_size, out_channels):
        super().__init__()
        self.linear = nn.Linear(hidden_size, patch_size * patch_size * out_channels, bias=True)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size, 2 * hidden_size, bias=True))

    def forward(self, x, c_act, act_dtype):
        n, t, d = x.shape
        lin = self.adaLN_modulation[1]
        mod = ops.linear(c_act, lin.weight, lin.bias, act_dtype=act_dtype, out_dtype=torch.float32)
        half = mod.shape[1] // 2
        xm = ops.LNModulateFn.apply(x, mod[:, :half], mod[:, half:], act_dtype)
        out = ops.linear(xm.view(n * t, d), self.linear.weight, self.linear.bias, act_dtype=act_dtype,
                         out_dtype=torch.float32)
        return out.view(n, t, -1)


# ----------------------------------------------------------------------------------------------------
# the model
# ----------------------------------------------------------------------------------------------------

class SiT(nn.Module):
    def __init__(self, path_type='edm', input_size=32, patch_size=2, in_channels=4, hidden_size=1152,
                 decoder_hidden_size=768, encoder_depth=8, encoder_depth_text=None, depth=28, num_heads=16,
                 mlp_ratio=4.0, class_dropout_prob=0.1, num_classes=1000, use_cfg=False, z_dims=[768],
                 z_types=['i'], projector_dim=2048, **block_kwargs):
        super().__init__()
        self.path_type = path_type
        self.in_channels = in_channels
        self.out_channels = in_channels
        self.patch_size = patch_size
        self.num_heads = num_heads
        self.use_cfg = use_cfg
        self.num_classes = num_classes
        self.z_dims = z_dims
        self.z_types = z_types
        self.encoder_depth = encoder_depth
        self.encoder_depth_text = encoder_depth_text
        self.reed_precision = None          # None: follow torch.autocast; or 'bf16' / 'fp32'

        self.x_embedder = PatchEmbed(input_size, patch_size, in_channels, hidden_size, bias=True)
        self.t_embedder = TimestepEmbedder(hidden_size)
        self.y_embedder = LabelEmbedder(num_classes, hidden_size, class_dropout_prob)
        self.pos_embed = nn.Parameter(torch.zeros(1, self.x_embedder.num_patches, hidden_size), requires_grad=False)
        self.blocks = nn.ModuleList(SiTBlock(hidden_size, num_heads, mlp_ratio=mlp_ratio, **block_kwargs)
                                    for _ in range(depth))
        self.projectors = nn.ModuleList(build_mlp(hidden_size, projector_dim, z) for z in z_dims)
        self.final_layer = FinalLayer(decoder_hidden_size, patch_size, self.out_channels)
        self.initialize_weights()

    # -- init: same RNG draws, in the same order, as sit.py:218-254 ---------------------------------
    def initialize_weights(self):
        for module in self.modules():               # leaves are visited in the same order as Module.apply
            if isinstance(module, nn.Linear):
                nn.init.xavier_uniform_(module.weight)
                if module.bias is not None:
                    nn.init.zeros_(module.bias)
        grid = int(self.x_embedder.num_patches ** 0.5)
        table = get_2d_sincos_pos_embed(self.pos_embed.shape[-1], grid)
        self.pos_embed.data.copy_(torch.from_numpy(table).float().unsqueeze(0))
        conv = self.x_embedder.proj
        nn.init.xavier_uniform_(conv.weight.data.view(conv.weight.shape[0], -1))
        nn.init.zeros_(conv.bias)
        nn.init.normal_(self.y_embedder.embedding_table.weight, std=0.02)
        for idx in (0, 2):
            nn.init.normal_(self.t_embedder.mlp[idx].weight, std=0.02)
        zeroed = [blk.adaLN_modulation[-1] for blk in self.blocks]
        zeroed += [self.final_layer.adaLN_modulation[-1], self.final_layer.linear]
        for lin in zeroed:
            nn.init.zeros_(lin.weight)
            nn.init.zeros_(lin.bias)

    def unpatchify(self, x, patch_size=None):
        """(N, T, p*p*C) -> (N, C, H, W); the 16-vector is ordered (p, q, c)."""
        c = self.out_channels
        p = self.x_embedder.patch_size[0] if patch_size is None else patch_size
        g = int(x.shape[1] ** 0.5)
        assert g * g == x.shape[1]
        return x.reshape(x.shape[0], g, g, p, p, c).permute(0, 5, 1, 3, 2, 4).reshape(x.shape[0], c, g * p, g * p)

    def _act_dtype(self, x):
        if self.reed_precision is not None:
            return {"bf16": torch.bfloat16, "fp32": torch.float32}[self.reed_precision]
        if torch.is_autocast_enabled():
            return torch.bfloat16          # fp16 autocast is served by the bf16 kernels (same tensor-core rate)
        return torch.float32

    def _project(self, k, tokens, act_dtype):
        """REED projector k on tokens ('i') or on the token mean ('t')  (sit.py:292-301)."""
        n, t, d = tokens.shape
        seq = self.projectors[k]
        if self.z_types[k] == 'i':
            h = ops.CastFn.apply(tokens.reshape(n * t, d), act_dtype)
        else:
            h = ops.TokenMeanFn.apply(tokens, act_dtype)
        h = ops.linear(h, seq[0].weight, seq[0].bias, act=ops.ACT_SILU, act_dtype=act_dtype)
        h = ops.linear(h, seq[2].weight, seq[2].bias, act=ops.ACT_SILU, act_dtype=act_dtype)
        z = ops.linear(h, seq[4].weight, seq[4].bias, act_dtype=act_dtype)
        return z.view(n, t, -1) if self.z_types[k] == 'i' else z

    def forward(self, x, t, y, inference=True):
        if not x.is_cuda:
            raise RuntimeError("reed_b200.SiT runs on CUDA (sm_100a) only; there is no CPU fallback. "
                               "Move the model and inputs to a B200 device.")
        act_dtype = self._act_dtype(x)
        with torch.autocast(device_type="cuda", enabled=False):
            return self._forward(x.float(), t, y, inference, act_dtype)

    def _forward(self, x, t, y, inference, act_dtype):
        n = x.shape[0]
        emb = self.x_embedder
        tok = ops.linear(emb.patches(x), emb.proj.weight, emb.proj.bias, act_dtype=act_dtype, out_dtype=torch.float32)
        tok = tok.view(n, emb.num_patches, -1) + self.pos_embed
        width = tok.shape[-1]

        c = self.t_embedder(t, act_dtype) + self.y_embedder(y, self.training)
        c_act, c_acc = ops.silu_cast(c, act_dtype)

        split = self.encoder_depth_text is not None and self.encoder_depth_text != self.encoder_depth
        zs = None
        z_img = z_txt = None
        # adaLN_modulation(c) of all blocks in one grouped GEMM (the weights stay separate parameters)
        ada_lin = [blk.adaLN_modulation[1] for blk in self.blocks]
        # (a trainer whose blocks receive their operands by per-block all-gathers under the forward turns this off: the
        # grouped GEMM would read every block's weights at the start of the step)
        grouped = getattr(self, "_reed_adaln_grouped", True) and ops.AdaLNAll.usable(c_act, ada_lin, act_dtype)
        ada_all = ops.AdaLNAll(c_act, ada_lin, act_dtype, c_acc) if grouped else None
        # consecutive blocks whose residual stream has no other consumer (no projector tap in between) share one fused
        # kernel for the later block's LayerNorm backward and the earlier block's gate backward (ops.BlockLink)
        taps = {self.encoder_depth, self.encoder_depth_text} if not inference else set()
        link_in = None
        depth = len(self.blocks)
        for i, blk in enumerate(self.blocks, start=1):
            link_out = (ops.BlockLink() if (ops._BLOCK_LINK and torch.is_grad_enabled() and tok.requires_grad and i < depth
                                             and i not in taps) else None)
            tok = blk(tok, c_act, act_dtype, c_acc, (ada_all, i - 1) if ada_all is not None else None, link_in, link_out)
            link_in = link_out
            if inference:
                continue
            if i == self.encoder_depth:
                if not split:
                    zs = [self._project(k, tok, act_dtype) for k in range(len(self.projectors))]
                else:
                    for k, kind in enumerate(self.z_types):
                        if kind == 'i':
                            z_img = self._project(k, tok, act_dtype)
            if split and i == self.encoder_depth_text:
                for k, kind in enumerate(self.z_types):
                    if kind == 't':
                        z_txt = self._project(k, tok, act_dtype)
        if not inference and split:
            zs = [z_img, z_txt]

        out = self.final_layer(tok, c_act, act_dtype)
        return self.unpatchify(out), zs


# ----------------------------------------------------------------------------------------------------
# 2-D sin-cos table (MAE convention; float64 maths; column index first)   sit.py:319-366
# ----------------------------------------------------------------------------------------------------

def get_1d_sincos_pos_embed_from_grid(embed_dim, pos):
    assert embed_dim % 2 == 0
    omega = 1.0 / 10000 ** (np.arange(embed_dim // 2, dtype=np.float64) / (embed_dim / 2.0))
    ang = np.asarray(pos).reshape(-1)[:, None] * omega[None, :]
    return np.concatenate([np.sin(ang), np.cos(ang)], axis=1)


def get_2d_sincos_pos_embed_from_grid(embed_dim, grid):
    assert embed_dim % 2 == 0
    return np.concatenate([get_1d_sincos_pos_embed_from_grid(embed_dim // 2, grid[0]),
                           get_1d_sincos_pos_embed_from_grid(embed_dim // 2, grid[1])], axis=1)


def get_2d_sincos_pos_embed(embed_dim, grid_size, cls_token=False, extra_tokens=0):
    axis = np.arange(grid_size, dtype=np.float32)
    grid = np.stack(np.meshgrid(axis, axis), axis=0).reshape(2, 1, grid_size, grid_size)   # [0] = column index
    table = get_2d_sincos_pos_embed_from_grid(embed_dim, grid)
    if cls_token and extra_tokens > 0:
        table = np.concatenate([np.zeros([extra_tokens, embed_dim]), table], axis=0)
    return table


# ----------------------------------------------------------------------------------------------------
# model zoo   sit.py:373-415
# ----------------------------------------------------------------------------------------------------

_FAMILY = {"XL": dict(depth=28, hidden_size=1152, num_heads=16), "L": dict(depth=24, hidden_size=1024, num_heads=16),
           "B": dict(depth=12, hidden_size=768, num_heads=12), "S": dict(depth=12, hidden_size=384, num_heads=6)}


def _make(family, patch):
    base = dict(_FAMILY[family], patch_size=patch)
    if family != "S":                       # the reference SiT-S constructors do not set decoder_hidden_size
        base["decoder_hidden_size"] = base["hidden_size"]

    def ctor(**kwargs):
        return SiT(**base, **kwargs)
    ctor.__name__ = f"SiT_{family}_{patch}"
    return ctor


SiT_models = {}
for _fam in ("XL", "L", "B", "S"):
    for _patch in (2, 4, 8):
        _fn = _make(_fam, _patch)
        globals()[_fn.__name__] = _fn
        SiT_models[f"SiT-{_fam}/{_patch}"] = _fn
