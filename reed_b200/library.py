"""``torch.library`` registration of the C-ABI kernels: ``torch.ops.reed.*`` custom ops with fake (meta) kernels and
autograd formulas, so that callers of the hot path can be traced (``torch.compile`` / ``torch.export`` see opaque ops with
known output shapes instead of ctypes calls) and checked with ``torch.library.opcheck``.

The ctypes C-ABI stays underneath: every op body is one or two calls into libreed_sm100.so through ``reed_b200.ops``.

  differentiable ops (register_autograd)            raw ops they are built from
  ------------------------------------------------  ----------------------------------------------------------------
  reed::velocity_mse(pred, x, eps, t, path)         reed::velocity_mse_bwd                (loss.py:175-189)
  reed::cosine_align(z_tilde, z) -> (align, stats)  reed::cosine_align_bwd                (loss.py:204-225)
  reed::ln_modulate(x, shift, scale, bf16)          reed::ln_modulate_bwd                 (sit.py:26-27,153-155)
  reed::attention(qkv, B, T, H, hd) -> (o, lse)     reed::attention_bwd                   (timm Attention, sit.py:13,134)
  reed::linear(x, w, bias, act, out_bf16)           reed::gemm_nt / reed::act_bwd         (every nn.Linear of sit.py)
  not differentiable (data in, data out)            reed::siloss_interp, reed::sampler_step, reed::sampler_cast

SiTBlockFn / LinearFn in ``ops.py`` stay ``torch.autograd.Function``s on purpose: in trainer mode their backward writes
weight gradients straight into the trainer's flat buckets and reads bf16 shadows that hang off the ``nn.Parameter`` objects
- Python-side state that an operator schema (tensors and scalars only) cannot carry.  Their arithmetic is the same kernels.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import ops

_DT = {0: torch.float32, 1: torch.bfloat16}


def _act_dtype(bf16: bool) -> torch.dtype:
    return torch.bfloat16 if bf16 else torch.float32


# ------------------------------------------------------------------------------------------------------------------
# SILoss pieces
# ------------------------------------------------------------------------------------------------------------------

@torch.library.custom_op("reed::siloss_interp", mutates_args=(), device_types="cuda")
def siloss_interp(x: Tensor, eps: Tensor, t: Tensor, path_type: int) -> Tensor:
    return ops._interpolate_raw(x, eps, t, path_type)


@siloss_interp.register_fake
def _(x, eps, t, path_type):
    return torch.empty_like(x, memory_format=torch.contiguous_format)


@torch.library.custom_op("reed::velocity_mse", mutates_args=(), device_types="cuda")
def velocity_mse(pred: Tensor, x: Tensor, eps: Tensor, t: Tensor, path_type: int) -> Tensor:
    return ops._mse_fwd_raw(pred, x, eps, t, path_type)


@velocity_mse.register_fake
def _(pred, x, eps, t, path_type):
    return pred.new_empty((pred.shape[0],), dtype=torch.float32)


@torch.library.custom_op("reed::velocity_mse_bwd", mutates_args=(), device_types="cuda")
def velocity_mse_bwd(pred: Tensor, x: Tensor, eps: Tensor, t: Tensor, g: Tensor, path_type: int) -> Tensor:
    return ops._mse_bwd_raw(pred, x, eps, t, g, path_type)


@velocity_mse_bwd.register_fake
def _(pred, x, eps, t, g, path_type):
    return torch.empty_like(pred, memory_format=torch.contiguous_format)


def _mse_setup(ctx, inputs, output):
    pred, x, eps, t, path_type = inputs
    ctx.save_for_backward(pred, x, eps, t)
    ctx.path_type = path_type


def _mse_backward(ctx, g):
    pred, x, eps, t = ctx.saved_tensors
    return torch.ops.reed.velocity_mse_bwd(pred, x, eps, t, g, ctx.path_type), None, None, None, None


velocity_mse.register_autograd(_mse_backward, setup_context=_mse_setup)


@torch.library.custom_op("reed::cosine_align", mutates_args=(), device_types="cuda")
def cosine_align(z_tilde: Tensor, z: Tensor) -> Tuple[Tensor, Tensor]:
    return ops._cos_fwd_raw(z_tilde, z)


@cosine_align.register_fake
def _(z_tilde, z):
    B, T, _Z = z_tilde.shape
    return z_tilde.new_empty((B,), dtype=torch.float32), z_tilde.new_empty((B * T, 3), dtype=torch.float32)


@torch.library.custom_op("reed::cosine_align_bwd", mutates_args=(), device_types="cuda")
def cosine_align_bwd(z_tilde: Tensor, z: Tensor, stats: Tensor, g: Tensor) -> Tensor:
    return ops._cos_bwd_raw(z_tilde, z, stats, g)


@cosine_align_bwd.register_fake
def _(z_tilde, z, stats, g):
    return torch.empty_like(z_tilde, memory_format=torch.contiguous_format)


def _cos_setup(ctx, inputs, output):
    z_tilde, z = inputs
    ctx.save_for_backward(z_tilde, z, output[1])


def _cos_backward(ctx, g_align, g_stats):
    z_tilde, z, stats = ctx.saved_tensors
    return torch.ops.reed.cosine_align_bwd(z_tilde, z, stats, g_align), None


cosine_align.register_autograd(_cos_backward, setup_context=_cos_setup)


# ------------------------------------------------------------------------------------------------------------------
# LayerNorm + modulate
# ------------------------------------------------------------------------------------------------------------------

@torch.library.custom_op("reed::ln_modulate", mutates_args=(), device_types="cuda")
def ln_modulate(x: Tensor, shift: Tensor, scale: Tensor, bf16: bool) -> Tuple[Tensor, Tensor, Tensor]:
    """x [B,T,D] fp32, shift / scale [B,D] fp32 -> (modulate(LayerNorm(x)) in the act dtype, row mean, row rstd)."""
    B, T, D = x.shape
    out, mean, rstd = ops.ln_modulate_fwd(x.contiguous().view(B * T, D), shift.contiguous(), scale.contiguous(), T,
                                          _act_dtype(bf16))
    return out.view(B, T, D), mean, rstd


@ln_modulate.register_fake
def _(x, shift, scale, bf16):
    B, T, D = x.shape
    return (x.new_empty((B, T, D), dtype=_act_dtype(bf16)), x.new_empty((B * T,), dtype=torch.float32),
            x.new_empty((B * T,), dtype=torch.float32))


@torch.library.custom_op("reed::ln_modulate_bwd", mutates_args=(), device_types="cuda")
def ln_modulate_bwd(dout: Tensor, x: Tensor, mean: Tensor, rstd: Tensor, scale: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    B, T, D = x.shape
    scale = scale.contiguous()
    dshift = torch.zeros((B, D), device=x.device, dtype=torch.float32)     # separate storages: op outputs may not alias
    dscale = torch.zeros((B, D), device=x.device, dtype=torch.float32)
    dx = ops.ln_modulate_bwd(dout.contiguous().view(B * T, D), x.contiguous().view(B * T, D), mean, rstd, scale, T, None,
                             dshift, dscale)
    return dx.view(B, T, D), dshift, dscale


@ln_modulate_bwd.register_fake
def _(dout, x, mean, rstd, scale):
    B, T, D = x.shape
    return (torch.empty_like(x, memory_format=torch.contiguous_format), x.new_empty((B, D), dtype=torch.float32),
            x.new_empty((B, D), dtype=torch.float32))


def _lnm_setup(ctx, inputs, output):
    x, shift, scale, bf16 = inputs
    ctx.save_for_backward(x, output[1], output[2], scale)


def _lnm_backward(ctx, dout, _dmean, _drstd):
    x, mean, rstd, scale = ctx.saved_tensors
    dx, dshift, dscale = torch.ops.reed.ln_modulate_bwd(dout, x, mean, rstd, scale)
    return dx, dshift, dscale, None


ln_modulate.register_autograd(_lnm_backward, setup_context=_lnm_setup)


# ------------------------------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------------------------------

@torch.library.custom_op("reed::attention", mutates_args=(), device_types="cuda")
def attention(qkv: Tensor, B: int, T: int, H: int, hd: int) -> Tuple[Tensor, Tensor]:
    """qkv [B*T, 3*H*hd] packed (3, H, hd) -> (context [B*T, H*hd], row log-sum-exp [B, H, T] fp32)."""
    return ops.attention_fwd(qkv.contiguous(), B, T, H, hd)


@attention.register_fake
def _(qkv, B, T, H, hd):
    return qkv.new_empty((B * T, H * hd)), qkv.new_empty((B, H, T), dtype=torch.float32)


@torch.library.custom_op("reed::attention_bwd", mutates_args=(), device_types="cuda")
def attention_bwd(qkv: Tensor, o: Tensor, d_o: Tensor, lse: Tensor, B: int, T: int, H: int, hd: int) -> Tensor:
    return ops.attention_bwd(qkv.contiguous(), o, d_o.contiguous(), lse, B, T, H, hd)


@attention_bwd.register_fake
def _(qkv, o, d_o, lse, B, T, H, hd):
    return torch.empty_like(qkv, memory_format=torch.contiguous_format)


def _attn_setup(ctx, inputs, output):
    qkv, B, T, H, hd = inputs
    ctx.save_for_backward(qkv, output[0], output[1])
    ctx.dims = (B, T, H, hd)


def _attn_backward(ctx, d_o, _dlse):
    qkv, o, lse = ctx.saved_tensors
    return torch.ops.reed.attention_bwd(qkv, o, d_o, lse, *ctx.dims), None, None, None, None


attention.register_autograd(_attn_backward, setup_context=_attn_setup)


# ------------------------------------------------------------------------------------------------------------------
# linear (functional form: the weight is given in the activation dtype)
# ------------------------------------------------------------------------------------------------------------------

@torch.library.custom_op("reed::gemm_nt", mutates_args=(), device_types="cuda")
def gemm_nt(a: Tensor, b: Tensor, bias: Optional[Tensor], a_mn: bool, b_mn: bool, out_bf16: bool) -> Tensor:
    """epi-free product out[M,N] = A . B^T (+ bias) in the C-ABI's operand conventions (reed_gemm)."""
    return ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, bias=bias, out_dtype=_act_dtype(out_bf16))


@gemm_nt.register_fake
def _(a, b, bias, a_mn, b_mn, out_bf16):
    M = a.shape[1] if a_mn else a.shape[0]
    N = b.shape[1] if b_mn else b.shape[0]
    return a.new_empty((M, N), dtype=_act_dtype(out_bf16))


@torch.library.custom_op("reed::linear", mutates_args=(), device_types="cuda")
def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor], act: int, out_bf16: bool) -> Tuple[Tensor, Tensor]:
    """y = act(x W^T + b) with act 0 none / 1 GELU(tanh) / 2 SiLU; also returns the pre-activation the backward needs
    (an empty tensor when act = 0).  x [M,K] and W [N,K] share a dtype (fp32 or bf16); bias fp32."""
    out_dtype = _act_dtype(out_bf16)
    if act == ops.ACT_NONE:
        return ops.gemm(x, weight, out_dtype=out_dtype, bias=bias), x.new_empty((0,))
    h = torch.empty((x.shape[0], weight.shape[0]), device=x.device, dtype=x.dtype)
    y = ops.gemm(x, weight, out_dtype=out_dtype, bias=bias, epilogue=ops.EPI_GELU if act == ops.ACT_GELU else ops.EPI_SILU, out2=h)
    return y, h


@linear.register_fake
def _(x, weight, bias, act, out_bf16):
    y = x.new_empty((x.shape[0], weight.shape[0]), dtype=_act_dtype(out_bf16))
    return y, (x.new_empty((0,)) if act == 0 else x.new_empty((x.shape[0], weight.shape[0])))


@torch.library.custom_op("reed::act_bwd", mutates_args=(), device_types="cuda")
def act_bwd(dy: Tensor, h: Tensor, act: int) -> Tensor:
    return ops.act_bwd(dy, h, act)


@act_bwd.register_fake
def _(dy, h, act):
    return torch.empty_like(dy, memory_format=torch.contiguous_format)


@torch.library.custom_op("reed::colsum", mutates_args=(), device_types="cuda")
def colsum(src: Tensor) -> Tensor:
    out = torch.zeros(src.shape[1], device=src.device, dtype=torch.float32)
    ops.colsum(src.contiguous(), out)
    return out


@colsum.register_fake
def _(src):
    return src.new_empty((src.shape[1],), dtype=torch.float32)


def _linear_setup(ctx, inputs, output):
    x, weight, bias, act, out_bf16 = inputs
    ctx.save_for_backward(x, weight, output[1])
    ctx.act, ctx.has_bias = act, bias is not None


def _linear_backward(ctx, dy, _dh):
    x, weight, h = ctx.saved_tensors
    bf16 = x.dtype == torch.bfloat16
    dy = dy.contiguous().to(x.dtype)
    if ctx.act != ops.ACT_NONE:
        dy = torch.ops.reed.act_bwd(dy, h, ctx.act)
    dx = torch.ops.reed.gemm_nt(dy, weight, None, False, True, bf16)                # dy W
    dw = torch.ops.reed.gemm_nt(dy, x, None, True, True, False).to(weight.dtype)    # dy^T x
    db = torch.ops.reed.colsum(dy) if ctx.has_bias else None
    return dx, dw, db, None, None


linear.register_autograd(_linear_backward, setup_context=_linear_setup)


# ------------------------------------------------------------------------------------------------------------------
# sampler step (samplers.py:61-104, 124-187): data in, data out
# ------------------------------------------------------------------------------------------------------------------

@torch.library.custom_op("reed::sampler_cast", mutates_args=(), device_types="cuda")
def sampler_cast(x64: Tensor, bf16: bool, dup: bool) -> Tensor:
    return ops.sampler_cast(x64.contiguous(), _act_dtype(bf16), dup)


@sampler_cast.register_fake
def _(x64, bf16, dup):
    shape = (x64.shape[0] * (2 if dup else 1),) + tuple(x64.shape[1:])
    return x64.new_empty(shape, dtype=_act_dtype(bf16))


OPS = ("siloss_interp", "velocity_mse", "velocity_mse_bwd", "cosine_align", "cosine_align_bwd", "ln_modulate", "ln_modulate_bwd",
       "attention", "attention_bwd", "gemm_nt", "linear", "act_bwd", "colsum", "sampler_cast")
