"""Tensor-level wrappers and autograd functions over the C-ABI kernels (libreed_sm100.so).

PyTorch is plumbing here: device memory (caching allocator), streams, autograd bookkeeping.  Every arithmetic
step of the hot path is one of the hand-written kernels; there is no PyTorch/CPU fallback.

Precision modes (`act dtype`):
  * torch.float32  - fp32 activations, SIMT fp32 GEMM/attention; parity bar 1e-5 relative.
  * torch.bfloat16 - bf16 activations and GEMM operands (tcgen05, fp32 accumulate), fp32 residual stream,
                     fp32 master weights with bf16 shadow copies; parity bar 2e-2 relative.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _cabi
from ._cabi import call

F32, BF16 = 0, 1
EPI_NONE, EPI_GELU, EPI_SILU, EPI_GATE_RES, EPI_DGELU, EPI_DSILU = range(6)
ACT_NONE, ACT_GELU, ACT_SILU = 0, 1, 2
BACKEND_AUTO, BACKEND_SIMT, BACKEND_TENSOR, BACKEND_TENSOR_CG1, BACKEND_TENSOR_CG2 = 0, 1, 2, 3, 4

import ctypes as _ctypes
import os as _os

# profiling knob: 0 = bias gradients of qkv / fc1 by a separate column-sum pass instead of the ones-column GEMM
_WGRAD_BIAS = _os.environ.get("REED_WGRAD_BIAS", "1") != "0"
# Weight-gradient GEMMs of a transformer block go to a second stream (trainer mode).  They depend only on dy and the saved
# activations, not on the dgrad chain, so their CTAs fill the SMs a dgrad GEMM leaves idle in its last partial round
# (N = 1152 shapes run 160 tiles on 74 CTA pairs: the third round is 16 % full).  Measured on the B200 in round 2:
# 909 -> 927 img/s on the XL/2 step (profiles/r02_bench_flags.txt); REED_WGRAD_STREAM=0 restores the single stream for A/B.
_WGRAD_STREAM = _os.environ.get("REED_WGRAD_STREAM", "1") != "0"
# The adaLN modulation linears of all blocks as one grouped GEMM (AdaLNAll); REED_ADALN_GROUPED=0 restores one GEMM per
# block for A/B runs.
_ADALN_GROUPED = _os.environ.get("REED_ADALN_GROUPED", "1") != "0"
_gemm_backend = BACKEND_AUTO
_attn_backend = BACKEND_AUTO
launch_count = 0     # kernels launched through this module (bench.py reports it as gpu_launches)


def set_backends(gemm: int = BACKEND_AUTO, attention: int = BACKEND_AUTO):
    """Test/debug knob: force the SIMT kernels or require the tensor-core kernels."""
    global _gemm_backend, _attn_backend
    _gemm_backend, _attn_backend = gemm, attention


def _code(dtype: torch.dtype) -> int:
    if dtype == torch.float32:
        return F32
    if dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"reed_b200 kernels take float32 or bfloat16 tensors, got {dtype}")


def _stream() -> int:
    """Raw cudaStream_t of torch's current stream on the current device (the stream being captured, under graph capture)."""
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("reed_b200 runs on CUDA (sm_100a) tensors only; there is no CPU fallback")


def _launch(name, *args, n=1):
    global launch_count
    launch_count += n
    call(name, *args)


# --------------------------------------------------------------------------------------------------
# weights: fp32 master -> bf16 shadow
# --------------------------------------------------------------------------------------------------

def cast(x: torch.Tensor, dtype: torch.dtype, op: int = 0) -> torch.Tensor:
    """dtype cast (op=0), SiLU+cast (op=1) or exact erf-GELU+cast (op=2) through reed_unary."""
    if op == 0 and x.dtype == dtype:
        return x
    x = x.contiguous()
    out = torch.empty_like(x, dtype=dtype)
    n = x.numel()
    if n % 4:
        raise ValueError("reed_unary needs a multiple of 4 elements")
    _launch("reed_unary", _p(x), _code(x.dtype), _p(out), _code(dtype), op, n, _stream())
    return out


# Bumped by whoever rewrites parameter storage through raw pointers (the fused clip/AdamW/EMA kernels, graph replays of
# the train step, checkpoint loads): tensor._version does not see those writes, so lazily created bf16 shadows - the EMA
# model's, above all - are also keyed on this counter.  Shadows the optimizer kernel itself refreshes are "managed".
weights_epoch = 0


def bump_weights_epoch():
    global weights_epoch
    weights_epoch += 1


def weight_for(p: torch.Tensor, act_dtype: torch.dtype) -> torch.Tensor:
    """The tensor the GEMM reads for parameter ``p``: the fp32 master, or its bf16 shadow (refreshed when stale)."""
    w = p.detach()
    if act_dtype == torch.float32:
        return w
    shadow = getattr(p, "_reed_shadow", None)
    if (shadow is not None and getattr(p, "_reed_shadow_version", -1) == p._version and shadow.device == p.device
            and (getattr(p, "_reed_shadow_managed", False) or getattr(p, "_reed_shadow_epoch", -1) == weights_epoch)):
        return shadow
    w = w.contiguous()
    if shadow is None or shadow.shape != w.shape or shadow.device != w.device:
        shadow = torch.empty_like(w, dtype=torch.bfloat16)
    _launch("reed_unary", _p(w), F32, _p(shadow), BF16, 0, w.numel(), _stream())
    p._reed_shadow = shadow
    p._reed_shadow_version = p._version
    p._reed_shadow_epoch = weights_epoch
    return shadow


def refresh_shadows(module) -> int:
    """Re-cast (in place, same storage) every stale bf16 shadow of ``module``'s parameters; returns how many were stale.
    CUDA graphs that baked shadow pointers in (generate.GraphedSiT) call this before a replay."""
    n = 0
    for p in module.parameters():
        if getattr(p, "_reed_shadow", None) is None:
            continue
        before = p._reed_shadow
        launched = launch_count
        assert weight_for(p, torch.bfloat16) is before, "a shadow changed storage: graphs holding its pointer are invalid"
        n += launch_count - launched
    return n


def _grad_target(p: torch.Tensor):
    """(buffer, accumulate) when a trainer owns flat gradient storage for ``p`` (see trainer.FlatState)."""
    main = getattr(p, "_reed_main_grad", None)
    if main is None:
        return None, False
    fresh = getattr(p, "_reed_grad_fresh", True)
    p._reed_grad_fresh = False
    return main, not fresh


# --------------------------------------------------------------------------------------------------
# raw kernel wrappers
# --------------------------------------------------------------------------------------------------

def gemm(a, b, *, a_mn=False, b_mn=False, out=None, out_dtype=torch.float32, epilogue=EPI_NONE, bias=None, aux=None,
         gate=None, rows_per_group=1, out2=None, accumulate=False):
    """out[M,N] = epi(A . B^T).  ``a`` is [M,K] (or [K,M] when a_mn), ``b`` is [N,K] (or [K,N] when b_mn)."""
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1, "operands must be row-major 2-D"
    assert a.dtype == b.dtype
    M, K = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
    assert K == Kb, f"contraction mismatch {K} vs {Kb}"
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=out_dtype)
    assert out.shape == (M, N) and out.stride(1) == 1
    _launch("reed_gemm", _code(a.dtype), _p(a), a.stride(0), int(a_mn), _p(b), b.stride(0), int(b_mn), _p(out),
            out.stride(0), _code(out.dtype), M, N, K, epilogue, _p(bias), _p(aux), aux.stride(0) if aux is not None else 0,
            _p(gate), gate.stride(0) if gate is not None else 0, rows_per_group, _p(out2),
            out2.stride(0) if out2 is not None else 0, int(accumulate), _gemm_backend, _stream())
    return out


def colsum(src: torch.Tensor, out: torch.Tensor):
    """out[n] += sum_m src[m, n]  (out fp32, pre-initialised)."""
    M, N = src.shape
    _launch("reed_colsum", _p(src), _code(src.dtype), src.stride(0), _p(out), M, N, _stream())


def act_bwd(dy, h, act):
    dy = dy.contiguous()
    dx = torch.empty_like(dy)
    _launch("reed_act_bwd", _p(dy), _code(dy.dtype), _p(h), _code(h.dtype), _p(dx), 1 if act == ACT_GELU else 2,
            dy.numel(), _stream())
    return dx


def ln_modulate_fwd(x, shift, scale, rows_per_group, act_dtype, eps=1e-6, ones_col=False):
    """ones_col (bf16 only): rows get spare elements with [1,0,...,0] at columns D..D+7 - the operand form wgrad_bias()
    contracts against.  The pitch grows by 64 elements (128 bytes) so that rows stay 128-byte aligned for the TMA boxes.
    Returns the [M, D] view; its storage is ``out._base``."""
    M, D = x.shape
    ones_col = ones_col and _WGRAD_BIAS and act_dtype == torch.bfloat16
    ld_out = D + 64 if ones_col else D
    buf = torch.empty((M, ld_out), device=x.device, dtype=act_dtype)
    mean = torch.empty((M,), device=x.device, dtype=torch.float32)
    rstd = torch.empty((M,), device=x.device, dtype=torch.float32)
    assert shift.stride(0) == scale.stride(0) and shift.stride(1) == 1 and scale.stride(1) == 1
    _launch("reed_ln_modulate_fwd", _p(x), _p(shift), _p(scale), shift.stride(0), rows_per_group, _p(buf), ld_out,
            _code(act_dtype), _p(mean), _p(rstd), M, D, eps, _stream())
    return (buf[:, :D] if ones_col else buf), mean, rstd


def wgrad_bias(dy2d, x_ext, k_in, dw, db, accumulate):
    """dW (+)= dy^T x and db += colsum(dy) in one GEMM; x_ext [tokens, >= K+8] carries the ones column at index K."""
    tokens, n_out = dy2d.shape
    assert x_ext.shape[1] >= k_in + 8
    assert dy2d.dtype == x_ext.dtype == torch.bfloat16 and dy2d.stride(1) == 1 and x_ext.stride(1) == 1
    assert dw.shape == (n_out, k_in) and dw.stride(1) == 1 and db.numel() == n_out and db.is_contiguous()
    _launch("reed_gemm_wgrad_bias", _p(dy2d), dy2d.stride(0), _p(x_ext), x_ext.stride(0), _p(dw), dw.stride(0), _p(db),
            n_out, k_in, tokens, int(accumulate), _stream())


def _ptr_table(tensors):
    """Host array of device pointers (reed_gemm_grouped reads it during the call only)."""
    return (_ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def gemm_grouped_fwd(a, weights, bias_all, out):
    """out[M, G*n] = a . [W_0; W_1; ...]^T + bias_all in one tensor-core launch (reed_gemm_grouped mode 0): G linears with
    separately stored [n, K] weights applied to the same bf16 input (the adaLN modulation of every block)."""
    M, K = a.shape
    n = weights[0].shape[0]
    assert a.dtype == torch.bfloat16 and a.stride(1) == 1 and out.dtype == torch.float32 and out.stride(1) == 1
    assert all(w.dtype == torch.bfloat16 and w.shape == (n, K) and w.stride(1) == 1 and w.stride(0) == weights[0].stride(0)
               for w in weights)
    assert out.shape == (M, n * len(weights)) and bias_all.numel() == n * len(weights) and bias_all.dtype == torch.float32
    _launch("reed_gemm_grouped", 0, _ptr_table([a]), a.stride(0), _ptr_table(weights), weights[0].stride(0), len(weights), n,
            _p(out), out.stride(0), M, n * len(weights), K, _p(bias_all), 0, _stream())
    return out


def gemm_grouped_dgrad(dys, weights, out, accumulate):
    """out[M, K] (+)= sum_g dy_g[M, n] . W_g[n, K]  (reed_gemm_grouped mode 1): the gradient of the shared input of G linears."""
    M, n = dys[0].shape
    K = weights[0].shape[1]
    assert all(d.dtype == torch.bfloat16 and d.shape == (M, n) and d.stride(1) == 1 and d.stride(0) == dys[0].stride(0) for d in dys)
    assert all(w.dtype == torch.bfloat16 and w.shape == (n, K) and w.stride(1) == 1 and w.stride(0) == weights[0].stride(0)
               for w in weights)
    assert out.shape == (M, K) and out.dtype == torch.float32 and out.stride(1) == 1 and len(dys) == len(weights)
    _launch("reed_gemm_grouped", 1, _ptr_table(dys), dys[0].stride(0), _ptr_table(weights), weights[0].stride(0), len(weights), n,
            _p(out), out.stride(0), M, K, n * len(weights), None, int(accumulate), _stream())
    return out


def ln_modulate_bwd(dout, x, mean, rstd, scale, rows_per_group, dres, dshift, dscale):
    """Returns dx = dres + LN'(dout); accumulates into dshift/dscale (two views with one row stride)."""
    M, D = x.shape
    dx = torch.empty_like(x)
    assert dshift.stride(0) == dscale.stride(0)
    _launch("reed_ln_modulate_bwd", _p(dout), _code(dout.dtype), _p(x), _p(mean), _p(rstd), _p(scale), scale.stride(0),
            dshift.stride(0), rows_per_group, _p(dres), _p(dx), _p(dshift), _p(dscale), M, D, _stream())
    return dx


def ln_modulate_gate_bwd(dout, x, mean, rstd, scale, rows_per_group, dres, dshift, dscale, y, gate, dgate, dbias):
    """ln_modulate_bwd followed by gate_bwd on its result, in one pass.  Returns (dx fp32, dy act dtype)."""
    M, D = x.shape
    dx = torch.empty_like(x)
    dy = torch.empty_like(y)
    ld = scale.stride(0)
    assert ld == gate.stride(0) and dshift.stride(0) == dscale.stride(0) == dgate.stride(0)
    _launch("reed_ln_modulate_gate_bwd", _p(dout), _code(dout.dtype), _p(x), _p(mean), _p(rstd), _p(scale), ld,
            dshift.stride(0), rows_per_group, _p(dres), _p(dx), _p(dshift), _p(dscale), _p(y), _p(gate), _p(dy), _p(dgate), _p(dbias), M, D,
            _stream())
    return dx, dy


def gate_bwd(dxn, y, gate, rows_per_group, dgate, dbias):
    M, D = dxn.shape
    dy = torch.empty_like(y)
    _launch("reed_gate_bwd", _p(dxn), _p(y), _code(y.dtype), _p(gate), gate.stride(0), dgate.stride(0), rows_per_group, _p(dy),
            _p(dgate),
            _p(dbias), M, D, _stream())
    return dy


def qk_norm_fwd(qkv, wq, bq, wk, bk, rows, H, hd, eps=1e-5):
    """LayerNorm(head_dim) of the q and k thirds of the packed qkv (timm Attention q_norm/k_norm); v copied through."""
    out = torch.empty_like(qkv)
    stats = torch.empty((rows, H, 2, 2), device=qkv.device, dtype=torch.float32)
    _launch("reed_qk_norm_fwd", _p(qkv), _code(qkv.dtype), _p(wq), _p(bq), _p(wk), _p(bk), _p(out), _p(stats), rows, H, hd,
            eps, _stream())
    return out, stats


def qk_norm_bwd(dout, qkv, stats, wq, wk, dwq, dbq, dwk, dbk, rows, H, hd):
    dqkv = torch.empty_like(qkv)
    _launch("reed_qk_norm_bwd", _p(dout), _code(dout.dtype), _p(qkv), _p(stats), _p(wq), _p(wk), _p(dqkv), _p(dwq), _p(dbq),
            _p(dwk), _p(dbk), rows, H, hd, _stream())
    return dqkv


def attention_fwd(qkv, B, T, H, hd):
    o = torch.empty((B * T, H * hd), device=qkv.device, dtype=qkv.dtype)
    lse = torch.empty((B, H, T), device=qkv.device, dtype=torch.float32)
    _launch("reed_attn_fwd", _code(qkv.dtype), _p(qkv), _p(o), _p(lse), B, T, H, hd, _attn_backend, _stream())
    return o, lse


def attention_bwd(qkv, o, d_o, lse, B, T, H, hd):
    dqkv = torch.empty_like(qkv)
    delta = torch.empty((B, H, T), device=qkv.device, dtype=torch.float32)
    _launch("reed_attn_bwd", _code(qkv.dtype), _p(qkv), _p(o), _p(d_o), _p(lse), _p(dqkv), _p(delta), B, T, H, hd,
            _attn_backend, _stream(), n=2)
    return dqkv


# --------------------------------------------------------------------------------------------------
# autograd functions
# --------------------------------------------------------------------------------------------------

class _SideStream:
    """Second stream for the weight-gradient GEMMs of one block backward (REED_WGRAD_STREAM=1, trainer mode only)."""
    _streams = {}

    def __init__(self, device):
        key = (device.type, device.index)
        if key not in _SideStream._streams:
            _SideStream._streams[key] = torch.cuda.Stream(device=device)
        self.stream = _SideStream._streams[key]
        self.keep = []            # operands stay referenced until the join: the allocator must not hand them out meanwhile
        self.used = False

    def run(self, fn, *operands):
        """Enqueue fn() on the side stream behind everything the current stream has enqueued so far."""
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        self.stream.wait_event(ready)
        with torch.cuda.stream(self.stream):
            out = fn()
        self.keep.extend(operands)
        self.used = True
        return out

    def join(self):
        if self.used:
            done = torch.cuda.Event()
            done.record(self.stream)
            torch.cuda.current_stream().wait_event(done)
        self.keep.clear()
        self.used = False


def _flat_grads(*params):
    """True when a trainer owns flat gradient storage for every given parameter (the kernels write there, autograd sees None)."""
    return all(p is None or getattr(p, "_reed_main_grad", None) is not None for p in params)


def _weight_grad(p, dy2d, x2d):
    """dW[N,K] = dy^T x written into the trainer's flat gradient when present; returns what autograd should see."""
    main, acc = _grad_target(p)
    if main is not None:
        gemm(dy2d, x2d, a_mn=True, b_mn=True, out=main.view(dy2d.shape[1], x2d.shape[1]), accumulate=acc)
        return None
    return gemm(dy2d, x2d, a_mn=True, b_mn=True, out_dtype=torch.float32).view(p.shape)


def _weight_and_bias_grad(w, b, dy2d, x2d, x_ext):
    """Gradients of y = x W^T + b.  With a trainer's flat gradient storage and an activation matrix that carries the ones
    column, both come out of one GEMM; otherwise the weight-gradient GEMM plus a column-sum pass."""
    wmain, wacc = (getattr(w, "_reed_main_grad", None), False)
    bmain = getattr(b, "_reed_main_grad", None) if b is not None else None
    if x_ext is not None and wmain is not None and bmain is not None and dy2d.dtype == torch.bfloat16:
        wmain, wacc = _grad_target(w)
        bmain, bacc = _grad_target(b)
        if not bacc:
            _zero_fresh(b, bmain)
        wgrad_bias(dy2d, x_ext, x2d.shape[1], wmain.view(dy2d.shape[1], x2d.shape[1]), bmain, wacc)
        return None, None
    db = _bias_grad(b, dy2d) if b is not None else None
    dw = _weight_grad(w, dy2d, x2d)
    return dw, db


def _zero_fresh(p, main):
    """First touch of an accumulate-into gradient this step: zero it - together with every other 1-D gradient of the
    same bucket when the trainer laid them out contiguously (one fill per block instead of five)."""
    grp = getattr(p, "_reed_zero_group", None)
    if grp is None:
        main.zero_()
    elif not grp.zeroed:
        grp.view.zero_()
        grp.zeroed = True


def _bias_grad(p, dy2d):
    if p is None:
        return None
    main, acc = _grad_target(p)
    if main is not None:
        if not acc:
            _zero_fresh(p, main)
        colsum(dy2d, main)
        return None
    out = torch.zeros(dy2d.shape[1], device=dy2d.device, dtype=torch.float32)
    colsum(dy2d, out)
    return out


OUTER_WGRAD_MAX_ROWS = 64      # reed_outer_wgrad stages the whole batch in shared memory


def outer_wgrad(dy, x, dw, db, accumulate):
    """dw[N,K] (+)= dy^T x and db[N] += colsum(dy) for a handful of rows (the batch is the contraction): reed_outer_wgrad.
    dy [B,N] fp32 or bf16, x [B,K]; db may be None."""
    B, N = dy.shape
    K = x.shape[1]
    assert x.shape[0] == B and dw.shape == (N, K) and dw.dtype == torch.float32 and dw.stride(1) == 1
    assert dy.stride(1) == 1 and x.stride(1) == 1 and (db is None or (db.dtype == torch.float32 and db.is_contiguous()))
    _launch("reed_outer_wgrad", _p(dy), _code(dy.dtype), dy.stride(0), _p(x), _code(x.dtype), x.stride(0), _p(dw), dw.stride(0),
            _p(db), N, K, B, int(accumulate), _stream())


def _outer_weight_and_bias_grad(w, b, dy, x):
    """Gradients of a Linear applied to a [B <= 64, K] input, both from one outer-product pass over the fp32 dy."""
    wmain, wacc = _grad_target(w)
    dw = wmain.view(w.shape) if wmain is not None else torch.empty(w.shape, device=dy.device, dtype=torch.float32)
    bmain, bacc = _grad_target(b)
    if bmain is not None:
        if not bacc:
            _zero_fresh(b, bmain)
        dbuf = bmain
    else:
        dbuf = torch.zeros(b.shape, device=dy.device, dtype=torch.float32)
    outer_wgrad(dy, x, dw, dbuf, wacc)
    return (None if wmain is not None else dw), (None if bmain is not None else dbuf)


class LinearFn(torch.autograd.Function):
    """y = act(x W^T + b).  x: [M,K] fp32 or act dtype; W: fp32 master [N,K]; output dtype selectable."""

    @staticmethod
    def forward(ctx, x, weight, bias, act, act_dtype, out_dtype):
        _require_cuda(x, weight)
        x_in_dtype = x.dtype
        xa = cast(x, act_dtype)
        w2 = weight_for(weight, act_dtype).view(weight.shape[0], -1)
        b = bias.detach() if bias is not None else None
        if act == ACT_NONE:
            y = gemm(xa, w2, out_dtype=out_dtype, bias=b)
            h = None
        else:
            h = torch.empty((xa.shape[0], w2.shape[0]), device=x.device, dtype=act_dtype)
            y = gemm(xa, w2, out_dtype=out_dtype, bias=b, epilogue=EPI_GELU if act == ACT_GELU else EPI_SILU, out2=h)
        ctx.save_for_backward(xa, h)
        ctx.params = (weight, bias)     # python objects: carry the shadow / flat-gradient attributes
        ctx.act, ctx.act_dtype, ctx.x_in_dtype = act, act_dtype, x_in_dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        xa, h = ctx.saved_tensors
        weight, bias = ctx.params
        act_dtype = ctx.act_dtype
        dy = cast(dy.contiguous(), act_dtype)
        if ctx.act != ACT_NONE:
            dy = act_bwd(dy, h, ctx.act)
        w2 = weight_for(weight, act_dtype).view(weight.shape[0], -1)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = gemm(dy, w2, b_mn=True, out_dtype=ctx.x_in_dtype)
        dw = _weight_grad(weight, dy, xa) if ctx.needs_input_grad[1] else None
        db = _bias_grad(bias, dy) if (bias is not None and ctx.needs_input_grad[2]) else None
        return dx, dw, db, None, None, None


def linear(x, weight, bias, *, act=ACT_NONE, act_dtype, out_dtype=None):
    return LinearFn.apply(x, weight, bias, act, act_dtype, out_dtype if out_dtype is not None else act_dtype)


class _GradAccumulator:
    """fp32 side buffer the consumers of a shared tensor add their input gradients into (one split-K GEMM each),
    instead of handing autograd 28 separate bf16 gradients to sum.  `deferred`: work to run right before the buffer is
    read (the grouped input gradient of the blocks' modulation linears, AdaLNAll)."""

    def __init__(self):
        self.buf = None
        self.deferred = []

    def get(self, like, dtype=torch.float32):
        if self.buf is None:
            self.buf = torch.zeros(like.shape, device=like.device, dtype=dtype)
        return self.buf


ADALN_GROUP_MAX = 32       # kMaxGroups of reed_gemm_grouped
ADALN_GROUP_MAX_ROWS = 1024


class AdaLNAll:
    """adaLN_modulation(c) of EVERY transformer block (sit.py:125-133) as one grouped GEMM: the blocks' [6D, D] weights are
    separate parameters, silu(c) is shared.  forward: mod_all[B, L*6D] = c_act [W_0; ...; W_{L-1}]^T + [b_0; ...]; block i reads
    columns i*6D..(i+1)*6D.  backward: every block leaves its dmod (act dtype) in `dmods[i]` and computes its own weight / bias
    gradients; the input gradient sum_i dmod_i W_i is one more grouped GEMM, run when the gradient of silu(c) is collected
    (SiluCastFn.backward).  Per block this replaces a 54-CTA forward GEMM (14 us for a 16 MB weight stream) and a split-K
    dgrad GEMM (12 us) by 1/28 of two launches that stream the 446 MB of weights at HBM speed."""

    def __init__(self, c_act, linears, act_dtype, c_acc):
        self.c_act = c_act
        self.weights = [weight_for(l.weight, act_dtype) for l in linears]
        self.width = linears[0].weight.shape[0]
        bias_all = torch.cat([l.bias.detach() for l in linears])
        self.mod_all = torch.empty((c_act.shape[0], self.width * len(linears)), device=c_act.device, dtype=torch.float32)
        gemm_grouped_fwd(c_act, self.weights, bias_all, self.mod_all)
        self.dmod_all = None                 # fp32 [L, B, 6D], zeroed by the first block that runs backward
        self.done = []                       # blocks whose dmod is complete
        self.c_acc = c_acc
        if c_acc is not None:
            c_acc.deferred.append(self._input_grad)

    @staticmethod
    def usable(c_act, linears, act_dtype):
        w = linears[0].weight
        return (act_dtype == torch.bfloat16 and c_act.dtype == torch.bfloat16 and 1 <= len(linears) <= ADALN_GROUP_MAX
                and c_act.shape[0] <= ADALN_GROUP_MAX_ROWS and w.shape[0] % 256 == 0 and w.shape[1] % 8 == 0 and w.shape[1] >= 64
                and _ADALN_GROUPED and (_gemm_backend & 7) != BACKEND_SIMT
                and all(l.weight.shape == w.shape and l.bias is not None for l in linears))

    def mod(self, i):
        return self.mod_all[:, i * self.width:(i + 1) * self.width]

    def dmod(self, i):
        """Block i's gradient buffer for its modulation vectors (the row-wise backward kernels add into it)."""
        if self.dmod_all is None:
            self.dmod_all = torch.zeros((len(self.weights), self.c_act.shape[0], self.width), device=self.c_act.device,
                                        dtype=torch.float32)
        return self.dmod_all[i]

    def _input_grad(self):
        if not self.done:
            return
        dmod_a = cast(self.dmod_all, self.c_act.dtype)          # one cast for all blocks
        done = sorted(self.done)
        gemm_grouped_dgrad([dmod_a[i] for i in done], [self.weights[i] for i in done], self.c_acc.get(self.c_act), True)
        self.done, self.dmod_all = [], None


class SiluCastFn(torch.autograd.Function):
    """silu(c) emitted in the act dtype (the shared input of every adaLN linear, sit.py:125-133,148-154)."""

    @staticmethod
    def forward(ctx, c, act_dtype):
        _require_cuda(c)
        ctx.save_for_backward(c)
        out = cast(c, act_dtype, op=1)
        ctx.acc = _GradAccumulator()
        return out

    @staticmethod
    def backward(ctx, dy):
        (c,) = ctx.saved_tensors
        dy = cast(dy.contiguous(), torch.float32)
        for fn in ctx.acc.deferred:          # the blocks' grouped adaLN input gradient lands in the side buffer
            fn()
        ctx.acc.deferred = []
        if ctx.acc.buf is not None:          # gradients the transformer blocks accumulated on the side
            total = torch.empty_like(dy)
            _launch("reed_add_f32", _p(dy), _p(ctx.acc.buf), _p(total), dy.numel(), _stream())
            dy = total
            ctx.acc.buf = None
        return act_bwd(dy, c, ACT_SILU), None


def silu_cast(c, act_dtype):
    """Returns (silu(c) in act dtype, accumulator the blocks add their dL/d silu(c) into)."""
    out = SiluCastFn.apply(c, act_dtype)
    acc = out.grad_fn.acc if out.grad_fn is not None and hasattr(out.grad_fn, "acc") else None
    return out, acc


class CastFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dtype):
        ctx.in_dtype = x.dtype
        return cast(x, dtype)

    @staticmethod
    def backward(ctx, dy):
        return cast(dy.contiguous(), ctx.in_dtype), None


class TokenMeanFn(torch.autograd.Function):
    """x.mean(dim=1) for the text projector (sit.py:292,301); fp32 in, act dtype out."""

    @staticmethod
    def forward(ctx, x, act_dtype):
        _require_cuda(x)
        B, T, D = x.shape
        x = x.contiguous()
        out = torch.empty((B, D), device=x.device, dtype=act_dtype)
        _launch("reed_group_mean_fwd", _p(x), _p(out), _code(act_dtype), B, T, D, _stream())
        ctx.shape = (B, T, D)
        return out

    @staticmethod
    def backward(ctx, dy):
        B, T, D = ctx.shape
        dy = cast(dy.contiguous(), torch.float32)
        dx = torch.empty((B, T, D), device=dy.device, dtype=torch.float32)
        _launch("reed_group_mean_bwd", _p(dy), _p(dx), B, T, D, 0, _stream())
        return dx, None


def _lib():
    """torch.ops.reed namespace (reed_b200.library registers the custom ops on first use)."""
    from . import library  # noqa: F401
    return torch.ops.reed


class LNModulateFn:
    """modulate(LayerNorm(x), shift, scale) with x [B,T,D] fp32, shift/scale [B,D] fp32 views (sit.py:26-27,153-155):
    the differentiable custom op ``reed::ln_modulate`` (kept under the round-1 name for its callers)."""

    @staticmethod
    def apply(x, shift, scale, act_dtype):
        _require_cuda(x)
        return _lib().ln_modulate(x, shift, scale, act_dtype == torch.bfloat16)[0]


class BlockLink:
    """Hand-over between the backward passes of two consecutive blocks whose residual stream has no other consumer.

    Block i+1's backward ends with the LayerNorm backward that produces dL/dx (= block i's incoming gradient); block i's
    backward starts with the gate backward of its MLP branch on exactly that tensor.  The two are one pass of the fused
    LN-backward -> gate-backward kernel (the same one a block uses internally for its attention branch): block i+1 runs it with
    block i's saved branch output and gate, leaves dy / the bias-gradient handle / the modulation-gradient buffer here, and
    block i picks them up instead of launching its own gate backward (one launch and one 38 MB re-read of the residual
    gradient less per block).  Should block i receive a different gradient after all (its output had another consumer, e.g.
    a user hook that taps the residual stream), the gate backward of the difference is added - the operation is linear in
    the incoming gradient.  REED_BLOCK_LINK=0 turns it off for A/B runs."""
    __slots__ = ("y", "gate", "bias", "ada", "dy", "db_buf", "db_ret", "dmod", "dx")

    def __init__(self):
        self.y = self.gate = self.bias = self.ada = self.dy = self.db_buf = self.db_ret = self.dmod = self.dx = None


_BLOCK_LINK = _os.environ.get("REED_BLOCK_LINK", "1") != "0"


class SiTBlockFn(torch.autograd.Function):
    """One adaLN-Zero transformer block (sit.py:125-137 + timm Attention/Mlp): 7 kernels forward, 16-17 backward.

    x: [B,T,D] fp32 residual stream; c_act: [B,D] = silu(c) in the act dtype (shared by all blocks).
    mod = c_act W_ada^T + b_ada = (shift_a, scale_a, gate_a, shift_m, scale_m, gate_m), fp32 [B,6D].
    """

    @staticmethod
    def forward(ctx, x, c_act, w_ada, b_ada, w_qkv, b_qkv, w_proj, b_proj, w_fc1, b_fc1, w_fc2, b_fc2, num_heads,
                act_dtype, after_backward, c_acc=None, qn_w=None, qn_b=None, kn_w=None, kn_b=None, ada=None, link_in=None,
                link_out=None):
        _require_cuda(x, c_act)
        B, T, D = x.shape
        M = B * T
        H, hd = num_heads, D // num_heads
        x0 = x.contiguous().view(M, D)
        c_act = c_act.contiguous()
        W = lambda p: weight_for(p, act_dtype)

        # ada = (AdaLNAll, block index): the modulation vectors of every block came out of one grouped GEMM
        mod = ada[0].mod(ada[1]) if ada is not None else gemm(c_act, W(w_ada), out_dtype=torch.float32, bias=b_ada.detach())
        sh_a, sc_a, g_a, sh_m, sc_m, g_m = (mod[:, i * D:(i + 1) * D] for i in range(6))
        xm1, mean1, rstd1 = ln_modulate_fwd(x0, sh_a, sc_a, T, act_dtype, ones_col=True)
        qkv_raw = gemm(xm1, W(w_qkv), out_dtype=act_dtype, bias=b_qkv.detach())
        qk_stats = None
        qkv = qkv_raw
        if qn_w is not None:      # timm Attention(qk_norm=True): LayerNorm(head_dim) on q and k before the softmax
            qkv, qk_stats = qk_norm_fwd(qkv_raw, qn_w.detach().float(), qn_b.detach().float(), kn_w.detach().float(),
                                        kn_b.detach().float(), M, H, hd)
        o, lse = attention_fwd(qkv, B, T, H, hd)
        y1 = torch.empty((M, D), device=x.device, dtype=act_dtype)
        x1 = gemm(o, W(w_proj), out_dtype=torch.float32, bias=b_proj.detach(), epilogue=EPI_GATE_RES, aux=x0, gate=g_a,
                  rows_per_group=T, out2=y1)
        xm2, mean2, rstd2 = ln_modulate_fwd(x1, sh_m, sc_m, T, act_dtype, ones_col=True)
        h = torch.empty((M, w_fc1.shape[0]), device=x.device, dtype=act_dtype)
        a = gemm(xm2, W(w_fc1), out_dtype=act_dtype, bias=b_fc1.detach(), epilogue=EPI_GELU, out2=h)
        y2 = torch.empty((M, D), device=x.device, dtype=act_dtype)
        x2 = gemm(a, W(w_fc2), out_dtype=torch.float32, bias=b_fc2.detach(), epilogue=EPI_GATE_RES, aux=x1, gate=g_m,
                  rows_per_group=T, out2=y2)

        # xm1 / xm2 are [M, D] views of wider rows (ones column): save the storage, re-slice in backward
        ctx.save_for_backward(x0, c_act, mod, mean1, rstd1, xm1._base if xm1._base is not None else xm1, qkv, o, lse, y1,
                              x1, mean2, rstd2, xm2._base if xm2._base is not None else xm2, h, a, y2,
                              qkv_raw if qk_stats is not None else None, qk_stats)
        ctx.params = (w_ada, b_ada, w_qkv, b_qkv, w_proj, b_proj, w_fc1, b_fc1, w_fc2, b_fc2)
        ctx.qk_params = (qn_w, qn_b, kn_w, kn_b)
        ctx.dims = (B, T, D, H, hd)
        ctx.act_dtype = act_dtype
        ctx.after_backward = after_backward
        ctx.c_acc = c_acc
        ctx.ada = ada
        # link_out: this block's output feeds the next block only, whose backward will run this block's MLP gate backward;
        # link_in: the same for the previous block, seen from this one
        if link_out is not None:
            link_out.y, link_out.gate, link_out.bias, link_out.ada = y2, g_m, b_fc2, ada
        ctx.link_in, ctx.link_out = link_in, link_out
        return x2.view(B, T, D)

    @staticmethod
    def backward(ctx, dx2):
        (x0, c_act, mod, mean1, rstd1, xm1s, qkv, o, lse, y1, x1, mean2, rstd2, xm2s, h, a, y2, qkv_raw,
         qk_stats) = ctx.saved_tensors
        w_ada, b_ada, w_qkv, b_qkv, w_proj, b_proj, w_fc1, b_fc1, w_fc2, b_fc2 = ctx.params
        B, T, D, H, hd = ctx.dims
        xm1_ext = xm1s if xm1s.shape[1] > D else None
        xm2_ext = xm2s if xm2s.shape[1] > D else None
        xm1, xm2 = xm1s[:, :D], xm2s[:, :D]
        M = B * T
        act_dtype = ctx.act_dtype
        W = lambda p: weight_for(p, act_dtype)
        dx2 = dx2.contiguous().view(M, D)
        if dx2.dtype != torch.float32:
            dx2 = cast(dx2, torch.float32)
        sh_a, sc_a, g_a, sh_m, sc_m, g_m = (mod[:, i * D:(i + 1) * D] for i in range(6))
        handed = ctx.link_out if (ctx.link_out is not None and ctx.link_out.dy is not None) else None
        if handed is not None:
            # the next block's backward already ran this block's MLP gate backward on the gradient it produced
            dmod = handed.dmod
        else:
            dmod = ctx.ada[0].dmod(ctx.ada[1]) if ctx.ada is not None else torch.zeros_like(mod)
        dsh_a, dsc_a, dg_a, dsh_m, dsc_m, dg_m = (dmod[:, i * D:(i + 1) * D] for i in range(6))

        def bias_buffer(p):
            main, acc = _grad_target(p)
            if main is not None:
                if not acc:
                    _zero_fresh(p, main)
                return main, None
            buf = torch.zeros(p.shape, device=p.device, dtype=torch.float32)
            return buf, buf

        # ---- MLP branch:  x2 = x1 + g_m * (gelu(xm2 W1^T + b1) W2^T + b2)
        if handed is not None:
            dy2, db2 = handed.dy, handed.db_ret
            if dx2.data_ptr() != handed.dx.data_ptr():
                # another consumer contributed to this block's output gradient: gate backward of the rest (linear in dx)
                dy2 = dy2 + gate_bwd((dx2 - handed.dx).contiguous(), y2, g_m, T, dg_m, handed.db_buf)
            handed.dy = handed.db_buf = handed.db_ret = handed.dmod = handed.dx = None
        else:
            db2_buf, db2 = bias_buffer(b_fc2)
            dy2 = gate_bwd(dx2, y2, g_m, T, dg_m, db2_buf)
        side = _SideStream(dx2.device) if (_WGRAD_STREAM and _flat_grads(w_fc2, w_fc1, b_fc1, w_proj, w_qkv, b_qkv, w_ada)) else None

        def off_stream(fn, *operands):
            return side.run(fn, *operands) if side is not None else fn()

        dw2 = off_stream(lambda: _weight_grad(w_fc2, dy2, a), dy2, a)
        dh = gemm(dy2, W(w_fc2), b_mn=True, out_dtype=act_dtype, epilogue=EPI_DGELU, aux=h)
        dw1, db1 = off_stream(lambda: _weight_and_bias_grad(w_fc1, b_fc1, dh, xm2, xm2_ext), dh, xm2s)
        dxm2 = gemm(dh, W(w_fc1), b_mn=True, out_dtype=act_dtype)
        # ---- attention branch:  x1 = x0 + g_a * (attn(xm1) Wp^T + bp); its gate backward rides on the LN backward
        dbp_buf, dbp = bias_buffer(b_proj)
        dx1, dy1 = ln_modulate_gate_bwd(dxm2, x1, mean2, rstd2, sc_m, T, dx2, dsh_m, dsc_m, y1, g_a, dg_a, dbp_buf)
        dwp = off_stream(lambda: _weight_grad(w_proj, dy1, o), dy1, o)
        d_o = gemm(dy1, W(w_proj), b_mn=True, out_dtype=act_dtype)
        dqkv = attention_bwd(qkv, o, d_o, lse, B, T, H, hd)
        qk_grads = (None, None, None, None)
        if qk_stats is not None:
            qn_w, qn_b, kn_w, kn_b = ctx.qk_params
            bufs, rets = [], []
            for prm in (qn_w, qn_b, kn_w, kn_b):
                buf, ret = bias_buffer(prm)
                bufs.append(buf)
                rets.append(ret)
            dqkv = qk_norm_bwd(dqkv, qkv_raw, qk_stats, qn_w.detach().float(), kn_w.detach().float(), bufs[0], bufs[1],
                               bufs[2], bufs[3], M, H, hd)
            qk_grads = tuple(rets)
        dwqkv, dbqkv = off_stream(lambda: _weight_and_bias_grad(w_qkv, b_qkv, dqkv, xm1, xm1_ext), dqkv, xm1s)
        dxm1 = gemm(dqkv, W(w_qkv), b_mn=True, out_dtype=act_dtype)
        L = ctx.link_in
        if L is not None and L.y is not None:
            # ... and the previous block's MLP gate backward rides on this block's first LayerNorm backward
            dmod_prev = L.ada[0].dmod(L.ada[1]) if L.ada is not None else torch.zeros_like(mod)
            dbq_buf, dbq_ret = bias_buffer(L.bias)
            dx0, dy_prev = ln_modulate_gate_bwd(dxm1, x0, mean1, rstd1, sc_a, T, dx1, dsh_a, dsc_a, L.y, L.gate,
                                                dmod_prev[:, 5 * D:6 * D], dbq_buf)
            L.dy, L.db_buf, L.db_ret, L.dmod, L.dx = dy_prev, dbq_buf, dbq_ret, dmod_prev, dx0
        else:
            dx0 = ln_modulate_bwd(dxm1, x0, mean1, rstd1, sc_a, T, dx1, dsh_a, dsc_a)

        # ---- adaLN linear:  mod = c_act W_ada^T + b_ada
        dc = None
        if ctx.ada is not None:
            # weight and bias gradient in one outer-product pass over the fp32 dmod; the block's share of dL/d silu(c) is
            # part of one grouped GEMM over all blocks, run when that gradient is collected (AdaLNAll._input_grad)
            if dmod.shape[0] <= OUTER_WGRAD_MAX_ROWS:
                dw_ada, db_ada = off_stream(lambda: _outer_weight_and_bias_grad(w_ada, b_ada, dmod, c_act), dmod, c_act)
            else:                         # a larger batch: the contraction is long enough for the tensor-core weight gradient
                dmod_a = cast(dmod, act_dtype)
                db_ada = _bias_grad(b_ada, dmod)
                dw_ada = off_stream(lambda: _weight_grad(w_ada, dmod_a, c_act), dmod_a, c_act)
            ctx.ada[0].done.append(ctx.ada[1])
            if side is not None:
                side.join()
            if ctx.after_backward is not None:
                ctx.after_backward()
            return (dx0.view(B, T, D), dc, dw_ada, db_ada, dwqkv, dbqkv, dwp, dbp, dw1, db1, dw2, db2, None, None, None,
                    None) + qk_grads + (None, None, None)
        dmod_a = cast(dmod, act_dtype)
        db_ada = _bias_grad(b_ada, dmod)
        dw_ada = off_stream(lambda: _weight_grad(w_ada, dmod_a, c_act), dmod_a, c_act)
        if ctx.needs_input_grad[1]:
            if ctx.c_acc is not None:     # [B, 6D] x [6D, D]: few rows, long reduction -> split-K slices add into the side buffer
                gemm(dmod_a, W(w_ada), b_mn=True, out=ctx.c_acc.get(c_act), accumulate=True)
            else:
                dc = gemm(dmod_a, W(w_ada), b_mn=True, out_dtype=act_dtype)

        if side is not None:
            side.join()                       # the bucket's gradients are complete on the current stream from here on
        if ctx.after_backward is not None:
            ctx.after_backward()
        return (dx0.view(B, T, D), dc, dw_ada, db_ada, dwqkv, dbqkv, dwp, dbp, dw1, db1, dw2, db2, None, None, None, None) + qk_grads + (None, None, None)


# --------------------------------------------------------------------------------------------------
# SILoss pieces (loss.py:175-186, 204-225)
# --------------------------------------------------------------------------------------------------

def _interpolate_raw(x, eps, t, path_type):
    x, eps = x.contiguous(), eps.contiguous()
    out = torch.empty_like(x)
    B = x.shape[0]
    _launch("reed_siloss_interp", _p(x), _p(eps), _p(t), _p(out), B, x.numel() // B, path_type, _stream())
    return out


def interpolate(x, eps, t, path_type):
    """x_t = alpha_t x + sigma_t eps with per-sample t (no autograd: inputs are data); custom op ``reed::siloss_interp``."""
    _require_cuda(x, eps, t)
    return _lib().siloss_interp(x, eps, t, path_type)


def _mse_fwd_raw(pred, x, eps, t, path_type):
    pred = pred.contiguous()
    B = pred.shape[0]
    out = torch.empty((B,), device=pred.device, dtype=torch.float32)
    _launch("reed_siloss_mse_fwd", _p(pred), _p(x), _p(eps), _p(t), _p(out), B, pred.numel() // B, path_type, _stream())
    return out


def _mse_bwd_raw(pred, x, eps, t, g, path_type):
    pred = pred.contiguous()
    B = pred.shape[0]
    g = g.contiguous().float()
    dpred = torch.empty_like(pred)
    _launch("reed_siloss_mse_bwd", _p(pred), _p(x), _p(eps), _p(t), _p(g), _p(dpred), B, pred.numel() // B, path_type, _stream())
    return dpred


class VelocityMSEFn:
    """mean_flat((pred - (dalpha x + dsigma eps))^2) -> (B,): the differentiable custom op ``reed::velocity_mse``."""

    @staticmethod
    def apply(pred, x, eps, t, path_type):
        _require_cuda(pred, x, eps, t)
        return _lib().velocity_mse(pred, x.contiguous(), eps.contiguous(), t, path_type)


def _cos_fwd_raw(zt, z):
    zt, z = zt.contiguous(), z.contiguous()
    B, T, Z = zt.shape
    stats = torch.empty((B * T, 3), device=zt.device, dtype=torch.float32)
    align = torch.zeros((B,), device=zt.device, dtype=torch.float32)
    _launch("reed_siloss_cos_fwd", _p(zt), _code(zt.dtype), _p(z), _code(z.dtype), _p(stats), _p(align), B, T, Z, _stream())
    return align, stats


def _cos_bwd_raw(zt, z, stats, g):
    zt, z = zt.contiguous(), z.contiguous()
    B, T, Z = zt.shape
    g = g.contiguous().float()
    dzt = torch.empty_like(zt)
    _launch("reed_siloss_cos_bwd", _p(zt), _code(zt.dtype), _p(z), _code(z.dtype), _p(stats), _p(g), _p(dzt), B, T, Z, _stream())
    return dzt


class CosineAlignFn:
    """-(normalize(z) . normalize(z~)).sum(-1).mean(-1) -> (B,);  z~: [B,T,Z] (grad), z: [B,T,Z] target: the differentiable
    custom op ``reed::cosine_align``."""

    @staticmethod
    def apply(zt, z):
        _require_cuda(zt, z)
        return _lib().cosine_align(zt, z)[0]


# --------------------------------------------------------------------------------------------------
# sampler step (samplers.py:61-104, 124-187)
# --------------------------------------------------------------------------------------------------

def sampler_cast(x64, model_dtype, dup):
    n = x64.numel()
    shape = (x64.shape[0] * (2 if dup else 1),) + tuple(x64.shape[1:])
    out = torch.empty(shape, device=x64.device, dtype=model_dtype)
    _launch("reed_sampler_cast", _p(x64), _p(out), _code(model_dtype), n, int(dup), _stream())
    return out


def sampler_step(x_cur, v, *, eps=None, d_prev=None, want_slope=False, next_dup=None, guided=False, cfg=1.0, t_cur=0.0,
                 dt=0.0, sde=False, path_type=0):
    """One fused update.  Returns (x_next fp64, slope fp64 or None, next model input or None)."""
    _require_cuda(x_cur, v)
    n = x_cur.numel()
    v = v.contiguous()
    assert v.numel() == n * (2 if guided else 1)
    x_next = torch.empty_like(x_cur)
    slope = torch.empty_like(x_cur) if want_slope else None
    x_model = None
    if next_dup is not None:
        shape = (x_cur.shape[0] * (2 if next_dup else 1),) + tuple(x_cur.shape[1:])
        x_model = torch.empty(shape, device=x_cur.device, dtype=v.dtype)
    _launch("reed_sampler_step", _p(x_cur), _p(v), _code(v.dtype), _p(eps), _p(d_prev), _p(slope), _p(x_next), _p(x_model),
            n, int(guided), int(bool(next_dup)), int(sde), path_type, float(cfg), float(t_cur), float(dt), _stream())
    return x_next, slope, x_model


def tcgen05_gemm_launches() -> int:
    """Running count of tcgen05 GEMM kernel launches made by the library (reed_gemm and reed_gemm_wgrad_bias)."""
    import ctypes
    n = ctypes.c_longlong(0)
    call("reed_gemm_tcgen05_launches", ctypes.byref(n))
    return int(n.value)


def device_check():
    import ctypes
    buf = ctypes.create_string_buffer(128)
    call("reed_device_check", buf, 128)
    return buf.value.decode()
