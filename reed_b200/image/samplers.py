"""Drop-in replacement for the reference ``image/samplers.py`` with fused fp64 step kernels.

Kept from /root/reference/image/samplers.py: ``expand_t_like_x`` (5-13), ``get_score_from_velocity`` (15-39),
``compute_diffusion`` (42-43), and the signatures, time grids, CFG window rule (tested on t_cur, also for the Heun
corrector), hard-coded null class 1000, fp64 state and fp64 return of ``euler_sampler`` (46-104) and
``euler_maruyama_sampler`` (107-187).  Each step is ONE kernel (reed_sampler_step): cast-in of the model output,
score/drift, CFG combine, update, and cast-out (batch-duplicated when the next evaluation is guided) of the next
model input.  The SDE noise is drawn with the same ``torch.randn_like(fp64 state)`` call per step.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops

_PATH_CODE = {"linear": 0, "cosine": 1}


def expand_t_like_x(t, x_cur):
    return t.view(t.size(0), *([1] * (x_cur.dim() - 1)))


def get_score_from_velocity(vt, xt, t, path_type="linear"):
    """PyTorch helper kept for API parity; the samplers use the fused kernel."""
    t = expand_t_like_x(t, xt)
    if path_type == "linear":
        alpha_t, d_alpha_t = 1 - t, -torch.ones_like(xt)
        sigma_t, d_sigma_t = t, torch.ones_like(xt)
    elif path_type == "cosine":
        ang = t * np.pi / 2
        alpha_t, sigma_t = torch.cos(ang), torch.sin(ang)
        d_alpha_t, d_sigma_t = -np.pi / 2 * torch.sin(ang), np.pi / 2 * torch.cos(ang)
    else:
        raise NotImplementedError
    ratio = alpha_t / d_alpha_t
    return (ratio * vt - xt) / (sigma_t ** 2 - ratio * d_sigma_t * sigma_t)


def compute_diffusion(t_cur):
    return 2 * t_cur


def _require_cuda(latents):
    if not latents.is_cuda:
        raise RuntimeError("reed_b200 samplers run on CUDA (sm_100a) tensors only; there is no CPU fallback")


def _guided(cfg_scale, t, lo, hi):
    return bool(cfg_scale > 1.0 and t <= hi and t >= lo)


def _evaluate(model, x_model, t_scalar, y, y_null, guided, dtype):
    rows = x_model.shape[0]
    labels = torch.cat([y, y_null], dim=0) if guided else y
    t_in = torch.full((rows,), float(t_scalar), dtype=torch.float64, device=x_model.device).to(dtype)
    return model(x_model, t_in, y=labels)[0]


def euler_sampler(model, latents, y, num_steps=20, heun=False, cfg_scale=1.0, guidance_low=0.0, guidance_high=1.0,
                  path_type="linear"):
    _require_cuda(latents)
    y_null = torch.tensor([1000] * y.size(0), device=y.device) if cfg_scale > 1.0 else None
    dtype = latents.dtype
    model_dtype = dtype if dtype in (torch.float32, torch.bfloat16) else torch.float32
    t_steps = torch.linspace(1, 0, num_steps + 1, dtype=torch.float64)
    x = latents.to(torch.float64).contiguous()
    with torch.no_grad():
        guided = _guided(cfg_scale, t_steps[0], guidance_low, guidance_high) if num_steps > 0 else False
        x_model = ops.sampler_cast(x, model_dtype, guided)
        for i in range(num_steps):
            t_cur, t_next = float(t_steps[i]), float(t_steps[i + 1])
            guided = _guided(cfg_scale, t_steps[i], guidance_low, guidance_high)
            two_stage = heun and i < num_steps - 1
            nxt_guided = guided if two_stage else (
                _guided(cfg_scale, t_steps[i + 1], guidance_low, guidance_high) if i + 1 < num_steps else None)
            d = _evaluate(model, x_model.to(dtype), t_cur, y, y_null, guided, dtype).to(model_dtype)
            x_e, slope, x_model = ops.sampler_step(x, d, want_slope=two_stage, next_dup=nxt_guided, guided=guided,
                                                   cfg=cfg_scale, t_cur=t_cur, dt=t_next - t_cur)
            if two_stage:
                after = _guided(cfg_scale, t_steps[i + 1], guidance_low, guidance_high) if i + 1 < num_steps else None
                d2 = _evaluate(model, x_model.to(dtype), t_next, y, y_null, guided, dtype).to(model_dtype)
                x_e, _, x_model = ops.sampler_step(x, d2, d_prev=slope, next_dup=after, guided=guided, cfg=cfg_scale,
                                                   t_cur=t_next, dt=t_next - t_cur)
            x = x_e
    return x


def euler_maruyama_sampler(model, latents, y, num_steps=20, heun=False, cfg_scale=1.0, guidance_low=0.0,
                           guidance_high=1.0, path_type="linear"):
    _require_cuda(latents)
    if path_type not in _PATH_CODE:
        raise NotImplementedError
    path = _PATH_CODE[path_type]
    y_null = torch.tensor([1000] * y.size(0), device=y.device) if cfg_scale > 1.0 else None
    dtype = latents.dtype
    model_dtype = dtype if dtype in (torch.float32, torch.bfloat16) else torch.float32
    t_steps = torch.cat([torch.linspace(1., 0.04, num_steps, dtype=torch.float64), torch.tensor([0.], dtype=torch.float64)])
    x = latents.to(torch.float64).contiguous()
    total = num_steps            # num_steps - 1 stochastic steps + the deterministic last step

    def run_step(i, stochastic):
        nonlocal x, x_model
        t_cur, t_next = float(t_steps[i]), float(t_steps[i + 1])
        guided = _guided(cfg_scale, t_steps[i], guidance_low, guidance_high)
        nxt = _guided(cfg_scale, t_steps[i + 1], guidance_low, guidance_high) if i + 1 < total else None
        eps = torch.randn_like(x) if stochastic else None          # fp64 normals, same call as samplers.py:142
        v = _evaluate(model, x_model.to(dtype), t_cur, y, y_null, guided, dtype).to(model_dtype)
        x, _, x_model = ops.sampler_step(x, v, eps=eps, next_dup=nxt, guided=guided, cfg=cfg_scale, t_cur=t_cur,
                                         dt=t_next - t_cur, sde=True, path_type=path)

    with torch.no_grad():
        x_model = ops.sampler_cast(x, model_dtype, _guided(cfg_scale, t_steps[0], guidance_low, guidance_high))
        for i in range(num_steps - 1):
            run_step(i, True)
    # the reference runs its last (mean) step outside no_grad (samplers.py:158-187); the result carries no graph
    # here either way because the state tensors are produced by raw kernels.
    with torch.no_grad():
        run_step(num_steps - 1, False)
    return x
