"""Functional CPU restatement of the reference samplers.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/image/samplers.py: score-from-velocity 15-39, diffusion 42-43,
euler_sampler 46-104, euler_maruyama_sampler 107-187.  State is fp64; the model is evaluated in the
dtype of the incoming latents.  SDE noise can be injected (``noises``) to replay a device run.
"""
from __future__ import annotations

import math
from typing import Callable, Optional, Sequence

import torch

NULL_CLASS = 1000   # samplers.py:59,120 (hard-coded)


def _eval_velocity(model, x64, t_scalar, y, dtype, guided, cfg_scale, post=None):
    n = x64.shape[0]
    if guided:
        xin = torch.cat([x64, x64], dim=0)
        yin = torch.cat([y, torch.full_like(y, NULL_CLASS)], dim=0)
    else:
        xin, yin = x64, y
    tin = torch.full((xin.shape[0],), float(t_scalar), dtype=torch.float64, device=x64.device)
    v = model(xin.to(dtype), tin.to(dtype), y=yin)[0].to(torch.float64)
    if post is not None:
        v = post(v, xin, tin)
    if guided:
        cond, uncond = v[:n], v[n:]
        v = uncond + cfg_scale * (cond - uncond)
    return v


def _in_window(cfg_scale, t, lo, hi):
    return cfg_scale > 1.0 and lo <= float(t) <= hi


def euler(model: Callable, latents, y, num_steps=20, heun=False, cfg_scale=1.0, guidance_low=0.0,
          guidance_high=1.0, path_type="linear"):
    dtype = latents.dtype
    ts = torch.linspace(1, 0, num_steps + 1, dtype=torch.float64)
    x = latents.to(torch.float64)
    with torch.no_grad():
        for i in range(num_steps):
            t0, t1 = ts[i], ts[i + 1]
            guided = _in_window(cfg_scale, t0, guidance_low, guidance_high)
            d0 = _eval_velocity(model, x, t0, y, dtype, guided, cfg_scale)
            x_e = x + (t1 - t0) * d0
            if heun and i < num_steps - 1:
                d1 = _eval_velocity(model, x_e, t1, y, dtype, guided, cfg_scale)   # window tested on t0 (samplers.py:85)
                x_e = x + (t1 - t0) * (0.5 * d0 + 0.5 * d1)
            x = x_e
    return x


def score_from_velocity(v, x, t, path_type="linear"):
    tb = t.view(-1, *([1] * (x.dim() - 1)))
    if path_type == "linear":
        a, da, s, ds = 1 - tb, -torch.ones_like(x), tb, torch.ones_like(x)
    elif path_type == "cosine":
        h = math.pi / 2
        a, s = torch.cos(tb * h), torch.sin(tb * h)
        da, ds = -h * torch.sin(tb * h), h * torch.cos(tb * h)
    else:
        raise NotImplementedError
    ratio = a / da
    var = s ** 2 - ratio * ds * s
    return (ratio * v - x) / var


def euler_maruyama(model: Callable, latents, y, num_steps=20, heun=False, cfg_scale=1.0, guidance_low=0.0,
                   guidance_high=1.0, path_type="linear", noises: Optional[Sequence[torch.Tensor]] = None):
    dtype = latents.dtype
    ts = torch.cat([torch.linspace(1.0, 0.04, num_steps, dtype=torch.float64), torch.zeros(1, dtype=torch.float64)])
    x = latents.to(torch.float64)

    def drift_of(t_scalar):
        w = 2 * t_scalar                                            # compute_diffusion
        return lambda v, xin, tin: v - 0.5 * w * score_from_velocity(v, xin, tin, path_type)

    with torch.no_grad():
        for i in range(num_steps - 1):
            t0, t1 = ts[i], ts[i + 1]
            dt = t1 - t0
            guided = _in_window(cfg_scale, t0, guidance_low, guidance_high)
            eps = noises[i] if noises is not None else torch.randn_like(x)
            d = _eval_velocity(model, x, t0, y, dtype, guided, cfg_scale, post=drift_of(t0))
            x = x + d * dt + torch.sqrt(2 * t0) * (eps * torch.sqrt(torch.abs(dt)))
    t0, t1 = ts[-2], ts[-1]
    guided = _in_window(cfg_scale, t0, guidance_low, guidance_high)
    d = _eval_velocity(model, x, t0, y, dtype, guided, cfg_scale, post=drift_of(t0))
    return x + (t1 - t0) * d
