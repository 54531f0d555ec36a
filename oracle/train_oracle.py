"""CPU restatement of the reference train-step glue.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/image/train.py: sample_posterior 84-91, update_ema 94-105, loss mix 396-398,
clip 402-407, AdamW 253-259/408, EMA 411-412.  Operates on dicts of leaf tensors (name -> tensor).
"""
from __future__ import annotations

import math
from typing import Dict

import torch


def sample_posterior(moments, noise, scale=0.18215, bias=0.0):
    mean, std = moments.chunk(2, dim=1)
    return (mean + std * noise) * scale + bias


def mix_losses(loss_dict, proj_coeff=0.5, diffusion_decay=1.0, repa_decay=1.0):
    return loss_dict["denoising_loss"].mean() * diffusion_decay + loss_dict["proj_loss"] * proj_coeff * repa_decay


def global_grad_norm(grads: Dict[str, torch.Tensor]) -> torch.Tensor:
    return torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()


def clip_coefficient(norm, max_norm=1.0):
    """torch.nn.utils.clip_grad_norm_: coef = max_norm / (norm + 1e-6), clamped to <= 1."""
    return torch.clamp(max_norm / (norm + 1e-6), max=1.0)


def adamw_ema_step(params, grads, exp_avg, exp_avg_sq, ema, step, *, lr=1e-4, beta1=0.9, beta2=0.999,
                   eps=1e-8, weight_decay=0.0, max_norm=1.0, ema_decay=0.9999):
    """One clipped AdamW + EMA update in place on ``params`` / moments / ``ema``; ``step`` is 1-based.

    ``ema`` covers every named parameter (including the frozen pos_embed, train.py:99-105); params
    without a gradient are only EMA-averaged.
    """
    norm = global_grad_norm(grads)
    coef = clip_coefficient(norm, max_norm)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    with torch.no_grad():
        for name, p in params.items():
            if name in grads:
                g = grads[name] * coef
                p.mul_(1 - lr * weight_decay)
                exp_avg[name].mul_(beta1).add_(g, alpha=1 - beta1)
                exp_avg_sq[name].mul_(beta2).addcmul_(g, g, value=1 - beta2)
                denom = exp_avg_sq[name].sqrt() / math.sqrt(bc2) + eps
                p.addcdiv_(exp_avg[name], denom, value=-lr / bc1)
            ema[name].mul_(ema_decay).add_(p, alpha=1 - ema_decay)
    return norm


def curriculum(global_step, *, repa_weight_decay="constant", repa_steps=400000, start_diffusion_steps=0,
               diffusion_warm_up_steps=50000, diffusion_decay="constant", max_train_steps=400000):
    """(diffusion_loss_decay, repa_weight_decay) of train.py:363-385 for one step, numpy arithmetic as there.

    The cosine diffusion decay keeps the reference's operator precedence (``/ max_train_steps - top_steps``).
    """
    import numpy as np
    if repa_weight_decay == "constant":
        repa = 1.0
    elif repa_weight_decay == "linear":
        repa = max(1.0 - global_step / repa_steps, 0.)
    elif repa_weight_decay == "cosine":
        repa = max((1.0 + np.cos(np.pi * global_step / repa_steps)) / 2, 0.)
    else:
        raise NotImplementedError
    top = diffusion_warm_up_steps + start_diffusion_steps
    if global_step < start_diffusion_steps:
        diff = 0.0
    elif global_step < top:
        diff = (global_step - start_diffusion_steps) / diffusion_warm_up_steps
    elif diffusion_decay == "constant":
        diff = 1.0
    elif diffusion_decay == "linear":
        diff = 1.0 - (global_step - top) / (max_train_steps - top)
    elif diffusion_decay == "cosine":
        diff = (1.0 + np.cos(np.pi * (global_step - top) / max_train_steps - top)) / 2
    else:
        raise NotImplementedError
    return float(diff), float(repa)
