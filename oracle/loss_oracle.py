"""Functional CPU restatement of the reference SILoss.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/image/loss.py: interpolant 49-64, time_weight 118-151, __call__ 153-237.
Random draws are arguments (``t``, ``noise``) so a CUDA run's draws can be replayed; ``draw_time``
restates the CPU-generator draw of loss.py:158-170.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional, Sequence

import torch
import torch.nn.functional as F

IMAGE_ENCODER_NAMES = ("dinov2", "mocov3", "clip", "mae", "jepa")   # loss.py:5


def draw_time(batch: int, weighting: str = "uniform", path_type: str = "linear") -> torch.Tensor:
    """(B,1,1,1) fp32 on CPU, from the CPU default generator (loss.py:158-168)."""
    if weighting == "uniform":
        return torch.rand((batch, 1, 1, 1))
    if weighting == "lognormal":
        sigma = torch.randn((batch, 1, 1, 1)).exp()
        if path_type == "linear":
            return sigma / (1 + sigma)
        if path_type == "cosine":
            return 2 / math.pi * torch.atan(sigma)
    raise NotImplementedError(weighting)


def path_coefficients(t: torch.Tensor, path_type: str):
    """alpha, sigma, d_alpha, d_sigma (loss.py:49-64)."""
    if path_type == "linear":
        return 1 - t, t, -1.0, 1.0
    if path_type == "cosine":
        half_pi = math.pi / 2
        return torch.cos(t * half_pi), torch.sin(t * half_pi), -half_pi * torch.sin(t * half_pi), half_pi * torch.cos(t * half_pi)
    raise NotImplementedError(path_type)


def schedule_weight(t: torch.Tensor, base: float, schedule: str, cutoffs=(0.0, 1.0)) -> torch.Tensor:
    """Keeps t's shape (B,1,1,1) (loss.py:118-151)."""
    if schedule == "linear":
        s = 1 - t
    elif schedule == "cosine":
        s = 0.5 * (1 + torch.cos(math.pi * t))
    elif schedule == "sigmoid":
        s = 1 / (1 + torch.exp((t - 0.5) * 10))
    elif schedule == "constant":
        s = torch.ones_like(t)
    elif schedule == "loglinear":
        s = 1 - torch.log(t + 1)
    elif schedule == "cutoff":
        s = torch.ones_like(t)
        s = torch.where((t < cutoffs[0]) | (t > cutoffs[1]), torch.zeros_like(s), s)
    else:
        raise ValueError(schedule)
    return base * s


def si_loss(model: Callable, images: torch.Tensor, t: torch.Tensor, noise: torch.Tensor,
            zs: Sequence[torch.Tensor], *, enc_names: Sequence[str], loss_weights: Dict[str, float],
            model_kwargs: Optional[dict] = None, path_type: str = "linear", time_schedule: str = "constant",
            cutoffs=(0.0, 1.0)) -> dict:
    """``t`` is (B,1,1,1).  Returns the reference's dict (loss.py:233-237) plus ``per_sample_align``."""
    kw = dict(model_kwargs or {})
    kw["inference"] = False
    t = t.to(device=images.device, dtype=images.dtype)
    a, s, da, ds = path_coefficients(t, path_type)
    x_t = a * images + s * noise
    target = da * images + ds * noise
    pred, z_model = model(x_t, t.flatten(), **kw)
    denoise = ((pred - target) ** 2).flatten(1).mean(dim=1)

    total = 0.0
    sums = {"image": [0.0, 0], "text": [0.0, 0]}
    per_sample = []
    for z_ref, z_hat, name in zip(zs, z_model, enc_names):
        base = loss_weights.get(name, 1.0)
        w = schedule_weight(t, base, time_schedule, cutoffs)
        kind = "image" if (name in IMAGE_ENCODER_NAMES or len(enc_names) == 1) else "text"
        zh = F.normalize(z_hat, dim=-1)
        zr = F.normalize(z_ref, dim=-1)
        if zr.ndim == 2:
            assert kind == "text" and zh.ndim == 2
            zr, zh = zr[:, None, :], zh[:, None, :]
        if base == 0.0:
            w = torch.ones_like(w)
        align = -(zr * zh).sum(dim=-1).mean(dim=-1)                      # (B,)
        # (B,) * (B,1,1,1) broadcasts to (B,1,1,B): mean == mean(align) * mean(w)   (loss.py:221-222)
        total = total + (align * w).mean()
        sums[kind][0] = sums[kind][0] + align.mean()
        sums[kind][1] += 1
        per_sample.append(align)
    return {
        "denoising_loss": denoise,
        "proj_loss": total,
        "img_proj_loss": sums["image"][0] / max(1, sums["image"][1]),
        "text_proj_loss": sums["text"][0] / max(1, sums["text"][1]),
        "per_sample_align": per_sample,
    }
