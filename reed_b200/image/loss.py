"""Drop-in replacement for the reference ``image/loss.py`` (SILoss) on fused sm_100a kernels.

Kept from /root/reference/image/loss.py: constructor kwargs (22-35), ``interpolant`` (49-64), ``encoder_weight``
(66-116, never called by the reference either), ``time_weight`` (118-151), and ``__call__(model, images,
model_kwargs=None, zs=None, **kwargs)`` (153-237) including
  * the RNG draw order: t from the CPU generator, then ``randn_like(images)`` on the device, then the model's
    label-dropout draw;
  * ``model_kwargs['inference'] = False`` written into the caller's dict;
  * ``(B,) * (B,1,1,1)`` broadcasting in the time-weighted projection loss (mean(curr) * mean(wts));
  * ``text_proj_loss`` being the Python float 0.0 when there is no text encoder.
The interpolant, the velocity-target MSE and the negative-cosine alignment run as fused kernels
(reed_siloss_*), forward and backward; only (B,)-sized bookkeeping stays in PyTorch.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import ops

IMAGE_ENCODERS = ['dinov2', 'mocov3', 'clip', 'mae', 'jepa']
_PATH_CODE = {"linear": 0, "cosine": 1}


def mean_flat(x):
    return x.flatten(1).mean(dim=1) if x.dim() > 1 else x


def sum_flat(x):
    return x.flatten(1).sum(dim=1) if x.dim() > 1 else x


class SILoss:
    def __init__(self, prediction='v', path_type="linear", weighting="uniform", encoders=[], enc_names=[],
                 loss_weights={"dinov2": 1.0, "t5": 1.0}, time_schedule="constant", cutoffs=[0.0, 1.0],
                 accelerator=None, latents_scale=None, latents_bias=None):
        self.prediction = prediction
        self.weighting = weighting
        self.path_type = path_type
        self.encoders = encoders
        self.enc_names = enc_names
        self.accelerator = accelerator
        self.latents_scale = latents_scale
        self.latents_bias = latents_bias
        self.loss_weights = loss_weights
        self.time_schedule = time_schedule
        self.cutoffs = cutoffs
        assert len(loss_weights) == len(enc_names), "Loss weights must be provided for each encoder."

    # -- small host/PyTorch helpers kept for API parity ------------------------------------------------
    def interpolant(self, t):
        if self.path_type == "linear":
            return 1 - t, t, -1, 1
        if self.path_type == "cosine":
            ang = t * np.pi / 2
            return torch.cos(ang), torch.sin(ang), -np.pi / 2 * torch.sin(ang), np.pi / 2 * torch.cos(ang)
        raise NotImplementedError()

    def encoder_weight(self, base_weight: float, current_step: int, total_steps: int, schedule: str = "linear",
                       focus: str = "text", transition_point: float = 0.5, sharpness: float = 10) -> float:
        progress = current_step / total_steps
        if schedule == "linear":
            toward_image = progress
        elif schedule == "cosine":
            toward_image = 0.5 * (1 - math.cos(math.pi * progress))
        elif schedule == "sigmoid":
            toward_image = 1 - 1 / (1 + math.exp((progress - transition_point) * sharpness))
        else:
            raise ValueError("Invalid schedule. Choose from 'linear', 'cosine', 'sigmoid'.")
        scale = toward_image if focus == "image" else 1 - toward_image
        return base_weight * scale

    def time_weight(self, t: torch.Tensor, base_weight: float = 1.0, schedule: str = "constant",
                    cutoffs: list = [0.0, 1.0]) -> torch.Tensor:
        if schedule == "constant":
            scale = torch.ones_like(t)
        elif schedule == "linear":
            scale = 1 - t
        elif schedule == "cosine":
            scale = 0.5 * (1 + torch.cos(math.pi * t))
        elif schedule == "sigmoid":
            scale = 1 / (1 + torch.exp((t - 0.5) * 10))
        elif schedule == "loglinear":
            scale = 1 - torch.log(t + 1)
        elif schedule == "cutoff":
            # same values as the reference's masked assignments (loss.py:145-147), without index_put: that needs a
            # device sync and cannot be captured in a CUDA graph
            scale = ((t >= cutoffs[0]) & (t <= cutoffs[1])).to(t.dtype)
        else:
            raise ValueError("Invalid schedule. Choose from 'linear', 'cosine', 'sigmoid'.")
        return base_weight * scale

    def _sample_time(self, batch):
        """CPU-generator draw, shape (B,1,1,1)  (loss.py:158-168)."""
        if self.weighting == "uniform":
            return torch.rand((batch, 1, 1, 1))
        if self.weighting == "lognormal":
            sigma = torch.randn((batch, 1, 1, 1)).exp()
            if self.path_type == "linear":
                return sigma / (1 + sigma)
            if self.path_type == "cosine":
                return 2 / np.pi * torch.atan(sigma)
        raise NotImplementedError(f"weighting={self.weighting!r} path_type={self.path_type!r}")

    # -- the loss ------------------------------------------------------------------------------------------
    def __call__(self, model, images, model_kwargs=None, zs=None, **kwargs):
        if model_kwargs == None:  # noqa: E711  (same truthiness rule as the reference)
            model_kwargs = {}
        if not images.is_cuda:
            raise RuntimeError("reed_b200.SILoss runs on CUDA (sm_100a) tensors only; there is no CPU fallback")
        if self.path_type not in _PATH_CODE:
            raise NotImplementedError()
        if self.prediction != 'v':
            raise NotImplementedError()
        path = _PATH_CODE[self.path_type]

        # `time_input=` (a (B,1,1,1) tensor) replaces the CPU-generator draw: the reference swallows unknown **kwargs,
        # so this is an extension, used by ReedTrainer to keep host RNG out of a captured CUDA graph
        time_input = kwargs.get("time_input")
        if time_input is None:
            time_input = self._sample_time(images.shape[0])
        time_input = time_input.to(device=images.device, dtype=images.dtype)
        noises = torch.randn_like(images)
        t32 = time_input.flatten().float().contiguous()
        x32 = images.float().contiguous()
        n32 = noises.float().contiguous()
        model_input = ops.interpolate(x32, n32, t32, path).to(images.dtype)

        model_kwargs['inference'] = False
        model_output, zs_tilde = model(model_input, time_input.flatten(), **model_kwargs)
        denoising_loss = ops.VelocityMSEFn.apply(model_output.float(), x32, n32, t32, path)

        proj_loss = 0.
        acc = {"image": [0., 0], "text": [0., 0]}
        save = kwargs.get("save_projloss", False)
        bsz = zs[0].shape[0]
        if save:
            loss_saver = {"image": torch.zeros(bsz, device=images.device), "text": torch.zeros(bsz, device=images.device),
                          "time": time_input.flatten()}
        for z, z_tilde, enc_name in zip(zs, zs_tilde, self.enc_names):
            base = self.loss_weights.get(enc_name, 1.0)
            wts = self.time_weight(time_input, base, self.time_schedule, self.cutoffs)
            key = "image" if enc_name in IMAGE_ENCODERS or len(self.enc_names) == 1 else "text"
            if z.ndim == 2:
                assert key == "text", "Only text encoders should have 2D embeddings."
                assert z_tilde.ndim == 2, "Pooling to 2D to align with text embeddings."
                z, z_tilde = z.unsqueeze(1), z_tilde.unsqueeze(1)
            if base == 0.0:
                wts = torch.ones_like(wts)
            if z.dtype not in (torch.float32, torch.bfloat16):
                z = z.float()
            curr_loss = ops.CosineAlignFn.apply(z_tilde, z)                      # (B,)
            weighted_loss = (curr_loss * wts).mean()      # (B,)*(B,1,1,1) -> (B,1,1,B): reference broadcasting kept
            proj_loss += weighted_loss
            acc[key][0] += curr_loss.mean()
            acc[key][1] += 1
            if save:
                loss_saver[key] += curr_loss
        img_proj_loss = acc["image"][0] / max(1, acc["image"][1])
        text_proj_loss = acc["text"][0] / max(1, acc["text"][1])
        out = {"denoising_loss": denoising_loss, "proj_loss": proj_loss, "img_proj_loss": img_proj_loss,
               "text_proj_loss": text_proj_loss}
        if save:
            out["loss_saver"] = loss_saver
        return out
