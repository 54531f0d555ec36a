"""Sampling driver for the Euler / Heun / Euler-Maruyama samplers: CUDA-graph replay of the
model evaluation, and the VAE-free part of the reference's ``generate.py`` loop.

Restates /root/reference/image/generate.py:
  * 48-51     per-rank seed ``global_seed * world_size + rank``
  * 57-85     model construction for sampling; EMA weights with the ``projectors.*`` keys removed, ``strict=False``
  * 105-121   how many iterations each rank runs (``ceil(num_fid_samples / global_batch) * global_batch`` samples)
  * 122-149   per iteration: ``z = randn(n, C, S, S)``, ``y = randint(0, num_classes, (n,))`` on the device, then the
              ODE (``euler_sampler``) or SDE (``euler_maruyama_sampler``) sampler, result cast to fp32
  * 164-165   sample ``i`` of an iteration has global index ``i * world_size + rank + total``
The VAE decode and PNG / .npz writers (150-166, 19-34) are outside this path: ``sample_latents`` returns / stores the
fp32 latents, which is what the decode consumes.

Why a graph: one sampler step is ONE fused update kernel plus one model evaluation of ~240 kernel launches (SiT-XL/2),
and BASELINE configs[4] runs 250 of them per batch.  At the small per-rank batches sampling uses, enqueueing the
evaluation from Python takes longer than the kernels run.  ``GraphedSiT`` records the evaluation once per input shape
(conditional batch n, and 2n inside the classifier-free-guidance window) and replays it; the per-step scalars (t, dt,
cfg) stay outside the graph as arguments of the fused update kernel, so no re-capture is needed along the time grid.
"""
from __future__ import annotations

import math
import os
from typing import Dict, Optional

import numpy as np
import torch

from .. import ops
from .samplers import euler_maruyama_sampler, euler_sampler


class GraphedSiT:
    """An eval-mode SiT whose ``model(x, t, y=labels)`` calls replay a captured CUDA graph per input shape.

    Used in place of the model in ``euler_sampler`` / ``euler_maruyama_sampler`` (same call signature, returns
    ``(pred, None)``).  The returned ``pred`` is the graph's static output buffer: it is overwritten by the next call
    with the same input shape, which is exactly the samplers' access pattern (the fused update consumes it at once).
    """

    def __init__(self, model, warmup: int = 2):
        if model.training:
            raise ValueError("GraphedSiT wraps an eval-mode model (label dropout draws device randomness in train mode)")
        self.model = model
        self.warmup = max(1, warmup)
        self._graphs: Dict[tuple, dict] = {}
        self.replays = 0

    def __getattr__(self, name):           # in_channels, projectors, ... (generate.py:114,124)
        return getattr(self.__dict__["model"], name)

    def eval(self):
        return self

    def _capture(self, x, t, y):
        g = {"x": x.clone(), "t": t.clone(), "y": y.clone()}
        side = torch.cuda.Stream(device=x.device)
        side.wait_stream(torch.cuda.current_stream(x.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(self.warmup):          # sizes the allocator pools, creates bf16 weight shadows / caches
                self.model(g["x"], g["t"], y=g["y"])
        torch.cuda.current_stream(x.device).wait_stream(side)
        torch.cuda.synchronize(x.device)
        graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(graph):
            g["pred"] = self.model(g["x"], g["t"], y=g["y"])[0]
        g["graph"] = graph
        g["epoch"] = ops.weights_epoch
        return g

    def __call__(self, x, t, y=None, inference=True):
        if not inference:
            raise ValueError("GraphedSiT serves inference evaluations only (the samplers' calls)")
        if not x.is_cuda:
            raise RuntimeError("reed_b200 runs on CUDA (sm_100a) tensors only; there is no CPU fallback")
        key = (tuple(x.shape), x.dtype, t.dtype, y.dtype, torch.is_autocast_enabled(), self.model.reed_precision)
        g = self._graphs.get(key)
        if g is None:
            g = self._graphs[key] = self._capture(x, t, y)
        if g["epoch"] != ops.weights_epoch:
            # the weights were rewritten since the capture (a train step updated the EMA this model is): the graph
            # reads the bf16 shadows through baked-in pointers, so re-cast them in place before replaying
            ops.refresh_shadows(self.model)
            g["epoch"] = ops.weights_epoch
        g["x"].copy_(x)
        g["t"].copy_(t)
        g["y"].copy_(y)
        g["graph"].replay()
        self.replays += 1
        return g["pred"], None


def rank_seed(global_seed: int, rank: int, world_size: int) -> int:
    """generate.py:50."""
    return global_seed * world_size + rank


def sampling_plan(num_fid_samples: int, per_proc_batch_size: int, world_size: int):
    """generate.py:105-118 -> (total_samples, iterations per rank)."""
    global_batch = per_proc_batch_size * world_size
    total = int(math.ceil(num_fid_samples / global_batch) * global_batch)
    assert total % world_size == 0, "total_samples must be divisible by world_size"
    per_rank = total // world_size
    assert per_rank % per_proc_batch_size == 0, "samples_needed_this_gpu must be divisible by the per-GPU batch size"
    return total, per_rank // per_proc_batch_size


def sample_indices(n: int, rank: int, world_size: int, total_so_far: int):
    """Global index of each of the n samples of one iteration on one rank (generate.py:164)."""
    return [i * world_size + rank + total_so_far for i in range(n)]


def load_legacy_checkpoints(state_dict, encoder_depth):
    """/root/reference/image/utils.py:207-219: early checkpoints kept the blocks after the encoder tap under
    ``decoder_blocks.{i}``; they are ``blocks.{i + encoder_depth}`` of the current model."""
    renamed = {}
    for key, value in state_dict.items():
        if "decoder_blocks" in key:
            parts = key.split(".")
            parts[0], parts[1] = "blocks", str(int(parts[1]) + encoder_depth)
            key = ".".join(parts)
        renamed[key] = value
    return renamed


def load_sampling_weights(model, state_dict, strict_backbone: bool = True, legacy: bool = False,
                          encoder_depth: Optional[int] = None):
    """generate.py:77-85: load EMA weights without the projector heads (they are not evaluated at inference);
    ``legacy`` applies ``load_legacy_checkpoints`` first (generate.py:80-83)."""
    if legacy:
        state_dict = load_legacy_checkpoints(state_dict, model.encoder_depth if encoder_depth is None else encoder_depth)
    sd = {k: v for k, v in state_dict.items() if "projectors" not in k}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    if strict_backbone:
        bad = [k for k in missing if "projectors" not in k] + list(unexpected)
        if bad:
            raise KeyError(f"checkpoint does not match the model: {bad[:8]}")
    return model


@torch.no_grad()
def sample_latents(model, *, num_fid_samples: int, per_proc_batch_size: int = 32, latent_size: int = 32,
                   num_classes: int = 1000, mode: str = "ode", num_steps: int = 50, heun: bool = False,
                   cfg_scale: float = 1.5, guidance_low: float = 0.0, guidance_high: float = 1.0,
                   path_type: str = "linear", global_seed: int = 0, rank: int = 0, world_size: int = 1,
                   device: Optional[torch.device] = None, graphed: bool = True, out_dir: Optional[str] = None,
                   seed_rng: bool = True):
    """The sampling loop of generate.py:119-167 up to the VAE decode.

    Returns ``(latents [k, C, S, S] fp32 on the CPU, labels [k], global indices [k])`` for this rank; with ``out_dir``
    each iteration is also written as ``<out_dir>/latents-rank{r}-{iteration:05d}.npz``.
    """
    assert cfg_scale >= 1.0, "In almost all cases, cfg_scale be >= 1.0"           # generate.py:88
    if mode not in ("ode", "sde"):
        raise NotImplementedError(mode)
    device = device if device is not None else next(model.parameters()).device
    if seed_rng:
        torch.manual_seed(rank_seed(global_seed, rank, world_size))
    _, iterations = sampling_plan(num_fid_samples, per_proc_batch_size, world_size)
    runner = GraphedSiT(model) if graphed and not isinstance(model, GraphedSiT) else model
    sampler = euler_maruyama_sampler if mode == "sde" else euler_sampler
    n = per_proc_batch_size
    lat, lab, idx = [], [], []
    total = 0
    if out_dir is not None:
        os.makedirs(out_dir, exist_ok=True)
    for it in range(iterations):
        z = torch.randn(n, model.in_channels, latent_size, latent_size, device=device)
        y = torch.randint(0, num_classes, (n,), device=device)
        samples = sampler(runner, z, y, num_steps=num_steps, heun=heun, cfg_scale=cfg_scale, guidance_low=guidance_low,
                          guidance_high=guidance_high, path_type=path_type).to(torch.float32)
        where = sample_indices(n, rank, world_size, total)
        lat.append(samples.cpu())
        lab.append(y.cpu())
        idx.extend(where)
        if out_dir is not None:
            np.savez(os.path.join(out_dir, f"latents-rank{rank}-{it:05d}.npz"), latents=lat[-1].numpy(),
                     labels=lab[-1].numpy(), indices=np.asarray(where, dtype=np.int64))
        total += n * world_size
    return torch.cat(lat), torch.cat(lab), torch.tensor(idx, dtype=torch.int64)
