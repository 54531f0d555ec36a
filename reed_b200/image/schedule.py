"""Curriculum scalars of the REED train step (BASELINE configs[2]: "REPA loss + curriculum weighting").

Restates the two host-side schedules the reference computes inline at every step,
/root/reference/image/train.py:363-385, from the arguments declared at train.py:513,544-548:

  * ``repa_weight_decay``     weight of the representation-alignment term over ``repa_steps``
  * ``diffusion_loss_decay``  weight of the denoising term: zero during the alignment-only pre-training stage
                              (``start_diffusion_steps``), linear warm-up over ``diffusion_warm_up_steps``, then the
                              chosen decay until ``max_train_steps``

They are plain Python floats; ``ReedTrainer.train_step(_graphed)`` takes them as ``repa_decay`` / ``diffusion_decay``
and feeds them to the captured step as device scalars, so the CUDA graph never has to be re-recorded.

Kept quirk: the reference's cosine diffusion decay reads ``np.pi * (step - top) / max_train_steps - top`` (the
subtraction binds after the division, train.py:383).  ``cos`` of that is what the reference trains with, so it is what
this returns; ``Curriculum(strict_reference=False)`` evaluates the evidently intended ``/(max_train_steps - top)``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

_KINDS = ("constant", "linear", "cosine")


def repa_weight_decay(global_step: int, kind: str = "constant", repa_steps: int = 400000) -> float:
    """train.py:363-370."""
    if kind == "constant":
        return 1.0
    if kind == "linear":
        return max(1.0 - global_step / repa_steps, 0.0)
    if kind == "cosine":
        return max((1.0 + math.cos(math.pi * global_step / repa_steps)) / 2, 0.0)
    raise NotImplementedError(kind)


def diffusion_loss_decay(global_step: int, kind: str = "constant", start_diffusion_steps: int = 0,
                         diffusion_warm_up_steps: int = 50000, max_train_steps: int = 400000,
                         strict_reference: bool = True) -> float:
    """train.py:372-385."""
    top = diffusion_warm_up_steps + start_diffusion_steps
    if global_step < start_diffusion_steps:
        return 0.0
    if start_diffusion_steps <= global_step < top:
        return (global_step - start_diffusion_steps) / diffusion_warm_up_steps
    if kind == "constant":
        return 1.0
    if kind == "linear":
        return 1.0 - (global_step - top) / (max_train_steps - top)
    if kind == "cosine":
        if strict_reference:
            return (1.0 + math.cos(math.pi * (global_step - top) / max_train_steps - top)) / 2
        return (1.0 + math.cos(math.pi * (global_step - top) / (max_train_steps - top))) / 2
    raise NotImplementedError(kind)


@dataclass
class Curriculum:
    """The pair of schedules with the reference's argument names and defaults (train.py:513,544-548)."""
    repa_weight_decay: str = "constant"
    repa_steps: int = 400000
    start_diffusion_steps: int = 0
    diffusion_warm_up_steps: int = 50000
    diffusion_decay: str = "constant"
    max_train_steps: int = 400000
    strict_reference: bool = True

    def __post_init__(self):
        for kind in (self.repa_weight_decay, self.diffusion_decay):
            if kind not in _KINDS:
                raise NotImplementedError(kind)

    def __call__(self, global_step: int):
        """-> (diffusion_decay, repa_decay) for the step about to run, in ``ReedTrainer.train_step`` order."""
        return (diffusion_loss_decay(global_step, self.diffusion_decay, self.start_diffusion_steps,
                                     self.diffusion_warm_up_steps, self.max_train_steps, self.strict_reference),
                repa_weight_decay(global_step, self.repa_weight_decay, self.repa_steps))
