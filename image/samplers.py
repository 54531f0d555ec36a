"""Drop-in for the reference's ``image/samplers.py``: ``from samplers import euler_sampler, euler_maruyama_sampler``
(generate.py:17) with cwd = image/."""
import _reed_path  # noqa: F401

from reed_b200.image.samplers import (  # noqa: F401
    compute_diffusion, euler_maruyama_sampler, euler_sampler, expand_t_like_x, get_score_from_velocity)
