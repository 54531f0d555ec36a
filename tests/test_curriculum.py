"""Curriculum scalars (train.py:363-385): oracle and product schedule against values computed by the reference source."""
import pytest

from oracle import train_oracle
from reed_b200.image.schedule import Curriculum, diffusion_loss_decay, repa_weight_decay

_ARGS = ("repa_weight_decay", "repa_steps", "start_diffusion_steps", "diffusion_warm_up_steps", "diffusion_decay",
         "max_train_steps")


def test_oracle_curriculum_matches_reference(golden):
    cases = golden("curriculum.pt")
    assert len(cases) == 3 * 3 * 2 * 2 * 17
    for c in cases:
        diff, repa = train_oracle.curriculum(c["global_step"], **{k: c[k] for k in _ARGS})
        assert diff == c["diffusion"] and repa == c["repa"], c


def test_product_curriculum_matches_reference(golden):
    for c in golden("curriculum.pt"):
        diff, repa = Curriculum(**{k: c[k] for k in _ARGS})(c["global_step"])
        assert diff == pytest.approx(c["diffusion"], rel=1e-12, abs=1e-15), c
        assert repa == pytest.approx(c["repa"], rel=1e-12, abs=1e-15), c


def test_curriculum_shape_and_errors():
    assert repa_weight_decay(10, "constant") == 1.0
    assert repa_weight_decay(5000, "linear", 4000) == 0.0                       # clamped at zero past repa_steps
    assert diffusion_loss_decay(10, "constant", start_diffusion_steps=100) == 0.0    # alignment-only stage
    assert diffusion_loss_decay(150, "constant", 100, 100) == 0.5                # warm-up
    assert diffusion_loss_decay(200, "linear", 100, 100, 1200) == 1.0
    assert diffusion_loss_decay(1200, "linear", 100, 100, 1200) == 0.0
    # the intended cosine (strict_reference=False) goes 1 -> 0; the reference's own expression does not
    assert diffusion_loss_decay(1200, "cosine", 100, 100, 1200, strict_reference=False) == pytest.approx(0.0, abs=1e-12)
    with pytest.raises(NotImplementedError):
        repa_weight_decay(1, "exp")
    with pytest.raises(NotImplementedError):
        Curriculum(diffusion_decay="exp")
