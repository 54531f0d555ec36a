"""Sweep of the CPU oracle against the UNMODIFIED reference executed live (/root/reference/image with oracle/timm_shim), over
architectures the committed fixtures do not hold: patch sizes 2/4/8, uneven head counts, qk_norm, image-only / image+text heads
with and without a separate text tap, both path types and weightings, every time schedule.  Runs only where the reference
checkout exists (the build container); the GPU box relies on the committed fixtures."""
import os

import pytest
import torch

from oracle import loss_oracle, samplers_oracle, sit_oracle
from oracle.fixtures import random_batch, random_state
from oracle.sit_oracle import ArchSpec

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/image"), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    from oracle import make_golden
    return make_golden._import_reference()


def _rel(a, b):
    a, b = torch.as_tensor(a).detach().double(), torch.as_tensor(b).detach().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


SWEEP = [
    # (spec kwargs, enc_names, loss_weights, path_type, weighting, time_schedule)
    (dict(input_size=16, patch_size=4, hidden_size=96, decoder_hidden_size=96, depth=3, num_heads=3, encoder_depth=2,
          z_dims=[48], z_types=["i"], projector_dim=64), ["dinov2"], {"dinov2": 1.0}, "linear", "uniform", "constant"),
    (dict(input_size=32, patch_size=8, hidden_size=64, decoder_hidden_size=64, depth=2, num_heads=4, encoder_depth=1,
          z_dims=[32], z_types=["i"], projector_dim=48, qk_norm=True), ["clip"], {"clip": 0.3}, "cosine", "uniform", "sigmoid"),
    (dict(input_size=8, patch_size=2, hidden_size=80, decoder_hidden_size=80, depth=4, num_heads=5, encoder_depth=2,
          encoder_depth_text=2, z_dims=[24, 40], z_types=["i", "t"], projector_dim=32),
     ["mae", "text_embeds_qwenvl_7b_layer_15"], {"mae": 1.0, "text_embeds_qwenvl_7b_layer_15": 0.5}, "linear", "lognormal",
     "cosine"),
    (dict(input_size=8, patch_size=2, hidden_size=64, decoder_hidden_size=64, depth=4, num_heads=2, encoder_depth=1,
          encoder_depth_text=3, z_dims=[16, 56], z_types=["i", "t"], projector_dim=32, class_dropout_prob=0.5),
     ["jepa", "t5"], {"jepa": 0.7, "t5": 0.0}, "cosine", "lognormal", "cutoff"),
    (dict(input_size=16, patch_size=2, hidden_size=48, decoder_hidden_size=48, depth=2, num_heads=1, encoder_depth=2,
          z_dims=[16], z_types=["i"], projector_dim=32, class_dropout_prob=0.0), ["mocov3"], {"mocov3": 2.0}, "linear",
     "uniform", "loglinear"),
]


@pytest.mark.parametrize("case", range(len(SWEEP)))
def test_loss_and_gradients_against_live_reference(ref, case):
    from oracle.make_golden import _ref_model, _replay_draws
    ref_sit, ref_loss, _, _ = ref
    kw, enc_names, weights, path_type, weighting, schedule = SWEEP[case]
    spec = ArchSpec(num_classes=1000, **kw)
    sd = random_state(spec, 100 + case)
    data = random_batch(spec, 3, 200 + case)
    model = _ref_model(ref_sit, spec, sd).train()
    t, noise, drop = _replay_draws(300 + case, data["x"], spec.class_dropout_prob, weighting, path_type)
    fn = ref_loss.SILoss(prediction="v", path_type=path_type, weighting=weighting, enc_names=list(enc_names),
                         loss_weights=dict(weights), time_schedule=schedule, cutoffs=[0.2, 0.8])
    torch.manual_seed(300 + case)
    want = fn(model, data["x"], dict(y=data["y"]), zs=data["zs"])
    (want["denoising_loss"].mean() + 0.5 * want["proj_loss"]).backward()

    leaves = {k: v.clone().requires_grad_(k != "pos_embed") for k, v in sd.items()}
    got = loss_oracle.si_loss(sit_oracle.as_model(leaves, spec, training=True, drop_mask=drop), data["x"], t, noise,
                              data["zs"], enc_names=enc_names, loss_weights=weights, model_kwargs=dict(y=data["y"]),
                              path_type=path_type, time_schedule=schedule, cutoffs=[0.2, 0.8])
    (got["denoising_loss"].mean() + 0.5 * got["proj_loss"]).backward()
    assert _rel(got["denoising_loss"], want["denoising_loss"]) < 5e-6
    for key in ("proj_loss", "img_proj_loss", "text_proj_loss"):
        w = torch.as_tensor(want[key]).double()
        assert float((torch.as_tensor(got[key]).double() - w).abs()) < 5e-6 * max(1.0, float(w.abs())), key
    for name, p in model.named_parameters():
        if p.grad is None:
            assert leaves[name].grad is None or float(leaves[name].grad.abs().max()) == 0.0, name
            continue
        g = leaves[name].grad
        assert g is not None, name
        denom = float(p.grad.abs().max())
        assert float((g - p.grad).abs().max()) <= 2e-5 * max(denom, 1e-6), name


@pytest.mark.parametrize("patch", [2, 4])
def test_samplers_against_live_reference(ref, patch):
    from oracle.make_golden import _ref_model
    ref_sit, _, ref_samplers, _ = ref
    spec = ArchSpec(input_size=8 * patch, patch_size=patch, hidden_size=64, decoder_hidden_size=64, depth=2, num_heads=2,
                    encoder_depth=1, z_dims=[16], z_types=["i"], projector_dim=32, num_classes=1000)
    sd = random_state(spec, 7 + patch)
    model = _ref_model(ref_sit, spec, sd).eval()
    g = torch.Generator().manual_seed(patch)
    z = torch.randn(2, 4, spec.input_size, spec.input_size, generator=g)
    y = torch.randint(0, 1000, (2,), generator=g)
    mine = sit_oracle.as_model(sd, spec)
    want = ref_samplers.euler_sampler(model, z, y, num_steps=4, heun=True, cfg_scale=1.7, guidance_low=0.1, guidance_high=0.9)
    got = samplers_oracle.euler(mine, z, y, num_steps=4, heun=True, cfg_scale=1.7, guidance_low=0.1, guidance_high=0.9)
    assert got.dtype == want.dtype == torch.float64 and float((got - want).abs().max()) < 1e-5
    torch.manual_seed(3)
    noises = [torch.randn(z.shape, dtype=torch.float64) for _ in range(4)]
    torch.manual_seed(3)
    want = ref_samplers.euler_maruyama_sampler(model, z, y, num_steps=5, cfg_scale=2.0, guidance_high=0.6, path_type="cosine")
    got = samplers_oracle.euler_maruyama(mine, z, y, num_steps=5, cfg_scale=2.0, guidance_high=0.6, path_type="cosine",
                                         noises=noises)
    assert float((got - want).abs().max()) < 1e-5
