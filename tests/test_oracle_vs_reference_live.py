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
        assert float((torch.as_tensor(got[key]).detach().double() - w.detach()).abs()) < 5e-6 * max(1.0, float(w.detach().abs())), key
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


def test_product_host_helpers_against_live_reference(ref):
    """The pure-PyTorch/host pieces of the drop-in modules (they run on CPU tensors in both implementations): SILoss
    constructor state, interpolant, time_weight, encoder_weight, the CPU time draw; sampler helper functions; model zoo."""
    ref_sit, ref_loss, ref_samplers, _ = ref
    from reed_b200.image import loss as my_loss, samplers as my_samplers
    from reed_b200.image.models import sit as my_sit
    assert my_loss.IMAGE_ENCODERS == ref_loss.IMAGE_ENCODERS
    t = torch.rand(6, 1, 1, 1, generator=torch.Generator().manual_seed(0))
    for path_type in ("linear", "cosine"):
        kw = dict(path_type=path_type, enc_names=["dinov2", "t5"], loss_weights={"dinov2": 1.0, "t5": 0.5},
                  time_schedule="cosine", cutoffs=[0.1, 0.9])
        a, b = my_loss.SILoss(**kw), ref_loss.SILoss(**kw)
        for attr in ("prediction", "weighting", "path_type", "enc_names", "loss_weights", "time_schedule", "cutoffs"):
            assert getattr(a, attr) == getattr(b, attr), attr
        for x, y in zip(a.interpolant(t), b.interpolant(t)):
            assert torch.equal(torch.as_tensor(x), torch.as_tensor(y))
        for sched in ("linear", "cosine", "sigmoid", "constant", "loglinear", "cutoff"):
            assert torch.equal(a.time_weight(t, 0.7, sched, [0.25, 0.75]), b.time_weight(t, 0.7, sched, [0.25, 0.75])), sched
        for sched in ("linear", "cosine", "sigmoid"):
            for focus in ("text", "image"):
                assert a.encoder_weight(1.5, 3, 10, sched, focus) == b.encoder_weight(1.5, 3, 10, sched, focus)
    # the CPU-generator time draw of __call__ (loss.py:158-168): same stream, same values
    for weighting, path_type in (("uniform", "linear"), ("lognormal", "linear"), ("lognormal", "cosine")):
        mine = my_loss.SILoss(weighting=weighting, path_type=path_type, enc_names=["dinov2"], loss_weights={"dinov2": 1.0})
        torch.manual_seed(11)
        got = mine._sample_time(5)
        torch.manual_seed(11)
        want = loss_oracle.draw_time(5, weighting, path_type)
        assert torch.equal(got, want)
    # sampler helpers (samplers.py:5-43)
    x = torch.randn(3, 4, 8, 8, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    v = torch.randn(3, 4, 8, 8, dtype=torch.float64, generator=torch.Generator().manual_seed(2))
    tt = torch.tensor([0.9, 0.5, 0.04], dtype=torch.float64)
    assert torch.equal(my_samplers.expand_t_like_x(tt, x), ref_samplers.expand_t_like_x(tt, x))
    assert my_samplers.compute_diffusion(0.3) == ref_samplers.compute_diffusion(0.3)
    for path_type in ("linear", "cosine"):
        got = my_samplers.get_score_from_velocity(v, x, tt, path_type)
        want = ref_samplers.get_score_from_velocity(v, x, tt, path_type)
        assert float((got - want).abs().max()) <= 1e-12 * float(want.abs().max())
    with pytest.raises(NotImplementedError):
        my_samplers.get_score_from_velocity(v, x, tt, "edm")
    # zoo, helper functions and the sin-cos table
    assert sorted(my_sit.SiT_models) == sorted(ref_sit.SiT_models)
    import numpy as np
    for dim, grid in ((64, 4), (1152, 16), (384, 32)):
        assert np.array_equal(my_sit.get_2d_sincos_pos_embed(dim, grid), ref_sit.get_2d_sincos_pos_embed(dim, grid))
    xs, sh, sc = torch.randn(2, 5, 8), torch.randn(2, 8), torch.randn(2, 8)
    assert torch.equal(my_sit.modulate(xs, sh, sc), ref_sit.modulate(xs, sh, sc))
    proj_a = my_sit.build_mlp(16, 32, 24)
    proj_b = ref_sit.build_mlp(16, 32, 24)
    assert [tuple(p.shape) for p in proj_a.parameters()] == [tuple(p.shape) for p in proj_b.parameters()]
    assert list(proj_a.state_dict()) == list(proj_b.state_dict())


def test_product_model_tree_against_live_reference(ref):
    """Constructing the drop-in SiT consumes the RNG like the reference and yields the same state_dict (keys, shapes, VALUES)
    for several zoo entries and keyword combinations; unpatchify agrees."""
    ref_sit, _, _, _ = ref
    from reed_b200.image.models import sit as my_sit
    combos = [("SiT-S/2", dict(decoder_hidden_size=384, z_dims=[768], z_types=["i"], encoder_depth=8, qk_norm=False)),
              ("SiT-S/4", dict(z_dims=[64, 32], z_types=["i", "t"], encoder_depth=4, encoder_depth_text=6, qk_norm=True)),
              ("SiT-B/8", dict(z_dims=[128], z_types=["i"], encoder_depth=2, qk_norm=False, class_dropout_prob=0.0,
                               num_classes=10))]
    for name, kw in combos:
        torch.manual_seed(5)
        mine = my_sit.SiT_models[name](input_size=32, use_cfg=True, fused_attn=True, **kw)
        after_mine = torch.rand(1)
        torch.manual_seed(5)
        theirs = ref_sit.SiT_models[name](input_size=32, use_cfg=True, fused_attn=True, **kw)
        after_theirs = torch.rand(1)
        assert torch.equal(after_mine, after_theirs), name                    # same number of RNG draws
        a, b = mine.state_dict(), theirs.state_dict()
        assert list(a) == list(b), name
        for k in a:
            assert a[k].shape == b[k].shape and torch.equal(a[k], b[k]), (name, k)
        assert [n for n, p in mine.named_parameters() if not p.requires_grad] == \
            [n for n, p in theirs.named_parameters() if not p.requires_grad]
        tokens = mine.x_embedder.num_patches
        y = torch.randn(2, tokens, mine.patch_size ** 2 * mine.out_channels)
        assert torch.equal(mine.unpatchify(y), theirs.unpatchify(y)), name


@pytest.mark.parametrize("hyper", [dict(lr=3e-3, betas=(0.8, 0.95), eps=1e-6, weight_decay=0.05, max_norm=0.3, ema_decay=0.99),
                                   dict(lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_norm=1.0, ema_decay=0.9999)])
def test_optimizer_glue_against_torch_adamw_and_reference_ema(ref, hyper):
    """train.py:94-105,253-259,402-412: clip_grad_norm_ -> torch.optim.AdamW -> the reference's update_ema, several steps,
    non-default hyper-parameters included (the committed train_glue fixture holds the defaults)."""
    from oracle import train_oracle
    _, _, _, ns = ref
    g = torch.Generator().manual_seed(0)
    shapes = {"a.weight": (7, 5), "a.bias": (7,), "b.weight": (3, 7), "frozen": (4,)}
    mod = torch.nn.ParameterDict({k.replace(".", "_"): torch.nn.Parameter(torch.randn(s, generator=g)) for k, s in shapes.items()})
    mod["frozen"].requires_grad_(False)
    import copy
    ema_mod = copy.deepcopy(mod)
    opt = torch.optim.AdamW(mod.parameters(), lr=hyper["lr"], betas=hyper["betas"], eps=hyper["eps"],
                            weight_decay=hyper["weight_decay"])
    params = {k: v.detach().clone() for k, v in mod.items()}
    ema = {k: v.clone() for k, v in params.items()}
    m1 = {k: torch.zeros_like(v) for k, v in params.items()}
    m2 = {k: torch.zeros_like(v) for k, v in params.items()}
    for step in range(1, 6):
        grads = {k: torch.randn(v.shape, generator=g) * (10.0 if step == 2 else 0.1) for k, v in params.items() if k != "frozen"}
        for k, p in mod.items():
            p.grad = grads[k].clone() if k in grads else None
        want_norm = torch.nn.utils.clip_grad_norm_(mod.parameters(), hyper["max_norm"])
        opt.step()
        ns["update_ema"](ema_mod, mod, decay=hyper["ema_decay"])
        got_norm = train_oracle.adamw_ema_step(params, grads, m1, m2, ema, step, lr=hyper["lr"], beta1=hyper["betas"][0],
                                               beta2=hyper["betas"][1], eps=hyper["eps"], weight_decay=hyper["weight_decay"],
                                               max_norm=hyper["max_norm"], ema_decay=hyper["ema_decay"])
        assert _rel(got_norm, want_norm) < 1e-6
        for k in params:
            assert float((params[k] - mod[k].detach()).abs().max()) < 2e-6, (step, k)
            assert float((ema[k] - ema_mod[k].detach()).abs().max()) < 2e-6, (step, k)


def test_legacy_checkpoint_renaming_against_live_reference():
    import ast
    from reed_b200.image.generate import load_legacy_checkpoints
    src = open("/root/reference/image/utils.py").read()
    ns = {}
    for node in ast.parse(src).body:                      # utils.py itself needs timm / torchvision models: take the function only
        if isinstance(node, ast.FunctionDef) and node.name == "load_legacy_checkpoints":
            exec(compile(ast.Module([node], []), "utils.py", "exec"), ns)
    sd = {"pos_embed": 1, "blocks.0.attn.qkv.weight": 2, "decoder_blocks.0.mlp.fc1.bias": 3, "decoder_blocks.19.attn.proj.weight": 4,
          "final_layer.linear.weight": 5, "projectors.0.0.weight": 6}
    for depth in (8, 1):
        assert load_legacy_checkpoints(sd, depth) == ns["load_legacy_checkpoints"](sd, depth)
        assert list(load_legacy_checkpoints(sd, depth)) == list(ns["load_legacy_checkpoints"](sd, depth))
