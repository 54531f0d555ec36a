"""GPU tests of gradient accumulation, the preprocessing kernel, the other BASELINE configs' size-independent properties
and the SiT-*/4, */8 geometries.  Written at the end of round 1, first run on a B200 in round 2 (profiles/r02_tests_gpu.txt)."""
import pytest
import torch

from oracle.fixtures import random_batch, random_state
from oracle.sit_oracle import ArchSpec

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _tiny(precision, seed=11):
    from reed_b200.image.loss import SILoss
    from reed_b200.image.models.sit import SiT
    from reed_b200.image.trainer import ReedTrainer
    spec = ArchSpec(input_size=16, hidden_size=128, decoder_hidden_size=128, depth=2, num_heads=2, encoder_depth=1,
                    z_dims=[64], projector_dim=128)
    m = SiT(path_type="linear", use_cfg=True, input_size=16, hidden_size=128, decoder_hidden_size=128, depth=2, num_heads=2,
            encoder_depth=1, z_dims=[64], z_types=["i"], projector_dim=128, num_classes=1000, fused_attn=True, qk_norm=False)
    m.load_state_dict(random_state(spec, seed))
    m = m.to(DEV).train()
    return spec, ReedTrainer(m, SILoss(enc_names=["dinov2"], loss_weights={"dinov2": 1.0}), precision=precision)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_gradient_accumulation_averages_micro_batch_gradients(precision):
    """train.py:362 `accelerator.accumulate`: the accumulated flat gradient of k micro-batches = mean of their gradients."""
    spec, tr = _tiny(precision)
    to_dev = lambda d: (d["x"].to(DEV), d["y"].to(DEV), [z.to(DEV) for z in d["zs"]])
    micro = [to_dev(random_batch(spec, 4, 70 + i)) for i in range(3)]
    singles = []
    for seed, batch in zip((1, 2, 3), micro):
        torch.manual_seed(seed)
        tr.state.begin_step()
        loss, _ = tr.compute_loss(*batch)
        loss.backward()
        tr.state.finish_backward()
        singles.append([b.grad.clone() for b in tr.state.buckets])
    want = [sum(gs) / 3 for gs in zip(*singles)]
    seeds = iter((1, 2, 3))
    plain_loss = tr.compute_loss

    def seeded(*a, **k):
        torch.manual_seed(next(seeds))
        return plain_loss(*a, **k)

    tr.compute_loss = seeded
    tr.optimizer_step = lambda **k: None                      # keep the weights: look at the accumulated gradient only
    loss, outs = tr.train_step_accumulated(micro)
    assert len(outs) == 3 and torch.isfinite(loss)
    for b, w in zip(tr.state.buckets, want):
        err = float((b.grad - w).abs().max() / w.abs().max().clamp_min(1e-20))
        assert err < (1e-5 if precision == "fp32" else 1e-2), (b.name, err)
        assert float(torch.nn.functional.cosine_similarity(b.grad, w, dim=0)) > 0.9999


def test_single_micro_batch_accumulated_step_equals_plain_step():
    spec, a = _tiny("fp32")
    _, b = _tiny("fp32")
    batch = random_batch(spec, 4, 80)
    dev = (batch["x"].to(DEV), batch["y"].to(DEV), [z.to(DEV) for z in batch["zs"]])
    torch.manual_seed(4)
    la, _ = a.train_step(*dev, diffusion_decay=0.8, repa_decay=0.6)
    torch.manual_seed(4)
    lb, _ = b.train_step_accumulated([dev], diffusion_decay=0.8, repa_decay=0.6)
    assert abs(float(la) - float(lb)) <= 1e-6 * max(1.0, abs(float(la)))
    for ba, bb in zip(a.state.buckets, b.state.buckets):
        assert float((ba.param - bb.param).abs().max()) <= 2e-5
        assert float((ba.ema - bb.ema).abs().max()) <= 2e-5
    assert a.step_count == b.step_count == 1


def test_preprocess_raw_image_matches_reference_golden(golden):
    """train.py:53-74 on the GPU kernel: against outputs of the reference function, the tap-level oracle and torch's own ops."""
    import torch.nn.functional as F
    from oracle import preprocess_oracle
    from reed_b200.image.preprocess import preprocess_raw_image
    for name, case in golden("preprocess.pt").items():
        r = case["resolution"]
        x = torch.randint(0, 256, (2, 3, r, r), generator=torch.Generator().manual_seed(case["seed"]), dtype=torch.uint8)
        y = preprocess_raw_image(x.to(DEV), case["enc_type"])
        assert tuple(y.shape) == case["shape"] and str(y.dtype) == case["dtype"], name
        if case["enc_type"] == "siglip":
            assert torch.equal(y.cpu(), x)
            continue
        assert float((y[..., ::7, ::5].cpu() - case["sample"]).abs().max()) < 1e-5, name
        assert float((y.cpu() - preprocess_oracle.preprocess_raw_image(x, case["enc_type"])).abs().max()) < 1e-5, name
    # the same sequence of torch ops on the device (what the reference runs there), float input, bf16 output
    xf = torch.rand(3, 3, 256, 256, device=DEV) * 255
    mean = torch.tensor(preprocess_oracle.IMAGENET_DEFAULT_MEAN, device=DEV).view(1, 3, 1, 1)
    std = torch.tensor(preprocess_oracle.IMAGENET_DEFAULT_STD, device=DEV).view(1, 3, 1, 1)
    want = F.interpolate((xf / 255. - mean) / std, 224, mode="bicubic")
    assert float((preprocess_raw_image(xf, "dinov2-vit-b") - want).abs().max()) < 1e-5
    got16 = preprocess_raw_image(xf, "dinov2-vit-b", out_dtype=torch.bfloat16)
    assert got16.dtype == torch.bfloat16 and float((got16.float() - want).abs().max()) < 2e-2
    with pytest.raises(ValueError):
        preprocess_raw_image(torch.zeros(1, 3, 128, 128, device=DEV), "dinov2")


def _rel(a, b):
    a, b = torch.as_tensor(a).detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("config", ["multimodal", "imagenet512"])
def test_full_size_properties_other_baseline_configs(config):
    """BASELINE configs[3] (SiT-XL/2 with image + caption heads, text head tapped at block 16) and configs[4] (4x64x64 latents,
    1024 tokens) at full width and a small batch: size-independent properties, as test_full_size_properties_xl2_bf16 does for
    configs[2]."""
    from reed_b200.image.loss import SILoss
    from reed_b200.image.models.sit import SiT_models
    torch.manual_seed(0)
    if config == "multimodal":
        size, tokens = 32, 256
        model = SiT_models["SiT-XL/2"](input_size=32, num_classes=1000, use_cfg=True, z_dims=[768, 3584], z_types=["i", "t"],
                                       encoder_depth=8, encoder_depth_text=16, fused_attn=True, qk_norm=False)
        names = ["dinov2", "text_embeds_qwenvl_7b_layer_15"]
        fn = SILoss(enc_names=names, loss_weights={names[0]: 1.0, names[1]: 0.5})
        zs = [torch.randn(2, tokens, 768, device=DEV), torch.randn(2, 3584, device=DEV)]
    else:
        size, tokens = 64, 1024
        model = SiT_models["SiT-XL/2"](input_size=64, num_classes=1000, use_cfg=True, z_dims=[768], z_types=["i"],
                                       encoder_depth=8, fused_attn=True, qk_norm=False)
        fn = SILoss(enc_names=["dinov2"], loss_weights={"dinov2": 1.0})
        zs = [torch.randn(2, tokens, 768, device=DEV)]
    model = model.to(DEV).train()
    model.reed_precision = "bf16"
    assert model.x_embedder.num_patches == tokens
    x = torch.randn(2, 4, size, size, device=DEV)
    y = torch.randint(0, 1000, (2,), device=DEV)
    # reference init: every gate is zero -> the prediction is exactly 0 and the denoising loss is mean(target^2)
    torch.manual_seed(1)
    out = fn(model, x, dict(y=y), zs=zs)
    torch.manual_seed(1)
    torch.rand(2, 1, 1, 1)
    noise = torch.randn_like(x)
    assert _rel(out["denoising_loss"], ((noise - x) ** 2).flatten(1).mean(1)) < 1e-5
    assert -1.5 <= float(out["proj_loss"]) <= 1.5
    if config == "multimodal":
        assert torch.is_tensor(out["text_proj_loss"]) and -1.0 <= float(out["text_proj_loss"]) <= 1.0
        assert _rel(out["proj_loss"], out["img_proj_loss"] * 1.0 + out["text_proj_loss"] * 0.5) < 1e-4
    else:
        assert out["text_proj_loss"] == 0.0
    (out["denoising_loss"].mean() + 0.5 * out["proj_loss"]).backward()
    g = model.final_layer.linear.weight.grad
    assert g is not None and torch.isfinite(g).all() and float(g.abs().max()) > 0
    for proj in model.projectors:                              # both heads receive a gradient through their taps
        assert float(proj[4].weight.grad.abs().max()) > 0
    # the projection loss reaches the trunk up to its tap (block 8 / block 16), not beyond: gates are zero, so blocks are
    # the identity and only the residual stream carries the gradient to the patch embedding
    assert float(model.x_embedder.proj.weight.grad.abs().max()) > 0
    # batch-permutation equivariance at inference with non-trivial gates
    model.eval()
    with torch.no_grad():
        for lin in [b.adaLN_modulation[1] for b in model.blocks] + [model.final_layer.adaLN_modulation[1], model.final_layer.linear]:
            lin.weight.normal_(0, 0.02)
            lin.bias.normal_(0, 0.02)
        tt = torch.rand(2, device=DEV)
        p1, z1 = model(x, tt, y=y)
        p2, _ = model(x.flip(0), tt.flip(0), y=y.flip(0))
    assert z1 is None and p1.shape == x.shape and torch.isfinite(p1).all()
    assert float((p1.flip(0) - p2).abs().max()) < 1e-5 * max(1.0, float(p1.abs().max()))


@pytest.mark.parametrize("precision,loss_tol,cos_min", [("fp32", 1e-5, 0.99999), ("bf16", 2e-2, 0.999)])
@pytest.mark.parametrize("variant", ["patch4_heads3", "patch8_qknorm"])
def test_other_patch_sizes_match_oracle(variant, precision, loss_tol, cos_min):
    """SiT-*/4 and */8 geometries (64 / 16 tokens, 64- / 256-wide patch vectors, head_dim 32 / 16) through the public API against
    the CPU oracle - which tests/test_oracle_vs_reference_live.py holds to the live reference on the same architectures."""
    import torch.nn.functional as F
    from oracle import loss_oracle, sit_oracle
    from reed_b200.image.loss import SILoss
    from test_parity_gpu import _Replay, _build          # pytest puts tests/ on sys.path (rootdir-relative "prepend" import mode)
    if variant == "patch4_heads3":
        spec = ArchSpec(input_size=32, patch_size=4, hidden_size=96, decoder_hidden_size=96, depth=3, num_heads=3,
                        encoder_depth=2, z_dims=[48], z_types=["i"], projector_dim=64, num_classes=1000)
    else:
        spec = ArchSpec(input_size=32, patch_size=8, hidden_size=64, decoder_hidden_size=64, depth=2, num_heads=4,
                        encoder_depth=1, z_dims=[32], z_types=["i"], projector_dim=48, num_classes=1000, qk_norm=True)
    sd = random_state(spec, 31)
    data = random_batch(spec, 3, 32)
    g = torch.Generator().manual_seed(33)
    t = torch.rand((3, 1, 1, 1), generator=g)
    noise = torch.randn(data["x"].shape, generator=g)
    drop = torch.tensor([False, True, False])
    model = _build(spec, sd, precision).train()
    fn = SILoss(enc_names=["dinov2"], loss_weights={"dinov2": 1.0}, time_schedule="linear")
    with _Replay(fn, model, t, noise, drop):
        out = fn(model, data["x"].to(DEV), dict(y=data["y"].to(DEV)), zs=[z.to(DEV) for z in data["zs"]])
    (out["denoising_loss"].mean() + 0.5 * out["proj_loss"]).backward()
    leaves = {k: v.clone().requires_grad_(k != "pos_embed") for k, v in sd.items()}
    ref = loss_oracle.si_loss(sit_oracle.as_model(leaves, spec, training=True, drop_mask=drop), data["x"], t, noise, data["zs"],
                              enc_names=["dinov2"], loss_weights={"dinov2": 1.0}, model_kwargs=dict(y=data["y"]),
                              time_schedule="linear")
    (ref["denoising_loss"].mean() + 0.5 * ref["proj_loss"]).backward()
    assert _rel(out["denoising_loss"], ref["denoising_loss"]) < loss_tol
    assert _rel(out["proj_loss"], ref["proj_loss"]) < loss_tol
    for name, p in model.named_parameters():
        if p.grad is None:
            continue
        if float(leaves[name].grad.norm()) < 1e-6:
            # exactly zero in exact arithmetic (k_norm.bias shifts every logit of a row by the same q.b): noise only
            assert float(p.grad.norm()) < (1e-5 if precision == "fp32" else 1e-2), name
            continue
        cos = float(F.cosine_similarity(p.grad.flatten().double().cpu(), leaves[name].grad.flatten().double(), dim=0))
        assert cos >= cos_min, (name, cos)
