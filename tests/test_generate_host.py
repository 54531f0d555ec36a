"""Host logic of the sampling driver (generate.py:48-51,77-85,105-118,164): CPU only."""
import pytest
import torch


def test_sampling_plan_and_indices():
    from reed_b200.image.generate import rank_seed, sample_indices, sampling_plan
    assert sampling_plan(50_000, 32, 8) == (50_176, 196)        # rounded up to a multiple of the global batch 256
    assert sampling_plan(64, 32, 2) == (64, 1)
    assert rank_seed(3, 5, 8) == 29
    # across ranks and iterations the indices tile 0..total-1 exactly once
    world, n = 4, 3
    total, iters = sampling_plan(20, n, world)
    seen, so_far = [], 0
    for _ in range(iters):
        for rank in range(world):
            seen += sample_indices(n, rank, world, so_far)
        so_far += n * world
    assert sorted(seen) == list(range(total))


def test_load_sampling_weights_drops_projectors():
    from reed_b200.image.generate import load_sampling_weights
    from reed_b200.image.models.sit import SiT
    kw = dict(input_size=8, hidden_size=32, decoder_hidden_size=32, depth=2, num_heads=2, encoder_depth=1,
              projector_dim=32, num_classes=10, qk_norm=False)
    torch.manual_seed(0)
    trained = SiT(z_dims=[16], **kw)
    torch.manual_seed(1)
    sampler_model = SiT(z_dims=[24], **kw)                      # generate.py builds the heads from its own flags
    before = sampler_model.projectors[0][0].weight.clone()
    load_sampling_weights(sampler_model, trained.state_dict())
    assert torch.equal(sampler_model.blocks[1].mlp.fc1.weight, trained.blocks[1].mlp.fc1.weight)
    assert torch.equal(sampler_model.projectors[0][0].weight, before)
    bad = {k: v for k, v in trained.state_dict().items() if "blocks.1.mlp" not in k}
    with pytest.raises(KeyError):
        load_sampling_weights(sampler_model, bad)


def test_graphed_model_refuses_train_mode_and_cpu():
    from reed_b200.image.generate import GraphedSiT
    from reed_b200.image.models.sit import SiT
    m = SiT(input_size=8, hidden_size=32, decoder_hidden_size=32, depth=1, num_heads=2, encoder_depth=1, z_dims=[16],
            projector_dim=32, num_classes=10, qk_norm=False)
    with pytest.raises(ValueError):
        GraphedSiT(m.train())
    g = GraphedSiT(m.eval())
    assert g.in_channels == 4 and g.projectors is m.projectors
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        g(torch.zeros(1, 4, 8, 8), torch.zeros(1), y=torch.zeros(1, dtype=torch.long))


def test_legacy_checkpoint_keys():
    from reed_b200.image.generate import load_legacy_checkpoints, load_sampling_weights
    from reed_b200.image.models.sit import SiT
    kw = dict(input_size=8, hidden_size=32, decoder_hidden_size=32, depth=3, num_heads=2, encoder_depth=1, z_dims=[16],
              projector_dim=32, num_classes=10, qk_norm=False)
    torch.manual_seed(0)
    trained = SiT(**kw)
    legacy = {}
    for k, v in trained.state_dict().items():            # write blocks 1.. the way early checkpoints named them
        parts = k.split(".")
        if parts[0] == "blocks" and int(parts[1]) >= 1:
            k = ".".join(["decoder_blocks", str(int(parts[1]) - 1)] + parts[2:])
        legacy[k] = v
    assert any(k.startswith("decoder_blocks.1.") for k in legacy)
    assert list(load_legacy_checkpoints(legacy, 1)) == list(trained.state_dict())
    torch.manual_seed(1)
    fresh = SiT(**kw)
    load_sampling_weights(fresh, legacy, legacy=True)
    assert torch.equal(fresh.blocks[2].attn.qkv.weight, trained.blocks[2].attn.qkv.weight)
    with pytest.raises(KeyError):
        load_sampling_weights(SiT(**kw), legacy)          # without the renaming the backbone keys do not match
