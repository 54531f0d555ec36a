"""Frozen DINOv2 target-encoder forward on the B200 against the CPU oracle restatement
(oracle/dinov2_oracle.py; parity unpinned - the hub model is third-party code absent from the reference tree), with random
weights: fp32 mode to 1e-4, bf16 mode to the 2e-2 bar on the tokens the REED loss consumes (x_norm_patchtokens)."""
import pytest
import torch

from oracle import dinov2_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _randomise(model, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("gamma"):
                p.copy_(0.5 + 0.5 * torch.rand(p.shape, generator=g))         # LayerScale far from its 1e-5 init
            elif name.endswith("norm1.weight") or name.endswith("norm2.weight") or name == "norm.weight":
                p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g))
            elif p.dim() == 1:
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
            elif name in ("cls_token", "pos_embed", "register_tokens", "mask_token"):
                p.copy_(0.2 * torch.randn(p.shape, generator=g))
            else:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) * fan_in ** -0.5)
    return model


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("bf16", 2e-2)])
@pytest.mark.parametrize("registers", [0, 4])
def test_small_vit_matches_oracle(precision, tol, registers):
    from reed_b200.image.encoders import DinoV2
    m = _randomise(DinoV2(128, 3, 2, img_size=56, num_register_tokens=registers, precision=precision), 1)
    x = torch.randn(3, 3, 56, 56, generator=torch.Generator().manual_seed(2))
    ref = dinov2_oracle.forward_features(m.state_dict(), x, 2)
    out = m.to(DEV).eval().forward_features(x.to(DEV))
    assert out["x_norm_patchtokens"].shape == (3, 16, 128) and out["x_norm_regtokens"].shape == (3, registers, 128)
    for k in ("x_norm_patchtokens", "x_norm_clstoken", "x_prenorm"):
        assert _rel(out[k].float(), ref[k]) < tol, k


def test_vit_b14_full_width_bf16_and_preprocessed_input():
    """dinov2-vit-b at the shapes of the train step: 256-pixel uint8 images -> preprocess_raw_image -> 224 pixels ->
    257 tokens x 768, 12 blocks (train.py:348-360)."""
    from oracle import preprocess_oracle
    from reed_b200.image.encoders import build_dinov2
    from reed_b200.image.preprocess import preprocess_raw_image
    m = _randomise(build_dinov2("b", resolution=256), 3)
    raw = torch.randint(0, 256, (2, 3, 256, 256), dtype=torch.uint8, generator=torch.Generator().manual_seed(4))
    x_ref = preprocess_oracle.preprocess_raw_image(raw.float(), "dinov2-vit-b")
    ref = dinov2_oracle.forward_features(m.state_dict(), x_ref, 12)
    m = m.to(DEV).eval()
    x = preprocess_raw_image(raw.to(DEV), "dinov2-vit-b")
    assert x.shape == (2, 3, 224, 224)
    z = m.forward_features(x)["x_norm_patchtokens"]
    assert z.shape == (2, 256, 768) and z.dtype == torch.bfloat16
    assert _rel(z.float(), ref["x_norm_patchtokens"]) < 3e-2
    cos = torch.nn.functional.cosine_similarity(z.float().cpu().flatten(1), ref["x_norm_patchtokens"].flatten(1), dim=1)
    assert float(cos.min()) > 0.9995
