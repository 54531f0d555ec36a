"""Parity at the scale the metric is quoted on: BASELINE.json's own architectures at FULL width and depth, through the
public API (SILoss.__call__ -> SiT.forward -> C-ABI kernels), against the CPU oracle on the same weights, inputs and
random draws.  Zero-initialised parameters of the reference init are replaced by a fully random state
(oracle.fixtures.random_state), otherwise every block is the identity and most gradients vanish.

  configs[1]  SiT-B/2   depth 12, D 768,  head_dim 64, T 256
  configs[2]  SiT-XL/2  depth 28, D 1152, head_dim 72, T 256
  configs[3]  SiT-XL/2  + second projector head (3584-d caption embedding, tapped at block 16), loss weights 1.0 / 0.5
  configs[4]  SiT-XL/2  4x64x64 latents = 1024 tokens

Bars (BASELINE.json north_star): fp32 losses <= 1e-5 relative, bf16 <= 2e-2 relative, per-parameter gradient cosine
>= 0.999 (fp32: >= 0.99999).  Reference lines: /root/reference/image/loss.py:153-237, image/models/sit.py:271-311,400-407.
"""
import time

import pytest
import torch
import torch.nn.functional as F

from oracle import loss_oracle, sit_oracle
from oracle.fixtures import random_batch, random_state
from oracle.sit_oracle import zoo_spec

pytestmark = pytest.mark.gpu
DEV = "cuda"

_TEXT = "text_embeds_qwenvl_7b_layer_15"
CONFIGS = {
    # name: (zoo name, spec overrides, batch, enc_names, loss_weights)
    "b2": ("SiT-B/2", dict(), 4, ["dinov2"], {"dinov2": 1.0}),
    "xl2": ("SiT-XL/2", dict(), 4, ["dinov2"], {"dinov2": 1.0}),
    "xl2_mm": ("SiT-XL/2", dict(z_dims=[768, 3584], z_types=["i", "t"], encoder_depth_text=16), 4,
               ["dinov2", _TEXT], {"dinov2": 1.0, _TEXT: 0.5}),
    "xl2_512": ("SiT-XL/2", dict(input_size=64), 2, ["dinov2"], {"dinov2": 1.0}),
}
_oracle_cache = {}


def _oracle(name):
    """Oracle losses + per-parameter gradients of one config (computed once per session: fp32 and bf16 share it)."""
    if name in _oracle_cache:
        return _oracle_cache[name]
    zoo, over, batch, enc_names, weights = CONFIGS[name]
    spec = zoo_spec(zoo, **over)
    sd = random_state(spec, 101)
    data = random_batch(spec, batch, 102)
    data["drop"] = torch.arange(batch) % 3 == 1                # at least one dropped label, deterministic
    t0 = time.time()
    # targets correlated with the projector outputs (as they become in training): random targets give alignments of
    # O(1e-3), against which a RELATIVE bar on proj_loss would only measure cancellation noise.  Noise of 3 sigma gives
    # per-token cosines near 0.3; much closer targets make the alignment gradient a small difference of bf16-rounded
    # unit vectors, and the REFERENCE ITSELF under torch.autocast(bfloat16) then drops to 0.996 cosine against its own
    # fp32 gradients on the projector biases (measured on the CPU with this oracle at 0.8 sigma; 0.9996 at 3 sigma)
    with torch.no_grad():
        tt = data["t"]
        _, z_hat = sit_oracle.as_model(sd, spec, training=True, drop_mask=data["drop"])(
            (1 - tt) * data["x"] + tt * data["noise"], tt.flatten(), y=data["y"], inference=False)
    g = torch.Generator().manual_seed(103)
    data["zs"] = [zh + 3.0 * zh.std() * torch.randn(zh.shape, generator=g) for zh in z_hat]
    leaves = {k: v.clone().requires_grad_(k != "pos_embed") for k, v in sd.items()}
    ref = loss_oracle.si_loss(sit_oracle.as_model(leaves, spec, training=True, drop_mask=data["drop"]), data["x"], data["t"],
                              data["noise"], data["zs"], enc_names=enc_names, loss_weights=weights,
                              model_kwargs=dict(y=data["y"]), time_schedule="linear")
    (ref["denoising_loss"].mean() + 0.5 * ref["proj_loss"]).backward()
    print(f"[{name}] CPU oracle forward+backward: {time.time() - t0:.1f} s")
    grads = {k: v.grad for k, v in leaves.items() if v.grad is not None}
    ref = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in ref.items()}
    _oracle_cache[name] = (spec, sd, data, ref, grads)
    return _oracle_cache[name]


def _rel(a, b):
    a, b = torch.as_tensor(a).detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("precision,loss_tol,cos_min", [("fp32", 1e-5, 0.99999), ("bf16", 2e-2, 0.999)])
@pytest.mark.parametrize("name", list(CONFIGS))
def test_baseline_config_matches_oracle_at_full_width(name, precision, loss_tol, cos_min):
    from reed_b200.image.loss import SILoss
    from test_parity_gpu import _Replay, _build
    spec, sd, data, ref, ref_grads = _oracle(name)
    _, _, batch, enc_names, weights = CONFIGS[name]
    model = _build(spec, sd, precision).train()
    assert len(model.blocks) == spec.depth and model.x_embedder.num_patches == spec.tokens
    fn = SILoss(enc_names=enc_names, loss_weights=weights, time_schedule="linear")
    with _Replay(fn, model, data["t"], data["noise"], data["drop"]):
        out = fn(model, data["x"].to(DEV), dict(y=data["y"].to(DEV)), zs=[z.to(DEV) for z in data["zs"]])
    (out["denoising_loss"].mean() + 0.5 * out["proj_loss"]).backward()
    torch.cuda.synchronize()
    errs = {k: _rel(out[k], ref[k]) for k in ("denoising_loss", "proj_loss", "img_proj_loss")}
    if len(enc_names) > 1:
        errs["text_proj_loss"] = _rel(out["text_proj_loss"], ref["text_proj_loss"])
    else:
        assert out["text_proj_loss"] == 0.0
    got = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
    assert set(got) == set(ref_grads)
    worst, worst_name = 1.0, None
    for k, g in ref_grads.items():
        cos = float(F.cosine_similarity(got[k].flatten().double().cpu(), g.flatten().double(), dim=0))
        if cos < worst:
            worst, worst_name = cos, k
    print(f"[{name} {precision}] loss rel.err {errs}; worst per-parameter gradient cosine {worst:.6f} ({worst_name}) "
          f"over {len(ref_grads)} parameters")
    for k, e in errs.items():
        assert e < loss_tol, (k, e)
    assert worst >= cos_min, (worst_name, worst)
    if precision == "fp32":
        for k, g in ref_grads.items():
            assert _rel(got[k], g) < 1e-3, k
