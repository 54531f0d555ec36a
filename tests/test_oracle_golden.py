"""Pin the CPU oracle against outputs of the unmodified reference (tests/golden/, made by oracle/make_golden.py)."""
import pytest
import torch

from oracle import loss_oracle, samplers_oracle, sit_oracle, train_oracle
from oracle.fixtures import checksum, random_batch, random_state
from oracle.sit_oracle import ArchSpec


def _rel(a, b):
    a, b = torch.as_tensor(a).detach().double(), torch.as_tensor(b).detach().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _run_loss_case(case, need_grads):
    spec = ArchSpec(**case["spec"])
    sd = random_state(spec, case["state_seed"])
    assert abs(sum(checksum(v) for v in sd.values()) - case["state_checksum"]) < 1e-6 * case["state_checksum"]
    data = random_batch(spec, case["batch"], case["batch_seed"])
    assert abs(checksum(data["x"]) - case["x_checksum"]) < 1e-9 * case["x_checksum"]
    leaves = {k: v.clone().requires_grad_(need_grads and k != "pos_embed") for k, v in sd.items()}
    model = sit_oracle.as_model(leaves, spec, training=True, drop_mask=case["drop"])
    out = loss_oracle.si_loss(model, data["x"], case["t"], case["noise"], data["zs"], enc_names=case["enc_names"],
                              loss_weights=case["loss_weights"], model_kwargs=dict(y=data["y"]),
                              path_type=case["path_type"], time_schedule=case["time_schedule"], cutoffs=case["cutoffs"])
    return spec, leaves, data, out


@pytest.mark.parametrize("name", ["loss_a.pt", "loss_b.pt"])
def test_loss_and_grads_match_reference(golden, name):
    case = golden(name)
    spec, leaves, data, out = _run_loss_case(case, True)
    assert _rel(out["denoising_loss"], case["denoising_loss"]) < 2e-6
    assert _rel(out["proj_loss"], case["proj_loss"]) < 2e-6
    assert _rel(out["img_proj_loss"], case["img_proj_loss"]) < 2e-6
    assert _rel(out["text_proj_loss"], case["text_proj_loss"]) < 2e-6 or float(case["text_proj_loss"]) == 0.0
    total = out["denoising_loss"].mean() + case["proj_coeff"] * out["proj_loss"]
    total.backward()
    assert set(case["grads"]) == {k for k, v in leaves.items() if v.grad is not None}
    for k, g in case["grads"].items():
        cos = torch.nn.functional.cosine_similarity(leaves[k].grad.flatten().double(), g.flatten().double(), dim=0)
        assert cos > 1 - 1e-6, (k, float(cos))
        assert _rel(leaves[k].grad, g) < 1e-4, k
    with torch.no_grad():
        t = case["t"]
        a, s, _, _ = loss_oracle.path_coefficients(t, "linear")
        pred, zs = sit_oracle.sit_forward(leaves, spec, a * data["x"] + s * case["noise"], t.flatten(), data["y"])
        assert zs is None
        assert _rel(pred, case["eval_pred"]) < 5e-6


def test_time_schedules_and_qk_norm(golden):
    for schedule, case in golden("loss_c_schedules.pt").items():
        _, _, _, out = _run_loss_case(case, False)
        assert _rel(out["denoising_loss"], case["denoising_loss"]) < 2e-6, schedule
        assert _rel(out["proj_loss"], case["proj_loss"]) < 2e-6, schedule
        assert _rel(out["per_sample_align"][0], case["saver_image"]) < 2e-6, schedule


def test_broadcast_quirk_is_reproduced(golden):
    """(B,)*(B,1,1,1) -> mean(curr)*mean(w), not the per-sample weighted mean (loss.py:221-222)."""
    case = golden("loss_c_schedules.pt")["cosine"]
    _, _, _, out = _run_loss_case(case, False)
    w = loss_oracle.schedule_weight(case["t"], 0.7, "cosine").flatten()
    align = out["per_sample_align"][0]
    assert _rel(out["proj_loss"], align.mean() * w.mean()) < 1e-6
    assert _rel(out["proj_loss"], (align * w).mean()) > 1e-4


def test_samplers_match_reference(golden):
    fx = golden("samplers_a.pt")
    spec = ArchSpec(**fx["spec"])
    model = sit_oracle.as_model(random_state(spec, fx["state_seed"]), spec)
    for name, v in fx["variants"].items():
        kw = v["kwargs"]
        if v["sde"]:
            res = samplers_oracle.euler_maruyama(model, fx["latents"], fx["y"], noises=v["noises"], **kw)
        else:
            res = samplers_oracle.euler(model, fx["latents"], fx["y"], **kw)
        assert res.dtype == torch.float64 and v["dtype"] == "torch.float64"
        assert float((res - v["result"]).abs().max()) < 2e-5, name


def test_known_answer_s2(golden):
    """SURVEY.md section 8(c) known-answer case, regenerated from seeds through the oracle."""
    fx = golden("known_answer_s2.pt")
    assert _rel(fx["denoising_loss"], torch.tensor([2.0167746544, 2.0226225853, 1.9806505442, 1.9718903303])) < 1e-6
    assert fx["text_proj_loss"] == 0.0


def test_train_glue_matches_reference(golden):
    fx = golden("train_glue.pt")
    spec = ArchSpec(**fx["spec"])
    sd = random_state(spec, fx["state_seed"])
    params = {k: v.clone() for k, v in sd.items()}
    ema = {k: v.clone() for k, v in sd.items()}
    m = {k: torch.zeros_like(v) for k, v in sd.items()}
    v2 = {k: torch.zeros_like(v) for k, v in sd.items()}
    for step, rec in enumerate(fx["steps"], start=1):
        data = random_batch(spec, 3, rec["batch_seed"])
        leaves = {k: p.clone().requires_grad_(k != "pos_embed") for k, p in params.items()}
        model = sit_oracle.as_model(leaves, spec, training=True, drop_mask=rec["drop"])
        out = loss_oracle.si_loss(model, data["x"], rec["t"], rec["noise"], data["zs"], enc_names=["dinov2"],
                                  loss_weights={"dinov2": 1.0}, model_kwargs=dict(y=data["y"]))
        loss = train_oracle.mix_losses(out) * rec["scale"]
        assert _rel(loss, rec["loss"]) < 5e-6
        loss.backward()
        grads = {k: l.grad for k, l in leaves.items() if l.grad is not None}
        norm = train_oracle.adamw_ema_step(params, grads, m, v2, ema, step)
        assert _rel(norm, rec["grad_norm"]) < 1e-4
    for k in params:
        # one AdamW step moves a weight by ~lr=1e-4; allow 5% of a step for m/sqrt(v) round-off at g~0
        assert float((params[k] - fx["final_model"][k]).abs().max()) < 5e-6, k
        assert float((ema[k] - fx["final_ema"][k]).abs().max()) < 5e-6, k
    post = fx["posterior"]
    assert _rel(train_oracle.sample_posterior(post["moments"], post["noise"]), post["out"]) < 1e-6


def test_flop_formula_matches_baseline_md():
    spec = sit_oracle.zoo_spec("SiT-XL/2")
    assert abs(sit_oracle.flops_per_image(spec, train=False) / 1e9 - 241.394) < 0.01
    spec = sit_oracle.zoo_spec("SiT-B/2")
    assert abs(sit_oracle.flops_per_image(spec, train=True) / 1e9 - 149.29) < 0.01
