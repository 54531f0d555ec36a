"""Generate tests/golden/*.pt by RUNNING THE UNMODIFIED REFERENCE.  TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
The reference's sit.py/loss.py/samplers.py are imported as they lie under /root/reference/image with
oracle/timm_shim standing in for the absent timm; update_ema/sample_posterior are exec'd from the
reference train.py source (train.py itself cannot be imported: accelerate/diffusers are absent).
Nothing in tests/, smoke() or bench.py reads /root/reference at run time - only these fixtures.
"""
from __future__ import annotations

import ast
import os
import sys
from collections import OrderedDict
from dataclasses import asdict

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/image"
OUT = os.environ.get("REED_GOLDEN_OUT") or os.path.join(os.path.dirname(HERE), "tests", "golden")


def _import_reference():
    sys.path.insert(0, os.path.join(HERE, "timm_shim"))
    sys.path.insert(0, REF)
    import models.sit as ref_sit          # noqa
    import loss as ref_loss               # noqa
    import samplers as ref_samplers       # noqa
    src = open(os.path.join(REF, "train.py")).read()
    ns = {"torch": torch, "OrderedDict": OrderedDict}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in ("update_ema", "sample_posterior"):
            exec(compile(ast.Module([node], []), "train.py", "exec"), ns)
    return ref_sit, ref_loss, ref_samplers, ns


def _ref_model(ref_sit, spec, sd):
    kw = asdict(spec)
    qk = kw.pop("qk_norm")
    m = ref_sit.SiT(path_type="linear", use_cfg=True, fused_attn=False, qk_norm=qk, **kw)
    m.load_state_dict(sd, strict=True)
    return m


def _replay_draws(seed, images, p, weighting="uniform", path_type="linear"):
    from .loss_oracle import draw_time
    torch.manual_seed(seed)
    t = draw_time(images.shape[0], weighting, path_type)
    noise = torch.randn_like(images)
    drop = torch.rand(images.shape[0]) < p
    return t, noise, drop


def loss_case(ref_sit, ref_loss, spec, *, state_seed, batch_seed, draw_seed, batch, enc_names, loss_weights,
              path_type="linear", weighting="uniform", time_schedule="constant", cutoffs=(0.0, 1.0),
              proj_coeff=0.5, with_grads=True):
    from .fixtures import random_state, random_batch, checksum
    sd = random_state(spec, state_seed)
    data = random_batch(spec, batch, batch_seed)
    model = _ref_model(ref_sit, spec, sd).train()
    t, noise, drop = _replay_draws(draw_seed, data["x"], spec.class_dropout_prob, weighting, path_type)
    fn = ref_loss.SILoss(prediction="v", path_type=path_type, weighting=weighting, enc_names=list(enc_names),
                         loss_weights=dict(loss_weights), time_schedule=time_schedule, cutoffs=list(cutoffs))
    torch.manual_seed(draw_seed)
    out = fn(model, data["x"], dict(y=data["y"]), zs=data["zs"], save_projloss=True)
    total = out["denoising_loss"].mean() + proj_coeff * out["proj_loss"]
    case = dict(
        spec=asdict(spec), state_seed=state_seed, batch_seed=batch_seed, batch=batch, enc_names=list(enc_names),
        loss_weights=dict(loss_weights), path_type=path_type, weighting=weighting, time_schedule=time_schedule,
        cutoffs=list(cutoffs), proj_coeff=proj_coeff,
        state_checksum=sum(checksum(v) for v in sd.values()), x_checksum=checksum(data["x"]),
        t=t, noise=noise, drop=drop,
        denoising_loss=out["denoising_loss"].detach().clone(), proj_loss=out["proj_loss"].detach().clone(),
        img_proj_loss=torch.as_tensor(out["img_proj_loss"]).detach().clone(),
        text_proj_loss=torch.as_tensor(out["text_proj_loss"]).detach().clone(),
        saver_image=out["loss_saver"]["image"].detach().clone(), saver_text=out["loss_saver"]["text"].detach().clone(),
        total=total.detach().clone(),
    )
    if with_grads:
        total.backward()
        case["grads"] = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    # plain forward (eval, inference) on the interpolated input for a direct model check
    model.eval()
    with torch.no_grad():
        x_t = (1 - t) * data["x"] + t * noise
        case["eval_pred"] = model(x_t, t.flatten(), y=data["y"])[0].clone()
    return case


def sampler_cases(ref_sit, ref_samplers, spec, state_seed):
    from .fixtures import random_state
    sd = random_state(spec, state_seed)
    model = _ref_model(ref_sit, spec, sd).eval()
    g = torch.Generator().manual_seed(77)
    z = torch.randn(3, spec.in_channels, spec.input_size, spec.input_size, generator=g)
    y = torch.randint(0, spec.num_classes, (3,), generator=g)
    variants = {
        "euler": dict(fn="euler_sampler", num_steps=6),
        "heun_cfg": dict(fn="euler_sampler", num_steps=5, heun=True, cfg_scale=1.5, guidance_low=0.2, guidance_high=0.8),
        "em": dict(fn="euler_maruyama_sampler", num_steps=6),
        "em_cfg_cosine": dict(fn="euler_maruyama_sampler", num_steps=5, cfg_scale=2.0, guidance_low=0.0,
                              guidance_high=0.7, path_type="cosine"),
    }
    out = dict(spec=asdict(spec), state_seed=state_seed, latents=z, y=y, variants={})
    for name, kw in variants.items():
        kw = dict(kw)
        fn = getattr(ref_samplers, kw.pop("fn"))
        torch.manual_seed(5)
        noises = [torch.randn(z.shape, dtype=torch.float64) for _ in range(kw["num_steps"] - 1)]
        torch.manual_seed(5)
        res = fn(model, z, y, **kw)
        out["variants"][name] = dict(kwargs=kw, sde="maruyama" in fn.__name__, noises=noises, result=res.clone(),
                                     dtype=str(res.dtype))
    return out


def init_case(ref_sit, name, seed=0, **kw):
    torch.manual_seed(seed)
    m = ref_sit.SiT_models[name](input_size=32, num_classes=1000, use_cfg=True, z_dims=[768], z_types=["i"],
                                 encoder_depth=8, fused_attn=True, qk_norm=False, **kw)
    stats = OrderedDict()
    for k, v in m.state_dict().items():
        flat = v.flatten().double()
        stats[k] = dict(shape=tuple(v.shape), sum=float(flat.sum()), abs_sum=float(flat.abs().sum()), head=flat[:4].tolist())
    return m, dict(name=name, seed=seed, kwargs=kw, tensors=stats,
                   n_params=sum(p.numel() for p in m.parameters()),
                   param_names=[n for n, _ in m.named_parameters()],
                   trainable=[n for n, p in m.named_parameters() if p.requires_grad])


def known_answer_s2(ref_sit, ref_loss):
    m, init = init_case(ref_sit, "SiT-S/2", seed=0, decoder_hidden_size=384)
    m.train()
    x = torch.randn(4, 4, 32, 32)
    y = torch.randint(0, 1000, (4,))
    zs = [torch.randn(4, 256, 768)]
    torch.manual_seed(123)
    out = ref_loss.SILoss(enc_names=["dinov2"], loss_weights={"dinov2": 1.0})(m, x, dict(y=y), zs=zs)
    return dict(init=init, denoising_loss=out["denoising_loss"].detach().clone(), proj_loss=out["proj_loss"].detach().clone(),
                text_proj_loss=out["text_proj_loss"])


def train_glue_case(ref_sit, ref_loss, ns, spec, state_seed, steps=3):
    """clip_grad_norm_(1.0) -> AdamW(lr 1e-4, wd 0) -> update_ema(0.9999), as train.py:253-259,402-412."""
    from copy import deepcopy
    from .fixtures import random_state, random_batch
    sd = random_state(spec, state_seed)
    model = _ref_model(ref_sit, spec, sd).train()
    ema = deepcopy(model)
    for p in ema.parameters():
        p.requires_grad_(False)
    ns["update_ema"](ema, model, decay=0)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, betas=(0.9, 0.999), weight_decay=0.0, eps=1e-8)
    fn = ref_loss.SILoss(enc_names=["dinov2"], loss_weights={"dinov2": 1.0})
    # a large lr variant would hide nothing; amplify grads instead so the clip is active on some steps
    records = []
    for s in range(steps):
        data = random_batch(spec, 3, 500 + s)
        t, noise, drop = _replay_draws(900 + s, data["x"], spec.class_dropout_prob)
        torch.manual_seed(900 + s)
        out = fn(model, data["x"], dict(y=data["y"]), zs=data["zs"])
        scale = 40.0 if s == 1 else 1.0
        loss = (out["denoising_loss"].mean() + 0.5 * out["proj_loss"]) * scale
        loss.backward()
        norm = torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
        ns["update_ema"](ema, model)
        records.append(dict(batch_seed=500 + s, t=t, noise=noise, drop=drop, loss=loss.detach().clone(), scale=scale,
                            grad_norm=norm.detach().clone()))
    g = torch.Generator().manual_seed(3)
    moments = torch.randn(2, 8, 4, 4, generator=g)
    torch.manual_seed(11)
    post_noise = torch.randn(2, 4, 4, 4)
    torch.manual_seed(11)
    post = ns["sample_posterior"](moments, latents_scale=0.18215, latents_bias=0.0)
    return dict(spec=asdict(spec), state_seed=state_seed, steps=records,
                final_model={k: v.detach().clone() for k, v in model.state_dict().items()},
                final_ema={k: v.detach().clone() for k, v in ema.state_dict().items()},
                posterior=dict(moments=moments, noise=post_noise, out=post))


def curriculum_case():
    """train.py:363-385 executed from the reference source (the three statements that open the
    ``with accelerator.accumulate(model)`` block) over a grid of steps and schedule arguments."""
    import itertools
    import types
    import numpy as np
    src = open(os.path.join(REF, "train.py")).read()
    body = None
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.With) and "accumulate" in ast.get_source_segment(src, node.items[0].context_expr):
            body = node.body[:3]
    assert body is not None and [type(s).__name__ for s in body] == ["If", "Assign", "If"]
    code = compile(ast.Module(body, []), "train.py", "exec")
    cases = []
    grids = dict(repa_weight_decay=["constant", "linear", "cosine"], diffusion_decay=["constant", "linear", "cosine"],
                 start_diffusion_steps=[0, 300], diffusion_warm_up_steps=[50, 1000])
    for combo in itertools.product(*grids.values()):
        kw = dict(zip(grids.keys(), combo), repa_steps=4000, max_train_steps=5000)
        for step in (0, 1, 49, 50, 299, 300, 349, 350, 999, 1000, 1299, 1300, 2500, 3999, 4000, 4999, 5000):
            ns = {"args": types.SimpleNamespace(**kw), "global_step": step, "np": np}
            exec(code, ns)
            cases.append(dict(kw, global_step=step, diffusion=float(ns["_diffusion_loss_decay"]),
                              repa=float(ns["_repa_weight_decay"])))
    return cases


def preprocess_case():
    """train.py:53-74 executed from the reference source on a seeded uint8 batch; a strided sample of every output plus
    whole-tensor checksums (the full 3x224x224 outputs would be megabytes)."""
    from torchvision.transforms import Normalize
    src = open(os.path.join(REF, "train.py")).read()
    ns = {"torch": torch, "Normalize": Normalize,
          "IMAGENET_DEFAULT_MEAN": (0.485, 0.456, 0.406), "IMAGENET_DEFAULT_STD": (0.229, 0.224, 0.225)}   # timm.data constants
    for node in ast.parse(src).body:
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", "") in ("CLIP_DEFAULT_MEAN", "CLIP_DEFAULT_STD"):
            exec(compile(ast.Module([node], []), "train.py", "exec"), ns)
        if isinstance(node, ast.FunctionDef) and node.name == "preprocess_raw_image":
            exec(compile(ast.Module([node], []), "train.py", "exec"), ns)
    cases = {}
    for enc_type, res, seed in (("dinov2", 256, 1), ("dinov2", 512, 2), ("clip", 256, 3), ("mocov3", 256, 4),
                                ("jepa", 256, 5), ("dinov1", 256, 6), ("mae", 256, 7), ("siglip", 256, 8)):
        x = torch.randint(0, 256, (2, 3, res, res), generator=torch.Generator().manual_seed(seed), dtype=torch.uint8)
        y = ns["preprocess_raw_image"](x, enc_type)
        cases[f"{enc_type}/{res}"] = dict(enc_type=enc_type, resolution=res, seed=seed, shape=tuple(y.shape), dtype=str(y.dtype),
                                          sample=y[..., ::7, ::5].clone(), total=float(y.double().sum()),
                                          abs_total=float(y.double().abs().sum()))
    return cases


def main():
    from .sit_oracle import ArchSpec
    if "--only-preprocess" in sys.argv:
        torch.save(preprocess_case(), os.path.join(OUT, "preprocess.pt"))
        print("preprocess.pt", os.path.getsize(os.path.join(OUT, "preprocess.pt")))
        return
    if "--only-curriculum" in sys.argv:
        _import_reference()
        torch.save(curriculum_case(), os.path.join(OUT, "curriculum.pt"))
        print("curriculum.pt", os.path.getsize(os.path.join(OUT, "curriculum.pt")))
        return
    torch.set_num_threads(8)
    ref_sit, ref_loss, ref_samplers, ns = _import_reference()
    os.makedirs(OUT, exist_ok=True)

    spec_a = ArchSpec(input_size=16, hidden_size=128, decoder_hidden_size=128, depth=2, num_heads=2, encoder_depth=1,
                      z_dims=[64], z_types=["i"], projector_dim=128, num_classes=1000)
    spec_b = ArchSpec(input_size=16, hidden_size=144, decoder_hidden_size=144, depth=2, num_heads=2, encoder_depth=1,
                      encoder_depth_text=2, z_dims=[64, 96], z_types=["i", "t"], projector_dim=128, num_classes=1000)
    spec_c = ArchSpec(input_size=8, hidden_size=64, decoder_hidden_size=64, depth=2, num_heads=2, encoder_depth=2,
                      z_dims=[32], z_types=["i"], projector_dim=64, num_classes=1000, qk_norm=True)
    spec_t = ArchSpec(input_size=8, hidden_size=64, decoder_hidden_size=64, depth=2, num_heads=1, encoder_depth=1,
                      z_dims=[32], z_types=["i"], projector_dim=64, num_classes=1000)

    torch.save(loss_case(ref_sit, ref_loss, spec_a, state_seed=1, batch_seed=2, draw_seed=3, batch=3,
                         enc_names=["dinov2"], loss_weights={"dinov2": 1.0}), os.path.join(OUT, "loss_a.pt"))
    torch.save(loss_case(ref_sit, ref_loss, spec_b, state_seed=4, batch_seed=5, draw_seed=6, batch=4,
                         enc_names=["dinov2", "text_embeds_qwenvl_7b_layer_15"],
                         loss_weights={"dinov2": 1.0, "text_embeds_qwenvl_7b_layer_15": 0.5},
                         path_type="cosine", weighting="lognormal", time_schedule="linear"), os.path.join(OUT, "loss_b.pt"))
    sched = {}
    for schedule in ("cosine", "sigmoid", "loglinear", "cutoff", "constant"):
        sched[schedule] = loss_case(ref_sit, ref_loss, spec_c, state_seed=7, batch_seed=8, draw_seed=9, batch=5,
                                    enc_names=["mocov3"], loss_weights={"mocov3": 0.0 if schedule == "constant" else 0.7},
                                    time_schedule=schedule, cutoffs=(0.25, 0.75), with_grads=False)
    torch.save(sched, os.path.join(OUT, "loss_c_schedules.pt"))
    torch.save(sampler_cases(ref_sit, ref_samplers, spec_a, state_seed=1), os.path.join(OUT, "samplers_a.pt"))
    torch.save(known_answer_s2(ref_sit, ref_loss), os.path.join(OUT, "known_answer_s2.pt"))
    _, init_b = init_case(ref_sit, "SiT-B/2", seed=0)
    torch.save(init_b, os.path.join(OUT, "init_b2.pt"))
    torch.save(train_glue_case(ref_sit, ref_loss, ns, spec_t, state_seed=21), os.path.join(OUT, "train_glue.pt"))
    torch.save(curriculum_case(), os.path.join(OUT, "curriculum.pt"))
    torch.save(preprocess_case(), os.path.join(OUT, "preprocess.pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
