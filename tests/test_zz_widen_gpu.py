"""Rows of SURVEY 8(f) on the B200: the posterior-draw kernel, the pinned batch loader, checkpoint resume and the
CUDA-graphed sampler evaluation - each against the reference arithmetic (torch expressions of train.py:84-91, the CPU
oracle, golden outputs of the reference) or against the eager path of the same kernels."""
import io
import json
import os

import numpy as np
import pytest
import torch

from oracle import samplers_oracle, sit_oracle, train_oracle
from oracle.fixtures import random_batch, random_state
from oracle.sit_oracle import ArchSpec

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _reference_posterior(moments, latents_scale, latents_bias):
    """train.py:84-91, verbatim semantics, in PyTorch on the same device."""
    mean, std = torch.chunk(moments, 2, dim=1)
    z = mean + std * torch.randn_like(mean)
    return z * latents_scale + latents_bias


@pytest.mark.parametrize("shape", [(32, 8, 32, 32), (3, 8, 64, 64), (5, 8, 3, 3), (2, 6, 5, 1), (0, 8, 4, 4)])
def test_sample_posterior_is_bit_identical_to_the_reference_sequence(shape):
    from reed_b200.image.dataset import sample_posterior
    C = shape[1] // 2
    g = torch.Generator(device=DEV).manual_seed(1)
    moments = torch.randn(shape, device=DEV, generator=g)
    scale = torch.tensor([0.18215, 0.2, 0.5, 1.25][:C] + [0.7] * max(0, C - 4)).view(1, C, 1, 1).to(DEV)
    bias = -torch.tensor([0.0, 0.1, -0.3, 2.0][:C] + [0.01] * max(0, C - 4)).view(1, C, 1, 1).to(DEV)
    for s, b in ((scale, bias), (0.18215, 0.0), (1.0, 0.0), (torch.tensor(0.5), 0.25)):
        torch.manual_seed(77)
        want = _reference_posterior(moments, s, b)
        nxt_want = torch.rand(1, device=DEV)
        torch.manual_seed(77)
        got = sample_posterior(moments, latents_scale=s, latents_bias=b)
        nxt_got = torch.rand(1, device=DEV)
        assert got.shape == want.shape and got.dtype == torch.float32
        assert torch.equal(got, want)
        assert torch.equal(nxt_got, nxt_want)              # the device generator advanced by the same amount


def test_sample_posterior_matches_reference_golden(golden):
    from reed_b200.image.dataset import sample_posterior
    post = golden("train_glue.pt")["posterior"]             # sample_posterior of the reference's train.py, run on CPU
    got = sample_posterior(post["moments"].to(DEV), latents_scale=0.18215, latents_bias=0.0, noise=post["noise"].to(DEV))
    assert torch.equal(got.cpu(), post["out"])
    assert torch.equal(got.cpu(), train_oracle.sample_posterior(post["moments"], post["noise"]))
    with pytest.raises(ValueError):
        sample_posterior(post["moments"].to(DEV), noise=post["noise"][:1].to(DEV))
    with pytest.raises(ValueError):
        sample_posterior(torch.zeros(2, 7, 4, 4, device=DEV))


def _make_tree(root, n, size):
    rng = np.random.default_rng(0)
    os.makedirs(os.path.join(root, "images"))
    os.makedirs(os.path.join(root, "vae-sd"))
    labels = []
    for i in range(n):
        np.save(os.path.join(root, "images", f"img{i:06d}.npy"), np.zeros((3, 2, 2), dtype=np.uint8))
        np.save(os.path.join(root, "vae-sd", f"img-mean-std-{i:06d}.npy"),
                rng.standard_normal((1, 8, size, size)).astype(np.float32))
        labels.append([f"img-mean-std-{i:06d}.npy", int(rng.integers(0, 1000))])
    json.dump({"labels": labels}, open(os.path.join(root, "vae-sd", "dataset.json"), "w"))


def test_batch_loader_stages_through_pinned_memory(tmp_path):
    from reed_b200.image.dataset import CustomDataset, LatentBatchLoader, sample_posterior
    _make_tree(str(tmp_path), 22, 16)
    ds = CustomDataset(str(tmp_path), load_images=False)
    loader = LatentBatchLoader(ds, 4, DEV, shuffle=True, generator=torch.Generator().manual_seed(2), depth=2)
    order = LatentBatchLoader(ds, 4, DEV, shuffle=True, generator=torch.Generator().manual_seed(2)).epoch_indices()
    assert len(loader) == 5 and all(b.is_pinned() for s in loader._slots for b in s.host)
    sink = torch.zeros(4, 4, 16, 16, device=DEV)
    seen = 0
    for k, (moments, labels, text) in enumerate(loader):
        assert moments.is_cuda and moments.shape == (4, 8, 16, 16) and text is None
        want = torch.stack([ds[i][1][0] for i in order[4 * k:4 * k + 4]])
        assert torch.equal(moments.cpu(), want)
        assert labels.tolist() == [int(ds.labels[i]) for i in order[4 * k:4 * k + 4]]
        for _ in range(20):                                  # keep the consumer stream busy while the next batch stages
            sink += sample_posterior(moments, 0.18215, 0.0)
        seen += 1
    assert seen == 5 and len(loader._slots) == 2 and all(b.is_pinned() for s in loader._slots for b in s.host)
    assert torch.isfinite(sink).all()


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
def test_checkpoint_resume_continues_identically(precision):
    """steps 1-2, save, load into a trainer with other weights, step 3 on both: same loss, same weights / moments / EMA."""
    spec, a = _tiny(precision, seed=11)
    to_dev = lambda d: (d["x"].to(DEV), d["y"].to(DEV), [z.to(DEV) for z in d["zs"]])
    batches = [random_batch(spec, 4, 60 + i) for i in range(3)]
    torch.manual_seed(5)
    for i in range(2):
        a.train_step(*to_dev(batches[i]))
    buf = io.BytesIO()
    torch.save(a.checkpoint(args={"note": "resume test"}), buf)
    _, b = _tiny(precision, seed=12)
    ck = torch.load(io.BytesIO(buf.getvalue()), map_location=DEV, weights_only=False)
    assert b.load_checkpoint(ck) == 2 and b.step_count == 2
    # the saved optimizer state is a genuine torch.optim.AdamW state dict
    opt = torch.optim.AdamW(b.model.parameters(), lr=1e-4)
    opt.load_state_dict(ck["opt"])
    losses = []
    for tr in (a, b):
        torch.manual_seed(99)
        loss, _ = tr.train_step(*to_dev(batches[2]), diffusion_decay=0.7, repa_decay=0.9)
        losses.append(float(loss))
        assert tr.step_count == 3
    tol = 1e-6 if precision == "fp32" else 1e-5             # same kernels on the same data: only atomics ordering differs
    assert abs(losses[0] - losses[1]) <= tol * max(1.0, abs(losses[0])), losses
    ptol = 2e-5 if precision == "fp32" else 2.5e-4
    for ba, bb in zip(a.state.buckets, b.state.buckets):
        for name in ("param", "ema", "exp_avg"):
            assert float((getattr(ba, name) - getattr(bb, name)).abs().max()) <= ptol, (ba.name, name)
        assert torch.equal(bb.shadow, bb.param.bfloat16())


def _sampler_model(precision):
    from reed_b200.image.models.sit import SiT
    spec = ArchSpec(input_size=16, hidden_size=128, decoder_hidden_size=128, depth=2, num_heads=2, encoder_depth=1,
                    z_dims=[64], projector_dim=128)
    sd = random_state(spec, 1)
    m = SiT(path_type="linear", use_cfg=True, input_size=16, hidden_size=128, decoder_hidden_size=128, depth=2, num_heads=2,
            encoder_depth=1, z_dims=[64], z_types=["i"], projector_dim=128, num_classes=1000, fused_attn=True, qk_norm=False)
    m.load_state_dict(sd)
    m.reed_precision = precision
    return spec, sd, m.to(DEV).eval()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_graphed_evaluation_reproduces_the_eager_samplers(precision):
    from reed_b200.image.generate import GraphedSiT
    from reed_b200.image.samplers import euler_maruyama_sampler, euler_sampler
    spec, sd, model = _sampler_model(precision)
    g = torch.Generator(device=DEV).manual_seed(3)
    z = torch.randn(3, 4, 16, 16, device=DEV, generator=g)
    y = torch.tensor([1, 17, 999], device=DEV)
    graphed = GraphedSiT(model)
    cases = [(euler_sampler, dict(num_steps=5)),
             (euler_sampler, dict(num_steps=4, heun=True, cfg_scale=2.5, guidance_low=0.2, guidance_high=0.8)),
             (euler_maruyama_sampler, dict(num_steps=6, cfg_scale=1.8, guidance_high=0.7))]
    for fn, kw in cases:
        torch.manual_seed(21)
        want = fn(model, z, y, **kw)
        torch.manual_seed(21)
        got = fn(graphed, z, y, **kw)
        assert got.dtype == torch.float64 and torch.isfinite(got).all()
        assert float((got - want).abs().max()) <= 1e-6 * max(1.0, float(want.abs().max())), (fn.__name__, kw)
    assert len(graphed._graphs) == 2 and graphed.replays > 10       # batch n and the guided batch 2n, captured once each
    if precision == "fp32":                                          # and the graphed path meets the oracle bar directly
        torch.manual_seed(9)
        res = euler_maruyama_sampler(graphed, z, y, num_steps=4)
        torch.manual_seed(9)
        noises = [torch.randn_like(z.double()).cpu() for _ in range(3)]
        ref = samplers_oracle.euler_maruyama(sit_oracle.as_model(sd, spec), z.cpu(), y.cpu(), num_steps=4, noises=noises)
        assert float((res.cpu() - ref).abs().max()) < 1e-3


def test_sample_latents_driver(tmp_path):
    from reed_b200.image.generate import sample_latents
    _, _, model = _sampler_model("fp32")
    kw = dict(num_fid_samples=5, per_proc_batch_size=2, latent_size=16, num_steps=3, cfg_scale=1.5, global_seed=4,
              world_size=2)
    lat0, lab0, idx0 = sample_latents(model, rank=0, out_dir=str(tmp_path), **kw)
    lat1, lab1, idx1 = sample_latents(model, rank=1, graphed=False, **kw)
    assert lat0.shape == (4, 4, 16, 16) and lat0.dtype == torch.float32 and lab0.shape == (4,)
    assert sorted(idx0.tolist() + idx1.tolist()) == list(range(8))           # 5 samples rounded up to 2 x (2 x 2)
    assert not torch.equal(lat0, lat1)                                        # per-rank seeds
    again, _, _ = sample_latents(model, rank=0, graphed=False, **kw)          # same seed: graph replay == eager launches
    assert float((again - lat0).abs().max()) <= 1e-6 * max(1.0, float(lat0.abs().max()))
    files = sorted(os.listdir(tmp_path))
    assert files == ["latents-rank0-00000.npz", "latents-rank0-00001.npz"]
    first = np.load(os.path.join(tmp_path, files[0]))
    assert np.array_equal(first["latents"], lat0[:2].numpy()) and first["indices"].tolist() == [0, 2]


def _tiny_trainer(precision="bf16", seed=11):
    from reed_b200.image.loss import SILoss
    from reed_b200.image.models.sit import SiT
    from reed_b200.image.trainer import ReedTrainer
    spec = ArchSpec(input_size=16, hidden_size=128, decoder_hidden_size=128, depth=2, num_heads=2, encoder_depth=1,
                    z_dims=[64], projector_dim=128)
    m = SiT(path_type="linear", use_cfg=True, input_size=16, hidden_size=128, decoder_hidden_size=128, depth=2, num_heads=2,
            encoder_depth=1, z_dims=[64], z_types=["i"], projector_dim=128, num_classes=1000, fused_attn=True, qk_norm=False)
    m.load_state_dict(random_state(spec, seed))
    m = m.to(DEV).train()
    return spec, ReedTrainer(m, SILoss(enc_names=["dinov2"], loss_weights={"dinov2": 1.0}), precision=precision,
                             ema_decay=0.5)


@pytest.mark.parametrize("graphed_eval", [False, True])
def test_ema_inference_in_bf16_follows_the_training_steps(graphed_eval):
    """The fused optimizer updates the EMA through raw pointers; bf16 evaluations of the EMA model must see the weights of
    the LATEST step, not the bf16 copies made at its first evaluation (also through GraphedSiT's baked-in pointers)."""
    from reed_b200 import ops
    from reed_b200.image.generate import GraphedSiT
    spec, tr = _tiny_trainer("bf16")
    to_dev = lambda d: (d["x"].to(DEV), d["y"].to(DEV), [z.to(DEV) for z in d["zs"]])
    x = torch.randn(4, 4, 16, 16, device=DEV)
    t = torch.rand(4, device=DEV)
    y = torch.randint(0, 1000, (4,), device=DEV)
    runner = GraphedSiT(tr.ema) if graphed_eval else tr.ema
    with torch.no_grad():
        first = runner(x, t, y=y)[0].clone()
    for i in range(3):
        tr.train_step(*to_dev(random_batch(spec, 4, 50 + i)))
    with torch.no_grad():
        later = runner(x, t, y=y)[0].clone()
        # a fresh bf16 recast of the current fp32 EMA: what the evaluation has to equal
        fresh = type(tr.ema)(**{k: getattr(spec, k) for k in ("input_size", "hidden_size", "decoder_hidden_size", "depth",
                                                              "num_heads", "encoder_depth", "projector_dim")},
                             path_type="linear", use_cfg=True, z_dims=[64], z_types=["i"], num_classes=1000,
                             fused_attn=True, qk_norm=False).to(DEV).eval()
        fresh.load_state_dict(tr.ema.state_dict())
        fresh.reed_precision = "bf16"
        want = fresh(x, t, y=y)[0]
    assert float((later - first).abs().max()) > 1e-3           # ema_decay 0.5: three steps move the EMA visibly
    assert torch.equal(later, want)
    assert ops.weights_epoch >= 3


def test_graphed_step_staging_survives_host_run_ahead():
    """train_step_graphed called 8 times with no host/device sync in between: every replay must consume ITS OWN time draws
    and curriculum scalars (ring of pinned rows fenced by events), whatever the host writes meanwhile."""
    spec, tr = _tiny_trainer("bf16")
    to_dev = lambda d: (d["x"].to(DEV), d["y"].to(DEV), [z.to(DEV) for z in d["zs"]])
    batch = to_dev(random_batch(spec, 4, 60))
    tr.capture(*batch, warmup=2)
    g = tr._g
    # slow the device down so the host really is several replays ahead
    ballast = torch.randn(4096, 4096, device=DEV)
    torch.manual_seed(321)
    seen_t, seen_s = [], []
    for i in range(8):
        for _ in range(4):
            ballast = ballast @ ballast * 1e-4
        tr.train_step_graphed(*batch, diffusion_decay=0.1 * (i + 1), repa_decay=1.0 - 0.05 * i)
        seen_t.append(g["time"].clone())                        # stream-ordered: after replay i, before step i+1's copies
        seen_s.append(g["scalars"].clone())
    torch.cuda.synchronize()
    torch.manual_seed(321)
    for i in range(8):
        want_t = torch.rand((4, 1, 1, 1))
        assert torch.equal(seen_t[i].cpu(), want_t), i
        assert torch.allclose(seen_s[i].cpu(), torch.tensor([0.1 * (i + 1), 1.0 - 0.05 * i])), i
    assert len({id(r) for r in g["host"]}) == tr._STAGING_ROWS >= 2


def test_cutoff_schedule_is_graph_capturable():
    from reed_b200.image.loss import SILoss
    from reed_b200.image.trainer import ReedTrainer
    spec, tr = _tiny_trainer("bf16")
    tr.loss_fn = SILoss(enc_names=["dinov2"], loss_weights={"dinov2": 1.0}, time_schedule="cutoff", cutoffs=[0.2, 0.8])
    to_dev = lambda d: (d["x"].to(DEV), d["y"].to(DEV), [z.to(DEV) for z in d["zs"]])
    batch = to_dev(random_batch(spec, 4, 61))
    tr.capture(*batch, warmup=1)
    loss, _ = tr.train_step_graphed(*batch)
    assert torch.isfinite(loss)
