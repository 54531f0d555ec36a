"""CPU-side checks: drop-in API surface, init parity with the reference, C-ABI exports (no GPU compute)."""
import ctypes
import os
import re

import pytest
import torch

from oracle.fixtures import parameter_shapes
from oracle.sit_oracle import ArchSpec, zoo_spec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as entry
    entry.build()
    from reed_b200 import _cabi
    return _cabi


def test_library_exports_every_declared_symbol(built):
    header = open(os.path.join(ROOT, "include", "reed_b200.h")).read()
    declared = set(re.findall(r"\b(reed_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    lib = ctypes.CDLL(built.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/reed_b200.h but not exported"
    assert declared == set(built.EXPORTS), declared ^ set(built.EXPORTS)
    assert built.version() >= 100


def test_missing_library_fails_loudly(monkeypatch, built):
    monkeypatch.setattr(built, "_lib", None)
    monkeypatch.setattr(built, "LIB_PATH", "/nonexistent/libreed_sm100.so")
    with pytest.raises(built.ReedLibraryError):
        built.load()


def _model(name="SiT-S/2", **kw):
    from reed_b200.image.models.sit import SiT_models
    args = dict(input_size=32, num_classes=1000, use_cfg=True, z_dims=[768], z_types=["i"], encoder_depth=8,
                fused_attn=True, qk_norm=False)
    args.update(kw)
    return SiT_models[name](**args)


def test_zoo_and_constructor_quirks():
    from reed_b200.image.models.sit import SiT_models
    assert sorted(SiT_models) == sorted(f"SiT-{f}/{p}" for f in ("XL", "L", "B", "S") for p in (2, 4, 8))
    with pytest.raises(KeyError):                     # qk_norm is a required block kwarg (sit.py:115)
        SiT_models["SiT-S/2"](input_size=32)
    m = _model("SiT-S/2")                              # decoder_hidden_size stays 768 for S (sit.py:172,400-407)
    assert m.final_layer.linear.weight.shape == (16, 768)
    m = _model("SiT-S/2", decoder_hidden_size=384, some_unknown_flag=3)   # extra kwargs are swallowed
    assert m.final_layer.linear.weight.shape == (16, 384)
    for attr in ("path_type", "in_channels", "out_channels", "patch_size", "num_heads", "use_cfg", "num_classes",
                 "z_dims", "z_types", "encoder_depth", "encoder_depth_text", "projectors"):
        assert hasattr(m, attr)


def test_state_dict_layout_matches_reference():
    spec = zoo_spec("SiT-S/2", decoder_hidden_size=384)
    m = _model("SiT-S/2", decoder_hidden_size=384)
    want = parameter_shapes(spec)
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert list(got) == list(want)
    assert got == dict(want)
    assert not m.pos_embed.requires_grad and "pos_embed" in dict(m.named_parameters())
    spec2 = ArchSpec(input_size=16, hidden_size=64, decoder_hidden_size=64, depth=2, num_heads=2, z_dims=[32, 48],
                     z_types=["i", "t"], encoder_depth=1, encoder_depth_text=2, projector_dim=64, qk_norm=True)
    from reed_b200.image.models.sit import SiT
    m2 = SiT(input_size=16, hidden_size=64, decoder_hidden_size=64, depth=2, num_heads=2, z_dims=[32, 48],
             z_types=["i", "t"], encoder_depth=1, encoder_depth_text=2, projector_dim=64, qk_norm=True)
    assert {k: tuple(v.shape) for k, v in m2.state_dict().items()} == dict(parameter_shapes(spec2))


@pytest.mark.parametrize("fixture,name,kw", [("known_answer_s2.pt", "SiT-S/2", dict(decoder_hidden_size=384)),
                                             ("init_b2.pt", "SiT-B/2", {})])
def test_init_reproduces_reference_rng_stream(golden, fixture, name, kw):
    fx = golden(fixture)
    init = fx["init"] if "init" in fx else fx
    torch.manual_seed(init["seed"])
    m = _model(name, **kw)
    assert [n for n, _ in m.named_parameters()] == init["param_names"]
    assert [n for n, p in m.named_parameters() if p.requires_grad] == init["trainable"]
    assert sum(p.numel() for p in m.parameters()) == init["n_params"]
    for k, v in m.state_dict().items():
        ref = init["tensors"][k]
        flat = v.flatten().double()
        assert tuple(v.shape) == ref["shape"]
        assert abs(float(flat.sum()) - ref["sum"]) <= 1e-9 * max(1.0, ref["abs_sum"]), k
        assert abs(float(flat.abs().sum()) - ref["abs_sum"]) <= 1e-9 * max(1.0, ref["abs_sum"]), k
        assert flat[:4].tolist() == ref["head"], k


def test_cpu_tensors_are_rejected_not_silently_served():
    from reed_b200.image.loss import SILoss
    from reed_b200.image.samplers import euler_maruyama_sampler, euler_sampler
    m = _model("SiT-S/2", decoder_hidden_size=384)
    x = torch.randn(2, 4, 32, 32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(x, torch.rand(2), torch.zeros(2, dtype=torch.long))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        SILoss(enc_names=["dinov2"], loss_weights={"dinov2": 1.0})(m, x, dict(y=torch.zeros(2, dtype=torch.long)),
                                                                    zs=[torch.randn(2, 256, 768)])
    for fn in (euler_sampler, euler_maruyama_sampler):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            fn(m, x, torch.zeros(2, dtype=torch.long))


def test_siloss_host_side_contract():
    from reed_b200.image.loss import SILoss
    with pytest.raises(AssertionError):
        SILoss(enc_names=["a", "b"], loss_weights={"a": 1.0})
    fn = SILoss(enc_names=["dinov2"], loss_weights={"dinov2": 1.0})
    t = torch.rand(5, 1, 1, 1)
    assert fn.time_weight(t, 0.5, "linear").shape == (5, 1, 1, 1)
    with pytest.raises(ValueError):
        fn.time_weight(t, 1.0, "bogus")
    with pytest.raises(ValueError):
        fn.encoder_weight(1.0, 1, 10, schedule="bogus")
    assert abs(fn.encoder_weight(2.0, 5, 10, "linear", "text") - 1.0) < 1e-12
    bad = SILoss(path_type="edm", enc_names=["dinov2"], loss_weights={"dinov2": 1.0})
    with pytest.raises(NotImplementedError):
        bad.interpolant(t)
    from oracle import loss_oracle
    for sched in ("linear", "cosine", "sigmoid", "constant", "loglinear", "cutoff"):
        a = fn.time_weight(t, 0.7, sched, [0.25, 0.75])
        b = loss_oracle.schedule_weight(t, 0.7, sched, (0.25, 0.75))
        assert torch.allclose(a, b, atol=1e-7), sched


def test_gemm_reserve_sms_is_host_state():
    """reed_gemm_reserve_sms only edits the host-side GEMM planner (no GPU needed) and rejects nonsense."""
    from reed_b200 import _cabi
    _cabi.load()
    _cabi.call("reed_gemm_reserve_sms", 16)
    _cabi.call("reed_gemm_reserve_sms", 0)
    with pytest.raises(_cabi.ReedLibraryError):
        _cabi.call("reed_gemm_reserve_sms", 1000)


def test_argument_errors_are_reported_without_a_device(built):
    """Argument validation happens before any CUDA call: non-zero return + reed_last_error(), no launch, no device needed."""
    lib = built.load()
    cases = [
        ("reed_sample_posterior", (None, None, None, None, 1.0, 0.0, None, -1, 4, 16, None), "sample_posterior: bad shape"),
        ("reed_sampler_step", (None, None, 0, None, None, None, None, None, 8, 0, 0, 1, 7, 1.0, 0.5, -0.1, None),
         "path_type"),
        ("reed_gemm_reserve_sms", (65,), "out of range"),
        ("reed_unary", (None, 0, None, 1, 0, 6, None), "n % 4"),
        ("reed_adamw_ema", (None, None, None, None, None, None, 8, None, 1.0, 1.0, 1e-4, 0.9, 0.999, 1e-8, 0.0, 0, 0.9999,
                            None, None), "1-based"),
    ]
    for name, args, needle in cases:
        with pytest.raises(built.ReedLibraryError, match=name) as err:
            built.call(name, *args)
        assert needle in str(err.value), (name, str(err.value))
    assert getattr(lib, "reed_gemm_reserve_sms")(0) == 0          # and a valid host-only call succeeds
    # zero-sized work returns before touching the device as well
    built.call("reed_sample_posterior", None, None, None, None, 1.0, 0.0, None, 0, 4, 16, None)
    built.call("reed_grad_sumsq", None, 0, None, None)


def test_cutoff_schedule_equals_reference_masked_assignment():
    """time_weight('cutoff') without index_put (graph-capturable) gives the reference's values (loss.py:143-147)."""
    from reed_b200.image.loss import SILoss
    fn = SILoss(enc_names=["dinov2"], loss_weights={"dinov2": 1.0})
    t = torch.tensor([0.0, 0.1, 0.2, 0.5, 0.8, 0.81, 1.0]).view(-1, 1, 1, 1)
    want = torch.ones_like(t)
    want[t < 0.2] = 0
    want[t > 0.8] = 0
    got = fn.time_weight(t, 0.7, "cutoff", [0.2, 0.8])
    assert got.shape == t.shape and torch.equal(got, 0.7 * want)


def test_lazy_shadows_are_keyed_on_the_weights_epoch(monkeypatch):
    """A bf16 shadow made lazily by weight_for goes stale when a kernel rewrites the master through its raw pointer
    (tensor._version does not move): the weights epoch invalidates it; optimizer-managed shadows stay valid."""
    from reed_b200 import ops
    casts = []

    def fake_launch(name, *args, n=1):
        casts.append(name)
        ops.launch_count += n
    monkeypatch.setattr(ops, "_launch", fake_launch)
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    p = torch.nn.Parameter(torch.randn(4, 8))
    s1 = ops.weight_for(p, torch.bfloat16)
    assert casts == ["reed_unary"] and ops.weight_for(p, torch.bfloat16) is s1 and len(casts) == 1
    ops.bump_weights_epoch()
    assert ops.weight_for(p, torch.bfloat16) is s1 and len(casts) == 2          # re-cast, same storage
    p._reed_shadow_managed = True
    ops.bump_weights_epoch()
    assert ops.weight_for(p, torch.bfloat16) is s1 and len(casts) == 2          # managed: the optimizer keeps it fresh
    p._reed_shadow_managed = False
    m = torch.nn.Linear(8, 4)
    assert ops.refresh_shadows(m) == 0                                          # no shadows yet: nothing to do
    ops.weight_for(m.weight, torch.bfloat16)
    ops.bump_weights_epoch()
    assert ops.refresh_shadows(m) == 1 and ops.refresh_shadows(m) == 0
    assert ops.weight_for(p, torch.float32).data_ptr() == p.data_ptr()


def test_preprocess_plan_matches_reference_substring_rules():
    from reed_b200.image.preprocess import _plan, preprocess_raw_image
    assert _plan("dinov1-vit-b", 256) is not None and _plan("dinov1", 256)[2] == 256      # train.py:66 `'dinov1' in enc_type`
    assert _plan("dinov2-vit-b", 512)[2] == 448 and _plan("sam", 256) is None
    x = torch.zeros(1, 3, 8, 8)
    assert preprocess_raw_image(x, "sam") is x


def test_custom_ops_are_registered_with_fake_kernels():
    """torch.ops.reed.* exist, carry schemas, and their fake kernels give the output shapes/dtypes without a device run."""
    from torch._subclasses.fake_tensor import FakeTensorMode
    from reed_b200 import library
    for name in library.OPS:
        op = getattr(torch.ops.reed, name)
        assert str(op.default._schema).startswith(f"reed::{name}(")
    with FakeTensorMode():
        x, s = torch.empty(2, 16, 32, device="cuda"), torch.empty(2, 32, device="cuda")
        out, mean, rstd = torch.ops.reed.ln_modulate(x, s, s, True)
        assert out.shape == (2, 16, 32) and out.dtype == torch.bfloat16 and mean.shape == rstd.shape == (32,)
        qkv = torch.empty(2 * 256, 3 * 2 * 72, device="cuda", dtype=torch.bfloat16)
        o, lse = torch.ops.reed.attention(qkv, 2, 256, 2, 72)
        assert o.shape == (512, 144) and lse.shape == (2, 2, 256) and lse.dtype == torch.float32
        y, h = torch.ops.reed.linear(torch.empty(8, 64, device="cuda"), torch.empty(16, 64, device="cuda"), None, 2, False)
        assert y.shape == h.shape == (8, 16)
        assert torch.ops.reed.velocity_mse(torch.empty(3, 4, 8, 8, device="cuda"), torch.empty(3, 4, 8, 8, device="cuda"),
                                           torch.empty(3, 4, 8, 8, device="cuda"), torch.empty(3, device="cuda"), 0).shape == (3,)
