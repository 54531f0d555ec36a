"""torch.library registration of the kernels (BASELINE north_star: "registered as torch custom ops with autograd"):
opcheck (schema, fake kernel, autograd registration, AOT dispatch) on the B200, values against torch expressions, and a
torch.compile'd caller that traces through the ops."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def reed():
    from reed_b200 import library  # noqa: F401  (registers torch.ops.reed.*)
    return torch.ops.reed


def _r(*shape, dtype=torch.float32, seed=0, grad=False):
    g = torch.Generator(device=DEV).manual_seed(seed)
    t = torch.randn(*shape, device=DEV, generator=g).to(dtype)
    return t.requires_grad_(grad)


def _cases():
    B, T, H, hd = 2, 256, 2, 72
    return {
        "siloss_interp": (_r(3, 4, 8, 8), _r(3, 4, 8, 8, seed=1), torch.rand(3, device=DEV), 0),
        "velocity_mse": (_r(3, 4, 8, 8, grad=True), _r(3, 4, 8, 8, seed=1), _r(3, 4, 8, 8, seed=2), torch.rand(3, device=DEV), 1),
        "cosine_align": (_r(2, 16, 64, dtype=torch.bfloat16, grad=True), _r(2, 16, 64, seed=3)),
        "ln_modulate": (_r(2, 16, 128, grad=True), _r(2, 128, seed=1, grad=True), _r(2, 128, seed=2, grad=True), True),
        "attention": (_r(B * T, 3 * H * hd, dtype=torch.bfloat16, grad=True), B, T, H, hd),
        "linear": (_r(256, 128, dtype=torch.bfloat16, grad=True), _r(192, 128, dtype=torch.bfloat16, seed=1, grad=True),
                   _r(192, seed=2, grad=True), 1, True),
        "gemm_nt": (_r(256, 128, dtype=torch.bfloat16), _r(192, 128, dtype=torch.bfloat16, seed=1), None, False, False, True),
        "colsum": (_r(64, 96, dtype=torch.bfloat16),),
        "sampler_cast": (_r(2, 4, 8, 8).double(), True, True),
    }


@pytest.mark.parametrize("name", ["siloss_interp", "velocity_mse", "cosine_align", "ln_modulate", "attention", "linear", "gemm_nt",
                                  "colsum", "sampler_cast"])
def test_opcheck(reed, name):
    args = _cases()[name]
    # test_autograd_registration / test_faketensor / test_schema / test_aot_dispatch_dynamic
    torch.library.opcheck(getattr(reed, name).default, args, atol=3e-2, rtol=3e-2)


def test_values_and_gradients_against_torch(reed):
    x, sh, sc = _r(2, 16, 128, grad=True), _r(2, 128, seed=1, grad=True), _r(2, 128, seed=2, grad=True)
    out = reed.ln_modulate(x, sh, sc, False)[0]
    ref = F.layer_norm(x, (128,), eps=1e-6) * (1 + sc[:, None]) + sh[:, None]
    assert float((out - ref).abs().max()) < 1e-4
    w = _r(2, 16, 128, seed=5)
    g1 = torch.autograd.grad((out * w).sum(), (x, sh, sc))
    g2 = torch.autograd.grad((ref * w).sum(), (x, sh, sc))
    for a, b in zip(g1, g2):
        assert float((a - b).abs().max()) < 2e-4 * max(1.0, float(b.abs().max()))
    # linear with GELU against torch in fp32
    xl, wl, bl = _r(64, 128, grad=True), _r(96, 128, seed=1, grad=True), _r(96, seed=2, grad=True)
    y = reed.linear(xl, wl, bl, 1, False)[0]
    yr = F.gelu(F.linear(xl, wl, bl), approximate="tanh")
    assert float((y - yr).abs().max()) < 1e-4
    gy = _r(64, 96, seed=7)
    for a, b in zip(torch.autograd.grad((y * gy).sum(), (xl, wl, bl)), torch.autograd.grad((yr * gy).sum(), (xl, wl, bl))):
        assert float((a - b).abs().max()) < 1e-3 * max(1.0, float(b.abs().max()))


def test_compiled_caller_traces_through_the_ops(reed):
    """torch.compile sees reed::* as opaque ops with fake kernels: no graph break, same numbers as eager."""
    def fn(pred, x, eps, t, zt, z):
        return reed.velocity_mse(pred, x, eps, t, 0).mean() + 0.5 * reed.cosine_align(zt, z)[0].mean()

    args = (_r(3, 4, 8, 8, grad=True), _r(3, 4, 8, 8, seed=1), _r(3, 4, 8, 8, seed=2), torch.rand(3, device=DEV),
            _r(3, 16, 64, grad=True, seed=3), _r(3, 16, 64, seed=4))
    eager = fn(*args)
    g_e = torch.autograd.grad(eager, (args[0], args[4]))
    compiled = torch.compile(fn, fullgraph=True, backend="aot_eager")
    out = compiled(*args)
    g_c = torch.autograd.grad(out, (args[0], args[4]))
    assert float((out - eager).abs()) < 1e-6
    for a, b in zip(g_c, g_e):
        assert torch.allclose(a, b, atol=1e-6)
