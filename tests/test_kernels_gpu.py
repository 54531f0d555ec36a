"""Kernel-level parity on the B200: every C-ABI kernel against a plain PyTorch fp32/fp64 restatement of the same op."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    from reed_b200 import _cabi, ops as _ops
    _cabi.load()
    _ops.device_check()
    yield _ops
    _ops.set_backends()


def _rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _rand(*shape, dtype=torch.float32, scale=1.0, seed=None):
    g = torch.Generator(device="cpu").manual_seed(seed if seed is not None else (hash(shape) % 1000))
    return (torch.randn(*shape, generator=g) * scale).to(DEV).to(dtype)


# ---------------------------------------------------------------------------------------------------------
# GEMM
# ---------------------------------------------------------------------------------------------------------

GEMM_SHAPES = [
    # M, N, K
    (256, 384, 128),       # aligned, small
    (1000, 192, 320),      # ragged M
    (128, 2304, 768),      # B/2 qkv
    (512, 1152, 4608),     # XL fc2, long K
    (77, 72 * 8, 200),     # ragged everything (N % 8 == 0)
    (32, 6912, 1152),      # adaLN: M = batch
    (4096, 768, 64),       # one k-block
    (6912, 1152, 32),      # adaLN wgrad: reduction over the batch (half a k-block, TMA zero-fill)
    (256, 128, 16),        # shortest supported reduction
    (2048, 16, 384),       # final layer: 16 output columns (one mostly out-of-bounds tile)
    (16, 1152, 2048),      # its weight gradient: 16 output rows
]


def _operands(M, N, K, a_mn, b_mn, dtype):
    a = _rand(M, K, dtype=dtype, seed=1)
    b = _rand(N, K, dtype=dtype, scale=K ** -0.5, seed=2)
    a_store = a.t().contiguous() if a_mn else a
    b_store = b.t().contiguous() if b_mn else b
    return a, b, a_store, b_store


@pytest.mark.parametrize("cg", [0, 1, 2])
@pytest.mark.parametrize("layout", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("shape", GEMM_SHAPES)
def test_gemm_tcgen05_layouts(ops, shape, layout, cg):
    """cg: 0 = the planner's choice, 1 / 2 = force cta_group::1 / the cta_group::2 CTA-pair kernel."""
    M, N, K = shape
    a_mn, b_mn = layout
    if (a_mn and M % 8) or (b_mn and N % 8) or (not a_mn and K % 8) or (not b_mn and K % 8):
        pytest.skip("TMA needs 16-byte row pitches")
    ops.set_backends(gemm={0: ops.BACKEND_TENSOR, 1: ops.BACKEND_TENSOR_CG1, 2: ops.BACKEND_TENSOR_CG2}[cg])
    a, b, a_s, b_s = _operands(M, N, K, a_mn, b_mn, torch.bfloat16)
    ref = a.float() @ b.float().t()
    out = ops.gemm(a_s, b_s, a_mn=bool(a_mn), b_mn=bool(b_mn), out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert _rel(out, ref) < 2e-5, (shape, layout)          # same bf16 inputs, fp32 accumulation
    out_bf = ops.gemm(a_s, b_s, a_mn=bool(a_mn), b_mn=bool(b_mn), out_dtype=torch.bfloat16)
    assert _rel(out_bf.float(), ref) < 1e-2


@pytest.mark.parametrize("mode", ["fp32_simt", "bf16_simt", "bf16_tensor", "bf16_tensor_cg1", "bf16_tensor_cg2"])
def test_gemm_epilogues(ops, mode):
    dtype = torch.float32 if mode == "fp32_simt" else torch.bfloat16
    ops.set_backends(gemm={"fp32_simt": ops.BACKEND_AUTO, "bf16_simt": ops.BACKEND_SIMT, "bf16_tensor": ops.BACKEND_TENSOR,
                           "bf16_tensor_cg1": ops.BACKEND_TENSOR_CG1, "bf16_tensor_cg2": ops.BACKEND_TENSOR_CG2}[mode])
    tol = 2e-5 if dtype == torch.float32 else 1.5e-2
    B, T, N, K = 5, 64, 256, 192        # M = 320: a full and a ragged 128/256-row tile, groups straddling tiles
    M = B * T
    a, w, _, _ = _operands(M, N, K, 0, 0, dtype)
    bias = _rand(N, seed=3)
    acc = a.float() @ w.float().t() + bias
    # bias only, fp32 out; then accumulate on top
    out = ops.gemm(a, w, out_dtype=torch.float32, bias=bias)
    assert _rel(out, acc) < tol
    ops.gemm(a, w, out=out, accumulate=True)
    assert _rel(out, 2 * acc - bias) < tol
    # GELU / SiLU with saved pre-activation
    for epi, fn in ((ops.EPI_GELU, lambda v: F.gelu(v, approximate="tanh")), (ops.EPI_SILU, F.silu)):
        h = torch.empty(M, N, device=DEV, dtype=dtype)
        y = ops.gemm(a, w, out_dtype=dtype, bias=bias, epilogue=epi, out2=h)
        assert _rel(h.float(), acc) < tol
        assert _rel(y.float(), fn(h.float())) < tol
    # gate * y + residual
    res = _rand(M, N, seed=4)
    mod = _rand(B, 3 * N, seed=5)
    gate = mod[:, N:2 * N]
    y2 = torch.empty(M, N, device=DEV, dtype=dtype)
    x_new = ops.gemm(a, w, out_dtype=torch.float32, bias=bias, epilogue=ops.EPI_GATE_RES, aux=res, gate=gate,
                     rows_per_group=T, out2=y2)
    assert _rel(y2.float(), acc) < tol
    want = res + gate.repeat_interleave(T, dim=0) * y2.float()
    assert _rel(x_new, want) < 1e-5
    # dgrad with activation derivative: D = (dy W) * act'(h)
    dy = _rand(M, N, dtype=dtype, seed=6)
    hsave = _rand(M, K, dtype=dtype, seed=7)
    for epi, fn in ((ops.EPI_DGELU, lambda v: F.gelu(v, approximate="tanh")), (ops.EPI_DSILU, F.silu)):
        hv = hsave.float().requires_grad_(True)
        (grad,) = torch.autograd.grad(fn(hv).sum(), hv)
        want = (dy.float() @ w.float()) * grad
        got = ops.gemm(dy, w, b_mn=True, out_dtype=dtype, epilogue=epi, aux=hsave)
        assert _rel(got.float(), want) < tol


@pytest.mark.parametrize("bn", [0, 1, 2, 3])                   # planner's choice / tile width 128 / 192 / 256
def test_gemm_tma_epilogue_many_tiles(ops, bn):
    """bf16 outputs leave through the TMA epilogue: every persistent CTA walks several tiles, so the per-warp operand
    ring, the bias look-ahead and the store boxes carry over tile boundaries; M and N are ragged (N % 32 == 8)."""
    ops.set_backends(gemm=ops.BACKEND_TENSOR_CG2 + 8 * bn)
    # a last column tile at most 32 columns wide: half of the epilogue warps own no chunk of it
    a, w, _, _ = _operands(20000, 1056, 64, 0, 0, torch.bfloat16)
    bias = _rand(1056, seed=5)
    out = ops.gemm(a, w, out_dtype=torch.bfloat16, bias=bias)
    assert _rel(out.float(), a.float() @ w.float().t() + bias) < 1e-2
    M, N, K = 40000 + 24, 1000, 128
    a, w, _, _ = _operands(M, N, K, 0, 0, torch.bfloat16)
    bias = _rand(N, seed=3)
    acc = a.float() @ w.float().t() + bias
    out = ops.gemm(a, w, out_dtype=torch.bfloat16, bias=bias)
    assert _rel(out.float(), acc) < 1e-2
    h = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    y = ops.gemm(a, w, out_dtype=torch.bfloat16, bias=bias, epilogue=ops.EPI_GELU, out2=h)
    assert _rel(h.float(), acc) < 1e-2
    assert _rel(y.float(), F.gelu(h.float(), approximate="tanh")) < 1.5e-2
    if bn == 2:
        return                                             # MN-major B has no 192-wide pair tile (96 columns per CTA)
    # dgrad through the activation: D[M, K] = (dy[M, N] W[N, K]) * gelu'(hsave[M, K]); W read MN-major
    K2 = 768
    w2 = _rand(N, K2, dtype=torch.bfloat16, scale=N ** -0.5, seed=8)
    dy = _rand(M, N, dtype=torch.bfloat16, seed=6)
    hsave = _rand(M, K2, dtype=torch.bfloat16, seed=7)
    hv = hsave.float().requires_grad_(True)
    (grad,) = torch.autograd.grad(F.gelu(hv, approximate="tanh").sum(), hv)
    want = (dy.float() @ w2.float()) * grad
    got = ops.gemm(dy, w2, b_mn=True, out_dtype=torch.bfloat16, epilogue=ops.EPI_DGELU, aux=hsave)
    torch.cuda.synchronize()
    assert _rel(got.float(), want) < 1.5e-2


@pytest.mark.parametrize("shape", [(32, 1152, 28), (4, 128, 2), (128, 384, 5), (7, 768, 12), (256, 768, 12), (200, 128, 3)])   # (batch, width, blocks)
def test_gemm_grouped_adaln(ops, shape):
    """reed_gemm_grouped: adaLN_modulation(c) of all blocks (sit.py:125-133) as one launch over separately stored weights,
    and the gradient of the shared input as one split-K launch over the concatenated reduction."""
    B, D, L = shape
    c = _rand(B, D, dtype=torch.bfloat16, seed=1)
    ws = [_rand(6 * D, D, dtype=torch.bfloat16, scale=D ** -0.5, seed=10 + g) for g in range(L)]
    bias = _rand(L * 6 * D, seed=2)
    out = torch.empty(B, L * 6 * D, device=DEV)
    ops.gemm_grouped_fwd(c, ws, bias, out)
    want = torch.cat([c.float() @ w.float().t() for w in ws], dim=1) + bias
    assert _rel(out, want) < 2e-3
    dys = [_rand(B, 6 * D, dtype=torch.bfloat16, seed=50 + g) for g in range(L)]
    want = sum(d.float() @ w.float() for d, w in zip(dys, ws))
    dc = torch.full((B, D), 7.0, device=DEV)
    ops.gemm_grouped_dgrad(dys, ws, dc, accumulate=False)
    assert _rel(dc, want) < 2e-3
    ops.gemm_grouped_dgrad(dys[:max(1, L // 2)], ws[:max(1, L // 2)], dc, accumulate=True)
    want2 = want + sum(d.float() @ w.float() for d, w in zip(dys[:max(1, L // 2)], ws[:max(1, L // 2)]))
    assert _rel(dc, want2) < 2e-3


@pytest.mark.parametrize("shape", [(32, 6912, 1152), (4, 768, 128), (64, 200, 72), (7, 1536, 260)])   # (batch, n_out, k_in)
@pytest.mark.parametrize("dy_dtype", [torch.float32, torch.bfloat16])
def test_outer_wgrad_with_bias(ops, shape, dy_dtype):
    """reed_outer_wgrad: dW (+)= dy^T x and db += colsum(dy) for a batch-sized contraction (autograd of adaLN_modulation,
    sit.py:125-128), fp32 or bf16 dy against a bf16 input."""
    B, N, K = shape
    dy = _rand(B, N, dtype=dy_dtype, seed=1)
    x = _rand(B, K, dtype=torch.bfloat16, seed=2)
    dw = torch.full((N, K), 3.0, device=DEV)
    db = torch.full((N,), 0.5, device=DEV)
    ops.outer_wgrad(dy, x, dw, db, accumulate=False)
    want_w = dy.float().t() @ x.float()
    want_b = 0.5 + dy.float().sum(0)
    assert _rel(dw, want_w) < 1e-5
    assert _rel(db, want_b) < 1e-5
    ops.outer_wgrad(dy, x, dw, None, accumulate=True)
    assert _rel(dw, 2 * want_w) < 1e-5
    assert _rel(db, want_b) < 1e-5


@pytest.mark.parametrize("bn", [0, 1, 2, 3])                   # planner's choice / tile width 128 / 192 / 256
@pytest.mark.parametrize("cg", [1, 2])
def test_gemm_gate_residual_many_tiles(ops, bn, cg):
    """x + gate * (a W^T + b) over several tiles per persistent CTA (sit.py:134-135): the residual ring of the TMA
    epilogue, the bias / gate look-ahead and the in-place boxes carry over tile boundaries; ragged M and N."""
    ops.set_backends(gemm=(ops.BACKEND_TENSOR_CG1 if cg == 1 else ops.BACKEND_TENSOR_CG2) + 8 * bn)
    for B, T, N, K in ((157, 96, 1104, 128), (40, 256, 1152, 320), (3, 32, 48, 64)):
        M = B * T
        a, w, _, _ = _operands(M, N, K, 0, 0, torch.bfloat16)
        bias = _rand(N, seed=3)
        res = _rand(M, N, seed=4)
        gate = _rand(B, 3 * N, seed=5)[:, N:2 * N]
        y = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
        out = ops.gemm(a, w, out_dtype=torch.float32, bias=bias, epilogue=ops.EPI_GATE_RES, aux=res, gate=gate,
                       rows_per_group=T, out2=y)
        acc = a.float() @ w.float().t() + bias
        assert _rel(y.float(), acc) < 1e-2
        want = res + gate.repeat_interleave(T, dim=0) * y.float()
        assert _rel(out, want) < 1e-5
        out_b = ops.gemm(a, w, out_dtype=torch.float32, bias=bias, epilogue=ops.EPI_GATE_RES, aux=res, gate=gate,
                         rows_per_group=T)                      # without the saved y
        assert torch.equal(out_b, out)


@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("bn", [1, 2, 3])                      # tile width 128 / 192 / 256
@pytest.mark.parametrize("layout", [(0, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("shape", [(2560, 1152, 192), (1024, 1096, 320), (700, 320, 128), (512, 64, 256)])
def test_gemm_ragged_last_column_tile(ops, shape, layout, bn, cg):
    """N leaves a remainder of at most half a tile: the last column tile runs as a half-width MMA (each CTA of a pair
    supplies a quarter tile of B) and the tiles are dealt in snake order.  Every (cta_group, BN, layout) is pinned."""
    M, N, K = shape
    a_mn, b_mn = layout
    if (a_mn and M % 8) or (b_mn and N % 8):
        pytest.skip("TMA needs 16-byte row pitches")
    ops.set_backends(gemm={1: ops.BACKEND_TENSOR_CG1, 2: ops.BACKEND_TENSOR_CG2}[cg] + 8 * bn)
    a, b, a_s, b_s = _operands(M, N, K, a_mn, b_mn, torch.bfloat16)
    ref = a.float() @ b.float().t()
    bias = _rand(N, seed=3)
    out = ops.gemm(a_s, b_s, a_mn=bool(a_mn), b_mn=bool(b_mn), out_dtype=torch.float32, bias=bias)
    torch.cuda.synchronize()
    assert _rel(out, ref + bias) < 2e-5, (shape, layout, bn, cg)
    out_bf = ops.gemm(a_s, b_s, a_mn=bool(a_mn), b_mn=bool(b_mn), out_dtype=torch.bfloat16)
    assert _rel(out_bf.float(), ref) < 1e-2


@pytest.mark.parametrize("cg", [0, 1, 2])
@pytest.mark.parametrize("shape", [(1152, 1152, 8192), (3456, 1152, 4096), (1152, 4608, 2048), (384, 1536, 8192)])
def test_gemm_stream_k_wgrad(ops, shape, cg):
    """Weight-gradient shapes whose tile count does not fill the SMs take the split path: whole rounds data-parallel,
    the last partial round cut into k slices that are added with fp32 red.add into rows zeroed beforehand."""
    ops.set_backends(gemm={0: ops.BACKEND_TENSOR, 1: ops.BACKEND_TENSOR_CG1, 2: ops.BACKEND_TENSOR_CG2}[cg])
    M, N, K = shape                                        # dW[M=N_out, N=K_in] = dy^T x, reduction over K tokens
    dy = _rand(K, M, dtype=torch.bfloat16, seed=11)
    x = _rand(K, N, dtype=torch.bfloat16, scale=K ** -0.5, seed=12)
    ref = dy.float().t() @ x.float()
    out = torch.full((M, N), 7.0, device=DEV)              # stale contents must be overwritten, not accumulated
    ops.gemm(dy, x, a_mn=True, b_mn=True, out=out)
    torch.cuda.synchronize()
    assert _rel(out, ref) < 2e-5, shape
    ops.gemm(dy, x, a_mn=True, b_mn=True, out=out, accumulate=True)
    assert _rel(out, 2 * ref) < 2e-5, shape
    # a strided destination (a view into a flat gradient bucket row-block)
    big = torch.zeros((M, N + 64), device=DEV)
    ops.gemm(dy, x, a_mn=True, b_mn=True, out=big[:, :N])
    assert _rel(big[:, :N], ref) < 2e-5 and float(big[:, N:].abs().max()) == 0.0


@pytest.mark.parametrize("shape", [(1152, 1152, 4096), (3456, 1152, 2048), (384, 256, 1024), (512, 128, 48)])
def test_gemm_wgrad_with_bias_column(ops, shape):
    """dW = dy^T x and db = colsum(dy) from ONE GEMM: LN+modulate leaves a ones column behind the activation rows and the
    weight-gradient GEMM contracts against it (last shape: too few tokens for the tensor path -> GEMM + column sums)."""
    n_out, k_in, tokens = shape
    T = 16
    x = _rand(tokens, k_in, seed=21)
    mod = _rand(tokens // T, 2 * k_in, scale=0.3, seed=22)
    xm, _, _ = ops.ln_modulate_fwd(x, mod[:, :k_in], mod[:, k_in:], T, torch.bfloat16, ones_col=True)
    ext = xm._base
    assert ext.shape[1] >= k_in + 8 and xm.stride(0) == ext.shape[1] and (ext.shape[1] * 2) % 128 == 0
    torch.cuda.synchronize()
    assert bool((ext[:, k_in] == 1).all()) and bool((ext[:, k_in + 1:k_in + 8] == 0).all())
    dy = _rand(tokens, n_out, dtype=torch.bfloat16, scale=tokens ** -0.5, seed=23)
    dw = torch.full((n_out, k_in), 3.0, device=DEV)        # stale contents must be overwritten
    db = torch.zeros(n_out, device=DEV)
    ops.wgrad_bias(dy, ext, k_in, dw, db, accumulate=False)
    torch.cuda.synchronize()
    ref_w = dy.float().t() @ xm.float()
    ref_b = dy.float().sum(0)
    assert _rel(dw, ref_w) < 2e-5, shape
    assert _rel(db, ref_b) < 2e-5, shape
    ops.wgrad_bias(dy, ext, k_in, dw, db, accumulate=True)       # gradient accumulation: both add up
    torch.cuda.synchronize()
    assert _rel(dw, 2 * ref_w) < 2e-5 and _rel(db, 2 * ref_b) < 2e-5


def test_gemm_simt_skinny_and_split_k(ops):
    ops.set_backends()
    # final-layer shapes: N = 16 forward, and its wgrad with a long reduction (split-K path)
    M, D = 8192, 384
    x = _rand(M, D, seed=1)
    w = _rand(16, D, scale=D ** -0.5, seed=2)
    out = ops.gemm(x, w, out_dtype=torch.float32)
    assert _rel(out, x @ w.t()) < 2e-5
    dy = _rand(M, 16, seed=3)
    dw = ops.gemm(dy, x, a_mn=True, b_mn=True, out_dtype=torch.float32)
    assert _rel(dw, dy.t() @ x) < 5e-5
    prev = dw.clone()
    ops.gemm(dy, x, a_mn=True, b_mn=True, out=dw, accumulate=True)
    assert _rel(dw, 2 * prev) < 5e-5
    xb, dyb = x.bfloat16(), dy.bfloat16()
    dwb = ops.gemm(dyb, xb, a_mn=True, b_mn=True, out_dtype=torch.float32)
    assert _rel(dwb, dyb.float().t() @ xb.float()) < 5e-5


# ---------------------------------------------------------------------------------------------------------
# LayerNorm + modulate, gate backward, small elementwise
# ---------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("D", [384, 768, 1152])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_ln_modulate_fwd_bwd(ops, D, dtype):
    B, T = 3, 40
    x = _rand(B * T, D, seed=1) * 2 + 0.3
    mod = _rand(B, 6 * D, scale=0.3, seed=2)
    shift, scale = mod[:, :D], mod[:, D:2 * D]
    out, mean, rstd = ops.ln_modulate_fwd(x, shift, scale, T, dtype)
    xr = x.clone().requires_grad_(True)
    sh = shift.clone().requires_grad_(True)
    sc = scale.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (D,), eps=1e-6).view(B, T, D) * (1 + sc[:, None]) + sh[:, None]
    tol = 2e-6 if dtype == torch.float32 else 8e-3
    assert _rel(out.float().view(B, T, D), ref) < tol
    dout = _rand(B * T, D, dtype=dtype, seed=3)
    dres = _rand(B * T, D, seed=4)
    ref.backward(dout.float().view(B, T, D))
    dmod = torch.zeros_like(mod)
    dx = ops.ln_modulate_bwd(dout, x, mean, rstd, scale, T, dres, dmod[:, :D], dmod[:, D:2 * D])
    assert _rel(dx, xr.grad + dres) < 2e-5
    assert _rel(dmod[:, :D], sh.grad) < 2e-5
    assert _rel(dmod[:, D:2 * D], sc.grad) < 2e-5
    assert float(dmod[:, 2 * D:].abs().max()) == 0.0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gate_bwd_and_colsum(ops, dtype):
    B, T, D = 4, 48, 768
    dxn = _rand(B * T, D, seed=1)
    y = _rand(B * T, D, dtype=dtype, seed=2)
    mod = _rand(B, 6 * D, seed=3)
    gate = mod[:, 2 * D:3 * D]
    dmod = torch.zeros_like(mod)
    dbias = torch.zeros(D, device=DEV)
    dy = ops.gate_bwd(dxn, y, gate, T, dmod[:, 2 * D:3 * D], dbias)
    want_dy = (dxn * gate.repeat_interleave(T, dim=0)).to(dtype)
    assert _rel(dy.float(), want_dy.float()) < 1e-6 + (4e-3 if dtype == torch.bfloat16 else 0)
    assert _rel(dmod[:, 2 * D:3 * D], (dxn * y.float()).view(B, T, D).sum(1)) < 2e-5
    assert _rel(dbias, dy.float().sum(0)) < 2e-5
    out = torch.zeros(D, device=DEV)
    ops.colsum(y, out)
    assert _rel(out, y.float().sum(0)) < 2e-5


@pytest.mark.parametrize("D", [128, 384, 1024, 1152])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_ln_modulate_gate_bwd_fused(ops, D, dtype):
    """The fused LN-backward + gate-backward equals the two separate kernels' results."""
    B, T = 3, 40
    x = _rand(B * T, D, seed=1) * 2 + 0.3
    mod = _rand(B, 6 * D, scale=0.3, seed=2)
    scale, gate = mod[:, D:2 * D], mod[:, 2 * D:3 * D]
    _, mean, rstd = ops.ln_modulate_fwd(x, mod[:, :D], scale, T, dtype)
    dout = _rand(B * T, D, dtype=dtype, seed=3)
    dres = _rand(B * T, D, seed=4)
    y = _rand(B * T, D, dtype=dtype, seed=5)
    dmod_a, dmod_b = torch.zeros_like(mod), torch.zeros_like(mod)
    db_a, db_b = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
    dx_ref = ops.ln_modulate_bwd(dout, x, mean, rstd, scale, T, dres, dmod_a[:, :D], dmod_a[:, D:2 * D])
    dy_ref = ops.gate_bwd(dx_ref, y, gate, T, dmod_a[:, 2 * D:3 * D], db_a)
    dx, dy = ops.ln_modulate_gate_bwd(dout, x, mean, rstd, scale, T, dres, dmod_b[:, :D], dmod_b[:, D:2 * D], y, gate,
                                      dmod_b[:, 2 * D:3 * D], db_b)
    torch.cuda.synchronize()
    assert _rel(dx, dx_ref) < 1e-6
    assert _rel(dy.float(), dy_ref.float()) < 1e-6
    assert _rel(dmod_b, dmod_a) < 2e-5 and _rel(db_b, db_a) < 2e-5
    # against autograd: d/dx of LN-modulate, then the gate branch
    xr = x.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (D,), eps=1e-6).view(B, T, D) * (1 + scale[:, None])
    ref.backward(dout.float().view(B, T, D))
    assert _rel(dx, xr.grad + dres) < 2e-5


def test_unary_actbwd_tokenmean(ops):
    x = _rand(6, 512, seed=1)
    assert _rel(ops.cast(x, torch.bfloat16).float(), x.bfloat16().float()) == 0.0
    assert _rel(ops.cast(x, torch.float32, op=1), F.silu(x)) < 1e-6
    dy = _rand(6, 512, seed=2)
    for act, fn in ((ops.ACT_GELU, lambda v: F.gelu(v, approximate="tanh")), (ops.ACT_SILU, F.silu)):
        xv = x.clone().requires_grad_(True)
        fn(xv).backward(dy)
        assert _rel(ops.act_bwd(dy, x, act), xv.grad) < 2e-6
    tok = _rand(3, 64, 256, seed=3).requires_grad_(True)
    m = ops.TokenMeanFn.apply(tok, torch.float32)
    assert _rel(m, tok.mean(1)) < 1e-6
    m.backward(torch.ones_like(m))
    assert _rel(tok.grad, torch.full_like(tok, 1 / 64)) < 1e-6


# ---------------------------------------------------------------------------------------------------------
# attention
# ---------------------------------------------------------------------------------------------------------

def _attention_reference(qkv, B, T, H, hd):
    q, k, v = qkv.view(B, T, 3, H, hd).permute(2, 0, 3, 1, 4)
    p = torch.softmax(q @ k.transpose(-1, -2) * hd ** -0.5, dim=-1)
    return (p @ v).transpose(1, 2).reshape(B * T, H * hd)


@pytest.mark.parametrize("cfg", [(2, 256, 3, 64), (2, 64, 2, 72), (1, 1024, 2, 72)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_attention_fwd_bwd(ops, cfg, dtype):
    B, T, H, hd = cfg
    ops.set_backends()
    qkv = _rand(B * T, 3 * H * hd, dtype=dtype, seed=1)
    o, lse = ops.attention_fwd(qkv, B, T, H, hd)
    ref_in = qkv.float().requires_grad_(True)
    ref = _attention_reference(ref_in, B, T, H, hd)
    tol = 3e-6 if dtype == torch.float32 else 1.2e-2
    assert _rel(o.float(), ref) < tol
    d_o = _rand(B * T, H * hd, dtype=dtype, seed=2)
    ref.backward(d_o.float())
    dqkv = ops.attention_bwd(qkv, o, d_o, lse, B, T, H, hd)
    assert _rel(dqkv.float(), ref_in.grad) < (2e-5 if dtype == torch.float32 else 2e-2)


@pytest.mark.parametrize("backend", [3, 4])                    # 3 = mma.sync kernels, 4 = tcgen05/TMEM kernels
@pytest.mark.parametrize("cfg", [(2, 256, 3, 64), (3, 256, 2, 72), (2, 128, 2, 64), (1, 128, 3, 72), (32, 256, 16, 72),
                                 (2, 512, 2, 72), (1, 1024, 3, 64), (3, 1024, 2, 72), (160, 256, 1, 72)])
@pytest.mark.parametrize("gain", [1.0, 5.0])
def test_attention_tensor_core_kernels(ops, cfg, backend, gain):
    """Both tensor-core attention implementations against fp32 torch attention on the same bf16 inputs.  tcgen05 path:
    one / two / four / eight key blocks (clusters of 1..8 CTAs in the backward ring), more work items than resident
    clusters (160 heads), and - with ``gain`` - score rows whose maximum grows along the key blocks by more than the
    lazy-rescale threshold of the forward's online softmax."""
    B, T, H, hd = cfg
    qkv = _rand(B * T, 3 * H * hd, dtype=torch.float32, seed=1).view(B * T, 3, H * hd)
    if gain != 1.0:
        qkv[:, 0] *= gain
        qkv[:, 1] *= 0.6 * gain * torch.linspace(0.2, 1.0, T, device=DEV).repeat(B)[:, None]
    qkv = qkv.view(B * T, 3 * H * hd).bfloat16()
    d_o = _rand(B * T, H * hd, dtype=torch.bfloat16, seed=2)
    ops.set_backends(attention=backend)
    o, lse = ops.attention_fwd(qkv, B, T, H, hd)
    dqkv = ops.attention_bwd(qkv, o, d_o, lse, B, T, H, hd)
    torch.cuda.synchronize()
    ops.set_backends()
    ref_in = qkv.float().requires_grad_(True)
    ref = _attention_reference(ref_in, B, T, H, hd)
    assert _rel(o.float(), ref) < 1.2e-2
    ref.backward(d_o.float())
    assert _rel(dqkv.float(), ref_in.grad) < 2e-2
    # per-slice check so that a wrong dq / dk / dv block cannot hide behind the largest one
    g = ref_in.grad.view(B * T, 3, H * hd)
    got = dqkv.float().view(B * T, 3, H * hd)
    for i in range(3):
        assert _rel(got[:, i], g[:, i]) < 2e-2, i
    q, k = ref_in.detach().view(B, T, 3, H, hd)[:, :, 0], ref_in.detach().view(B, T, 3, H, hd)[:, :, 1]
    s = torch.einsum("bthd,bshd->bhts", q, k) * hd ** -0.5
    want = torch.logsumexp(s, dim=-1)
    assert float(((lse - want).abs() / want.abs().clamp_min(1.0)).max()) < 2e-3


# ---------------------------------------------------------------------------------------------------------
# SILoss kernels
# ---------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("path", [0, 1])
def test_siloss_interp_and_mse(ops, path):
    B = 5
    x, eps, pred = _rand(B, 4, 32, 32, seed=1), _rand(B, 4, 32, 32, seed=2), _rand(B, 4, 32, 32, seed=3)
    t = torch.rand(B, device=DEV)
    tb = t.view(B, 1, 1, 1)
    if path == 0:
        a, s, da, ds = 1 - tb, tb, -1.0, 1.0
    else:
        a, s = torch.cos(tb * math.pi / 2), torch.sin(tb * math.pi / 2)
        da, ds = -math.pi / 2 * torch.sin(tb * math.pi / 2), math.pi / 2 * torch.cos(tb * math.pi / 2)
    assert _rel(ops.interpolate(x, eps, t, path), a * x + s * eps) < 2e-6
    pr = pred.clone().requires_grad_(True)
    ref = ((pr - (da * x + ds * eps)) ** 2).flatten(1).mean(1)
    pg = pred.clone().requires_grad_(True)
    got = ops.VelocityMSEFn.apply(pg, x, eps, t, path)
    assert _rel(got, ref) < 2e-6
    w = torch.rand(B, device=DEV)
    (ref * w).sum().backward()
    (got * w).sum().backward()
    assert _rel(pg.grad, pr.grad) < 2e-6


@pytest.mark.parametrize("dt_pair", [(torch.float32, torch.float32), (torch.bfloat16, torch.float32),
                                     (torch.bfloat16, torch.bfloat16), (torch.float32, torch.bfloat16)])
@pytest.mark.parametrize("shape", [(4, 256, 768), (3, 1, 3584)])
def test_siloss_cosine(ops, shape, dt_pair):
    B, T, Z = shape
    zt = _rand(B, T, Z, dtype=dt_pair[0], seed=1)
    z = _rand(B, T, Z, dtype=dt_pair[1], seed=2)
    zr = zt.detach().float().clone().requires_grad_(True)
    ref = -(F.normalize(z.float(), dim=-1) * F.normalize(zr, dim=-1)).sum(-1).mean(-1)
    zg = zt.detach().clone().requires_grad_(True)
    got = ops.CosineAlignFn.apply(zg, z)
    assert _rel(got, ref) < 5e-6
    w = torch.rand(B, device=DEV)
    (ref * w).sum().backward()
    (got * w).sum().backward()
    assert _rel(zg.grad.float(), zr.grad) < (2e-5 if dt_pair[0] == torch.float32 else 8e-3)


# ---------------------------------------------------------------------------------------------------------
# sampler step, optimizer
# ---------------------------------------------------------------------------------------------------------

def test_sampler_step_formulas(ops):
    n = 3
    x = torch.randn(n, 4, 16, 16, device=DEV, dtype=torch.float64)
    v = torch.randn(2 * n, 4, 16, 16, device=DEV)
    eps = torch.randn_like(x)
    t, dt, cfg = 0.7, -0.05, 1.8
    # ODE, guided
    xn, slope, xm = ops.sampler_step(x, v, want_slope=True, next_dup=True, guided=True, cfg=cfg, t_cur=t, dt=dt)
    d = v.double()
    d = d[n:] + cfg * (d[:n] - d[n:])
    assert _rel(xn, x + dt * d) < 1e-14 and _rel(slope, d) < 1e-14
    assert torch.equal(xm[:n], xn.float()) and torch.equal(xm[n:], xn.float())
    # Heun corrector
    xh, _, _ = ops.sampler_step(x, v[:n], d_prev=slope, t_cur=t, dt=dt)
    assert _rel(xh, x + dt * (0.5 * slope + 0.5 * v[:n].double())) < 1e-14
    # SDE, both paths
    for path, name in ((0, "linear"), (1, "cosine")):
        from oracle.samplers_oracle import score_from_velocity
        tin = torch.full((2 * n,), t, dtype=torch.float64, device=DEV)
        vv = v.double()
        s = score_from_velocity(vv, torch.cat([x, x]), tin, name)
        dd = vv - 0.5 * (2 * t) * s
        dd = dd[n:] + cfg * (dd[:n] - dd[n:])
        want = x + dd * dt + math.sqrt(2 * t) * eps * math.sqrt(abs(dt))
        got, _, _ = ops.sampler_step(x, v, eps=eps, guided=True, cfg=cfg, t_cur=t, dt=dt, sde=True, path_type=path)
        assert _rel(got, want) < 1e-12, name


def test_fused_adamw_ema_matches_torch(ops):
    from reed_b200._cabi import call
    n = 10007 * 4 + 3
    g0 = torch.Generator().manual_seed(0)
    p = torch.randn(n, generator=g0).to(DEV)
    ref_p = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([ref_p], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    ema_ref = p.clone()
    m, v, ema = torch.zeros_like(p), torch.zeros_like(p), p.clone()
    shadow = torch.zeros(n, device=DEV, dtype=torch.bfloat16)
    norm_sq = torch.zeros(1, device=DEV, dtype=torch.float64)
    st = torch.cuda.current_stream().cuda_stream
    for step in range(1, 4):
        g = (torch.randn(n, generator=g0) * (5.0 if step == 2 else 0.001)).to(DEV)
        ref_p.grad = g.clone()
        total = torch.nn.utils.clip_grad_norm_([ref_p], 1.0)
        opt.step()
        ema_ref.mul_(0.99).add_(ref_p.data, alpha=0.01)
        norm_sq.zero_()
        call("reed_grad_sumsq", g.data_ptr(), n, norm_sq.data_ptr(), st)
        assert _rel(norm_sq.sqrt().float(), total) < 1e-5
        # odd steps pass the step as a host scalar, even steps through the device counter a CUDA-graph replay reads
        step_dev = torch.tensor([step], device=DEV, dtype=torch.int32)
        call("reed_adamw_ema", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), ema.data_ptr(), shadow.data_ptr(),
             n, norm_sq.data_ptr(), 1.0, 1.0, 1e-3, 0.9, 0.999, 1e-8, 0.01, step if step % 2 else 0, 0.99,
             None if step % 2 else step_dev.data_ptr(), st)
        assert float((p - ref_p.data).abs().max()) < 2e-6
        assert float((ema - ema_ref).abs().max()) < 2e-6
        assert torch.equal(shadow, p.bfloat16())
