"""Can the optimizer pass (HBM-bound) hide under the forward GEMMs (tensor-bound) of the next step?

Times (a) 28 blocks' forward GEMMs alone, (b) 29 AdamW+EMA bucket passes alone, (c) both at once on two streams.
REED_ADAMW_THREADS / REED_ADAMW_GRID shape the optimizer CTAs so they fit beside a resident GEMM CTA.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from reed_b200 import _cabi, ops  # noqa: E402

_cabi.load()
dev = "cuda"
D, M, T = 1152, 8192, 256
bf = torch.bfloat16
r = lambda *s, dt=bf: (torch.randn(*s, device=dev) * 0.05).to(dt)
x, x4 = r(M, D), r(M, 4 * D)
w_qkv, w_proj, w_fc1, w_fc2 = r(3 * D, D), r(D, D), r(4 * D, D), r(D, 4 * D)
b_qkv, b_d, b_4d = r(3 * D, dt=torch.float32), r(D, dt=torch.float32), r(4 * D, dt=torch.float32)
res, gate = r(M, D, dt=torch.float32), r(M // T, D, dt=torch.float32)
y_out, h_out = torch.empty(M, D, device=dev, dtype=bf), torch.empty(M, 4 * D, device=dev, dtype=bf)
qkv_out, a_out, x_out = torch.empty(M, 3 * D, device=dev, dtype=bf), torch.empty(M, 4 * D, device=dev, dtype=bf), torch.empty(M, D, device=dev)


def gemms(blocks=28):
    for _ in range(blocks):
        ops.gemm(x, w_qkv, out=qkv_out, bias=b_qkv)
        ops.gemm(x, w_proj, out=x_out, bias=b_d, epilogue=ops.EPI_GATE_RES, aux=res, gate=gate, rows_per_group=T, out2=y_out)
        ops.gemm(x, w_fc1, out=a_out, bias=b_4d, epilogue=ops.EPI_GELU, out2=h_out)
        ops.gemm(x4, w_fc2, out=x_out, bias=b_d, epilogue=ops.EPI_GATE_RES, aux=res, gate=gate, rows_per_group=T, out2=y_out)


n = 23_900_000 // 8 * 8
nb = 8      # distinct buckets (rotated) so the pass streams from HBM like the real one
P = [torch.zeros(n, device=dev) for _ in range(nb)]
G = [torch.randn(n, device=dev) * 1e-3 for _ in range(nb)]
M1 = [torch.zeros(n, device=dev) for _ in range(nb)]
M2 = [torch.zeros(n, device=dev) for _ in range(nb)]
E = [torch.zeros(n, device=dev) for _ in range(nb)]
S = [torch.zeros(n, device=dev, dtype=bf) for _ in range(nb)]
norm = torch.ones(1, device=dev, dtype=torch.float64)


def adam(stream, buckets=29):
    for i in range(buckets):
        j = i % nb
        ops._launch("reed_adamw_ema", P[j].data_ptr(), G[j].data_ptr(), M1[j].data_ptr(), M2[j].data_ptr(), E[j].data_ptr(),
                    S[j].data_ptr(), n, norm.data_ptr(), 1.0, 1.0, 1e-4, 0.9, 0.999, 1e-8, 0.0, 1, 0.9999, None,
                    stream.cuda_stream)


main = torch.cuda.current_stream()
side = torch.cuda.Stream(priority=0)
hi = torch.cuda.Stream(priority=-1)


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        fn()
        e1.record(main)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def g_graph(fn):
    """capture fn (which may fork to other streams from `main`) into a graph; returns a replay callable"""
    fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        fn()
    return gr.replay


def only_gemm():
    gemms()


def only_adam():
    adam(torch.cuda.current_stream())


def both():
    cur = torch.cuda.current_stream()
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        adam(side)
    gemms()
    cur.wait_stream(side)


for name, fn in (("gemm only", only_gemm), ("adam only", only_adam), ("both", both)):
    rp = g_graph(fn)
    print(f"{name:10s} {timed(rp):8.3f} ms   (threads={os.environ.get('REED_ADAMW_THREADS', '256')} grid={os.environ.get('REED_ADAMW_GRID', 'auto')})", flush=True)
