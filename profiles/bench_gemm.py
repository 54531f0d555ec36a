"""Times the tcgen05 GEMM on every shape/epilogue of one SiT block (fwd, dgrad, wgrad) with CUDA events.

    python profiles/bench_gemm.py [--model xl|b] [--tokens 8192] [--iters 20]

An L2-sized buffer is rewritten between iterations so every launch starts cold, like inside a train step.
Prints TFLOP/s per shape and the time one block's GEMMs take in total.
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from reed_b200 import _cabi, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="xl")
    ap.add_argument("--tokens", type=int, default=8192)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", default=None, help="substring filter on the case name")
    ap.add_argument("--no-flush", action="store_true", help="leave L2 warm between iterations")
    ap.add_argument("--bn", type=int, default=0, help="pin the tile width (128/192/256)")
    ap.add_argument("--cg", type=int, default=0, help="pin the cta_group (1 / 2); 0 = planner's choice")
    ap.add_argument("--cublas", action="store_true",
                    help="also time torch.matmul (cuBLAS, no epilogue) on the same operands: a yardstick, not the product path")
    args = ap.parse_args()
    _cabi.load()
    D = {"xl": 1152, "b": 768, "l": 1024, "s": 384}[args.model]
    M, T = args.tokens, 256
    dev = "cuda"
    ops.set_backends(gemm={0: ops.BACKEND_TENSOR, 1: ops.BACKEND_TENSOR_CG1, 2: ops.BACKEND_TENSOR_CG2}[args.cg]
                     + ({0: 0, 128: 1, 192: 2, 256: 3}[args.bn] << 3))
    bf = torch.bfloat16
    r = lambda *s, dt=bf: (torch.randn(*s, device=dev) * 0.05).to(dt)
    flush = torch.zeros(64 << 20, dtype=torch.int32, device=dev)
    x, x4 = r(M, D), r(M, 4 * D)
    w_qkv, w_proj, w_fc1, w_fc2 = r(3 * D, D), r(D, D), r(4 * D, D), r(D, 4 * D)
    b_qkv, b_d, b_4d = r(3 * D, dt=torch.float32), r(D, dt=torch.float32), r(4 * D, dt=torch.float32)
    res, gate = r(M, D, dt=torch.float32), r(M // T, D, dt=torch.float32)
    dy3, dyd, dy4 = r(M, 3 * D), r(M, D), r(M, 4 * D)
    h = r(M, 4 * D)
    y_out, h_out = torch.empty(M, D, device=dev, dtype=bf), torch.empty(M, 4 * D, device=dev, dtype=bf)
    g_qkv, g_proj = torch.zeros(3 * D, D, device=dev), torch.zeros(D, D, device=dev)
    g_fc1, g_fc2 = torch.zeros(4 * D, D, device=dev), torch.zeros(D, 4 * D, device=dev)
    cases = [
        ("qkv fwd   +bias", M, 3 * D, D, lambda: ops.gemm(x, w_qkv, out_dtype=bf, bias=b_qkv)),
        ("proj fwd  gate+res", M, D, D, lambda: ops.gemm(x, w_proj, out_dtype=torch.float32, bias=b_d, epilogue=ops.EPI_GATE_RES,
                                                         aux=res, gate=gate, rows_per_group=T, out2=y_out)),
        ("fc1 fwd   gelu", M, 4 * D, D, lambda: ops.gemm(x, w_fc1, out_dtype=bf, bias=b_4d, epilogue=ops.EPI_GELU, out2=h_out)),
        ("fc2 fwd   gate+res", M, D, 4 * D, lambda: ops.gemm(x4, w_fc2, out_dtype=torch.float32, bias=b_d, epilogue=ops.EPI_GATE_RES,
                                                             aux=res, gate=gate, rows_per_group=T, out2=y_out)),
        ("fc2 dgrad dgelu", M, 4 * D, D, lambda: ops.gemm(dyd, w_fc2, b_mn=True, out_dtype=bf, epilogue=ops.EPI_DGELU, aux=h)),
        ("fc1 dgrad", M, D, 4 * D, lambda: ops.gemm(dy4, w_fc1, b_mn=True, out_dtype=bf)),
        ("proj dgrad", M, D, D, lambda: ops.gemm(dyd, w_proj, b_mn=True, out_dtype=bf)),
        ("qkv dgrad", M, D, 3 * D, lambda: ops.gemm(dy3, w_qkv, b_mn=True, out_dtype=bf)),
        ("fc2 wgrad", D, 4 * D, M, lambda: ops.gemm(dyd, x4, a_mn=True, b_mn=True, out=g_fc2)),
        ("fc1 wgrad", 4 * D, D, M, lambda: ops.gemm(dy4, x, a_mn=True, b_mn=True, out=g_fc1)),
        ("proj wgrad", D, D, M, lambda: ops.gemm(dyd, x, a_mn=True, b_mn=True, out=g_proj)),
        ("qkv wgrad", 3 * D, D, M, lambda: ops.gemm(dy3, x, a_mn=True, b_mn=True, out=g_qkv)),
    ]
    # the plain contraction of each case as cuBLAS sees it (bf16 out, no bias / epilogue / accumulation)
    cublas = {
        "qkv fwd   +bias": lambda: torch.matmul(x, w_qkv.t()), "proj fwd  gate+res": lambda: torch.matmul(x, w_proj.t()),
        "fc1 fwd   gelu": lambda: torch.matmul(x, w_fc1.t()), "fc2 fwd   gate+res": lambda: torch.matmul(x4, w_fc2.t()),
        "fc2 dgrad dgelu": lambda: torch.matmul(dyd, w_fc2), "fc1 dgrad": lambda: torch.matmul(dy4, w_fc1),
        "proj dgrad": lambda: torch.matmul(dyd, w_proj), "qkv dgrad": lambda: torch.matmul(dy3, w_qkv),
        "fc2 wgrad": lambda: torch.matmul(dyd.t(), x4), "fc1 wgrad": lambda: torch.matmul(dy4.t(), x),
        "proj wgrad": lambda: torch.matmul(dyd.t(), x), "qkv wgrad": lambda: torch.matmul(dy3.t(), x),
    }

    def median_us(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(args.iters):
            if not args.no_flush:
                flush.sum()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) * 1e3)
        ts.sort()
        return ts[len(ts) // 2]

    total_us, total_fl, total_cb = 0.0, 0.0, 0.0
    for name, m, n, k, fn in cases:
        if args.only and args.only not in name:
            continue
        if args.cublas:
            cb = median_us(cublas[name])
            total_cb += cb
            print(f"{name:20s} cuBLAS plain matmul      {cb:8.1f} us  {2.0 * m * n * k / cb / 1e6:7.1f} TFLOP/s")
        for _ in range(3):
            fn()
        ts = []
        for _ in range(args.iters):
            if not args.no_flush:
                flush.sum()        # read-only L2 flush: leaves clean lines, no write-back competing with the timed kernel
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) * 1e3)
        ts.sort()
        us = ts[len(ts) // 2]
        fl = 2.0 * m * n * k
        total_us += us
        total_fl += fl
        print(f"{name:20s} M={m:5d} N={n:5d} K={k:5d}  {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s")
    print(f"block total {total_us:8.1f} us  {total_fl / total_us / 1e6:7.1f} TFLOP/s")
    if args.cublas and total_cb:
        print(f"cuBLAS plain total {total_cb:8.1f} us  {total_fl / total_cb / 1e6:7.1f} TFLOP/s")


if __name__ == "__main__":
    main()
