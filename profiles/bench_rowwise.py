"""Times the HBM-bound row kernels of one SiT block at the bench shape and reports achieved GB/s (algorithmic bytes)."""
import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from reed_b200 import _cabi, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--D", type=int, default=1152)
    ap.add_argument("--B", type=int, default=32)
    ap.add_argument("--T", type=int, default=256)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    _cabi.load()
    dev, bf = "cuda", torch.bfloat16
    B, T, D = a.B, a.T, a.D
    M = B * T
    x = torch.randn(M, D, device=dev)
    mod = torch.randn(B, 6 * D, device=dev) * 0.1
    dmod = torch.zeros_like(mod)
    sl = lambda t, i: t[:, i * D:(i + 1) * D]
    dout = torch.randn(M, D, device=dev).to(bf)
    dres = torch.randn(M, D, device=dev)
    y = torch.randn(M, D, device=dev).to(bf)
    dbias = torch.zeros(D, device=dev)
    out, mean, rstd = ops.ln_modulate_fwd(x, sl(mod, 0), sl(mod, 1), T, bf)
    flush = torch.zeros(64 << 20, dtype=torch.int32, device=dev)
    n = M * D
    cases = [
        ("ln_modulate_fwd", 6 * n, lambda: ops.ln_modulate_fwd(x, sl(mod, 0), sl(mod, 1), T, bf)),
        ("ln_modulate_bwd", 14 * n, lambda: ops.ln_modulate_bwd(dout, x, mean, rstd, sl(mod, 1), T, dres, sl(dmod, 0), sl(dmod, 1))),
        ("gate_bwd", 8 * n, lambda: ops.gate_bwd(dres, y, sl(mod, 2), T, sl(dmod, 2), dbias)),
        ("ln_modulate_gate_bwd", 18 * n, lambda: ops.ln_modulate_gate_bwd(dout, x, mean, rstd, sl(mod, 1), T, dres, sl(dmod, 0),
                                                                            sl(dmod, 1), y, sl(mod, 2), sl(dmod, 2), dbias)),
    ]
    for name, nbytes, fn in cases:
        for _ in range(3):
            fn()
        ts = []
        for _ in range(a.iters):
            flush.sum()            # read-only L2 flush: leaves clean lines, no write-back competing with the timed kernel
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) * 1e3)
        ts.sort()
        us = ts[len(ts) // 2]
        print(f"{name:24s} {us:8.1f} us  {nbytes / us / 1e3:8.1f} GB/s  ({nbytes / 1e6:.1f} MB)")


if __name__ == "__main__":
    main()
