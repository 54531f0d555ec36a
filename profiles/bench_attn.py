"""Times the attention kernels (tcgen05 vs mma.sync) on the SiT-XL/2 shape with CUDA events, cold L2 between launches.

    python profiles/bench_attn.py [--batch 32] [--tokens 256] [--heads 16] [--hd 72]
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
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--tokens", type=int, default=256)
    ap.add_argument("--heads", type=int, default=16)
    ap.add_argument("--hd", type=int, default=72)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    _cabi.load()
    B, T, H, hd = args.batch, args.tokens, args.heads, args.hd
    dev = "cuda"
    qkv = torch.randn(B * T, 3 * H * hd, device=dev).bfloat16()
    d_o = torch.randn(B * T, H * hd, device=dev).bfloat16()
    flush = torch.zeros(64 << 20, dtype=torch.int32, device=dev)
    fl_fwd = 4.0 * B * H * T * T * hd
    for name, backend in (("mma.sync", 3), ("tcgen05", 4)):
        ops.set_backends(attention=backend)
        try:
            o, lse = ops.attention_fwd(qkv, B, T, H, hd)
        except Exception as e:  # unsupported shape for this backend
            print(name, "unsupported:", str(e)[:80])
            continue
        res = {}
        for what, fn in (("fwd", lambda: ops.attention_fwd(qkv, B, T, H, hd)),
                         ("bwd", lambda: ops.attention_bwd(qkv, o, d_o, lse, B, T, H, hd))):
            for _ in range(3):
                fn()
            ts = []
            for _ in range(args.iters):
                flush.sum()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                fn()
                e.record()
                torch.cuda.synchronize()
                ts.append(s.elapsed_time(e) * 1e3)
            ts.sort()
            res[what] = ts[len(ts) // 2]
        print(f"{name:9s} B={B} T={T} H={H} hd={hd}: fwd {res['fwd']:7.1f} us ({fl_fwd / res['fwd'] / 1e6:6.1f} TFLOP/s)  "
              f"bwd {res['bwd']:7.1f} us ({2.5 * fl_fwd / res['bwd'] / 1e6:6.1f} TFLOP/s)")
    ops.set_backends()


if __name__ == "__main__":
    main()
