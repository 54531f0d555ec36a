"""Correctness + timing of the flash-style tcgen05 attention kernels (backend 5) against an fp32 torch reference, on shapes
that cover both head dims, one / two / many key blocks and rows whose running maximum has to move (the lazy-rescale path).

    python profiles/check_attn.py [--bwd] [--time]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from reed_b200 import _cabi, ops  # noqa: E402

FA = 5


def reference(qkv, d_o, B, T, H, hd):
    x = qkv.float().view(B, T, 3, H, hd).permute(2, 0, 3, 1, 4).contiguous().requires_grad_(True)
    q, k, v = x[0], x[1], x[2]
    s = (q @ k.transpose(-1, -2)) * hd ** -0.5
    lse = torch.logsumexp(s, dim=-1)
    o = torch.softmax(s, dim=-1) @ v                      # [B, H, T, hd]
    o2 = o.transpose(1, 2).reshape(B * T, H * hd)
    dqkv = None
    if d_o is not None:
        (gx,) = torch.autograd.grad(o2, x, d_o.float())
        dqkv = gx.permute(1, 3, 0, 2, 4).reshape(B * T, 3 * H * hd)
    return o2.detach(), lse.detach(), dqkv


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bwd", action="store_true")
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="time with a warm L2 (no 256 MB flush between launches)")
    args = ap.parse_args()
    _cabi.load()
    dev = "cuda"
    cases = [  # B, T, H, hd, q gain, k gain
        (2, 256, 3, 72, 1.0, 1.0), (2, 256, 2, 64, 1.0, 1.0), (3, 128, 2, 72, 1.0, 1.0), (2, 512, 2, 72, 1.0, 1.0),
        (1, 1024, 2, 72, 1.0, 1.0), (2, 1024, 3, 64, 1.0, 1.0), (2, 256, 2, 72, 6.0, 3.0), (1, 1024, 2, 72, 4.0, 4.0),
        (20, 256, 16, 72, 1.0, 1.0),
    ]
    ok = True
    for B, T, H, hd, gq, gk in cases:
        g = torch.Generator(device=dev).manual_seed(B * 1000 + T + hd)
        qkv = torch.randn(B * T, 3, H * hd, device=dev, generator=g)
        qkv[:, 0] *= gq
        # keys whose scores grow along the sequence: later key blocks raise the row maximum (rescale path)
        qkv[:, 1] *= gk * torch.linspace(0.2, 1.0, T, device=dev).repeat(B)[:, None]
        qkv = qkv.view(B * T, 3 * H * hd).bfloat16()
        d_o = torch.randn(B * T, H * hd, device=dev, generator=g).bfloat16()
        o_ref, lse_ref, dqkv_ref = reference(qkv, d_o if args.bwd else None, B, T, H, hd)
        ops.set_backends(attention=FA)
        o, lse = ops.attention_fwd(qkv, B, T, H, hd)
        torch.cuda.synchronize()
        eo = float((o.float() - o_ref).abs().max() / o_ref.abs().max())
        el = float((lse - lse_ref).abs().max())
        line = f"B={B:2d} T={T:4d} H={H:2d} hd={hd} gain {gq}/{gk}: o rel.err {eo:.2e}  lse abs.err {el:.2e}"
        good = eo < 2e-2 and el < 2e-2
        if args.bwd:
            dqkv = ops.attention_bwd(qkv, o, d_o, lse, B, T, H, hd)
            torch.cuda.synchronize()
            parts = dqkv.float().view(B * T, 3, H * hd), dqkv_ref.view(B * T, 3, H * hd)
            errs = [float((parts[0][:, i] - parts[1][:, i]).abs().max() / parts[1][:, i].abs().max()) for i in range(3)]
            line += "  dq/dk/dv rel.err " + "/".join(f"{e:.2e}" for e in errs)
            good = good and max(errs) < 3e-2
        print(line, "" if good else "   <-- MISMATCH")
        ok = ok and good
    ops.set_backends()
    print("CHECK OK" if ok else "CHECK FAILED")
    if args.time:
        flush = torch.zeros(64 << 20, dtype=torch.int32, device=dev)
        for B, T, H, hd in ((32, 256, 16, 72), (8, 1024, 16, 72), (32, 256, 12, 64)):
            qkv = torch.randn(B * T, 3 * H * hd, device=dev).bfloat16()
            d_o = torch.randn(B * T, H * hd, device=dev).bfloat16()
            fl = 4.0 * B * H * T * T * hd
            for name, backend in (("round-1", 0), ("fa", FA)):
                ops.set_backends(attention=backend)
                o, lse = ops.attention_fwd(qkv, B, T, H, hd)
                res = {}
                todo = [("fwd", lambda: ops.attention_fwd(qkv, B, T, H, hd))]
                if args.bwd:
                    todo.append(("bwd", lambda: ops.attention_bwd(qkv, o, d_o, lse, B, T, H, hd)))
                for what, fn in todo:
                    for _ in range(3):
                        fn()
                    ts = []
                    for _ in range(15):
                        if not args.no_flush:
                            flush.sum()
                        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        s.record()
                        fn()
                        e.record()
                        torch.cuda.synchronize()
                        ts.append(s.elapsed_time(e) * 1e3)
                    ts.sort()
                    res[what] = ts[len(ts) // 2]
                msg = f"{name:8s} B={B} T={T} H={H} hd={hd}: fwd {res['fwd']:7.1f} us ({fl / res['fwd'] / 1e6:6.1f} TFLOP/s)"
                if "bwd" in res:
                    msg += f"  bwd {res['bwd']:7.1f} us ({2.5 * fl / res['bwd'] / 1e6:6.1f} TFLOP/s)"
                print(msg)
        ops.set_backends()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
