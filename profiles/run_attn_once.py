"""Launch the attention kernels a few times on one shape (the command ncu wraps):
    python profiles/run_attn_once.py [--backend 5] [--batch 32] [--tokens 256] [--heads 16] [--hd 72] [--bwd] [--reps 3]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reed_b200 import _cabi, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--backend", type=int, default=5)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--tokens", type=int, default=256)
ap.add_argument("--heads", type=int, default=16)
ap.add_argument("--hd", type=int, default=72)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--bwd", action="store_true")
a = ap.parse_args()
_cabi.load()
B, T, H, hd = a.batch, a.tokens, a.heads, a.hd
qkv = torch.randn(B * T, 3 * H * hd, device="cuda").bfloat16()
d_o = torch.randn(B * T, H * hd, device="cuda").bfloat16()
ops.set_backends(attention=a.backend)
for _ in range(a.reps):
    o, lse = ops.attention_fwd(qkv, B, T, H, hd)
    if a.bwd:
        ops.attention_bwd(qkv, o, d_o, lse, B, T, H, hd)
torch.cuda.synchronize()
print("done")
