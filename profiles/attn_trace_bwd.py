"""Debug: clock64 trace of CTA 0 of the attention dK/dV kernel (build with make EXTRA=-DREED_ATTN_TRACE)."""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from reed_b200 import _cabi, ops
lib = _cabi.load()
B, T, H, hd = 32, 256, 16, 72
qkv = torch.randn(B * T, 3 * H * hd, device="cuda").bfloat16()
d_o = torch.randn(B * T, H * hd, device="cuda").bfloat16()
ops.set_backends(attention=4)
o, lse = ops.attention_fwd(qkv, B, T, H, hd)
for _ in range(2):
    ops.attention_bwd(qkv, o, d_o, lse, B, T, H, hd)
torch.cuda.synchronize()
buf = np.zeros((4, 2048), dtype=np.uint64)
lib.reed_debug_trace.argtypes = [ctypes.c_void_p]
assert lib.reed_debug_trace(buf.ctypes.data) == 0
names = {0: "grp0", 1: "grp1", 2: "mma", 3: "tma"}
ev = []
for r in range(3):
    for v in buf[r][:12]:
        v = int(v)
        if v == 0:
            break
        ev.append((v & ((1 << 56) - 1), names[r], v >> 56))
ev.sort()
t0 = ev[0][0]
for t, who, tag in ev:
    print(f"{t - t0:8d} {who} {tag}")
