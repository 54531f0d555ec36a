"""Debug: clock64 timeline of CTA 0 of the single-pass attention backward (build: make -C reed_b200/csrc EXTRA=-DREED_ATTN_TRACE
BUILD=build_dbg OUT=../libreed_sm100_dbg.so; run with REED_LIB=.../libreed_sm100_dbg.so).
Tags - ew0/ew1: 1 wait scores, 2 scores ready, 3 P^T arrived, 4 dS^T tiles free, 5 dS^T arrived;  mma: 1/2 P^T half 0/1 ready,
3 dS^T ready, 4 dQ may be issued, 5 step issued;  drain: 1 wait dQ, 2 dQ ready, 3 message ready, 4 TMEM drained, 6 item done;
tma: 1 item start, 2 Q/dO slot free."""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from reed_b200 import _cabi, ops
lib = _cabi.load()
B, T, H, hd = 32, int(sys.argv[1]) if len(sys.argv) > 1 else 256, 16, 72
qkv = torch.randn(B * T, 3 * H * hd, device="cuda").bfloat16()
d_o = torch.randn(B * T, H * hd, device="cuda").bfloat16()
o, lse = ops.attention_fwd(qkv, B, T, H, hd)
for _ in range(2):
    ops.attention_bwd(qkv, o, d_o, lse, B, T, H, hd)
torch.cuda.synchronize()
buf = np.zeros((5, 2048), dtype=np.uint64)
lib.reed_debug_trace.argtypes = [ctypes.c_void_p]
assert lib.reed_debug_trace(buf.ctypes.data) == 0
names = {0: "ew0", 1: "ew1", 2: "mma", 3: "drain", 4: "tma"}
ev = []
for r in range(5):
    for v in buf[r][:64]:
        v = int(v)
        if v == 0:
            break
        ev.append((v & ((1 << 56) - 1), names[r], v >> 56))
ev.sort()
t0 = ev[0][0]
cols = {"tma": 0, "mma": 1, "ew0": 2, "ew1": 3, "drain": 4}
print("   clock " + "".join(f"{n:>8s}" for n in cols))
for t, who, tag in ev:
    if t - t0 > 60000:
        break
    print(f"{t - t0:8d} " + "        " * cols[who] + f"{tag:8d}")
