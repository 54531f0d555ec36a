"""Runs torch.matmul (cuBLAS) once on the 12 contractions of one SiT-XL/2 block, for an ncu launch list that shows which
kernels / tile shapes / cluster sizes the library picks (a yardstick for the planner, not a product path).

    ncu --metrics gpu__time_duration.sum,launch__cluster_size --clock-control none --csv \
        --log-file gpurun_out/cublas_kernels.csv python profiles/cublas_kernels.py
"""
import torch

D, M = 1152, 8192
bf = torch.bfloat16
r = lambda *s: (torch.randn(*s, device="cuda") * 0.05).to(bf)
x, x4 = r(M, D), r(M, 4 * D)
w_qkv, w_proj, w_fc1, w_fc2 = r(3 * D, D), r(D, D), r(4 * D, D), r(D, 4 * D)
dy3, dyd, dy4 = r(M, 3 * D), r(M, D), r(M, 4 * D)
cases = [
    ("qkv fwd", lambda: torch.matmul(x, w_qkv.t())), ("proj fwd", lambda: torch.matmul(x, w_proj.t())),
    ("fc1 fwd", lambda: torch.matmul(x, w_fc1.t())), ("fc2 fwd", lambda: torch.matmul(x4, w_fc2.t())),
    ("fc2 dgrad", lambda: torch.matmul(dyd, w_fc2)), ("fc1 dgrad", lambda: torch.matmul(dy4, w_fc1)),
    ("proj dgrad", lambda: torch.matmul(dyd, w_proj)), ("qkv dgrad", lambda: torch.matmul(dy3, w_qkv)),
    ("fc2 wgrad", lambda: torch.matmul(dyd.t(), x4)), ("fc1 wgrad", lambda: torch.matmul(dy4.t(), x)),
    ("proj wgrad", lambda: torch.matmul(dyd.t(), x)), ("qkv wgrad", lambda: torch.matmul(dy3.t(), x)),
]
torch.cuda.synchronize()
for name, fn in cases:
    torch.cuda.nvtx.range_push(name)
    fn()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
print("done")
