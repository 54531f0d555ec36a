"""Host (Python + ctypes launch) time per train step vs device time: is the step launch-bound?"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from reed_b200 import ops  # noqa: E402

cfg = dict(bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "xl2"])
dev = torch.device("cuda", 0)
trainer, spec = bench.build_trainer(cfg, dev)
batches = bench.make_batches(cfg, spec, dev, 2)
for i in range(3):
    trainer.train_step(*batches[i % 2][1])
torch.cuda.synchronize()
n = 5
l0 = ops.launch_count
t0 = time.perf_counter()
for i in range(n):
    trainer.train_step(*batches[i % 2][1])
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3 * (t1 - t0) / n:.2f} ms/step, wall {1e3 * (t2 - t0) / n:.2f} ms/step, reed launches/step {(ops.launch_count - l0) / n:.0f}")
# phases
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
trainer.train_step(*batches[0][1])
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
