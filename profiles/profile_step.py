"""One train step bracketed by cudaProfilerStart/Stop, for `ncu --profile-from-start off`.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python profiles/profile_step.py --config xl2
    ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -c 3 \
        -o gpurun_out/prof python profiles/profile_step.py --config xl2

Numbers printed by a run under ncu are never bench values.
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="xl2", choices=sorted(bench.CONFIGS))
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--batch", type=int, default=None)
    args = ap.parse_args()
    cfg = dict(bench.CONFIGS[args.config])
    if args.batch:
        cfg["local_batch"] = args.batch
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    trainer, spec = bench.build_trainer(cfg, dev)
    batches = bench.make_batches(cfg, spec, dev, 2)
    for i in range(args.warmup):
        trainer.train_step(*batches[i % 2][1])
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    for i in range(args.steps):
        trainer.train_step(*batches[i % 2][1])
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("profiled", args.steps, "step(s)")


if __name__ == "__main__":
    main()
