"""Sampling throughput of BASELINE configs[4] (SiT-XL/2, 4x64x64 latents = 1024 tokens, Euler-Maruyama SDE):
eager launches vs generate.GraphedSiT (CUDA-graph replay of the model evaluation).  Not yet run on a B200 (written
after round 1's GPU minutes were spent) - first thing to run in round 2:

    python profiles/bench_sampler.py [--size 64] [--batch 8] [--steps 50] [--cfg 1.8] [--precision bf16]

Prints model evaluations/s and images/s for both paths, and the FLOP rate from SURVEY 8(d)
(1049 GFLOP per evaluation and sample at T=1024, 241.4 at T=256; doubled inside the CFG window)."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="SiT-XL/2")
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--cfg", type=float, default=1.0)
    ap.add_argument("--mode", default="sde", choices=["sde", "ode"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    args = ap.parse_args()
    from oracle.sit_oracle import flops_per_image, zoo_spec
    from reed_b200 import ops
    from reed_b200.image.generate import GraphedSiT
    from reed_b200.image.models.sit import SiT_models
    from reed_b200.image.samplers import euler_maruyama_sampler, euler_sampler

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = SiT_models[args.model](input_size=args.size, num_classes=1000, use_cfg=True, z_dims=[768], z_types=["i"],
                                   encoder_depth=8, fused_attn=True, qk_norm=False)
    with torch.no_grad():
        g = torch.Generator().manual_seed(1)
        for p in model.parameters():
            if p.requires_grad and float(p.abs().max()) == 0.0:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    model = model.to(dev).eval()
    model.reed_precision = args.precision
    spec = zoo_spec(args.model, input_size=args.size, z_dims=[], z_types=[], encoder_depth=8, encoder_depth_text=None)   # no projector at inference
    fwd_flops = flops_per_image(spec, train=False)
    sampler = euler_maruyama_sampler if args.mode == "sde" else euler_sampler
    z = torch.randn(args.batch, 4, args.size, args.size, device=dev)
    y = torch.randint(0, 1000, (args.batch,), device=dev)
    kw = dict(num_steps=args.steps, cfg_scale=args.cfg)

    def run(runner, label):
        sampler(runner, z, y, num_steps=3, cfg_scale=args.cfg)            # warm-up / capture
        torch.cuda.synchronize()
        l0 = ops.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out = sampler(runner, z, y, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms, wall = e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3
        rows = args.batch * (2 if args.cfg > 1.0 else 1)
        print(f"{label:8s} {ms:9.1f} ms device ({wall:9.1f} ms wall)  {args.steps / ms * 1e3:8.1f} evals/s  "
              f"{args.batch / ms * 1e3:7.2f} img/s  {rows * args.steps * fwd_flops / ms / 1e9:7.1f} TFLOP/s  "
              f"launches from Python: {ops.launch_count - l0}")
        return out

    a = run(model, "eager")
    b = run(GraphedSiT(model), "graphed")
    print("max |eager - graphed| =", float((a - b).abs().max()), "(different noise draws if sde: compare with --mode ode)")


if __name__ == "__main__":
    main()
