"""Benchmark of the REED image hot path: SiT train step (SILoss fwd + bwd + grad all-reduce + clip + AdamW + EMA).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config xl2|b2|xl2_mm|xl2_512|s2] [--impl reference]

N>1 is launched by torchrun (one rank per GPU, NCCL).  Prints ONE JSON line on rank 0 (contract in the task spec):
value = whole-job images/s with inputs resident in HBM, device-timed (CUDA events, max over ranks);
e2e   = the same through the public API with pinned-host inputs copied H2D and the loss read back D2H every step;
roofline = achieved FLOP/s of the dominant kernel (the tcgen05 GEMM) from CUDA events around its launches;
cpu_baseline = the CPU oracle port of the same step on this box's host cores, on a bounded sample.
``--impl reference`` times the reference's own algorithm on CPU (the oracle port: the reference is pure Python and
cannot travel to the GPU box) and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TRAFFIC_FILE = "r02_gemm_traffic.json"     # ncu DRAM-traffic capture of the GEMM launches of one xl2 step (profiles/gemm_traffic.py)

CONFIGS = {
    # name: (zoo name, input_size, local batch, z_dims, z_types, enc_depth, enc_depth_text, enc_names, weights, cpu batch)
    "xl2": dict(model="SiT-XL/2", input_size=32, local_batch=32, z_dims=[768], z_types=["i"], encoder_depth=8,
                encoder_depth_text=None, enc_names=["dinov2"], loss_weights={"dinov2": 1.0}, cpu_batch=4,
                workload="SiT-XL/2 REED train step, 4x32x32 latents (T=256), DINOv2-B 768-d targets, "
                         "local batch 32 per GPU = BASELINE configs[2] (global 256 on 8 GPUs)"),
    "b2": dict(model="SiT-B/2", input_size=32, local_batch=256, z_dims=[768], z_types=["i"], encoder_depth=8,
               encoder_depth_text=None, enc_names=["dinov2"], loss_weights={"dinov2": 1.0}, cpu_batch=8,
               workload="SiT-B/2 REED train step, batch 256 on one GPU = BASELINE configs[1]"),
    "xl2_mm": dict(model="SiT-XL/2", input_size=32, local_batch=32, z_dims=[768, 3584], z_types=["i", "t"],
                   encoder_depth=8, encoder_depth_text=16, enc_names=["dinov2", "text_embeds_qwenvl_7b_layer_15"],
                   loss_weights={"dinov2": 1.0, "text_embeds_qwenvl_7b_layer_15": 0.5}, cpu_batch=4,
                   workload="SiT-XL/2 multimodal REED (image + caption embedding heads) = BASELINE configs[3]"),
    "xl2_512": dict(model="SiT-XL/2", input_size=64, local_batch=8, z_dims=[768], z_types=["i"], encoder_depth=8,
                    encoder_depth_text=None, enc_names=["dinov2"], loss_weights={"dinov2": 1.0}, cpu_batch=2,
                    workload="SiT-XL/2 ImageNet-512 train step, 4x64x64 latents (T=1024) = BASELINE configs[4] (train)"),
    "s2": dict(model="SiT-S/2", input_size=32, local_batch=64, z_dims=[768], z_types=["i"], encoder_depth=8,
               encoder_depth_text=None, enc_names=["dinov2"], loss_weights={"dinov2": 1.0}, cpu_batch=4,
               extra=dict(decoder_hidden_size=384), workload="SiT-S/2 REED train step = BASELINE configs[0] shapes"),
}


def spec_for(cfg):
    from oracle.sit_oracle import zoo_spec
    return zoo_spec(cfg["model"], input_size=cfg["input_size"], z_dims=cfg["z_dims"], z_types=cfg["z_types"],
                    encoder_depth=cfg["encoder_depth"], encoder_depth_text=cfg["encoder_depth_text"],
                    **cfg.get("extra", {}))


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference step (also the cpu_baseline leg of the GPU arm)
# ---------------------------------------------------------------------------------------------------------

def cpu_reference_steps(cfg, steps, warmup, batch=None, budget_s=None):
    """Times clip+AdamW+EMA train steps of the oracle (torch CPU, fp32, all host threads). Returns (img/s, s/step, info).
    budget_s bounds the sample: no new step is started once the wall clock since the first timed step exceeds it (at least one
    timed step always runs); info["steps"] is the number actually timed."""
    from oracle import loss_oracle, sit_oracle, train_oracle
    from oracle.fixtures import random_batch, random_state
    torch.set_num_threads(os.cpu_count() or 1)
    spec = spec_for(cfg)
    batch = batch or cfg["cpu_batch"]
    sd = random_state(spec, 0)
    params = {k: v.clone() for k, v in sd.items()}
    ema = {k: v.clone() for k, v in sd.items()}
    m1 = {k: torch.zeros_like(v) for k, v in sd.items()}
    m2 = {k: torch.zeros_like(v) for k, v in sd.items()}
    times = []
    started = None
    for it in range(warmup + steps):
        if it >= warmup:
            started = started if started is not None else time.perf_counter()
            if budget_s is not None and times and time.perf_counter() - started > budget_s:
                break
        data = random_batch(spec, batch, 100 + it)
        t0 = time.perf_counter()
        leaves = {k: p.detach().requires_grad_(k != "pos_embed") for k, p in params.items()}
        model = sit_oracle.as_model(leaves, spec, training=True)
        out = loss_oracle.si_loss(model, data["x"], loss_oracle.draw_time(batch), torch.randn_like(data["x"]),
                                  data["zs"], enc_names=cfg["enc_names"], loss_weights=cfg["loss_weights"],
                                  model_kwargs=dict(y=data["y"]))
        train_oracle.mix_losses(out).backward()
        grads = {k: l.grad for k, l in leaves.items() if l.grad is not None}
        train_oracle.adamw_ema_step(params, grads, m1, m2, ema, it + 1)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    per_step = sum(times) / len(times)
    info = dict(kind="port", cores=torch.get_num_threads(), steps=len(times),
                sample=f"{len(times)} timed step(s) of the CPU oracle port (reference algorithm, torch fp32), batch {batch}, "
                       f"after {warmup} warm-up")
    return batch / per_step, per_step, info


def run_reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the driver passes the GPU arm's --steps/--warmup: honour them up to a wall-clock budget (a CPU step of SiT-XL/2 at
    # batch 4 takes seconds), and report the number of steps actually timed
    warmup = max(1, min(args.warmup, 2))
    ips, per_step, info = cpu_reference_steps(cfg, max(1, min(args.steps, 20)), warmup, budget_s=90.0)
    steps = info["steps"]
    info["value"] = ips
    info["unit"] = "images/s"
    line = {
        "impl": "reference", "metric": "SiT REED train images/sec", "value": ips, "unit": "images/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["workload"] + f" - CPU reference arm: the same step at batch {cfg['cpu_batch']} per step "
                   "(images/s is batch-size independent on the CPU: one image's FLOPs already fill every core)",
                   "cpu_batch": cfg["cpu_batch"], "gpu_local_batch": cfg["local_batch"]},
        "cpu_baseline": info,
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------

class ClockSampler(threading.Thread):
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.samples[0][1]),
                "power_w_max": max(float(s[2]) for s in self.samples), "reasons": reasons, "samples": len(self.samples)}


def build_trainer(cfg, dev, precision="bf16"):
    """Model (reference init, zero-init layers perturbed), SILoss and the flat-buffer trainer for one config."""
    from reed_b200.image.loss import SILoss
    from reed_b200.image.models.sit import SiT_models
    from reed_b200.image.trainer import ReedTrainer
    spec = spec_for(cfg)
    torch.manual_seed(0)          # same init on every rank; the trainer also broadcasts rank 0's weights
    model = SiT_models[cfg["model"]](input_size=cfg["input_size"], num_classes=1000, use_cfg=True, z_dims=cfg["z_dims"],
                                     z_types=cfg["z_types"], encoder_depth=cfg["encoder_depth"],
                                     encoder_depth_text=cfg["encoder_depth_text"], fused_attn=True, qk_norm=False,
                                     **cfg.get("extra", {}))
    with torch.no_grad():          # un-zero the adaLN / output layers so no kernel sees degenerate all-zero operands
        g = torch.Generator().manual_seed(1)
        for p in model.parameters():
            if p.requires_grad and float(p.abs().max()) == 0.0:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    model = model.to(dev).train()
    loss_fn = SILoss(enc_names=cfg["enc_names"], loss_weights=cfg["loss_weights"])
    # With more than one rank the trainer shards the optimizer and exchanges gradients / operands with its own NVSwitch
    # multicast kernels (profiles/r02_sharded_check.txt).  A/B knobs: REED_NVLS=0 -> NCCL reduce-scatter / all-gather,
    # REED_SHARD_OPT=0 -> replicated optimizer behind an NCCL all-reduce (round 1's data path).
    shard = os.environ.get("REED_SHARD_OPT", "1") != "0"
    return ReedTrainer(model, loss_fn, precision=precision, comm_sms=int(os.environ.get("REED_COMM_SMS", "16")),
                       shard_optimizer=shard if shard is False else None,
                       nvls=False if (not shard or os.environ.get("REED_NVLS", "1") == "0") else None), spec


def make_batches(cfg, spec, dev, n_buf):
    """n_buf synthetic batches: [(pinned host tensors, device-resident copies)]."""
    B, S, T = cfg["local_batch"], cfg["input_size"], spec.tokens
    out = []
    for _ in range(n_buf):
        x = torch.randn(B, 4, S, S).pin_memory()
        y = torch.randint(0, 1000, (B,)).pin_memory()
        zs = [(torch.randn(B, T, z) if k == "i" else torch.randn(B, z)).bfloat16().pin_memory()
              for z, k in zip(cfg["z_dims"], cfg["z_types"])]
        out.append(((x, y, zs), (x.to(dev), y.to(dev), [z.to(dev) for z in zs])))
    return out


def run_gpu_arm(args, cfg):
    import torch.distributed as dist
    from reed_b200 import _cabi, ops
    from oracle.sit_oracle import flops_per_image

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # the all-reduce kernels share the GPU with the backward GEMMs: cap their CTAs to the SMs the trainer reserves
        os.environ.setdefault("NCCL_MAX_CTAS", os.environ.get("REED_COMM_SMS", "16"))
        dist.init_process_group("nccl", device_id=dev)
    _cabi.load()
    ops.device_check()

    trainer, spec = build_trainer(cfg, dev)
    B, S, T = cfg["local_batch"], cfg["input_size"], spec.tokens
    n_buf = 4
    torch.manual_seed(1234 + rank)     # per-rank data/RNG streams (train.py:176)
    batches = make_batches(cfg, spec, dev, n_buf)
    host = [h for h, _ in batches]
    resident = [r for _, r in batches]
    h2d_bytes = sum(t.numel() * t.element_size() for t in (host[0][0], host[0][1], *host[0][2]))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    # ---- device-resident steps --------------------------------------------------------------------------------
    # The whole step (SILoss forward, backward with the per-block all-reduces, clip, AdamW, EMA) is captured once as a
    # CUDA graph and replayed; --eager launches the same kernels one by one from Python instead.
    graphed = not args.eager
    launches_per_step = None
    if graphed:
        l0 = ops.launch_count
        trainer.capture(*resident[0], warmup=2)
        launches_per_step = (ops.launch_count - l0) // 3     # 2 warm-up steps + the recorded one
    step_fn = trainer.train_step_graphed if graphed else trainer.train_step

    def resident_step(i):
        x, y, zs = resident[i % n_buf]
        step_fn(x, y, zs)

    n_warm = max(10, args.warmup)      # graph replays are cheap: let clocks and power settle before the timed region
    for i in range(n_warm):
        resident_step(i)
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ops.launch_count
    ms = timed(resident_step, args.steps)
    launches = launches_per_step * args.steps if graphed else ops.launch_count - launches0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    value = world * B * args.steps / (ms / 1e3)

    # ---- end to end: pinned host -> device every step, loss read back every step ---------------------------------
    copy_stream = torch.cuda.Stream()
    staged = {}

    def stage(i):
        x, y, zs = host[i % n_buf]
        with torch.cuda.stream(copy_stream):
            staged[i] = (x.to(dev, non_blocking=True), y.to(dev, non_blocking=True),
                         [z.to(dev, non_blocking=True) for z in zs], torch.cuda.Event())
            staged[i][3].record(copy_stream)

    losses = []

    def e2e_step(i):
        if i not in staged:
            stage(i)
        stage(i + 1)                                     # prefetch the next batch while this one computes
        x, y, zs, ev = staged.pop(i)
        torch.cuda.current_stream().wait_event(ev)
        loss, _ = step_fn(x, y, zs)
        losses.append(float(loss))                       # D2H read of the step's loss (host sync, like train.py:456-466)

    for i in range(2):
        e2e_step(i)
    staged.clear()
    e2e_ms = timed(e2e_step, args.steps)
    e2e_value = world * B * args.steps / (e2e_ms / 1e3)

    # ---- dominant kernel: every tcgen05 GEMM launch of one step, re-issued back to back inside a CUDA graph and timed
    # with CUDA events on the launching stream (an eager step is host-bound, so events around single launches would
    # time Python, not the kernel).  achieved = sum of algorithmic 2*M*N*K / summed launch time.
    roof = None
    records = []          # (algorithmic flops, replay closure) of every GEMM call that may take the tcgen05 path
    real_gemm, real_wgrad_bias = ops.gemm, ops.wgrad_bias

    def recording(a, b, **kw):
        before = ops.tcgen05_gemm_launches()
        out = real_gemm(a, b, **kw)
        if ops.tcgen05_gemm_launches() > before:          # the library took the tensor-core path for this call
            M, K = (a.shape[1], a.shape[0]) if kw.get("a_mn") else (a.shape[0], a.shape[1])
            N = b.shape[1] if kw.get("b_mn") else b.shape[0]
            kw2 = dict(kw)
            kw2["out"] = out
            records.append((2.0 * M * N * K, lambda a=a, b=b, kw2=kw2: real_gemm(a, b, **kw2)))
        return out

    def recording_wgrad_bias(dy2d, x_ext, k_in, dw, db, accumulate):
        before = ops.tcgen05_gemm_launches()
        real_wgrad_bias(dy2d, x_ext, k_in, dw, db, accumulate)
        if ops.tcgen05_gemm_launches() > before:
            tokens, n_out = dy2d.shape                     # dW[n_out, k_in] and db[n_out]: k_in + 1 useful output columns
            records.append((2.0 * tokens * n_out * (k_in + 1),
                            lambda: real_wgrad_bias(dy2d, x_ext, k_in, dw, db, accumulate)))

    real_grouped_fwd, real_grouped_dgrad = ops.gemm_grouped_fwd, ops.gemm_grouped_dgrad

    def recording_grouped_fwd(a, weights, bias_all, out):          # adaLN modulation of all blocks: one launch
        real_grouped_fwd(a, weights, bias_all, out)
        records.append((2.0 * a.shape[0] * out.shape[1] * a.shape[1], lambda: real_grouped_fwd(a, weights, bias_all, out)))
        return out

    def recording_grouped_dgrad(dys, weights, out, accumulate):
        real_grouped_dgrad(dys, weights, out, accumulate)
        records.append((2.0 * out.shape[0] * out.shape[1] * dys[0].shape[1] * len(dys),
                        lambda: real_grouped_dgrad(dys, weights, out, accumulate)))
        return out

    def eager_step(i):
        x, y, zs = resident[i % n_buf]
        trainer.train_step(x, y, zs)

    if world > 1:
        dist.barrier()
    ops.gemm, ops.wgrad_bias = recording, recording_wgrad_bias
    ops.gemm_grouped_fwd, ops.gemm_grouped_dgrad = recording_grouped_fwd, recording_grouped_dgrad
    tc_before = ops.tcgen05_gemm_launches()
    try:
        eager_step(0)                      # every rank runs it so the collectives stay matched
        torch.cuda.synchronize()
    finally:
        ops.gemm, ops.wgrad_bias = real_gemm, real_wgrad_bias
        ops.gemm_grouped_fwd, ops.gemm_grouped_dgrad = real_grouped_fwd, real_grouped_dgrad
    tc_launches_in_step = ops.tcgen05_gemm_launches() - tc_before
    assert len(records) == tc_launches_in_step, (len(records), tc_launches_in_step)   # the sample is the whole population
    if rank == 0 and records:
        reps = 3
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _, replay in records[:8]:
                replay()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _, replay in records:
                replay()
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        tms = e0.elapsed_time(e1) / reps
        flops = sum(r[0] for r in records)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("bf16_tflops_sustained") or 1400.0
        # DRAM bytes per GEMM launch, from the committed ncu capture of the same step on the same build (xl2 workload only;
        # profiles/gemm_traffic.py writes it and records the launch count it saw: it must equal launches_timed)
        traffic, traffic_launches = None, None
        try:
            if args.config == "xl2":
                tj = json.load(open(os.path.join(ROOT, "profiles", TRAFFIC_FILE)))
                traffic, traffic_launches = tj["dram_bytes_per_launch"], tj.get("launches")
        except Exception:
            pass
        achieved = flops / (tms * 1e-3) / 1e12 if tms > 0 else 0.0
        roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": f"profiles/{TRAFFIC_FILE} (ncu dram bytes, mean per launch over "
                f"{traffic_launches} launches)" if traffic is not None else None, "kernel": "gemm_tcgen05_kernel",
                "launches_timed": len(records), "tcgen05_launches_per_step": tc_launches_in_step,
                "avg_launch_us": tms * 1e3 / len(records),
                "method": "all tcgen05 GEMM launches of one train step replayed back to back in a CUDA graph, CUDA events",
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
                if peaks else "fallback 1400 (sustained, B200_PROFILING.md)",
                "gemm_share_of_step": tms / (ms / args.steps) if ms > 0 else None}
        del graph
    records.clear()

    if rank == 0:
        train_flops = flops_per_image(spec, train=True)
        peaks_burst = 1650.2
        try:
            peaks_burst = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]
        except Exception:
            pass
        steps_cpu, warm_cpu = 8, 1          # bounded sample: stops after ~15 s of timed CPU work
        cpu_ips, _, cpu_info = (cpu_reference_steps(cfg, steps_cpu, warm_cpu, budget_s=15.0)
                                if world == 1 and not args.skip_cpu else (None, None, None))
        line = {
            "metric": "SiT REED train images/sec", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": n_warm, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": cfg["workload"], "model": cfg["model"], "local_batch": B, "global_batch": B * world,
                       "tokens": T, "parallelism": f"dp{world}", "optimizer": ("sharded over ranks, NVLS multicast exchange" if trainer.nvls is not None else "sharded over ranks, NCCL")
                       if trainer.shard else "replicated", "l2": "per-step working set (activations, weights) far exceeds the 126 MB L2; "
                       f"{n_buf} rotating input batches", "precision": "bf16 GEMM operands, fp32 accumulate/residual/master weights",
                       "launch": "CUDA graph replay of the whole step" if graphed else "eager (one launch per kernel from Python)"},
            "model_flops_per_image": train_flops,
            "tensor_peak_frac_of_measured_burst": value / world * train_flops / 1e12 / peaks_burst,
            "tensor_peak_frac_of_nominal_2250": value / world * train_flops / 1e12 / 2250.0,
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches,
            "clocks": sampler.summary(),
            "roofline": roof,
            "last_loss": losses[-1] if losses else None,
        }
        if cpu_info is not None:
            cpu_info.update(value=cpu_ips, unit="images/s")
            line["cpu_baseline"] = cpu_info
        print(json.dumps(line), flush=True)
    if world > 1:
        # every rank has finished its collectives; NCCL teardown with a captured graph still alive can block, and
        # the JSON line is already out, so leave without the communicator destructor
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    if os.environ.get("REED_BENCH_WATCHDOG"):       # debugging aid: dump every thread's Python stack and exit after N s
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["REED_BENCH_WATCHDOG"]), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="xl2", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="reed", choices=["reed", "reference"])
    ap.add_argument("--skip-cpu", action="store_true", help="omit the cpu_baseline leg")
    ap.add_argument("--eager", action="store_true", help="launch kernels from Python instead of replaying the CUDA graph")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference_arm(args, cfg)
    else:
        run_gpu_arm(args, cfg)


if __name__ == "__main__":
    main()
