"""Sharded optimizer (ReedTrainer(shard_optimizer=True)): host-side logic on CPU over gloo, world size 2.

The CUDA kernels are replaced by torch restatements of the same formulas (clip coefficient, AdamW, EMA, bf16 shadow);
what is under test is the partitioning: bucket layout, reduce(-scatter), the global norm from per-rank slices, slice-wise
updates, the operand all-gather ahead of each block's forward, gather_state / checkpoint.  The invariant: after a
step the sharded trainer holds - on every rank, once gathered - exactly what the replicated trainer holds."""
import math
import os
import socket
import zlib

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _tiny_model():
    from reed_b200.image.models.sit import SiT
    torch.manual_seed(0)
    return SiT(input_size=8, hidden_size=32, decoder_hidden_size=32, depth=3, num_heads=2, encoder_depth=1, z_dims=[16],
               projector_dim=32, num_classes=10, qk_norm=False)


def _emulate_kernels(tr):
    """torch stand-ins for reed_grad_sumsq / reed_adamw_ema (optim.cu) on CPU tensors."""
    def sumsq(flat, out):
        out += flat.double().pow(2).sum()

    def adamw(b, lo, n, device_step):
        sl = slice(lo, lo + n)
        scale = tr.reducer.grad_scale
        coef = scale * min(tr.max_grad_norm / (math.sqrt(float(tr._norm_sq)) * scale + 1e-6), 1.0)
        g = b.grad[sl] * coef
        step = tr.step_count
        b.param[sl].mul_(1 - tr.lr * tr.weight_decay)
        b.exp_avg[sl].mul_(tr.betas[0]).add_(g, alpha=1 - tr.betas[0])
        b.exp_avg_sq[sl].mul_(tr.betas[1]).addcmul_(g, g, value=1 - tr.betas[1])
        denom = b.exp_avg_sq[sl].sqrt() / math.sqrt(1 - tr.betas[1] ** step) + tr.eps
        b.param[sl].addcdiv_(b.exp_avg[sl], denom, value=-tr.lr / (1 - tr.betas[0] ** step))
        b.ema[sl].mul_(tr.ema_decay).add_(b.param[sl], alpha=1 - tr.ema_decay)
        b.shadow[sl].copy_(b.param[sl])

    tr._k_sumsq, tr._k_adamw = sumsq, adamw
    tr.state.frozen = []                      # the frozen-parameter EMA kernel is not part of what is tested here


def _replicated_step(tr):
    """What ReedTrainer.optimizer_step does with the real kernels, through the same stand-ins."""
    tr.step_count += 1
    tr._norm_sq.zero_()
    for b in tr.state.buckets:
        tr._k_sumsq(b.grad, tr._norm_sq)
    for b in tr.state.buckets:
        tr._k_adamw(b, 0, b.numel, False)


def _fill_grads(tr, rank, step):
    """Rank- and step-dependent 'local gradients', identical for the two trainers; keyed by parameter name because the
    two layouts place parameters at different offsets."""
    for b in tr.state.buckets:
        b.grad.zero_()
        for name, p, off in zip(b.names, b.params, b.offsets):
            g = torch.Generator().manual_seed(zlib.crc32(f"{name}/{rank}/{step}".encode()))
            b.grad[off:off + p.numel()].copy_(torch.randn(p.numel(), generator=g) * 3.0)


def _by_name(tr, field):
    out = {}
    for b in tr.state.buckets:
        for name, p, off in zip(b.names, b.params, b.offsets):
            out[name] = getattr(b, field)[off:off + p.numel()].clone()
    return out


class _FakeMulticast:
    """Stands in for nvls.NvlsExchange on CPU: same calls, torch collectives instead of multimem loads / stores."""

    class _Done:
        def wait(self):
            pass

    def __init__(self, trainer):
        self.tr = trainer
        self.opened = 0

    def reduce_scatter(self, bucket, rank, world, norm_sq):
        dist.all_reduce(bucket.grad)                                  # ld_reduce: the owned slice arrives summed ...
        lo, n = bucket.shard(rank, world)
        norm_sq += bucket.grad[lo:lo + n].double().pow(2).sum()       # ... and its squares are added in the same pass
        return self._Done()

    def open_step(self, state):                                       # multimem.st: every rank's slice lands everywhere
        self.opened += 1
        for b in state.buckets:
            if b.sharded:
                self.tr.reducer.gather(b, b.shadow)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from reed_b200.image.trainer import ReedTrainer
        rep = ReedTrainer(_tiny_model(), None, precision="bf16", shard_optimizer=False)
        shd = ReedTrainer(_tiny_model(), None, precision="bf16", shard_optimizer=True)
        for tr in (rep, shd):
            _emulate_kernels(tr)
        ok = shd.shard and not rep.shard
        # the NCCL form awaits each block's operands in that block's forward: nothing may read all blocks' weights at step
        # start, so the grouped adaLN GEMM (ops.AdaLNAll) is off there and on for the replicated trainer
        ok &= shd.model._reed_adaln_grouped is False and rep.model._reed_adaln_grouped is True
        # layout: block buckets hold 2-D weights only, in `world` equal 16-byte-aligned slices; 1-D parameters are replicated
        for b in shd.state.buckets:
            if b.name == "outer":
                ok &= not b.sharded and any(n.startswith("blocks.") and n.endswith(".bias") for n in b.names)
            else:
                ok &= b.sharded and b.numel % (8 * world) == 0 and all(p.dim() == 2 for p in b.params)
                lo, n = b.shard(rank, world)
                ok &= (lo, n) == (rank * b.numel // world, b.numel // world)
        ok &= sum(p.numel() for b in shd.state.buckets for p in b.params) == \
            sum(p.numel() for b in rep.state.buckets for p in b.params)
        for step in (1, 2, 3):
            for tr in (rep, shd):
                _fill_grads(tr, rank, step)
            # replicated: all-reduce everything, update everything
            rep.reducer.finish()
            _replicated_step(rep)
            # sharded: the block buckets go out early (backward hook), the outer bucket in finish(); slices only are updated
            shd.reducer.launch(shd.state.bucket_of_block(2))
            shd.reducer.finish()
            shd.optimizer_step()
            ok &= shd._operands_stale and not shd._state_complete
            ok &= abs(float(shd._norm_sq) - float(rep._norm_sq)) <= 1e-9 * float(rep._norm_sq)
            # before the gather, foreign slices of the shadows are out of date ...
            blk = shd.state.bucket_of_block(1)
            lo, n = blk.shard(1 - rank, world)
            ok &= not torch.equal(blk.shadow[lo:lo + n].float(), _flat_like(rep, blk, "shadow")[lo:lo + n].float())
            # ... the forward pre-hook of each block waits for exactly its bucket
            shd._gather_operands()
            ok &= all(b.gather_work is not None for b in shd.state.buckets if b.sharded)
            shd._await_operands(blk)
            ok &= blk.gather_work is None and shd.state.bucket_of_block(0).gather_work is not None
            for b in shd.state.buckets:
                shd._await_operands(b)
            want, got = _by_name(rep, "shadow"), _by_name(shd, "shadow")
            ok &= all(torch.equal(want[k], got[k]) for k in want)
            # fp32 masters of foreign slices are still old until gather_state()
            try:
                shd.checkpoint()
                ok = False
            except RuntimeError:
                pass
            shd.gather_state()
            for field in ("param", "ema", "exp_avg", "exp_avg_sq"):
                want, got = _by_name(rep, field), _by_name(shd, field)
                ok &= all(torch.equal(want[k], got[k]) for k in want)
        # the multicast-exchange wiring (norm partials from the reduce-scatter, one barrier instead of all-gathers)
        rep2 = ReedTrainer(_tiny_model(), None, precision="bf16", shard_optimizer=False)
        mc = ReedTrainer(_tiny_model(), None, precision="bf16", shard_optimizer=True)
        for tr in (rep2, mc):
            _emulate_kernels(tr)
        mc.nvls = _FakeMulticast(mc)
        mc.reducer.nvls, mc.reducer.norm_sq_shard = mc.nvls, mc._norm_sq_shard
        for step in (1, 2):
            mc._gather_operands()
            mc._begin_step()
            for tr in (rep2, mc):
                _fill_grads(tr, rank, step)
            rep2.reducer.finish()
            _replicated_step(rep2)
            mc.reducer.launch(mc.state.bucket_of_block(1))
            mc.reducer.finish()
            mc.optimizer_step()
            ok &= abs(float(mc._norm_sq) - float(rep2._norm_sq)) <= 1e-9 * float(rep2._norm_sq)
        ok &= mc.nvls.opened == 1 and mc._operands_stale
        mc._gather_operands()
        ok &= mc.nvls.opened == 2 and all(b.gather_work is None for b in mc.state.buckets)
        want, got = _by_name(rep2, "shadow"), _by_name(mc, "shadow")
        ok &= all(torch.equal(want[k], got[k]) for k in want)
        mc.gather_state()
        for field in ("param", "ema", "exp_avg_sq"):
            want, got = _by_name(rep2, field), _by_name(mc, field)
            ok &= all(torch.equal(want[k], got[k]) for k in want)
        ck = shd.checkpoint()
        ok &= ck["steps"] == 3 and all(torch.equal(v, rep.model.state_dict()[k]) for k, v in ck["model"].items())
        # defaults with more than one rank: sharded; the multicast kernels only on an NCCL group (this one is gloo)
        auto = ReedTrainer(_tiny_model(), None, precision="bf16")
        ok &= auto.shard and auto.nvls is None and all(b.sharded == (b.name != "outer") for b in auto.state.buckets)
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def _flat_like(rep, sharded_bucket, field):
    """The replicated trainer's values laid out like `sharded_bucket` (parameters sit at other offsets there)."""
    values = _by_name(rep, field)
    flat = torch.zeros(sharded_bucket.numel, dtype=getattr(sharded_bucket, field).dtype)
    for name, p, off in zip(sharded_bucket.names, sharded_bucket.params, sharded_bucket.offsets):
        flat[off:off + p.numel()] = values[name]
    return flat


def test_sharded_optimizer_matches_replicated_world2_gloo():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_single_process_ignores_the_flag():
    from reed_b200.image.trainer import ReedTrainer
    tr = ReedTrainer(_tiny_model(), None, shard_optimizer=True)
    assert not tr.shard and all(not b.sharded for b in tr.state.buckets)
    assert [b.name for b in tr.state.buckets] == ["outer", "blocks.0", "blocks.1", "blocks.2"]
    tr.gather_state()                                   # no-op
    assert tr.state.buckets[1].shard(0, 1) == (0, tr.state.buckets[1].numel)
