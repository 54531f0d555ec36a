"""Data-parallel train step for SiT + SILoss: the step glue of the reference ``image/train.py`` on flat buffers.

Restates /root/reference/image/train.py:
  * 220, 270      EMA = deepcopy(model), ``update_ema(ema, model, decay=0)`` at start
  * 253-259       AdamW(lr 1e-4, betas (0.9, 0.999), wd 0, eps 1e-8)
  * 396-398       loss = mean(denoising) * diffusion_decay + proj_loss * proj_coeff * repa_decay
  * 401           backward; under DDP one gradient all-reduce (average) per step, bucketed, overlapped with backward
  * 402-407       clip_grad_norm_(max_grad_norm)
  * 408-412       optimizer.step(); zero_grad; update_ema(ema, model, 0.9999)   (EMA covers the frozen pos_embed too)

B200 design: parameters, gradients, Adam moments, EMA weights and the bf16 GEMM shadows live in flat per-bucket
buffers (one bucket per transformer block + one for embedders/projectors/final layer).  Weight-gradient GEMMs write
straight into the flat gradient (no autograd accumulation pass); each block's bucket is all-reduced over NCCL
(NVLink 5 / NVSwitch) as soon as that block's backward finishes, on NCCL's own stream, overlapping the remaining
backward; the optimizer tail is one norm pass + one fused clip/AdamW/EMA/shadow-refresh kernel per bucket.
"""
from __future__ import annotations

import copy
import os
from typing import Dict, List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from .. import ops

_ALIGN = 8     # elements; keeps every parameter's bf16 shadow 16-byte aligned for TMA


class Bucket:
    def __init__(self, name: str):
        self.name = name
        self.params: List[nn.Parameter] = []
        self.names: List[str] = []
        self.offsets: List[int] = []
        self.numel = 0
        self.work = None
        self.sharded = False       # optimizer state of this bucket is partitioned across the data-parallel ranks
        self.gather_work = None    # in-flight all-gather of the bf16 shadows / fp32 masters (sharded optimizer)

    def shard(self, rank: int, world: int):
        """(first element, element count) of the slice rank ``rank`` owns; the whole bucket when it is not sharded."""
        if not self.sharded:
            return 0, self.numel
        size = self.numel // world
        return rank * size, size

    def add(self, name, p):
        self.names.append(name)
        self.params.append(p)
        self.offsets.append(self.numel)
        self.numel += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN


class _ZeroGroup:
    """The contiguous gradient region of a bucket's 1-D parameters and whether it was zeroed this step."""

    def __init__(self, view):
        self.view = view
        self.zeroed = False


class FlatState:
    """Re-homes a model's parameters into flat per-bucket storage (works on any device; kernels need CUDA)."""

    def __init__(self, model: nn.Module, ema: Optional[nn.Module] = None, with_shadow: bool = True,
                 shard_world: int = 1, alloc=None):
        """shard_world > 1 lays the buckets out for the sharded optimizer (ReedTrainer(shard_optimizer=True)): block
        buckets hold the 2-D weights only and are padded to ``shard_world`` equal 16-byte-aligned slices; every 1-D
        parameter (biases are read in fp32 by the GEMM epilogues on every rank) moves to the replicated outer bucket."""
        self.model = model
        self.shard_world = shard_world

        def new(kind, b, dtype):
            """``alloc(kind, bucket, numel, dtype, device)`` may supply a buffer (nvls.py: symmetric memory); else zeros."""
            t = alloc(kind, b, b.numel, dtype, b.params[0].device) if alloc is not None else None
            return t if t is not None else torch.zeros(b.numel, device=b.params[0].device, dtype=dtype)
        self.ema = ema
        self.buckets: List[Bucket] = []
        self.frozen: List[tuple] = []          # (param, ema_param) for requires_grad=False parameters
        ema_params = dict(ema.named_parameters()) if ema is not None else {}
        by_key: Dict[str, Bucket] = {}

        def bucket_for(name, p):
            in_block = name.startswith("blocks.") and not (shard_world > 1 and p.dim() == 1)
            key = ".".join(name.split(".")[:2]) if in_block else "outer"
            if key not in by_key:
                by_key[key] = Bucket(key)
                self.buckets.append(by_key[key])
            return by_key[key]

        # within a bucket the 1-D parameters (biases) come first and contiguous: their gradients are accumulated
        # into (column sums, atomics) and are zeroed with ONE fill per bucket on first touch (ops._zero_fresh)
        named = list(model.named_parameters())
        for name, p in [(n, q) for n, q in named if q.dim() == 1] + [(n, q) for n, q in named if q.dim() != 1]:
            if not p.requires_grad:
                self.frozen.append((p, ema_params.get(name)))
                continue
            bucket_for(name, p).add(name, p)

        if shard_world > 1:
            quantum = _ALIGN * shard_world
            for b in self.buckets:
                if b.name != "outer":
                    b.sharded = True
                    b.numel = (b.numel + quantum - 1) // quantum * quantum     # the pad stays zero in every buffer

        for b in self.buckets:
            dev = b.params[0].device
            b.param = torch.zeros(b.numel, device=dev, dtype=torch.float32)
            b.grad = new("grad", b, torch.float32)
            b.exp_avg = torch.zeros(b.numel, device=dev, dtype=torch.float32)
            b.exp_avg_sq = torch.zeros(b.numel, device=dev, dtype=torch.float32)
            b.ema = torch.zeros(b.numel, device=dev, dtype=torch.float32) if ema is not None else None
            b.shadow = new("shadow", b, torch.bfloat16) if with_shadow else None
            for name, p, off in zip(b.names, b.params, b.offsets):
                n = p.numel()
                view = b.param[off:off + n].view(p.shape)
                view.copy_(p.data)
                p.data = view
                gview = b.grad[off:off + n].view(p.shape)
                p.grad = gview
                # nn.Embedding gradients arrive through autograd (dense add into p.grad); everything else is
                # written by the wgrad / bias-gradient kernels directly
                p._reed_kernel_grad = not name.endswith("embedding_table.weight")
                if p._reed_kernel_grad:
                    p._reed_main_grad = gview
                    p._reed_grad_fresh = True
                if with_shadow:
                    sview = b.shadow[off:off + n].view(p.shape)
                    sview.copy_(p.data)
                    p._reed_shadow = sview
                    p._reed_shadow_version = p._version
                    p._reed_shadow_managed = True       # reed_adamw_ema rewrites it together with the master
                if ema is not None:
                    ep = ema_params[name]
                    eview = b.ema[off:off + n].view(p.shape)
                    eview.copy_(p.data)                      # update_ema(ema, model, decay=0)
                    ep.data = eview
            n1d = sum((p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN for p in b.params if p.dim() == 1)
            b.zero_group = _ZeroGroup(b.grad[:n1d])
            for p in b.params:
                if p.dim() == 1 and p._reed_kernel_grad:
                    p._reed_zero_group = b.zero_group
        for p, ep in self.frozen:
            if ep is not None:
                ep.data.copy_(p.data)
        self._block_bucket = {b.name: b for b in self.buckets}

    # -- gradient lifecycle ----------------------------------------------------------------------------------
    def begin_step(self):
        """Equivalent of optimizer.zero_grad(): kernel-written gradients are overwritten on their first write;
        gradients that arrive through autograd (the label-embedding table) are zeroed and accumulated into."""
        # the 1-D gradient regions of all buckets (kernels add into them): one multi-tensor fill instead of one fill per
        # bucket on first touch (29 launches per step on XL/2)
        views = [b.zero_group.view for b in self.buckets if b.zero_group.view.numel() > 0]
        if views and views[0].is_cuda:
            torch._foreach_zero_(views)
        for b in self.buckets:
            b.work = None
            b.zero_group.zeroed = bool(views) and views[0].is_cuda
            for p, off in zip(b.params, b.offsets):
                if p._reed_kernel_grad:
                    p._reed_grad_fresh = True
                    p.grad = None
                else:
                    g = b.grad[off:off + p.numel()].view(p.shape)
                    g.zero_()
                    p.grad = g

    def finish_backward(self):
        """Zero the gradient of any kernel-written parameter the backward pass never touched; expose .grad views."""
        for b in self.buckets:
            for p, off in zip(b.params, b.offsets):
                if p._reed_kernel_grad:
                    if p._reed_grad_fresh:
                        p._reed_main_grad.zero_()
                    p.grad = p._reed_main_grad

    def bucket_of_block(self, index: int) -> Optional[Bucket]:
        return self._block_bucket.get(f"blocks.{index}")

    def total_params(self):
        return sum(sum(p.numel() for p in b.params) for b in self.buckets)


class GradientReducer:
    """Bucketed gradient all-reduce (sum; the 1/world factor is folded into the optimizer kernel)."""

    def __init__(self, state: FlatState, group=None):
        self.state = state
        self.group = group
        live = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if live else 1
        self.rank = dist.get_rank(group) if live else 0
        # gloo (the CPU tests) has neither reduce-scatter nor an in-place all-gather: fall back to all-reduce / a staged copy
        self.nccl = live and dist.get_backend(group) == "nccl"
        self.enabled = True        # False while micro-batches of an accumulated step are still adding into the buckets
        self.nvls = None           # nvls.NvlsExchange: multicast reduce-scatter kernels instead of NCCL for sharded buckets
        self.norm_sq_shard = None  # ... which also accumulate the owned slices' sums of squares into this device scalar

    @property
    def grad_scale(self):
        return 1.0 / self.world

    def launch(self, bucket: Bucket):
        if self.world == 1 or bucket.work is not None or not self.enabled:
            return
        if bucket.sharded and self.nvls is not None:
            bucket.work = self.nvls.reduce_scatter(bucket, self.rank, self.world, self.norm_sq_shard)
            return
        if bucket.sharded and self.nccl:
            # sharded optimizer: every rank only needs the sum over ranks of its own slice (in place: NCCL's
            # recvbuff == sendbuff + rank * recvcount form) - half the NVLink traffic of the all-reduce
            lo, n = bucket.shard(self.rank, self.world)
            bucket.work = dist.reduce_scatter_tensor(bucket.grad[lo:lo + n], bucket.grad, op=dist.ReduceOp.SUM,
                                                     group=self.group, async_op=True)
            return
        bucket.work = dist.all_reduce(bucket.grad, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def gather(self, bucket: Bucket, flat: torch.Tensor, async_op=False):
        """All-gather a sharded bucket's flat buffer in place: every rank contributes the slice it owns."""
        lo, n = bucket.shard(self.rank, self.world)
        own = flat[lo:lo + n]
        if not self.nccl:
            own = own.clone()
        return dist.all_gather_into_tensor(flat, own, group=self.group, async_op=async_op)

    def finish(self):
        for b in self.state.buckets:
            self.launch(b)
        for b in self.state.buckets:
            if b.work is not None:
                b.work.wait()
                b.work = None


class ReedTrainer:
    """loss -> backward (+ overlapped all-reduce) -> clip -> AdamW -> EMA, one call per optimizer step."""

    _STAGING_ROWS = 4      # pinned rows for the graphed step's host-side inputs (how far the host may run ahead)

    def __init__(self, model: nn.Module, loss_fn, *, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                 max_grad_norm=1.0, ema_decay=0.9999, proj_coeff=0.5, precision: Optional[str] = "bf16", group=None,
                 with_ema=True, comm_sms: int = 16, shard_optimizer: Optional[bool] = None, nvls: Optional[bool] = None):
        self.model = model
        self.loss_fn = loss_fn
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.max_grad_norm, self.ema_decay, self.proj_coeff = max_grad_norm, ema_decay, proj_coeff
        if precision is not None:
            model.reed_precision = precision
        self.ema = None
        if with_ema:
            self.ema = copy.deepcopy(model)
            for p in self.ema.parameters():
                p.requires_grad_(False)
            self.ema.eval()
        # Sharded optimizer (ZeRO-1 over NVSwitch; the default when there is more than one rank): the clip/AdamW/EMA pass is
        # HBM-bound at 38 B/param and every rank would repeat it on identical data.  Each rank reduce-scatters the block
        # buckets, updates 1/world of every block's weights (masters, moments, EMA) and only the bf16 GEMM operands are
        # gathered - under the next forward, block by block.  Same arithmetic on every element (_optimizer_step_sharded);
        # validated against the replicated trainer on 2 x B200 (profiles/r02_sharded_check.txt: 1790 -> 1841 img/s, and
        # 1866 img/s with the multicast kernels below).
        world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        if shard_optimizer is None:
            shard_optimizer = world > 1
        if nvls is None:                       # our own multicast kernels need bf16 operands and an NCCL (NVSwitch) group
            nvls = bool(shard_optimizer) and world > 1 and precision == "bf16" and dist.get_backend(group) == "nccl"
            if nvls:                           # not chosen by the caller: fall back to NCCL collectives where the fabric has no multicast
                from .nvls import multicast_available
                nvls = multicast_available(group, next(model.parameters()).device)
                if not nvls and dist.get_rank(group) == 0:
                    import warnings
                    warnings.warn("reed_b200: NVSwitch multicast is not available to this process group; the sharded step uses "
                                  "NCCL reduce-scatter / all-gather instead of the multimem kernels")
        self.shard = bool(shard_optimizer) and world > 1
        self.precision = precision
        # nvls (needs shard_optimizer, bf16 operands, NCCL group on one NVSwitch domain): the reduce-scatter and the
        # operand all-gather become multicast loads / stores inside our own kernels (csrc/nvls.cu, image/nvls.py)
        self.nvls = None
        if self.shard and nvls and precision == "bf16" and dist.get_backend(group) == "nccl":
            from .nvls import NvlsExchange
            self.nvls = NvlsExchange(group, ctas=comm_sms or 16)
        self.state = FlatState(model, self.ema, with_shadow=True, shard_world=world if self.shard else 1,
                               alloc=self.nvls.alloc if self.nvls is not None else None)
        self.reducer = GradientReducer(self.state, group)
        # SMs left to the NCCL kernels while backward overlaps the bucket all-reduces (bench.py caps NCCL's CTAs to match)
        self.comm_sms = comm_sms if self.reducer.world > 1 else 0
        if self.reducer.world > 1:                      # DDP broadcasts rank 0's weights when it wraps the model
            for b in self.state.buckets:
                dist.broadcast(b.param, 0, group=group)
                if b.ema is not None:
                    b.ema.copy_(b.param)
                if b.shadow is not None:
                    b.shadow.copy_(b.param)
        self.step_count = 0
        dev = next(model.parameters()).device
        self._norm_sq = torch.zeros(1, device=dev, dtype=torch.float64)
        self._norm_sq_shard = torch.zeros(1, device=dev, dtype=torch.float64)
        if self.nvls is not None:
            self.nvls.attach(self.state)
            self.reducer.nvls, self.reducer.norm_sq_shard = self.nvls, self._norm_sq_shard
        self._operands_stale = False         # sharded optimizer: other ranks' slices of the GEMM operands are out of date
        self._state_complete = True          # ... and of the masters / EMA / moments (gather_state() completes them)
        self._step_dev = torch.zeros(1, device=dev, dtype=torch.int32)   # device copy of step_count (graph replays)
        self._graph = None
        # NCCL form of the sharded step: block i's operands arrive by an all-gather awaited in block i's forward pre-hook,
        # so nothing may read all blocks' weights up front (the multicast form opens the step with one barrier instead)
        model._reed_adaln_grouped = not (self.shard and self.nvls is None)
        # all-reduce each block's bucket as soon as that block's backward has produced its last gradient
        for i, blk in enumerate(model.blocks):
            bucket = self.state.bucket_of_block(i)
            blk._reed_after_backward = (lambda b=bucket: self._after_block_backward(b))
            if self.shard:                   # the block's operands arrive by all-gather: wait for this block's only
                blk.register_forward_pre_hook(lambda _mod, _inp, b=bucket: self._await_operands(b))

    def _begin_step(self):
        self.state.begin_step()
        if self.nvls is not None:            # the multicast reduce-scatter kernels add their slices' squares during backward
            self._norm_sq_shard.zero_()

    def _after_block_backward(self, bucket: Bucket):
        self.reducer.launch(bucket)

    def compute_loss(self, images, labels, zs, diffusion_decay=1.0, repa_decay=1.0, time_input=None):
        """diffusion_decay / repa_decay: Python floats or 0-d device tensors (the curriculum scalars of train.py:363-385)."""
        extra = {} if time_input is None else {"time_input": time_input}
        out = self.loss_fn(self.model, images, dict(y=labels), zs=zs, **extra)
        loss = out["denoising_loss"].mean() * diffusion_decay + out["proj_loss"] * (self.proj_coeff * repa_decay)
        return loss, out

    def optimizer_step(self, device_step=False):
        """device_step: read the step from the device counter (incremented here by a kernel) instead of passing the
        host integer - the form a CUDA graph can replay."""
        ops.bump_weights_epoch()               # masters and EMA change behind autograd's back: lazily made shadows are stale
        if self.shard:
            return self._optimizer_step_sharded(device_step)
        self.step_count += 1
        self._step_dev += 1
        st = ops._stream()
        self._norm_sq.zero_()
        for b in self.state.buckets:
            ops._launch("reed_grad_sumsq", b.grad.data_ptr(), b.numel, self._norm_sq.data_ptr(), st)
        clip = self.max_grad_norm is not None and self.max_grad_norm > 0
        for b in self.state.buckets:
            ops._launch("reed_adamw_ema", b.param.data_ptr(), b.grad.data_ptr(), b.exp_avg.data_ptr(),
                        b.exp_avg_sq.data_ptr(), b.ema.data_ptr() if b.ema is not None else b.param.data_ptr(),
                        b.shadow.data_ptr() if b.shadow is not None else None, b.numel,
                        self._norm_sq.data_ptr() if clip else None, float(self.max_grad_norm or 0.0),
                        self.reducer.grad_scale, self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                        0 if device_step else self.step_count, self.ema_decay if b.ema is not None else 0.0,
                        self._step_dev.data_ptr() if device_step else None, st)
        for p, ep in self.state.frozen:
            if ep is not None:
                ops._launch("reed_ema_update", p.data_ptr(), ep.data_ptr(), p.numel(), self.ema_decay, st)

    # -- sharded optimizer ---------------------------------------------------------------------------------------
    def _k_sumsq(self, flat, out):
        ops._launch("reed_grad_sumsq", flat.data_ptr(), flat.numel(), out.data_ptr(), ops._stream())

    def _k_adamw(self, b: Bucket, lo: int, n: int, device_step: bool):
        """clip + AdamW + EMA + shadow refresh on elements [lo, lo+n) of bucket b (lo is a multiple of 8 elements)."""
        clip = self.max_grad_norm is not None and self.max_grad_norm > 0
        ema = b.ema if b.ema is not None else b.param
        if b.sharded and self.nvls is not None:     # bf16 operands stored through the multicast address: no all-gather
            ops._launch("reed_adamw_ema_mc", b.param[lo:].data_ptr(), b.grad[lo:].data_ptr(), b.exp_avg[lo:].data_ptr(),
                        b.exp_avg_sq[lo:].data_ptr(), ema[lo:].data_ptr(), b.shadow_mc + 2 * lo, n,
                        self._norm_sq.data_ptr() if clip else None, float(self.max_grad_norm or 0.0),
                        self.reducer.grad_scale, self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                        0 if device_step else self.step_count, self.ema_decay if b.ema is not None else 0.0,
                        self._step_dev.data_ptr() if device_step else None, ops._stream())
            return
        ops._launch("reed_adamw_ema", b.param[lo:].data_ptr(), b.grad[lo:].data_ptr(), b.exp_avg[lo:].data_ptr(),
                    b.exp_avg_sq[lo:].data_ptr(), ema[lo:].data_ptr(),
                    b.shadow[lo:].data_ptr() if b.shadow is not None else None, n,
                    self._norm_sq.data_ptr() if clip else None, float(self.max_grad_norm or 0.0),
                    self.reducer.grad_scale, self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                    0 if device_step else self.step_count, self.ema_decay if b.ema is not None else 0.0,
                    self._step_dev.data_ptr() if device_step else None, ops._stream())

    def _optimizer_step_sharded(self, device_step=False):
        """After the reduce-scatters each rank holds the summed gradient of its slice of every block bucket (and, after
        the all-reduce, of the whole outer bucket).  Global norm = all-reduce of the slices' sums of squares + the outer
        bucket's; then the fused kernel runs on the owned slices only: 38 B/param over 1/world of the weights."""
        self.step_count += 1
        self._step_dev += 1
        r, w = self.reducer.rank, self.reducer.world
        self._norm_sq.zero_()
        if self.nvls is None:
            self._norm_sq_shard.zero_()
        for b in self.state.buckets:
            lo, n = b.shard(r, w)
            if not (b.sharded and self.nvls is not None):        # else: summed by reed_nvls_reduce_scatter_sumsq already
                self._k_sumsq(b.grad[lo:lo + n], self._norm_sq_shard if b.sharded else self._norm_sq)
        dist.all_reduce(self._norm_sq_shard, op=dist.ReduceOp.SUM, group=self.reducer.group)
        self._norm_sq += self._norm_sq_shard
        for b in self.state.buckets:
            lo, n = b.shard(r, w)
            self._k_adamw(b, lo, n, device_step)
        for p, ep in self.state.frozen:
            if ep is not None:
                ops._launch("reed_ema_update", p.data_ptr(), ep.data_ptr(), p.numel(), self.ema_decay, ops._stream())
        self._operands_stale = True
        self._state_complete = False

    def _operand_buffers(self, b: Bucket):
        """What the forward reads of a block bucket: bf16 shadows in bf16 mode, fp32 masters in fp32 mode (both when the
        mode follows torch.autocast)."""
        if self.precision == "bf16":
            return [b.shadow]
        if self.precision == "fp32":
            return [b.param]
        return [b.shadow, b.param]

    def _gather_operands(self, force=False):
        """Start the all-gathers that bring every rank's updated slices of the GEMM operands to this rank, in forward
        order on NCCL's stream; block i's forward waits for bucket i only (_await_operands)."""
        if not self.shard or not (self._operands_stale or force):
            return
        if self.nvls is not None:              # operands arrived by multicast stores: one barrier opens the step
            self.nvls.open_step(self.state)
            self._operands_stale = False
            return
        for b in self.state.buckets:
            if b.sharded:
                b.gather_work = [self.reducer.gather(b, flat, async_op=True) for flat in self._operand_buffers(b)]
        self._operands_stale = False

    def _await_operands(self, b: Bucket):
        if b.gather_work is not None:
            for work in b.gather_work:
                work.wait()
            b.gather_work = None

    def gather_state(self):
        """Make every rank's copy of the sharded buffers complete (masters, EMA, moments, shadows): call before reading
        ``model`` / ``ema`` weights outside the train step (sampling, evaluation, checkpoints).  No-op when not sharded."""
        if not self.shard:
            return
        for b in self.state.buckets:
            self._await_operands(b)
            if b.sharded:
                for flat in (b.param, b.exp_avg, b.exp_avg_sq, b.ema, b.shadow):
                    if flat is not None:
                        self.reducer.gather(b, flat)
        self._operands_stale = False
        self._state_complete = True

    def _backward(self, loss):
        """backward with the per-bucket all-reduces in flight; the GEMM grids shrink by comm_sms SMs meanwhile."""
        if self.comm_sms:
            ops.call("reed_gemm_reserve_sms", self.comm_sms)
        try:
            loss.backward()
            self.state.finish_backward()
            self.reducer.finish()
        finally:
            if self.comm_sms:
                ops.call("reed_gemm_reserve_sms", 0)

    def grad_norm(self) -> torch.Tensor:
        """Global gradient norm of the last step (device scalar, after the all-reduce average)."""
        return (self._norm_sq.sqrt() * self.reducer.grad_scale).float()

    # -- checkpoints in the reference's format (train.py:280-289 resume, 418-429 save) ---------------------------
    # {"model": state_dict, "ema": state_dict, "opt": torch.optim.AdamW state_dict, "args": ..., "steps": int}: a file
    # written here resumes under the reference's train.py / loads in its generate.py (ckpt['ema']), and the reverse.
    def _param_slots(self):
        """[(name, param, bucket or None, offset)] in ``model.parameters()`` order = AdamW's parameter indices."""
        where = {}
        for b in self.state.buckets:
            for name, off in zip(b.names, b.offsets):
                where[name] = (b, off)
        return [(name, p) + where.get(name, (None, 0)) for name, p in self.model.named_parameters()]

    def checkpoint(self, args=None, steps: Optional[int] = None) -> dict:
        """Snapshot as the dict the reference passes to ``torch.save`` (train.py:420-426).  Tensors are copies: the
        live parameters are views of the flat buckets, and ``torch.save`` of a view would write the whole bucket."""
        if self.shard and not self._state_complete:
            # gather_state() is a collective; the reference saves on the main process only (train.py:419), so it cannot
            # be hidden in here
            raise RuntimeError("sharded optimizer: call trainer.gather_state() on EVERY rank before checkpoint()")

        def copied(sd):
            return type(sd)((k, v.detach().clone()) for k, v in sd.items())
        groups = torch.optim.AdamW(self.model.parameters(), lr=self.lr, betas=tuple(self.betas), eps=self.eps,
                                   weight_decay=self.weight_decay).state_dict()["param_groups"]
        opt_state = {}
        if self.step_count > 0:
            for idx, (name, p, b, off) in enumerate(self._param_slots()):
                if b is None:                       # frozen (pos_embed): AdamW never creates state for it
                    continue
                n = p.numel()
                opt_state[idx] = {"step": torch.tensor(float(self.step_count)),
                                  "exp_avg": b.exp_avg[off:off + n].view(p.shape).clone(),
                                  "exp_avg_sq": b.exp_avg_sq[off:off + n].view(p.shape).clone()}
        return {"model": copied(self.model.state_dict()),
                "ema": copied(self.ema.state_dict()) if self.ema is not None else None,
                "opt": {"state": opt_state, "param_groups": groups},
                "args": args, "steps": self.step_count if steps is None else steps}

    def load_checkpoint(self, ckpt: dict, strict: bool = True) -> int:
        """Resume from a checkpoint dict (this trainer's or the reference's, train.py:281-289).  Returns ckpt['steps']."""
        self.model.load_state_dict(ckpt["model"], strict=strict)         # in-place copies: the flat views stay in place
        if self.ema is not None and ckpt.get("ema") is not None:
            self.ema.load_state_dict(ckpt["ema"], strict=strict)
        opt = ckpt.get("opt")
        step = 0
        slots = self._param_slots()
        for b in self.state.buckets:
            b.exp_avg.zero_()
            b.exp_avg_sq.zero_()
        if opt is not None:
            group = opt["param_groups"][0]
            if len(opt["param_groups"]) != 1 or len(group["params"]) != len(slots):
                raise ValueError(f"optimizer state covers {sum(len(g['params']) for g in opt['param_groups'])} parameters in "
                                 f"{len(opt['param_groups'])} group(s); this model has {len(slots)} in one group")
            self.lr, self.betas, self.eps = group["lr"], tuple(group["betas"]), group["eps"]
            self.weight_decay = group["weight_decay"]
            for idx, st in opt["state"].items():
                name, p, b, off = slots[group["params"].index(idx)]
                if b is None:
                    continue
                n = p.numel()
                b.exp_avg[off:off + n].view(p.shape).copy_(st["exp_avg"])
                b.exp_avg_sq[off:off + n].view(p.shape).copy_(st["exp_avg_sq"])
                step = max(step, int(float(st["step"])))
        self.step_count = step
        self._step_dev.fill_(step)
        for b in self.state.buckets:                    # the bf16 GEMM operands follow the loaded masters
            if b.shadow is not None:
                b.shadow.copy_(b.param)
            for p in b.params:
                if getattr(p, "_reed_shadow", None) is not None:
                    p._reed_shadow_version = p._version
        ops.bump_weights_epoch()
        return int(ckpt.get("steps", step))

    def train_step(self, images, labels, zs, diffusion_decay=1.0, repa_decay=1.0):
        self._gather_operands()
        self._begin_step()
        loss, out = self.compute_loss(images, labels, zs, diffusion_decay, repa_decay)
        self._backward(loss)
        self.optimizer_step()
        return loss.detach(), out

    def train_step_accumulated(self, micro_batches, diffusion_decay=1.0, repa_decay=1.0):
        """One optimizer step over several micro-batches: ``accelerator.accumulate(model)`` with
        ``--gradient-accumulation-steps k`` (train.py:151,362,401).  Every micro-batch's loss is divided by k before
        backward (what ``accelerator.backward`` does), the weight-gradient kernels add into the flat buckets from the
        second micro-batch on, and the bucket all-reduces start only in the last backward (DDP ``no_sync`` until then).
        ``micro_batches``: sequence of ``(images, labels, zs)``.  Returns (mean loss, list of per-micro-batch dicts)."""
        micro_batches = list(micro_batches)
        k = len(micro_batches)
        if k == 0:
            raise ValueError("train_step_accumulated needs at least one micro-batch")
        self._gather_operands()
        self._begin_step()
        total, outs = 0.0, []
        for i, (images, labels, zs) in enumerate(micro_batches):
            loss, out = self.compute_loss(images, labels, zs, diffusion_decay, repa_decay)
            outs.append(out)
            total = total + loss.detach()
            if i < k - 1:
                self.reducer.enabled = False
                try:
                    (loss / k).backward()
                finally:
                    self.reducer.enabled = True
            else:
                self._backward(loss / k)
        self.optimizer_step()
        return total / k, outs

    # -- CUDA-graph replay of the whole step -----------------------------------------------------------------------
    # The step launches ~1100 kernels; enqueueing them from Python costs about as long as they run on a B200.  The
    # graph holds loss forward, backward (with the per-block all-reduces), clip, AdamW, EMA and the shadow refresh.
    # Host randomness stays outside: the SILoss time draw (CPU generator, loss.py:159) is made here with the same
    # call and copied into the graph's static input; device randomness (noise, label dropout) is captured through
    # torch's graph-safe generator state.
    def _step_body(self, g):
        self._gather_operands(force=True)      # recorded in the graph: every replay starts by completing the operands
        self._begin_step()
        loss, out = self.compute_loss(g["images"], g["labels"], g["zs"], g["scalars"][0], g["scalars"][1],
                                      time_input=g["time"])
        self._backward(loss)
        self.optimizer_step(device_step=True)
        return loss.detach(), out

    def capture(self, images, labels, zs, warmup=2):
        """Record one train step on batches shaped like (images, labels, zs).  The warm-up steps are real optimizer
        steps on the given batch (they size the allocator pools and run every first-use initialisation)."""
        dev = images.device
        g = {"images": images.clone(), "labels": labels.clone(), "zs": [z.clone() for z in zs],
             "time": torch.zeros((images.shape[0], 1, 1, 1), device=dev, dtype=torch.float32),
             "scalars": torch.ones(2, device=dev, dtype=torch.float32),
             # ring of pinned staging rows for the per-step host values (time draws + 2 curriculum scalars): a row is
             # rewritten only after the H2D copies that read it have executed (event fence) - the host may run several
             # replays ahead of the device
             "host": [torch.zeros(images.shape[0] + 2, dtype=torch.float32).pin_memory() for _ in range(self._STAGING_ROWS)],
             "host_done": [None] * self._STAGING_ROWS, "host_next": 0}
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                g["time"].copy_(self.loss_fn._sample_time(images.shape[0]))
                self._step_body(g)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        steps_before = self.step_count
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            g["loss"], g["out"] = self._step_body(g)
        self.step_count = steps_before          # capture records, it does not run: undo the host-side counter bump
        self._step_dev.fill_(steps_before)
        self._graph, self._g = graph, g
        return self

    def train_step_graphed(self, images, labels, zs, diffusion_decay=1.0, repa_decay=1.0):
        """Replay the captured step on a new batch.  Returns (loss, out) as views of the graph's static outputs."""
        assert self._graph is not None, "call capture() first"
        g = self._g
        row = g["host_next"]
        g["host_next"] = (row + 1) % self._STAGING_ROWS
        if g["host_done"][row] is not None:
            g["host_done"][row].synchronize()                       # the copies issued _STAGING_ROWS replays ago have run
        host = g["host"][row]
        bsz = images.shape[0]
        host[:bsz] = self.loss_fn._sample_time(bsz).flatten()       # the draw SILoss would make (CPU generator)
        host[bsz] = diffusion_decay
        host[bsz + 1] = repa_decay
        g["time"].view(-1).copy_(host[:bsz], non_blocking=True)
        g["scalars"].copy_(host[bsz:], non_blocking=True)
        g["host_done"][row] = torch.cuda.Event()
        g["host_done"][row].record(torch.cuda.current_stream(images.device))
        g["images"].copy_(images, non_blocking=True)
        g["labels"].copy_(labels, non_blocking=True)
        for dst, src in zip(g["zs"], zs):
            dst.copy_(src, non_blocking=True)
        self._graph.replay()
        ops.bump_weights_epoch()
        self._state_complete = not self.shard
        self.step_count += 1
        return g["loss"], g["out"]
