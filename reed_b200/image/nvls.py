"""Host side of the NVSwitch-multicast gradient exchange (csrc/nvls.cu) for the sharded optimizer.

Default data path of ``ReedTrainer`` on more than one rank (bf16 mode, NCCL group whose GPUs share an NVSwitch multicast
domain); validated on 2 x B200 in round 2 against the replicated trainer (profiles/r02_sharded_check.txt).
``ReedTrainer(nvls=False)`` / ``REED_NVLS=0`` keeps NCCL collectives for the same sharded step.

What it replaces: in the sharded data-parallel step (trainer.py) the NCCL reduce-scatter of every block bucket, the
per-slice sum-of-squares pass and the NCCL all-gather of the bf16 GEMM operands (train.py:151,293,401 DDP all-reduce ->
402-412 clip / AdamW / EMA in the reference).  With this module

  * the gradient and bf16-operand buffers of the sharded buckets are symmetric-memory allocations
    (``torch.distributed._symmetric_memory``: same size on every rank, peer-mapped, with a multicast address);
  * after a block's backward, on a side stream: a cross-rank barrier, then ``reed_nvls_reduce_scatter_sumsq`` pulls the
    summed slice through the switch (``multimem.ld_reduce``) and adds its squares to the norm accumulator;
  * the optimizer kernel ``reed_adamw_ema_mc`` stores the new bf16 operands through the multicast address, so every
    rank's operand buffer is complete once all ranks passed the barrier that opens the next step.

PyTorch is plumbing here (allocation, rendezvous, the barrier kernel, streams and events); the data path is the two
kernels.  Hazards and what orders them:
  gradients   written by this rank's weight-gradient GEMMs  -> barrier(bucket)      -> read by every rank's ld_reduce
              read by the peers' ld_reduce                   -> norm all-reduce      -> overwritten by the next backward
  operands    read by this rank's forward/backward GEMMs     -> norm all-reduce      -> overwritten by the peers' multimem.st
              written by the peers' multimem.st              -> barrier(step start)  -> read by the next forward
(the all-reduce of the norm scalar completes on a rank only after every rank has enqueued - behind its own backward and
reduce-scatters - its contribution, so it doubles as the "everyone is done reading" fence).
"""
from __future__ import annotations

from typing import Optional

import torch

from .. import ops


class _Join:
    """Stands in for a c10d Work: wait() makes the current stream wait for the side stream's reduce-scatter."""

    def __init__(self, event):
        self.event = event

    def wait(self):
        torch.cuda.current_stream().wait_event(self.event)


def _enable_group(symm, group):
    """Needed by older 2.x releases, a deprecated no-op in newer ones."""
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            symm.enable_symm_mem_for_group(group.group_name)
        except Exception:
            pass


def multicast_available(group, device) -> bool:
    """Collective: can this process group map symmetric memory with an NVSwitch multicast address?  Every rank gets the
    same answer (the trainer picks its exchange path from it when the caller did not choose)."""
    import torch.distributed as dist
    ok = 1
    try:
        import torch.distributed._symmetric_memory as symm
        g = group if group is not None else dist.group.WORLD
        _enable_group(symm, g)
        probe = symm.empty(1024, dtype=torch.float32, device=device)
        handle = symm.rendezvous(probe, g)
        ok = 1 if handle.multicast_ptr else 0
    except Exception:
        ok = 0
    flag = torch.tensor([ok], device=device, dtype=torch.int32)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return bool(int(flag))


class NvlsExchange:
    def __init__(self, group=None, ctas: int = 16):
        import torch.distributed._symmetric_memory as symm
        self.symm = symm
        self.group = group if group is not None else torch.distributed.group.WORLD
        self.ctas = ctas
        self.stream: Optional[torch.cuda.Stream] = None
        self.handles = {}
        _enable_group(symm, self.group)

    # -- allocation hook for trainer.FlatState ---------------------------------------------------------------------
    def alloc(self, kind: str, bucket, numel: int, dtype, device):
        """Symmetric memory for the buffers that cross the switch (gradients, bf16 operands of sharded buckets)."""
        if not bucket.sharded or kind not in ("grad", "shadow"):
            return None
        t = self.symm.empty(numel, dtype=dtype, device=device)
        t.zero_()
        return t

    def attach(self, state):
        """Collective: exchange the handles of every sharded bucket's buffers; record the multicast addresses."""
        dev = state.buckets[0].param.device
        self.stream = torch.cuda.Stream(device=dev)
        for b in state.buckets:
            if not b.sharded:
                continue
            hg = self.symm.rendezvous(b.grad, self.group)
            hs = self.symm.rendezvous(b.shadow, self.group)
            if not hg.multicast_ptr or not hs.multicast_ptr:
                raise RuntimeError("NVLS multicast is not available for this process group (needs NVSwitch + fabric "
                                   "multicast support); construct the trainer with nvls=False (REED_NVLS=0) to use NCCL")
            b.grad_mc, b.shadow_mc = int(hg.multicast_ptr), int(hs.multicast_ptr)
            self.handles[b.name] = (hg, hs)
        torch.cuda.synchronize(dev)

    # -- per step ----------------------------------------------------------------------------------------------------
    def reduce_scatter(self, bucket, rank: int, world: int, norm_sq: torch.Tensor):
        """Enqueue, on the side stream, barrier + multicast reduce-scatter (+ sum of squares) of ``bucket``; everything the
        current stream has enqueued so far (the block's weight-gradient GEMMs) comes first."""
        lo, n = bucket.shard(rank, world)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            self.handles[bucket.name][0].barrier(channel=0)
            ops._launch("reed_nvls_reduce_scatter_sumsq", bucket.grad_mc, bucket.grad.data_ptr(), lo, n,
                        norm_sq.data_ptr(), self.ctas, ops._stream())
            done = torch.cuda.Event()
            done.record(self.stream)
        return _Join(done)

    def open_step(self, state):
        """Barrier on the current stream: every rank's optimizer kernel (multicast operand stores) has finished."""
        for b in state.buckets:
            if b.sharded:
                self.handles[b.name][1].barrier(channel=0)
                return
