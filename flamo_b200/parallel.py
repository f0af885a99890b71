"""Multi-GPU training of the sweep: one process per GPU (torchrun), torch.distributed for the plumbing.

The reference has no multi-device path.  Bins and batch items are independent units,
so the path shards with no data-path collective; the only exchange is ONE all-reduce per step of the
flat parameter-gradient buffer (a few hundred bytes to a few KB) with the per-criterion losses riding
in its tail.  Two sharding modes:

  shard="batch"  every rank sweeps all bins of its own batch items (weak scaling: per-GPU work fixed);
                 gradients are averaged.
  shard="bins"   every rank sweeps a contiguous bin range of the same batch (strong scaling); bin-mean
                 criteria are weighted by M_rank / M so the SUM over ranks is the global mean, and
                 parameter-only criteria (requires_model=True) are weighted 1/world.

Everything between the input copy and the loss read-back, including the NCCL all-reduce, is captured
in the step's CUDA graph when graph=True.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import sweep
from .optimize.trainer import Trainer


def bin_range(M: int, rank: int, world: int):
    """Contiguous, balanced partition of M bins."""
    base, rem = divmod(M, world)
    b0 = rank * base + min(rank, rem)
    return b0, b0 + base + (1 if rank < rem else 0)


class DataParallelTrainer(Trainer):
    def __init__(self, net, *args, shard: str = "batch", process_group=None, **kwargs):
        super().__init__(net, *args, **kwargs)
        assert shard in ("batch", "bins")
        self.shard, self.pg = shard, process_group
        self.world = dist.get_world_size(self.pg) if dist.is_initialized() else 1
        self.rank = dist.get_rank(self.pg) if dist.is_initialized() else 0
        self._flat = None

    # gradients live in one flat buffer so that a step needs exactly one collective
    def _setup_flat(self, n_vals):
        ps = [p for p in self.net.parameters() if p.requires_grad]
        total = sum(p.numel() for p in ps)
        dt = ps[0].dtype
        self._flat = torch.zeros(total + n_vals, dtype=dt, device=ps[0].device)
        off = 0
        for p in ps:
            p.grad = self._flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self._n_grad = total

    def _zero_grad(self):
        if self._flat is None:
            self._setup_flat(self.n_loss + 1)
        self._flat.zero_()

    def _zero_grad_captured(self):
        self._flat.zero_()

    def _loss_vector(self, inputs, targets):
        if self.shard == "batch" or self.world == 1:
            return super()._loss_vector(inputs, targets)
        M = self.net.nfft // 2 + 1
        b0, b1 = bin_range(M, self.rank, self.world)
        with sweep.bin_shard(b0, b1):
            est, done = self._predict(inputs, targets)  # a fused criterion slices the target itself
        tg = targets[:, b0:b1] if targets.shape[1] == M else targets
        weight = [1.0 / self.world if self.requires_model[i] else (b1 - b0) / M for i in range(len(self.criterion))]
        return self._criteria(est, done, tg, weight=weight)

    def _sync(self, vals):
        if self.world == 1:
            return vals
        self._flat[self._n_grad:] = vals.to(self._flat.dtype)
        dist.all_reduce(self._flat, op=dist.ReduceOp.SUM, group=self.pg)
        if self.shard == "batch":
            self._flat.div_(self.world)
        return self._flat[self._n_grad:].to(vals.dtype)
