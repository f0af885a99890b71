"""Multi-GPU training of the sweep: one process per GPU (torchrun), torch.distributed for the plumbing.

The reference has no multi-device path.  Bins and batch items are independent units,
so the path shards with no data-path collective; the only exchange is ONE all-reduce per step of the
flat parameter-gradient buffer (a few hundred bytes to a few KB) with the per-criterion losses riding
in its tail.  On GPUs that buffer lives in symmetric (peer-mapped) memory and the all-reduce is ONE
hand-written kernel per rank over NVLink (libfsweep fsweep_allreduce_p2p: signal, read every peer's
buffer, sum in rank order, scale, write back — bit-identical on every rank, capture safe); NCCL is
the fallback.  Two sharding modes:

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


class _AllGatherBins(torch.autograd.Function):
    """(B, M_rank, ...) on every rank -> (B, M, ...) on every rank (bin ranges of `bin_range`, concatenated in rank
    order).  Whatever consumes the gathered spectrum (a time-domain output layer and its criterion) is evaluated
    identically on every rank; its gradient reaches the parameters through THIS rank's bins only, times `world`,
    because the trainer weights such a replicated criterion by 1 / world before the gradients are summed over ranks."""

    @staticmethod
    def forward(ctx, x, M, rank, world, group):
        sizes = [bin_range(M, r, world) for r in range(world)]
        width = max(b - a for a, b in sizes)
        pad = x.new_zeros((x.shape[0], width) + tuple(x.shape[2:]))
        pad[:, :x.shape[1]] = x
        parts = [torch.empty_like(pad) for _ in range(world)]
        if x.is_complex():  # (gloo has no complex all_gather)
            real = [torch.view_as_real(p) for p in parts]
            dist.all_gather(real, torch.view_as_real(pad).contiguous(), group=group)
        else:
            dist.all_gather(parts, pad.contiguous(), group=group)
        ctx.range, ctx.world = sizes[rank], world
        return torch.cat([p[:, :b - a] for p, (a, b) in zip(parts, sizes)], dim=1)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.range
        return g[:, a:b] * ctx.world, None, None, None, None


class DataParallelTrainer(Trainer):
    def __init__(self, net, *args, shard: str = "batch", process_group=None, **kwargs):
        super().__init__(net, *args, **kwargs)
        assert shard in ("batch", "bins")
        self.shard, self.pg = shard, process_group
        self.world = dist.get_world_size(self.pg) if dist.is_initialized() else 1
        self.rank = dist.get_rank(self.pg) if dist.is_initialized() else 0
        self._flat = None

    # gradients live in one flat buffer so that a step needs exactly one collective
    def _p2p_buffer(self, numel, dtype, device):
        """The flat buffer in symmetric (peer-mapped) memory + what libfsweep's one-shot all-reduce kernel needs, or
        None (not CUDA + NCCL / not float32 / too large / symmetric memory unavailable / FLAMO_B200_P2P_ALLREDUCE=0): the
        collective is then NCCL's.  Measured (profiles/r01_notes.md): inside the captured step the one-shot kernel
        (scale fused, latency flat in the number of ranks) beats NCCL by 6 us on 2 GPUs and 33 us on 4."""
        import os

        from . import _lib

        if (self.world == 1 or device.type != "cuda" or dtype != torch.float32 or dist.get_backend(self.pg) != "nccl"
                or os.environ.get("FLAMO_B200_P2P_ALLREDUCE", "1") == "0"):
            return None
        try:
            if numel > _lib.lib().fsweep_allreduce_p2p_max_n() or len(self._ps) + 1 > 32:
                return None
            import torch.distributed._symmetric_memory as symm_mem

            # receive area of the push kernel: [2 (epoch parity)][world][numel] 8-byte slots {value, epoch}
            buf = symm_mem.empty(2 * 2 * self.world * numel, dtype=dtype, device=device)
            buf.zero_()
            hdl = symm_mem.rendezvous(buf, self.pg if self.pg is not None else dist.group.WORLD)
            if hdl.signal_pad_size < 2048 or hdl.world_size != self.world:
                return None
            # [epoch, sticky error flag] of the gradient exchange, then of the loss-value exchange: one tensor, one read
            self._flags = torch.zeros(4, dtype=torch.int32, device=device)
            self._p2p = (hdl, self._flags[0:2])
            # a second, tiny receive area for the step's loss VALUES: they exist before the backward pass, so a
            # captured step exchanges them early, on a side branch of its graph (_early_vals)
            n_vals = self.n_loss + 1
            vbuf = symm_mem.empty(2 * 2 * self.world * n_vals, dtype=dtype, device=device)
            vbuf.zero_()
            vhdl = symm_mem.rendezvous(vbuf, self.pg if self.pg is not None else dist.group.WORLD)
            self._p2p_vals = (vhdl, self._flags[2:4], vbuf, n_vals)
            torch.cuda.synchronize(device)
            dist.barrier(self.pg)
            return buf
        except Exception as e:  # pragma: no cover - depends on the box
            import warnings

            warnings.warn(f"peer-memory all-reduce unavailable ({type(e).__name__}: {e}); using NCCL")
            self._p2p = None
            return None

    def _setup_flat(self, n_vals):
        ps = [p for p in self.net.parameters() if p.requires_grad]
        total = sum(p.numel() for p in ps)
        dt = ps[0].dtype
        self._p2p = None
        self._p2p_vals = None
        self._ps = ps
        self._n_grad = total
        self._cap = total + n_vals
        self._recv = self._p2p_buffer(total + n_vals, dt, ps[0].device) if all(p.dtype == dt for p in ps) else None
        # NCCL fallback: gradients and loss values are packed into one flat buffer, one all-reduce
        self._flat = torch.zeros(total + n_vals, dtype=dt, device=ps[0].device) if self._recv is None else self._recv

    # backward writes ordinary per-parameter gradients; ONE cat packs them (and the loss values) into the flat
    # buffer right before the collective, and the parameters' .grad then become views of the reduced buffer
    def _zero_grad(self):
        if self._flat is None:
            self._setup_flat(self.n_loss + 1)
        for p in self._ps:
            p.grad = None

    def _zero_grad_captured(self):
        for p in self._ps:
            p.grad = None

    def _loss_vector(self, inputs, targets):
        if self.shard == "batch" or self.world == 1:
            return super()._loss_vector(inputs, targets)
        M = self.net.nfft // 2 + 1
        b0, b1 = bin_range(M, self.rank, self.world)
        gather = lambda t: _AllGatherBins.apply(t, M, self.rank, self.world, self.pg)  # noqa: E731
        with sweep.bin_shard(b0, b1, gather=gather):
            est, done = self._predict(inputs, targets)  # a fused criterion slices the target itself
        # a prediction that does not hold this rank's bins went through a layer that needs the whole spectrum (iFFT:
        # all-gathered, then evaluated identically on every rank): its criteria are replicated, weight 1 / world
        replicated = est is not None and est.shape[1] != b1 - b0
        tg = targets[:, b0:b1] if (targets.shape[1] == M and not replicated) else targets
        w_pred = 1.0 / self.world if replicated else (b1 - b0) / M
        weight = [1.0 / self.world if self.requires_model[i] else w_pred for i in range(len(self.criterion))]
        return self._criteria(est, done, tg, weight=weight)

    def _early_vals(self, vals):
        """Captured step, peer-memory exchange available: the loss values — final before the backward pass starts — are
        exchanged between the ranks and handed to the host (sweep.NOTIFY_SLOT) by their own launch of the push kernel
        on a SIDE BRANCH of the graph, concurrently with the adjoint maps.  The host has them ~10 us before the step
        ends instead of after the gradient exchange, which is what its round trip (poll -> return -> next enqueue)
        needs to stay off the critical path; `_sync` then exchanges the gradients only."""
        import os

        from . import _lib

        self._vals_exchanged = False
        slot = sweep.NOTIFY_SLOT
        pv = getattr(self, "_p2p_vals", None)
        if (self.world == 1 or slot is None or not slot.get("after_sync") or slot.get("used") is not None or pv is None
                or slot.get("side") is None or vals.dtype != torch.float32 or not vals.is_contiguous()
                or vals.numel() != pv[3] or os.environ.get("FLAMO_B200_EARLY_VALS", "1") == "0"):
            return
        hdl, epoch, _, n_vals = pv
        scale = 1.0 / self.world if self.shard == "batch" else 1.0
        v = vals.detach()
        segs = (_lib.Seg * 1)(_lib.Seg(v.data_ptr(), v.numel()))
        cur, side = torch.cuda.current_stream(v.device), slot["side"]
        side.wait_stream(cur)
        slot["used"] = (v.numel(), v.dtype)
        with torch.cuda.stream(side), torch.cuda.device(v.device):
            _lib.check(_lib.lib().fsweep_allreduce_push_notify(
                segs, 1, hdl.buffer_ptrs_dev, hdl.signal_pad_ptrs_dev, hdl.rank, hdl.world_size, n_vals, scale,
                epoch.data_ptr(), slot["host_vals"].data_ptr(), slot["host_seq"].data_ptr(), slot["counter"].data_ptr(),
                side.cuda_stream))
        v.record_stream(side)
        slot["join"] = True
        sweep.launch_count += 1
        self._vals_exchanged = True

    def _sync(self, vals):
        if self.world == 1:
            return vals
        dt = self._flat.dtype
        scale = 1.0 / self.world if self.shard == "batch" else 1.0
        if self._p2p is not None and vals.dtype == dt:
            # ONE kernel per rank over NVLink peer memory, in place on the tensors autograd left behind: gather the
            # gradients (and the loss values), push them into every peer's receive area, one flag round, local sum in
            # rank order, scale, scatter back (libfsweep fsweep_allreduce_push) — no cat, no views, no second launch
            from . import _lib

            for p in self._ps:
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
                elif not p.grad.is_contiguous():
                    p.grad = p.grad.contiguous()
            vals = vals.contiguous()
            early = getattr(self, "_vals_exchanged", False)  # the values went ahead on their own (_early_vals)
            self._vals_exchanged = False
            tensors = [p.grad for p in self._ps] + ([] if early else [vals])
            segs = (_lib.Seg * len(tensors))(*[_lib.Seg(t.data_ptr(), t.numel()) for t in tensors])
            hdl, epoch = self._p2p
            # captured step: the exchanged loss values go straight to the Trainer's pinned host buffer (sweep.NOTIFY_SLOT)
            slot = sweep.NOTIFY_SLOT
            notify = (not early and slot is not None and slot.get("used") is None and slot.get("after_sync")
                      and vals.dtype == torch.float32 and vals.numel() <= 8 and slot["counter"].device == vals.device)
            if notify:
                slot["used"] = (vals.numel(), vals.dtype)
            with torch.cuda.device(vals.device):
                _lib.check(_lib.lib().fsweep_allreduce_push_notify(
                    segs, len(tensors), hdl.buffer_ptrs_dev, hdl.signal_pad_ptrs_dev, hdl.rank, hdl.world_size, self._cap,
                    scale, epoch.data_ptr(), slot["host_vals"].data_ptr() if notify else None,
                    slot["host_seq"].data_ptr() if notify else None, slot["counter"].data_ptr() if notify else None,
                    torch.cuda.current_stream(vals.device).cuda_stream))
            sweep.launch_count += 1
            return vals
        pieces = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).to(dt) for p in self._ps]
        torch.cat(pieces + [vals.to(dt)], out=self._flat[:self._cap])
        off = 0
        for p in self._ps:
            p.grad = self._flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        flat = self._flat[:self._cap]
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.pg)
        if self.shard == "batch":
            flat.div_(self.world)
        return self._flat[self._n_grad:self._cap].to(vals.dtype)

    def check_exchange(self):
        """Raises if the peer-memory all-reduce ever timed out waiting for a peer (its results are then undefined).
        A host read: call it outside the captured step (train_step does, every `check_every` steps)."""
        if getattr(self, "_p2p", None) is not None:
            f = self._flags.tolist()
            flag = f[1] or f[3]
            if flag:
                raise RuntimeError(f"fsweep_allreduce_p2p: rank {flag - 1} did not arrive at the exchange (timeout); "
                                   "the gradients of that step are invalid")

    check_every = 256

    def _check_exchange_async(self):
        """The sticky error flags of the exchange kernels without a synchronize: every `check_every` steps a non-blocking
        copy brings them into pinned host memory, and the copy issued `check_every` steps EARLIER is looked at (it is
        long complete: later steps' losses have been seen since).  A blocking read here was a 60 - 140 us outlier every 64
        steps of a 50 us step (and two sliced copies were hardly better: the cost is host time inside train_step).  `check_exchange()` is the immediate, blocking form."""
        if getattr(self, "_p2p", None) is None:
            return
        host = getattr(self, "_err_host", None)
        if host is None:
            host = self._err_host = torch.zeros(4, dtype=torch.int32, pin_memory=True)
        else:
            f = host.tolist()
            flag = f[1] or f[3]
            if flag:
                raise RuntimeError(f"fsweep_allreduce_push: rank {flag - 1} did not arrive at the exchange (timeout); "
                                   "the gradients since the previous check are invalid")
        host.copy_(self._flags, non_blocking=True)

    def train_step(self, data):
        out = super().train_step(data)
        self._n_steps = getattr(self, "_n_steps", 0) + 1
        if self._n_steps % self.check_every == 1:
            self._check_exchange_async()
        return out
