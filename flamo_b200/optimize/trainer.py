"""Trainer with the interface of flamo.optimize.trainer.Trainer (reference trainer.py:9-313).

`train_step` is the unit BASELINE.json's metric is quoted on: zero_grad -> net(inputs) -> weighted
criteria -> backward -> Adam step.  On a CUDA device the whole step can be captured once into a CUDA
graph (`graph=True`, the default when the device is CUDA): the step then costs one graph launch, one
host<->device copy of the inputs if they live on the host, and ONE device->host read of the
per-criterion losses (the reference pays one sync per criterion, trainer.py:184-192).  The arithmetic
is identical in both modes.
"""
from __future__ import annotations

import os
import time
import warnings
from typing import Optional

import torch
import torch.nn as nn

try:
    from tqdm import trange
except Exception:  # pragma: no cover
    trange = range


class Trainer:
    def __init__(self, net: nn.Module, max_epochs: int = 10, lr: float = 1e-3, patience: int = 5,
                 patience_delta: float = 0.01, step_size: int = 50, step_factor: float = 0.1, log: bool = True,
                 train_dir: str = None, device: str = "cpu", graph: Optional[bool] = None,
                 fuse_criterion: bool = True):
        self.device = device
        self.fuse_criterion = fuse_criterion  # evaluate a lone MSE criterion inside the sweep kernel when possible
        self._unfusable = set()
        self.log = log
        self.net = net.to(device)
        self.max_epochs, self.lr = max_epochs, lr
        self.patience, self.patience_delta = patience, patience_delta
        self.min_val_loss = float("inf")
        on_cuda = torch.device(device).type == "cuda"
        self.use_graph = on_cuda if graph is None else (graph and on_cuda)
        params = list(self.net.parameters())
        self._all_params = params
        if self.use_graph:
            # a captured step needs `step` and `lr` on the device.  Default: optimize/adam.py, the update of ALL
            # parameters as ONE libfsweep launch (same arithmetic as torch.optim.Adam, tests/test_gpu_adam.py; 1.6 us
            # per captured config-2 step faster than torch's capturable fused Adam, which is two launches: a foreach
            # add on the step counters, then the multi-tensor update).  FLAMO_B200_SWEEP_ADAM=0 selects torch's.
            from .adam import SweepAdam

            lr_t = torch.tensor(float(lr), device=device, dtype=torch.float32)
            if os.environ.get("FLAMO_B200_SWEEP_ADAM", "1") == "1" and SweepAdam.supported(params):
                self.optimizer = SweepAdam([p for p in params], lr=lr_t)
            else:
                self.optimizer = torch.optim.Adam(params, lr=lr_t, capturable=True, fused=True)
        else:
            self.optimizer = torch.optim.Adam(params, lr=self.lr)
        self.n_loss = 0
        if self.log:
            assert os.path.isdir(train_dir), "The directory specified in train_dir does not exist."
        self.train_dir = train_dir
        self.criterion, self.alpha, self.requires_model = [], [], []
        self.scheduler = torch.optim.lr_scheduler.StepLR(self.optimizer, step_size=step_size, gamma=step_factor)
        self.train_loss_log, self.valid_loss_log = {}, {}
        self._graphs = {}
        self._warm = {}

    def register_criterion(self, criterion: nn.Module, alpha: int = 1, requires_model: bool = False):
        self.criterion.append(criterion.to(self.device))
        self.alpha.append(alpha)
        self.requires_model.append(requires_model)
        self.n_loss += 1
        self._graphs.clear()

    # -- loops -----------------------------------------------------------------------------------
    def train(self, train_dataset, valid_dataset):
        self.train_loss, self.valid_loss = [], []
        self.train_loss_log, self.valid_loss_log = {}, {}
        for c in self.criterion:
            self.train_loss_log[c.__class__.__name__] = []
            self.valid_loss_log[c.__class__.__name__] = []
        st = time.time()
        for epoch in trange(self.max_epochs, desc="Training"):
            st_epoch = time.time()
            total = 0
            for data in train_dataset:
                total += self.train_step(data)
            self.scheduler.step()
            self.train_loss.append(total / len(train_dataset))
            total = 0
            for data in valid_dataset:
                total += self.valid_step(data)
            self.valid_loss.append(total / len(valid_dataset))
            self.print_results(epoch, time.time() - st_epoch)
            if self.log:
                self.save_model(epoch)
            if self.early_stop():
                print("Early stopping at epoch: {}".format(epoch))
                break
        print("Training time: {:.3f}s".format(time.time() - st))

    def move_to_device(self, data):
        if isinstance(data, list):
            return [x.to(self.device) for x in data]
        return data.to(self.device)

    def _log(self, log, values):
        for c, v in zip(self.criterion, values):
            log.setdefault(c.__class__.__name__, []).append(v)

    # -- forward + criteria --------------------------------------------------------------------------
    def _fused_kind(self, crit):
        """ABI criterion kind when `crit` is one the sweep kernels evaluate themselves, else None."""
        from .._lib import CRIT_MSE, CRIT_MSE_CHSUM
        from .loss import mse_loss

        if type(crit) is mse_loss:
            return CRIT_MSE_CHSUM
        if type(crit) is nn.MSELoss and crit.reduction == "mean":
            return CRIT_MSE
        return None

    def _predict(self, inputs, targets):
        """Returns (est, done): the network output for the criteria and a dict {criterion index: loss tensor} of
        criteria already evaluated.  When exactly ONE registered criterion consumes the prediction and it is an MSE
        the kernels can fuse with the |.| output layer, est is None: that criterion comes out of the sweep itself,
        and the model-only criteria are evaluated first (they share the memoised parameter maps with the sweep)."""
        users = [i for i, c in enumerate(self.criterion) if getattr(c, "uses_prediction", True)]
        if self.fuse_criterion and len(users) == 1 and hasattr(self.net, "forward_loss"):
            i = users[0]
            kind = self._fused_kind(self.criterion[i])
            key = (i, tuple(inputs.shape), tuple(targets.shape))
            if kind is not None and not self.requires_model[i] and key not in self._unfusable:
                inval = getattr(self.net, "_invalidate_caches", None)
                if inval is not None:
                    inval()
                done = {j: (c(None, targets, self.net) if self.requires_model[j] else c(None, targets))
                        for j, c in enumerate(self.criterion) if j != i}
                loss = self.net.forward_loss(inputs, targets, kind, keep_caches=True)
                if loss is not None:
                    done[i] = loss
                    return None, done
                self._unfusable.add(key)
                return self.net(inputs), done
        return self.net(inputs), {}

    def _criteria(self, est, done, targets, weight=None):
        """Evaluate the remaining criteria and combine: returns the vector [part_0, ..., part_{n-1}, total] with
        total = sum_i alpha_i part_i (reference trainer.py:184-188).  `weight[i]` (multi-GPU shard weights) scales
        part_i.  On CUDA the combination is one kernel (sweep.WeightedTotal)."""
        from .. import sweep

        parts = []
        for i, (crit, needs_model) in enumerate(zip(self.criterion, self.requires_model)):
            if i in done:
                t = done[i]
            else:
                t = crit(est, targets, self.net) if needs_model else crit(est, targets)
            parts.append(t)
        scales = [1.0] * len(parts) if weight is None else [float(w) for w in weight]
        self._parts = None
        if sweep.WeightedTotal.supported(parts):
            if torch.is_grad_enabled() and all(p.requires_grad for p in parts):
                # the combination is linear with CONSTANT weights: evaluate it outside autograd and let _train_core
                # seed every criterion with its own constant d total / d part_i = alpha_i * scale_i (no backward
                # kernels for the combination, and a weight of 1 costs the criterion's backward nothing)
                self._parts = (parts, [float(a) * s for a, s in zip(self.alpha, scales)])
                with torch.no_grad():
                    return sweep.WeightedTotal.apply(tuple(self.alpha), tuple(scales), *[p.detach() for p in parts])
            return sweep.WeightedTotal.apply(tuple(self.alpha), tuple(scales), *parts)
        total = None
        for i, t in enumerate(parts):
            if scales[i] != 1.0:
                t = parts[i] = t * scales[i]
            term = t if self.alpha[i] == 1 else self.alpha[i] * t
            total = term if total is None else total + term
        return torch.stack([p.reshape(()) for p in parts] + [total.reshape(())])

    def _loss_vector(self, inputs, targets):
        """Forward + criteria: [per-criterion values ..., weighted total] as one tensor (differentiable)."""
        est, done = self._predict(inputs, targets)
        return self._criteria(est, done, targets)

    def _losses(self, inputs, targets):
        """(total, [per-criterion tensors]) — kept for callers of the earlier interface."""
        vals = self._loss_vector(inputs, targets)
        return vals[-1], list(vals[:-1].unbind(0))

    # -- one optimisation step, as a pure device-side function (captured or eager) --------------------
    def _zero_grad(self):
        self.optimizer.zero_grad(set_to_none=True)

    def _sync(self, vals):
        """Hook between backward and the optimizer step (multi-GPU trainers all-reduce here)."""
        return vals

    def _early_vals(self, vals):
        """Hook between the criteria and the backward pass (multi-GPU trainers may exchange the loss values here)."""

    def _train_core(self, inputs, targets):
        from .. import sweep

        vals = self._loss_vector(inputs, targets)
        self._early_vals(vals)
        if getattr(self, "_parts", None) is not None:
            parts, weights = self._parts
            self._parts = None
            seeds = [sweep._const_tensor((w,), p.dtype, p.device).reshape(p.shape) for p, w in zip(parts, weights)]
            torch.autograd.backward(parts, grad_tensors=seeds)
        else:
            # d total / d .: seed the backward pass on the vector itself with a constant one-hot (no select / fill kernels)
            seed = sweep._const_tensor((0.0,) * (vals.numel() - 1) + (1.0,), vals.dtype, vals.device)
            torch.autograd.backward(vals, grad_tensors=seed)
        vals = self._sync(vals.detach())
        if type(self)._sync is not Trainer._sync:
            sweep.notify_values(vals)  # (captured multi-GPU step: the exchanged values ride with the optimizer's launch)
        self.optimizer.step()
        sweep.flush_deferred_total()  # (a deferred criteria total nobody picked up)
        return vals

    def _eager_train_step(self, inputs, targets):
        self._zero_grad()
        vals = self._train_core(inputs, targets).tolist()
        self._log(self.train_loss_log, vals[:-1])
        return vals[-1]

    # -- captured step (CUDA) --------------------------------------------------------------------
    def _graph_key(self, inputs, targets):
        """Shapes + the version counters of the FROZEN parameters: their mapped coefficients are baked into a captured
        step (DSP._coef_cache), so assign_value / load_state_dict on a frozen module must lead to a new capture.
        (Trainable parameters are read from their own storage at every replay.)"""
        frozen = tuple(p._version for p in self._all_params if not p.requires_grad)
        return (tuple(inputs.shape), inputs.dtype, tuple(targets.shape), targets.dtype, frozen)

    def _graph_for(self, inputs, targets):
        key = self._graph_key(inputs, targets)
        g = self._graphs.get(key)
        if g is not None:
            return g
        n_warm = self._warm.get(key, 0)
        if n_warm < 3:  # eager warm-up: lazy state (Adam moments, plans, cuFFT plans, caches) must exist
            self._warm[key] = n_warm + 1
            return None
        from .. import sweep

        static_in, static_tg = inputs.clone(), targets.clone()
        graph = torch.cuda.CUDAGraph()
        self._zero_grad()
        n0 = sweep.launch_count
        # Loss read-back without a host round trip at the END of the step: the criteria-total kernel of the captured
        # step stores the losses straight into mapped pinned host memory with a sequence number
        # (fsweep_weighted_total_notify); train_step polls that number and returns the loss while the adjoint of the
        # maps and the optimizer are still running (stream-ordered for everything that follows).  A trainer whose _sync
        # exchanges the values between ranks gets them after the exchange, as a rider of the optimizer's launch
        # (sweep.notify_values): no copy node and no stream synchronize there either.
        slot = None
        if os.environ.get("FLAMO_B200_NOTIFY", "1") != "0":
            slot = {"after_sync": type(self)._sync is not Trainer._sync,"host_vals": torch.zeros(8 * 16, dtype=torch.uint8, pin_memory=True),
                    "host_seq": torch.zeros(1, dtype=torch.int32, pin_memory=True),
                    "counter": torch.zeros(1, dtype=torch.int32, device=static_in.device), "expected": 0,
                    "defer": os.environ.get("FLAMO_B200_DEFER_TOTAL", "0") != "0",
                    "side": torch.cuda.Stream(static_in.device)
                    if (os.environ.get("FLAMO_B200_TOTAL_BRANCH", "0") != "0"
                        or type(self)._sync is not Trainer._sync) else None}
        try:
            sweep.NOTIFY_SLOT = slot
            with torch.cuda.graph(graph):
                self._zero_grad_captured()
                out = self._train_core(static_in, static_tg)
                if slot is not None and slot.get("join"):
                    torch.cuda.current_stream(static_in.device).wait_stream(slot["side"])
                notified = slot is not None and slot.get("used") == (out.numel(), out.dtype)
                out_host = None
                if not notified:
                    # the step's single device->host read is a node of the graph: the losses land in pinned host memory
                    out_host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
                    out_host.copy_(out, non_blocking=True)
        except Exception as e:  # keep training eagerly (still on the CUDA sweep) if capture is impossible
            warnings.warn(f"CUDA-graph capture of the training step failed ({e}); continuing without graph.")
            self.use_graph = False
            torch.cuda.synchronize()
            return None
        finally:
            sweep.NOTIFY_SLOT = None
        if notified:
            slot["seq_np"] = slot["host_seq"].numpy()
            if out.dtype == torch.float32:  # {value, launch number} pairs, see fsweep_weighted_total_notify
                slot["vals_np"] = slot["host_vals"].view(torch.float32)[:2 * out.numel():2].numpy()
                slot["pair_np"] = slot["host_vals"].view(torch.int32)[1:2 * out.numel():2].numpy()
            else:
                slot["vals_np"] = slot["host_vals"].view(out.dtype)[:out.numel()].numpy()
                slot["pair_np"] = None
        else:
            slot = None
        g = (graph, static_in, static_tg, out_host, sweep.launch_count - n0, [None, None], slot)  # sweep kernels per replay
        self._graphs[key] = g
        return g

    _UPLOAD_MAX_BYTES = 4 << 20

    def _upload(self, pairs):
        """Pinned host tensors -> the captured step's static buffers as ONE kernel reading the host memory over PCIe
        (libfsweep fsweep_upload) instead of one DMA copy per tensor; False: the caller copies the ordinary way (pageable
        or non-contiguous sources, dtype / shape mismatches, big batches — the DMA engine is the better mover there)."""
        if os.environ.get("FLAMO_B200_UPLOAD_KERNEL", "1") == "0" or len(pairs) > 4:
            return False
        for dst, src in pairs:
            if (src.is_cuda or not src.is_pinned() or not src.is_contiguous() or not dst.is_contiguous()
                    or src.dtype != dst.dtype or src.shape != dst.shape
                    or src.numel() * src.element_size() > self._UPLOAD_MAX_BYTES or src.numel() == 0):
                return False
        import ctypes as C

        from .. import _lib, sweep

        n = len(pairs)
        srcs = (C.c_void_p * n)(*[s.data_ptr() for _, s in pairs])
        dsts = (C.c_void_p * n)(*[d.data_ptr() for d, _ in pairs])
        nbytes = (C.c_int64 * n)(*[s.numel() * s.element_size() for _, s in pairs])
        dev = pairs[0][0].device
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().fsweep_upload(srcs, dsts, nbytes, n, torch.cuda.current_stream(dev).cuda_stream))
        sweep.launch_count += 1
        return True

    def _await_losses(self, slot, device):
        """Spin on the sequence number the captured step writes behind its losses (pinned host memory)."""
        slot["expected"] += 1
        want, seq = slot["expected"], slot["seq_np"]
        spins = 0
        while seq[0] != want:
            spins += 1
            if spins == 2000000:  # ~ a second: something is wrong (or a debugger holds the GPU); stop spinning blind
                torch.cuda.current_stream(device).synchronize()
                if seq[0] != want:
                    raise RuntimeError(f"captured step: loss notification {int(seq[0])} != expected {want}")
        pair = slot["pair_np"]
        if pair is not None:  # float32: every value carries the launch number it belongs to (no fence on the device)
            while not (pair == want).all():
                spins += 1
                if spins > 4000000:
                    raise RuntimeError("captured step: loss values did not arrive")
        return slot["vals_np"].tolist()

    def _zero_grad_captured(self):
        """Inside the captured region: nothing to do when grads are (re)allocated by backward."""

    def train_step(self, data):
        inputs, targets = data
        if self.use_graph:
            # a captured step exists for these shapes: copy straight into its static buffers (from pinned host
            # memory this is ONE asynchronous H2D copy per tensor — no intermediate device tensor, no host sync)
            key = self._graph_key(inputs, targets)
            g = self._graphs.get(key)
            if g is None:
                inputs = self.move_to_device(inputs)
                targets = self.move_to_device(targets)
                g = self._graph_for(inputs, targets)
            if g is not None:
                from .. import sweep

                graph, static_in, static_tg, out_host, n_kernels, last, notify = g
                # A DEVICE tensor that is the very tensor (storage, version) copied in by the previous step is already
                # in the static buffer: a dataset resident in HBM costs no copy per step.  Host tensors are copied
                # every step (from pinned memory: ONE asynchronous H2D copy per tensor).
                pending = []
                for slot, (dst, src) in enumerate(((static_in, inputs), (static_tg, targets))):
                    tag = (src.data_ptr(), src._version, src.device) if src.is_cuda else None
                    if tag is None or last[slot] != tag:
                        pending.append((dst, src))
                        last[slot] = tag
                if pending and not self._upload(pending):
                    for dst, src in pending:
                        dst.copy_(src, non_blocking=True)
                graph.replay()
                sweep.launch_count += n_kernels
                inval = getattr(self.net, "_invalidate_caches", None)
                if inval is not None:  # parameters changed on the device without a version bump
                    inval()
                if notify is not None:
                    vals = self._await_losses(notify, static_in.device)  # the rest of the step is still in flight
                else:
                    torch.cuda.current_stream(static_in.device).synchronize()
                    vals = out_host.tolist()  # written by the graph's own device->host copy node
                self._log(self.train_loss_log, vals[:-1])
                return vals[-1]
            return self._eager_train_step(inputs, targets)
        return self._eager_train_step(self.move_to_device(inputs), self.move_to_device(targets))

    @torch.no_grad()
    def valid_step(self, data):
        inputs, targets = data
        inputs = self.move_to_device(inputs)
        targets = self.move_to_device(targets)
        vals = self._loss_vector(inputs, targets).tolist()
        self._log(self.valid_loss_log, vals[:-1])
        return vals[-1]

    # -- bookkeeping -----------------------------------------------------------------------------
    def print_results(self, e: int, e_time: float):
        print(get_str_results(epoch=e, train_loss=self.train_loss, valid_loss=self.valid_loss, time=e_time))

    def get_train_dir(self):
        if self.train_dir is not None:
            os.makedirs(self.train_dir, exist_ok=True)
        else:
            self.train_dir = os.path.join("output", time.strftime("%Y%m%d-%H%M%S"))
            os.makedirs(self.train_dir)

    def save_model(self, e: int):
        d = os.path.join(self.train_dir, "checkpoints")
        os.makedirs(d, exist_ok=True)
        torch.save(self.net.state_dict(), os.path.join(d, "model_e" + str(e) + ".pt"))

    def early_stop(self):
        if self.valid_loss[-1] < (self.min_val_loss - self.patience_delta):
            self.min_val_loss = self.valid_loss[-1]
            self.counter = 0
        elif (self.min_val_loss - self.patience_delta) < self.valid_loss[-1] < (self.min_val_loss + self.patience_delta):
            self.counter += 1
            if self.counter >= self.patience:
                return True
        return False


def get_str_results(epoch=None, train_loss=None, valid_loss=None, time=None):
    s = ""
    if epoch is not None:
        s += "epoch: {:3d} ".format(epoch)
    if train_loss is not None:
        s += "- train_loss: {:6.4f} ".format(train_loss[-1])
    if valid_loss is not None:
        s += "- test_loss: {:6.4f} ".format(valid_loss[-1])
    if time is not None:
        s += "- time: {:6.4f} s".format(time)
    return s
