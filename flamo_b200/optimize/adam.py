"""Adam with torch.optim.Adam's arithmetic (the reference Trainer's optimizer, flamo/optimize/trainer.py:42) as ONE
kernel launch for all parameters (libfsweep fsweep_adam_step).  Used by the Trainer for captured CUDA steps, where the
parameter-sized kernels are paid per launch; learning rate and step counters live on the device, so a captured step can
be replayed and lr schedulers (which `fill_` a tensor learning rate in place) keep working."""
from __future__ import annotations

import torch

from .. import _lib, sweep

MAX_NUMEL = 1 << 16  # one block per tensor: larger parameters are better served by torch's multi-tensor kernel


class SweepAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        params = list(params)
        dev = params[0].device
        if not torch.is_tensor(lr):
            lr = torch.tensor(float(lr), device=dev, dtype=torch.float32)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @staticmethod
    def supported(params) -> bool:
        params = [p for p in params if p.requires_grad]
        return (0 < len(params) <= _lib.ADAM_MAX_TENSORS and sweep._BACKEND.name == "cuda"
                and all(p.is_cuda and p.dtype == params[0].dtype and p.device == params[0].device
                        and p.is_contiguous() and p.numel() <= MAX_NUMEL for p in params)
                and params[0].dtype in (torch.float32, torch.float64))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            entries = []
            dtype = None
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                dtype = p.dtype
                entries.append(_lib.AdamTensor(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(),
                                               st["exp_avg_sq"].data_ptr(), st["step"].data_ptr(), p.numel()))
            if not entries:
                continue
            arr = (_lib.AdamTensor * len(entries))(*entries)
            lr = group["lr"]
            b1, b2 = group["betas"]
            import ctypes as C

            taken = None
            slot = sweep.NOTIFY_SLOT
            if slot is not None and slot.get("adam_rider"):  # exchanged loss values of a captured multi-GPU step
                taken = sweep.take_deferred_total(dtype, lr.device)
            sweep.flush_deferred_total()  # (a captured step's criteria total that no backward kernel carried)
            with torch.cuda.device(lr.device):
                _lib.check(_lib.lib().fsweep_adam_step_total(arr, len(entries),
                                                             _lib.C64 if dtype == torch.float32 else _lib.C128,
                                                             lr.data_ptr(), float(b1), float(b2), float(group["eps"]),
                                                             C.byref(taken[0]) if taken is not None else None,
                                                             torch.cuda.current_stream(lr.device).cuda_stream))
            sweep.launch_count += 1
        return loss
