"""Lowering of a flamo module tree to a flat sweep program + the autograd seam over libfsweep.

    module tree ──_lower()──► Program(ops, coefficient tensors) ──run()──► SweepFunction
                                                                          │ forward  → fsweep_forward
                                                                          │ backward → fsweep_backward

The coefficient tensors are produced by the modules' own `map`s in plain PyTorch (O(#params),
differentiable, user-overridable — SURVEY.md §7); everything that is O(M) happens inside the two
C-ABI calls.  Gradients come back w.r.t. the coefficients and autograd chains them through the maps.
"""
from __future__ import annotations

import os
import threading
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import (CRIT_MSE_CHSUM, EPI_ABS, EPI_NONE, F_GRAD, F_ISINT, OP_DELAY, OP_GAIN, OP_PDELAY, OP_PGAIN, OP_PSOS, OP_PTABLE,
                   OP_RECURSION, OP_SOS, OP_TABLE, Op, Plan)

MAX_OPS_PER_LAUNCH = 24
_DIAG = (OP_PGAIN, OP_PSOS, OP_PDELAY, OP_PTABLE)
_TABLE = (OP_TABLE, OP_PTABLE)

# ---------------------------------------------------------------------------------------------
# bin sharding (multi-GPU): inside `bin_shard(begin, end)` every sweep processes only that range
# of rFFT bins and returns tensors with end-begin bins.
_tls = threading.local()


class bin_shard:
    """`gather` (optional): callable that turns a tensor holding this rank's bins (B, end - begin, ...) into the full
    (B, M, ...) tensor on every rank — handed in by the multi-GPU trainer; the layers that need ALL bins (iFFT,
    iFFTAntiAlias: SURVEY.md §8e last row) call it through `gather_bins`."""

    def __init__(self, begin: int, end: int, gather=None):
        self.range = (int(begin), int(end))
        self.gather = gather

    def __enter__(self):
        self.prev = (getattr(_tls, "shard", None), getattr(_tls, "gather", None))
        _tls.shard, _tls.gather = self.range, self.gather
        return self

    def __exit__(self, *a):
        _tls.shard, _tls.gather = self.prev


def gather_bins(x: torch.Tensor, nfft: int) -> torch.Tensor:
    """Inside a bin shard, a bin-domain tensor that holds only this rank's bins becomes the full spectrum (all-gather
    over the ranks of the shard); anywhere else the tensor is returned as it is.  A layer that needs every bin and finds
    no way to get them raises instead of transforming a fragment of the spectrum."""
    shard = current_shard()
    M = nfft // 2 + 1
    if shard is None or x.shape[1] == M or x.shape[1] != shard[1] - shard[0]:
        return x
    gather = getattr(_tls, "gather", None)
    if gather is None:
        raise RuntimeError(f"a layer that needs all {M} bins (iFFT / iFFTAntiAlias) was given the {x.shape[1]} bins of "
                           f"the bin shard {shard}: run it under flamo_b200.parallel.DataParallelTrainer(shard='bins') "
                           "(which all-gathers the spectrum) or outside sweep.bin_shard")
    return gather(x)


def current_shard() -> Optional[Tuple[int, int]]:
    return getattr(_tls, "shard", None)


# ---------------------------------------------------------------------------------------------
class CudaBackend:
    """The product backend: ctypes calls into libfsweep.so on the current CUDA stream."""

    name = "cuda"

    def plan(self, ops: Sequence[tuple], nfft: int, alias_decay_db: float, dtype: int):
        return Plan([Op(*o) for o in ops], nfft, alias_decay_db, dtype)

    @staticmethod
    def _stream(t: torch.Tensor):
        return torch.cuda.current_stream(t.device).cuda_stream

    # Every library call runs with the TENSOR's device current: the kernels, the occupancy queries and the internal
    # memsets all act on the process's current device, which need not be the model's (Trainer(device="cuda:1")
    # without torch.cuda.set_device).  All GPUs of a node are identical, so a plan's cached launch geometry holds
    # for each of them.
    def forward(self, plan, ops, coefs, x, y, cols, bin_begin, epilogue):
        with torch.cuda.device(x.device):
            return self._forward(plan, ops, coefs, x, y, cols, bin_begin, epilogue)

    def backward(self, plan, ops, coefs, x, gy, grads, gx, cols, bin_begin, epilogue):
        with torch.cuda.device(x.device):
            return self._backward(plan, ops, coefs, x, gy, grads, gx, cols, bin_begin, epilogue)

    def loss(self, plan, ops, coefs, x, target, kind, scale, loss, grads, gx, bin_begin):
        with torch.cuda.device(x.device):
            return self._loss(plan, ops, coefs, x, target, kind, scale, loss, grads, gx, bin_begin)

    def _forward(self, plan, ops, coefs, x, y, cols, bin_begin, epilogue):
        B, nb = x.shape[0], x.shape[1]
        return plan.forward([c.data_ptr() for c in coefs], x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0), B,
                            cols, bin_begin, nb, epilogue, self._stream(x))

    def _backward(self, plan, ops, coefs, x, gy, grads, gx, cols, bin_begin, epilogue):
        B, nb = x.shape[0], x.shape[1]
        ws_bytes = plan.workspace_bytes(B, cols, nb)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        return plan.backward([c.data_ptr() for c in coefs], x.data_ptr(), x.stride(0), gy.data_ptr(), gy.stride(0),
                             [g.data_ptr() if g is not None else None for g in grads],
                             gx.data_ptr() if gx is not None else None, gx.stride(0) if gx is not None else 0, B,
                             cols, bin_begin, nb, epilogue, ws.data_ptr(), ws_bytes, self._stream(x))


    def _loss(self, plan, ops, coefs, x, target, kind, scale, loss, grads, gx, bin_begin):
        """Fused |.| + MSE criterion.  grads is None: loss only (one forward launch); else loss and its
        gradients from ONE backward launch (+ the finalize kernel)."""
        B, nb = x.shape[0], x.shape[1]
        ws_bytes = plan.workspace_bytes(B, 1, nb)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        crit = _lib.Criterion(kind, 0, target.data_ptr(), target.stride(0), float(scale), loss.data_ptr())
        cp = [c.data_ptr() for c in coefs]
        if grads is None:
            return plan.forward_loss(cp, x.data_ptr(), x.stride(0), crit, B, bin_begin, nb, ws.data_ptr(), ws_bytes,
                                     self._stream(x))
        return plan.backward_loss(cp, x.data_ptr(), x.stride(0), crit,
                                  [g.data_ptr() if g is not None else None for g in grads],
                                  gx.data_ptr() if gx is not None else None, gx.stride(0) if gx is not None else 0, B,
                                  bin_begin, nb, ws.data_ptr(), ws_bytes, self._stream(x))


_BACKEND = CudaBackend()  # tests/ may swap this for a CPU emulator to exercise the host logic without a GPU
_PLANS = {}
_PLANS_LOCK = threading.Lock()
launch_count = 0  # kernels enqueued by this process through the sweep (bench bookkeeping)


def _get_plan(ops: Tuple[tuple, ...], nfft: int, alias_decay_db: float, dtype: int):
    key = (_BACKEND.name, ops, int(nfft), float(alias_decay_db), dtype)
    with _PLANS_LOCK:
        p = _PLANS.get(key)
        if p is None:
            p = _BACKEND.plan(ops, nfft, alias_decay_db, dtype)
            _PLANS[key] = p
        return p


def _batch_view(t: torch.Tensor) -> torch.Tensor:
    """Kernel layout: dims 1.. contiguous, arbitrary batch stride."""
    if t.shape[0] == 1 or t[0].is_contiguous():
        if t.shape[0] == 1 and not t[0].is_contiguous():
            return t.contiguous()
        return t
    return t.contiguous()


class SweepFunction(torch.autograd.Function):
    """y = program(x); x: (B, n_bins, N_in, cols) complex, already restricted to the processed bins."""

    @staticmethod
    def forward(ctx, x, plan, ops, epilogue, bin_begin, n_out, *coefs):
        global launch_count
        x = _batch_view(x)
        B, nb, _, cols = x.shape
        real = torch.float32 if x.dtype == torch.complex64 else torch.float64
        y = torch.empty((B, nb, n_out, cols), dtype=real if epilogue == EPI_ABS else x.dtype, device=x.device)
        coefs = tuple(c.contiguous() for c in coefs)
        if nb > 0:
            launch_count += _BACKEND.forward(plan, ops, coefs, x, y, cols, bin_begin, epilogue) or 0
        ctx.plan, ctx.ops, ctx.epilogue, ctx.bin_begin = plan, ops, epilogue, bin_begin
        ctx.save_for_backward(x, *coefs)
        return y

    @staticmethod
    def backward(ctx, gy):
        global launch_count
        x, *coefs = ctx.saved_tensors
        B, nb, _, cols = x.shape
        gy = _batch_view(gy)
        need = ctx.needs_input_grad
        leaf_ops = [o for o in ctx.ops if o[0] != OP_RECURSION]
        grads: List[Optional[torch.Tensor]] = []
        for i, c in enumerate(coefs):
            if need[6 + i] and (leaf_ops[i][4] & F_GRAD):
                # TABLE gradients are only written inside the processed bin range
                grads.append(torch.zeros_like(c) if leaf_ops[i][0] in _TABLE else torch.empty_like(c))
            else:
                grads.append(None)
        gx = torch.empty_like(x, memory_format=torch.contiguous_format) if need[0] else None
        if gx is None and all(g is None for g in grads):
            # nothing to compute: e.g. an external integer-delay tensor that requires grad (its gradient through
            # `round` is zero; the reference returns zeros, autograd treats None the same way)
            return (None,) * (6 + len(coefs))
        if nb > 0:
            launch_count += _BACKEND.backward(ctx.plan, ctx.ops, coefs, x, gy, grads, gx, cols, ctx.bin_begin,
                                              ctx.epilogue) or 0
        else:
            grads = [None if g is None else torch.zeros_like(g) for g in grads]
        return (gx, None, None, None, None, None, *grads)


class SweepLossFunction(torch.autograd.Function):
    """loss = criterion(|program(x)|, target) with the output layer and the criterion evaluated inside the sweep
    kernel (include/fsweep.h, fsweep_*_loss).  When a gradient will be needed, the forward call already runs the
    backward kernel — it recomputes the forward states anyway — and keeps d loss / d coefficients; autograd's
    backward only scales them by the upstream gradient.  No forward sweep, no |Y| / dL/d|Y| round trip."""

    @staticmethod
    def forward(ctx, x, target, plan, ops, kind, scale, bin_begin, *coefs):
        global launch_count
        x = _batch_view(x)
        target = _batch_view(target)
        real = torch.float32 if x.dtype == torch.complex64 else torch.float64
        coefs = tuple(c.contiguous() for c in coefs)
        loss = torch.empty((), dtype=real, device=x.device)
        need = ctx.needs_input_grad
        leaf_ops = [o for o in ops if o[0] != OP_RECURSION]
        want = [need[7 + i] and bool(leaf_ops[i][4] & F_GRAD) for i in range(len(coefs))]
        if not (any(want) or need[0]):
            launch_count += _BACKEND.loss(plan, ops, coefs, x, target, kind, scale, loss, None, None, bin_begin) or 0
            return loss
        grads = [(torch.zeros_like(c) if leaf_ops[i][0] in _TABLE else torch.empty_like(c)) if want[i] else None
                 for i, c in enumerate(coefs)]
        gx = torch.empty_like(x, memory_format=torch.contiguous_format) if need[0] else None
        launch_count += _BACKEND.loss(plan, ops, coefs, x, target, kind, scale, loss, grads, gx, bin_begin) or 0
        ctx.grads, ctx.gx = grads, gx
        return loss

    @staticmethod
    def backward(ctx, g):
        if g.data_ptr() == _const_tensor((1.0,), g.dtype, g.device).data_ptr():
            # the Trainer seeds every criterion with the cached constant alpha_i: for alpha = 1 (the usual weight of the
            # prediction criterion) the kernel's gradients are already d total / d coefficients
            return (ctx.gx, None, None, None, None, None, None, *ctx.grads)
        live = [t for t in ctx.grads if t is not None] + ([ctx.gx] if ctx.gx is not None else [])
        by_dtype = {}
        for t in live:
            by_dtype.setdefault(t.dtype, []).append(t)
        scaled = {}
        for dt, ts in by_dtype.items():
            gs = g.to(dt) if not dt.is_complex else g.to(torch.float32 if dt == torch.complex64 else torch.float64)
            for t, r in zip(ts, torch._foreach_mul(ts, gs)):
                scaled[id(t)] = r
        pick = lambda t: None if t is None else scaled[id(t)]
        return (pick(ctx.gx), None, None, None, None, None, None, *[pick(t) for t in ctx.grads])


def _cast(coef: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """Coefficient in the dtype a launch reads.  Frozen coefficients (no autograd history: DSP._coef_cache keeps
    them across steps) remember their casts, so a frozen module still costs no launch per step."""
    if coef.dtype == dtype:
        return coef
    if coef.requires_grad:
        return coef.to(dtype)
    memo = coef.__dict__.setdefault("_fsweep_cast", {})
    t = memo.get(dtype)
    if t is None:
        t = memo[dtype] = coef.to(dtype)
    return t


def _wants_f64(ops) -> bool:
    """float32 models whose response float32 ARITHMETIC cannot hold to the 1e-4 magnitude bar are
    swept by the float64 kernels, with the signal converted on the way in and out (the coefficients come out of the
    maps in float64 anyway).  Two shapes qualify, both measured (tests/test_numerics_model_cpu.py):
      * a closed loop wider than 8 channels (and <= 32, the float64 kernels' limit): the output is a sum of > 8
        solved channels that cancel down to the metric's floor (config 4: 1.3e-4 in float32);
      * a dense cascade of many second-order sections feeding a wide sum (config 3: 16 x 30 sections per output):
        the float32 rounding of the packed section coefficients ALONE is 1.1e-4 of the floor there.
    FLAMO_B200_PRECISION=float32 keeps float32 arithmetic everywhere, =float64 promotes every launch."""
    mode = os.environ.get("FLAMO_B200_PRECISION", "auto")
    if mode == "float32":
        return False
    if mode == "float64":
        return all(o[1] <= 32 and o[2] <= 32 for o in ops)
    if any(o[1] > 32 or o[2] > 32 for o in ops):
        return False
    for o in ops:
        if o[0] == OP_RECURSION and o[1] > 8:
            return True
        if o[0] == OP_SOS and o[2] >= 8 and o[2] * o[3] >= 128:
            return True
    return False


# ---------------------------------------------------------------------------------------------
class Program:
    """Flat sweep program under construction.  Leaves are (kind, n_out, n_in, K, flags, 0, 0, 0) tuples
    paired with a coefficient tensor; a recursion is a marker tuple followed by its two chains."""

    def __init__(self, nfft: int, alias_decay_db: float, cdtype: torch.dtype, device):
        self.nfft, self.alias_decay_db = int(nfft), float(alias_decay_db)
        self.cdtype = cdtype
        self.real = torch.float32 if cdtype == torch.complex64 else torch.float64
        self.device = device
        self.items: List[tuple] = []  # ("leaf", op, coef) | ("rec", n_out, n_in, [leaf...], [leaf...]) | ("eager", fn)
        self._chain: Optional[list] = None

    # -- construction ------------------------------------------------------------------------
    def leaf(self, kind: int, n_out: int, n_in: int, coef: torch.Tensor, K: int = 0, isint: bool = False):
        want = kind in (OP_GAIN, OP_PGAIN, OP_SOS, OP_PSOS, OP_TABLE, OP_PTABLE) or not isint
        flags = (F_ISINT if isint else 0) | (F_GRAD if (coef.requires_grad and torch.is_grad_enabled() and want) else 0)
        # coefficients stay in the precision their map produced (float64 for every non-identity map); the cast to
        # the arithmetic type of the launch happens in flatten_segment, once that type is known (_wants_f64)
        if kind in (OP_DELAY, OP_PDELAY):
            coef = coef.to(torch.float64)
        item = ("leaf", (kind, int(n_out), int(n_in), int(K), flags, 0, 0, 0), coef)
        (self._chain if self._chain is not None else self.items).append(item)

    def leaf_items(self, op: tuple, coefs: torch.Tensor):
        """One op with ONE COEFFICIENT SET PER BATCH ITEM (fsweep_op_t::per_item; hyper-conditioning through a batched
        ext_param): `op` is the tuple leaf() built for a single set, coefs is (B, *set shape).  The launch runs the
        generic kernels with a grid slice per item; the gradient comes back per item, (B, *set shape)."""
        kind, n_out, n_in, K, flags = op[:5]
        want = kind in (OP_GAIN, OP_PGAIN, OP_SOS, OP_PSOS, OP_TABLE, OP_PTABLE) or not (flags & F_ISINT)
        flags = (flags & ~F_GRAD) | (F_GRAD if (coefs.requires_grad and torch.is_grad_enabled() and want) else 0)
        item = ("leaf", (kind, n_out, n_in, K, flags, 0, 0, 1), coefs)
        (self._chain if self._chain is not None else self.items).append(item)

    def items_target(self) -> list:
        """The list new leaves are appended to (the open recursion chain, or the top level)."""
        return self._chain if self._chain is not None else self.items

    def table_of(self, module, ext_param=None):
        """Fallback for a bin-wise LINEAR sub-system the flat program cannot express where it stands — a Recursion or a
        Parallel inside a Recursion path (the reference nests them freely, system.py:397-425 applies the feedback path to
        an identity signal and never looks inside).  Its response is measured the way the reference does it: the module
        is run on the identity signal (its own sweep launch: (1, M, n_in, n_in) in, (1, M, n_out, n_in) out) and enters
        this program as a streamed per-bin TABLE, gradients flowing back through that launch.  Costs one launch and
        8 M n_out n_in bytes of HBM per nested system; tables are indexed by absolute bin, so it is evaluated on all
        bins even inside a bin shard."""
        n_in, n_out = int(module.input_channels), int(module.output_channels)
        M = self.nfft // 2 + 1
        eye = torch.eye(n_in, dtype=self.cdtype, device=self.device).expand(1, M, n_in, n_in)
        prev = getattr(_tls, "shard", None)
        _tls.shard = None
        try:
            H = module(eye, ext_param) if ext_param is not None else module(eye)
        finally:
            _tls.shard = prev
        self.leaf(OP_TABLE, n_out, n_in, H[0])

    def eager(self, fn):
        """A module the sweep cannot express: run it in PyTorch between two launches."""
        if self._chain is not None:
            raise _lib.Unsupported(_lib.E_UNSUPPORTED, "non-DSP module inside a Recursion path")
        self.items.append(("eager", fn))

    def recursion(self, lower_ff, lower_fb):
        if self._chain is not None:  # (system.Recursion._lower routes a nested loop through table_of)
            raise _lib.Unsupported(_lib.E_UNSUPPORTED, "nested Recursion")
        self._chain = ff = []
        lower_ff()
        self._chain = fb = []
        lower_fb()
        self._chain = None
        if not ff or not fb:
            raise ValueError("Recursion needs non-empty feedforward and feedback paths")
        self.items.append(("rec", ff[-1][1][1], ff[0][1][2], ff, fb))

    # -- execution ---------------------------------------------------------------------------
    def _segments(self):
        seg, n_leaf, has_rec = [], 0, False
        for it in self.items:
            if it[0] == "eager":
                if seg:
                    yield ("sweep", seg)
                yield it
                seg, n_leaf, has_rec = [], 0, False
                continue
            n = 1 if it[0] == "leaf" else len(it[3]) + len(it[4])
            if seg and (n_leaf + n > MAX_OPS_PER_LAUNCH or (it[0] == "rec" and has_rec)):
                yield ("sweep", seg)
                seg, n_leaf, has_rec = [], 0, False
            seg.append(it)
            n_leaf += n
            has_rec |= it[0] == "rec"
        if seg:
            yield ("sweep", seg)

    @staticmethod
    def flatten_segment(payload, cdtype=None):
        """One launch worth of items -> (flat op tuples, coefficient tensors in slot order, n_out).  With `cdtype`
        (the complex arithmetic type of the launch) the coefficients are cast to what the kernels of that type read:
        real / complex of that precision, delays float64."""
        ops, coefs = [], []
        for it in payload:
            if it[0] == "leaf":
                ops.append(it[1])
                coefs.append(it[2])
            else:
                _, n_out, n_in, ff, fb = it
                ops.append((OP_RECURSION, n_out, n_in, 0, 0, len(ff), len(fb), 0))
                for l in ff + fb:
                    ops.append(l[1])
                    coefs.append(l[2])
        last = payload[-1]
        n_out = last[1][1] if last[0] == "leaf" else last[1]
        if cdtype is not None:
            real = torch.float32 if cdtype == torch.complex64 else torch.float64
            leaf_ops = [o for o in ops if o[0] != OP_RECURSION]
            coefs = [c if o[0] in (OP_DELAY, OP_PDELAY) else _cast(c, cdtype if o[0] in _TABLE else real)
                     for o, c in zip(leaf_ops, coefs)]
        return tuple(ops), coefs, n_out

    def plan_for(self, ops, cdtype=None):
        cdtype = cdtype or self.cdtype
        return _get_plan(ops, self.nfft, self.alias_decay_db, _lib.C64 if cdtype == torch.complex64 else _lib.C128)

    @staticmethod
    def _per_item_batch(ops, coefs, x4):
        """Per-item launches: every per-item slot must hold one set per batch item of the signal (a single-item signal
        is shared by all sets, like the reference examples' `z[0].unsqueeze(0)`).  Returns the signal to launch on."""
        sizes = {c.shape[0] for o, c in zip([o for o in ops if o[0] != OP_RECURSION], coefs) if o[7]}
        if not sizes:
            return x4
        if len(sizes) != 1:
            raise ValueError(f"per-item parameters with different batch sizes in one program: {sorted(sizes)}")
        n = sizes.pop()
        if x4.shape[0] == 1 and n > 1:
            return x4.expand(n, *x4.shape[1:])
        if x4.shape[0] != n:
            raise ValueError(f"{n} per-item parameter sets for a signal with {x4.shape[0]} batch items")
        return x4

    def _check_signal(self, ops, x4, bin_begin):
        """Backstop in front of every launch (the kernels trust the plan's widths): the signal must carry the channels
        the program's first op reads and only bins that exist."""
        if x4.shape[2] != ops[0][2]:
            raise ValueError(f"the sweep program expects {ops[0][2]} input channels, got a signal of shape "
                             f"{tuple(x4.shape)} (batch, bins, channels, columns)")
        if bin_begin < 0 or bin_begin + x4.shape[1] > self.nfft // 2 + 1:
            raise ValueError(f"bins [{bin_begin}, {bin_begin + x4.shape[1]}) outside the {self.nfft // 2 + 1} bins of "
                             f"nfft = {self.nfft}")

    def run(self, x: torch.Tensor, epilogue: int = EPI_NONE) -> torch.Tensor:
        if not x.is_complex():
            raise TypeError("sweep input must be complex (bin-domain) — put a dsp.FFT input layer in front")
        if x.device.type != "cuda" and _BACKEND.name == "cuda":
            raise RuntimeError(
                "flamo_b200 evaluates the frequency sweep with hand-written CUDA kernels only; got a "
                f"{x.device.type} tensor. Move the model and data to a CUDA device (no CPU fallback exists).")
        shard = current_shard()
        M = self.nfft // 2 + 1
        bin_begin = 0
        cur = x
        if shard is not None and cur.shape[1] == M:
            bin_begin = shard[0]
            cur = cur[:, shard[0]:shard[1]]
        elif shard is not None:
            bin_begin = shard[0]  # already restricted by an earlier launch of the same series
        segs = list(self._segments())
        for si, (tag, payload) in enumerate(segs):
            if tag == "eager":
                # the module's output is taken as returned: an iFFT / cropping Transform changes dim 1 (and the dtype),
                # exactly as when the reference runs its modules one after the other (system.py:279-301)
                cur = payload(cur)
                continue
            if not (torch.is_tensor(cur) and cur.is_complex() and cur.dim() >= 3):
                raise TypeError("a frequency-domain module follows a module whose output is not a bin-domain "
                                f"(complex, (B, bins, channels, ...)) tensor: got {tuple(cur.shape)} {cur.dtype}")
            trail = tuple(cur.shape[3:])
            cols = 1
            for d in trail:
                cols *= d
            x4 = cur.reshape(cur.shape[0], cur.shape[1], cur.shape[2], cols)
            ops, coefs, n_out = self.flatten_segment(payload)
            self._check_signal(ops, x4, bin_begin)
            epi = epilogue if si == len(segs) - 1 else EPI_NONE
            io_dtype = x4.dtype
            ex = torch.complex128 if (io_dtype == torch.complex128 or _wants_f64(ops)) else torch.complex64
            _, coefs, _ = self.flatten_segment(payload, ex)
            x4 = self._per_item_batch(ops, coefs, x4)
            code = _lib.C64 if ex == torch.complex64 else (_lib.C128 | (_lib.DT_GRAD32 if ex != io_dtype else 0))
            plan = _get_plan(ops, self.nfft, self.alias_decay_db, code)
            x4 = SweepFunction.apply(x4.to(ex), plan, ops, epi, bin_begin, n_out, *coefs)
            if ex != io_dtype:  # float32 signal, float64 arithmetic: hand back what a float32 model returns
                x4 = x4.to(torch.float32 if epi == EPI_ABS else io_dtype)
            cur = x4.reshape(x4.shape[:3] + trail)
        if epilogue == EPI_ABS and segs and segs[-1][0] == "eager":
            cur = torch.abs(cur)
        return cur

    def run_loss(self, x: torch.Tensor, target: torch.Tensor, kind: int) -> Optional[torch.Tensor]:
        """criterion(|program(x)|, target) through the fused kernels, or None when this program / these shapes
        cannot be fused (more than one launch, trailing columns, eager modules, a target that is not the plain
        (B, M[, N_out]) magnitude tensor) — the caller then takes the unfused path."""
        if not x.is_complex() or x.dim() != 3:
            return None
        if x.device.type != "cuda" and _BACKEND.name == "cuda":
            return None
        segs = list(self._segments())
        if len(segs) != 1 or segs[0][0] != "sweep":
            return None
        ops, _, n_out = self.flatten_segment(segs[0][1])
        if any(o[7] for o in ops if o[0] != OP_RECURSION):
            return None  # per-item parameter sets: the unfused path
        ex = torch.complex128 if (x.dtype == torch.complex128 or _wants_f64(ops)) else torch.complex64
        _, coefs, _ = self.flatten_segment(segs[0][1], ex)
        B, M = x.shape[0], x.shape[1]
        real = torch.float32 if x.dtype == torch.complex64 else torch.float64
        if kind == CRIT_MSE_CHSUM:  # mse_loss: MSE(sum_r |Y_r|, target.squeeze(-1))
            tgt = target.squeeze(-1) if target.dim() == 3 else target
            if tuple(tgt.shape) != (B, M):
                return None
        else:
            tgt = target
            if tuple(tgt.shape) != (B, M, n_out):
                return None
        if tgt.dtype != real or tgt.device != x.device:
            return None
        scale = 1.0 / tgt.numel()
        shard = current_shard()
        x4 = x.unsqueeze(-1)
        bin_begin = 0
        if shard is not None:
            if M != self.nfft // 2 + 1:
                return None
            bin_begin = shard[0]
            x4, tgt = x4[:, shard[0]:shard[1]], tgt[:, shard[0]:shard[1]]
            scale = 1.0 / tgt.numel()  # mean over the shard, as the unfused path computes it
        if x4.shape[1] == 0 or x4.shape[2] != ops[0][2] or bin_begin + x4.shape[1] > self.nfft // 2 + 1:
            return None  # (the unfused path raises the proper error)
        code = _lib.C64 if ex == torch.complex64 else (_lib.C128 | (_lib.DT_GRAD32 if ex != x.dtype else 0))
        plan = _get_plan(ops, self.nfft, self.alias_decay_db, code)
        if ex != x.dtype:  # float32 model, float64 arithmetic (_wants_f64)
            return SweepLossFunction.apply(x4.to(ex), tgt.to(torch.float64), plan, ops, kind, scale, bin_begin,
                                           *coefs).to(real)
        return SweepLossFunction.apply(x4, tgt, plan, ops, kind, scale, bin_begin, *coefs)


class OrthogonalMap(torch.autograd.Function):
    """(E, sp) = (exp(triu(P,1) - triu(P,1)^T), sparsity_loss(E)) on the device without a host read-back (libfsweep
    fsweep_expm_*_sp): the orthogonal map of dsp.Matrix (reference dsp.py:649) with the parameter-only criterion of the
    colorless-FDN examples (reference optimize/loss.py:36-63) riding along — sparsity_loss picks `sp` up instead of
    launching its own kernels, and both gradients go back through the map in ONE backward launch."""

    @staticmethod
    def supported(P: torch.Tensor) -> bool:
        return (P.is_cuda and P.dim() == 2 and P.shape[0] == P.shape[1] and P.dtype in (torch.float32, torch.float64)
                and P.shape[0] <= _lib.lib().fsweep_expm_max_n())

    @staticmethod
    def forward(ctx, P):
        global launch_count
        Pc = P.detach().contiguous()
        E = torch.empty_like(Pc)
        n = Pc.shape[0]
        sp = torch.empty((), dtype=Pc.dtype, device=Pc.device) if n >= 2 else torch.zeros((), dtype=Pc.dtype,
                                                                                               device=Pc.device)
        with torch.cuda.device(P.device):
            _lib.check(_lib.lib().fsweep_expm_forward_sp(Pc.data_ptr(), E.data_ptr(), n, 1, _real_code(Pc.dtype),
                                                          sp.data_ptr() if n >= 2 else None,
                                                          torch.cuda.current_stream(P.device).cuda_stream))
        launch_count += 1
        ctx.save_for_backward(Pc, E)
        if NOTIFY_SLOT is not None and ctx.needs_input_grad[0]:
            NOTIFY_SLOT["carrier"] = True  # the adjoint of this map will run behind the criteria: it can carry their total
        return E, sp

    @staticmethod
    def backward(ctx, G, gsp):
        global launch_count
        Pc, E = ctx.saved_tensors
        n = Pc.shape[0]
        if n < 2:
            gsp = None
        if G is None and gsp is None:
            return None
        Gc = G.to(Pc.dtype).contiguous() if G is not None else None
        gs = gsp.to(Pc.dtype).contiguous() if gsp is not None else None
        gP = torch.empty_like(Pc)
        import ctypes as C

        taken = take_deferred_total(Pc.dtype, Pc.device)  # a captured step's criteria total rides along (one extra block)
        with torch.cuda.device(Pc.device):
            _lib.check(_lib.lib().fsweep_expm_backward_sp_total(
                Pc.data_ptr(), Gc.data_ptr() if Gc is not None else None, gP.data_ptr(), n, 1, _real_code(Pc.dtype),
                E.data_ptr() if gs is not None else None, gs.data_ptr() if gs is not None else None,
                C.byref(taken[0]) if taken is not None else None, torch.cuda.current_stream(Pc.device).cuda_stream))
        launch_count += 1
        return gP


class BiquadDesign(torch.autograd.Function):
    """Raw Biquad parameter -> packed section coefficients of the SOS op in ONE launch each way (libfsweep
    fsweep_biquad_design: bounded map, RBJ low-/high-pass taps, Taylor packing, float64 inside) instead of the ~60
    parameter-sized PyTorch launches of map -> designer -> pack_sections (reference dsp.py:1494-1563)."""

    @staticmethod
    def forward(ctx, param, n_out, n_in, parallel, highpass):
        global launch_count
        p = param.detach().contiguous()
        K = p.shape[0]
        shape = (K, n_in, 2, 8) if parallel else (K, n_in, n_out, 2, 8)
        packed = torch.empty(shape, dtype=torch.float64, device=p.device)
        with torch.cuda.device(p.device):
            _lib.check(_lib.lib().fsweep_biquad_design(p.data_ptr(), K, n_out, n_in, int(parallel), int(highpass),
                                                        _real_code(p.dtype), packed.data_ptr(), None, None,
                                                        torch.cuda.current_stream(p.device).cuda_stream))
        launch_count += 1
        ctx.save_for_backward(p)
        ctx.meta = (n_out, n_in, parallel, highpass)
        return packed

    @staticmethod
    def backward(ctx, g):
        global launch_count
        (p,) = ctx.saved_tensors
        n_out, n_in, parallel, highpass = ctx.meta
        gc = g.to(torch.float64).contiguous()
        gp = torch.empty_like(p)
        with torch.cuda.device(p.device):
            _lib.check(_lib.lib().fsweep_biquad_design(p.data_ptr(), p.shape[0], n_out, n_in, int(parallel), int(highpass),
                                                        _real_code(p.dtype), None, gc.data_ptr(), gp.data_ptr(),
                                                        torch.cuda.current_stream(p.device).cuda_stream))
        launch_count += 1
        return gp, None, None, None, None


class SVFDesign(torch.autograd.Function):
    """Raw SVF parameter (general mixing) -> packed section coefficients in ONE launch each way (libfsweep
    fsweep_svf_design), as BiquadDesign."""

    @staticmethod
    def forward(ctx, param, n_out, n_in, parallel):
        global launch_count
        p = param.detach().contiguous()
        K = p.shape[1]
        shape = (K, n_in, 2, 8) if parallel else (K, n_in, n_out, 2, 8)
        packed = torch.empty(shape, dtype=torch.float64, device=p.device)
        with torch.cuda.device(p.device):
            _lib.check(_lib.lib().fsweep_svf_design(p.data_ptr(), K, n_out, n_in, int(parallel), _real_code(p.dtype),
                                                     packed.data_ptr(), None, None,
                                                     torch.cuda.current_stream(p.device).cuda_stream))
        launch_count += 1
        ctx.save_for_backward(p)
        ctx.meta = (n_out, n_in, parallel)
        return packed

    @staticmethod
    def backward(ctx, g):
        global launch_count
        (p,) = ctx.saved_tensors
        n_out, n_in, parallel = ctx.meta
        gc = g.to(torch.float64).contiguous()
        gp = torch.empty_like(p)
        with torch.cuda.device(p.device):
            _lib.check(_lib.lib().fsweep_svf_design(p.data_ptr(), p.shape[1], n_out, n_in, int(parallel),
                                                     _real_code(p.dtype), None, gc.data_ptr(), gp.data_ptr(),
                                                     torch.cuda.current_stream(p.device).cuda_stream))
        launch_count += 1
        return gp, None, None, None


def _real_code(dtype: torch.dtype) -> int:
    return _lib.C64 if dtype == torch.float32 else _lib.C128


_RFFT_TABLES = {}


def rfft(x: torch.Tensor, nfft: int, norm: str = "backward", envelope: torch.Tensor = None, force: bool = False):
    """torch.fft.rfft(x [* envelope.view(1, -1, 1)], n=nfft, dim=1, norm=norm) for the excitation of a step (reference
    dsp.py:69-93, :122-163).  A float32 (B, T, C) signal on the GPU that needs no gradient goes through libfsweep's
    two-launch four-step FFT (fsweep_rfft; cuFFT plans nfft = 96000 as five launches); everything else — CPU tensors,
    float64, inputs that require grad, sizes fsweep_rfft_supported refuses, and (unless `force`) the sizes where cuFFT is
    the faster one — stays on torch.fft / cuFFT."""
    global launch_count
    use = (x.is_cuda and x.dtype == torch.float32 and x.dim() == 3 and _BACKEND.name == "cuda"
           and not (x.requires_grad and torch.is_grad_enabled()) and norm in ("backward", "forward", "ortho")
           and (envelope is None or (envelope.is_cuda and envelope.dtype == torch.float32 and envelope.numel() == nfft
                                      and x.shape[1] == nfft and not envelope.requires_grad))
           and os.environ.get("FLAMO_B200_RFFT", "1") != "0"
           # a LATENCY design: it wins where cuFFT needs several launches for one short job (measured on a B200: nfft
           # 96000 8.7 vs 12.1 us warm, 192000 11.6 vs 13.7) and loses on small single-launch sizes and on big batches
           and (force or (nfft >= 16384 and x.shape[0] * x.shape[2] * nfft <= (1 << 20)))
           and _lib.lib().fsweep_rfft_supported(int(nfft))
           # the twiddle table is made at the first call, which must not be a captured one
           and ((int(nfft), x.device.index) in _RFFT_TABLES or not torch.cuda.is_current_stream_capturing()))
    if not use:
        if envelope is not None:
            x = x * envelope.view(1, -1, 1)
        return torch.fft.rfft(x, n=nfft, dim=1, norm=norm)
    x = x.detach()
    if x.stride(2) != 1 or x.stride(1) != x.shape[2] or (x.shape[0] > 1 and x.stride(0) < x.shape[1] * x.shape[2]):
        x = x.contiguous()
    B, T, Cn = x.shape
    stream = torch.cuda.current_stream(x.device).cuda_stream
    L = _lib.lib()
    with torch.cuda.device(x.device):
        key = (int(nfft), x.device.index)
        table = _RFFT_TABLES.get(key)
        if table is None:
            table = torch.empty(L.fsweep_rfft_table_entries(int(nfft)), dtype=torch.complex64, device=x.device)
            _lib.check(L.fsweep_rfft_table(table.data_ptr(), int(nfft), stream))
            torch.cuda.current_stream(x.device).synchronize()  # once per (nfft, device): later calls may use other streams
            _RFFT_TABLES[key] = table
        ws_bytes = L.fsweep_rfft_workspace_bytes(int(nfft), B * Cn)
        ws = torch.empty(ws_bytes // 8, dtype=torch.complex64, device=x.device)
        X = torch.empty((B, nfft // 2 + 1, Cn), dtype=torch.complex64, device=x.device)
        scale = {"backward": 1.0, "forward": 1.0 / nfft, "ortho": nfft ** -0.5}[norm]
        _lib.check(L.fsweep_rfft(x.data_ptr(), B, T, Cn, x.stride(0) if B > 1 else T * Cn, int(nfft), scale,
                                 envelope.data_ptr() if envelope is not None else None, table.data_ptr(),
                                 ws.data_ptr(), ws_bytes, X.data_ptr(), stream))
    launch_count += 2
    return X


class SparsityFunction(torch.autograd.Function):
    """sparsity_loss of a mapped (N, N) or (B, N, N) matrix (reference optimize/loss.py:36-63) as one launch each
    way (libfsweep fsweep_sparsity_*) instead of half a dozen parameter-sized PyTorch kernels."""

    @staticmethod
    def supported(A: torch.Tensor) -> bool:
        return (A.is_cuda and A.dim() in (2, 3) and A.shape[-1] == A.shape[-2] and A.shape[-1] >= 2
                and A.dtype in (torch.float32, torch.float64) and _BACKEND.name == "cuda")

    @staticmethod
    def forward(ctx, A):
        global launch_count
        Ac = A.detach().contiguous()
        loss = torch.empty((), dtype=Ac.dtype, device=Ac.device)
        n_mats = Ac.shape[0] if Ac.dim() == 3 else 1
        with torch.cuda.device(A.device):
            _lib.check(_lib.lib().fsweep_sparsity_forward(Ac.data_ptr(), n_mats, Ac.shape[-1], _real_code(Ac.dtype),
                                                           loss.data_ptr(),
                                                           torch.cuda.current_stream(A.device).cuda_stream))
        launch_count += 1
        ctx.save_for_backward(Ac)
        return loss

    @staticmethod
    def backward(ctx, g):
        global launch_count
        (Ac,) = ctx.saved_tensors
        gc = g.to(Ac.dtype).contiguous()
        gA = torch.empty_like(Ac)
        n_mats = Ac.shape[0] if Ac.dim() == 3 else 1
        with torch.cuda.device(g.device):
            _lib.check(_lib.lib().fsweep_sparsity_backward(Ac.data_ptr(), gc.data_ptr(), n_mats, Ac.shape[-1],
                                                            _real_code(Ac.dtype), gA.data_ptr(),
                                                            torch.cuda.current_stream(g.device).cuda_stream))
        launch_count += 1
        return gA


# Set by Trainer._graph_for around the capture of a step: {"host_vals": pinned uint8 buffer, "host_seq": pinned int32[1],
# "counter": device int32[1]}.  The ONE WeightedTotal launch of the captured step then also stores its values into the
# pinned buffer and bumps the sequence number behind them (fsweep_weighted_total_notify); "used" records (count, dtype).
NOTIFY_SLOT = None


def run_total_job(job, vals):
    """Launch a criteria-total job (a _lib.TotalJob) on its own."""
    global launch_count
    with torch.cuda.device(vals.device):
        _lib.check(_lib.lib().fsweep_weighted_total_notify(
            job.parts, job.alphas, job.scales, job.n, _real_code(vals.dtype), job.vals, job.host_vals, job.host_seq,
            job.seq_counter, torch.cuda.current_stream(vals.device).cuda_stream))
    launch_count += 1


def take_deferred_total(dtype=None, device=None):
    """The pending criteria-total job of the step being captured, if any (and if it matches dtype / device)."""
    slot = NOTIFY_SLOT
    if slot is None or slot.get("job") is None:
        return None
    job, parts, vals = slot["job"]
    if (dtype is not None and vals.dtype != dtype) or (device is not None and vals.device != device):
        return None
    slot["job"] = None
    return job, parts, vals


def notify_values(vals):
    """Captured multi-GPU step: `vals` (the loss values AFTER the exchange between the ranks) go to the Trainer's pinned
    host buffer as a rider of the optimizer's launch (fsweep_adam_step_total) — no copy node, no stream synchronize."""
    slot = NOTIFY_SLOT
    n = vals.numel()
    if (slot is None or slot.get("used") is not None or not slot.get("after_sync") or not vals.is_cuda
            or vals.dtype not in (torch.float32, torch.float64) or not 1 <= n <= _lib.MAX_CRITERIA
            or slot["counter"].device != vals.device):
        return
    vals = vals.contiguous()
    scratch = torch.empty(n + 1, dtype=vals.dtype, device=vals.device)
    job = _lib.TotalJob()
    for i in range(n):  # identity "total": scratch[i] = host[i] = vals[i]
        job.parts[i], job.alphas[i], job.scales[i] = vals.data_ptr() + i * vals.element_size(), 0.0, 1.0
    job.n, job.vals = n, scratch.data_ptr()
    job.host_vals, job.host_seq = slot["host_vals"].data_ptr(), slot["host_seq"].data_ptr()
    job.seq_counter = slot["counter"].data_ptr()
    slot["used"] = (n, vals.dtype)
    slot["adam_rider"] = True
    slot["job"] = (job, vals, scratch)


def flush_deferred_total():
    taken = take_deferred_total()
    if taken is not None:
        run_total_job(taken[0], taken[2])


class WeightedTotal(torch.autograd.Function):
    """vals = [s_0 part_0, ..., s_{n-1} part_{n-1}, sum_i alpha_i s_i part_i] in one launch (libfsweep
    fsweep_weighted_total): the Trainer's `loss += alpha * criterion` accumulation (reference trainer.py:184-188)
    and the vector it reads back once per step.  Backward is one elementwise kernel."""

    MAX = 8

    @staticmethod
    def supported(parts) -> bool:
        p0 = parts[0]
        if os.environ.get("FLAMO_B200_WEIGHTED_TOTAL", "1") == "0":
            return False
        return (_BACKEND.name == "cuda" and 1 <= len(parts) <= WeightedTotal.MAX and torch.is_tensor(p0) and p0.is_cuda
                and p0.dtype in (torch.float32, torch.float64)
                and all(torch.is_tensor(p) and p.is_cuda and p.dtype == p0.dtype and p.numel() == 1
                        and p.device == p0.device for p in parts))

    @staticmethod
    def forward(ctx, alphas, scales, *parts):
        global launch_count
        import ctypes as C

        n = len(parts)
        parts = [p.detach().reshape(1).contiguous() for p in parts]
        vals = torch.empty(n + 1, dtype=parts[0].dtype, device=parts[0].device)
        pp = (C.c_void_p * n)(*[p.data_ptr() for p in parts])
        aa = (C.c_double * n)(*[float(a) for a in alphas])
        ss = (C.c_double * n)(*[float(a) for a in scales])
        slot = NOTIFY_SLOT
        if (slot is not None and slot.get("used") is None and not slot.get("after_sync")
                and slot["counter"].device == vals.device):
            # the Trainer captures a step: the values also go straight to its pinned host buffer (see NOTIFY_SLOT)
            slot["used"] = (n + 1, vals.dtype)
            job = _lib.TotalJob()
            for i in range(n):
                job.parts[i], job.alphas[i], job.scales[i] = parts[i].data_ptr(), aa[i], ss[i]
            job.n, job.vals = n, vals.data_ptr()
            job.host_vals, job.host_seq = slot["host_vals"].data_ptr(), slot["host_seq"].data_ptr()
            job.seq_counter = slot["counter"].data_ptr()
            if slot.get("defer") and slot.get("carrier"):
                # nothing on the device reads the values: they ride along with the first parameter-sized launch of
                # the backward pass (OrthogonalMap.backward takes the job — early enough for the host to be back
                # before the step ends; Trainer._train_core launches it on its own if nobody did)
                slot["job"] = (job, parts, vals)
            elif slot.get("side") is not None:
                # on a parallel branch of the captured graph: off the critical path sweep -> adjoint maps -> optimizer
                cur, side = torch.cuda.current_stream(vals.device), slot["side"]
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    run_total_job(job, vals)
                vals.record_stream(side)
                slot["join"] = True
            else:
                run_total_job(job, vals)
        else:
            with torch.cuda.device(vals.device):
                _lib.check(_lib.lib().fsweep_weighted_total(pp, aa, ss, n, _real_code(vals.dtype), vals.data_ptr(),
                                                             torch.cuda.current_stream(vals.device).cuda_stream))
            launch_count += 1
        ctx.coef = (tuple(float(a) * float(c) for a, c in zip(alphas, scales)), tuple(float(c) for c in scales))
        for c in ctx.coef:  # created here, outside any later capture of the backward
            _const_tensor(c, vals.dtype, vals.device)
        ctx.shapes = [p.shape for p in parts]
        return vals

    @staticmethod
    def backward(ctx, g):
        n = len(ctx.coef[0])
        a_s = _const_tensor(ctx.coef[0], g.dtype, g.device)
        sc = _const_tensor(ctx.coef[1], g.dtype, g.device)
        if all(c == 1.0 for c in ctx.coef[1]):
            gp = torch.addcmul(g[:n], a_s, g[n:n + 1].expand(n))  # g_i + alpha_i g_total: one kernel
        else:
            gp = torch.addcmul(sc * g[:n], a_s, g[n:n + 1].expand(n))
        return (None, None, *[gp[i].reshape(()) for i in range(n)])


_CONST_TENSORS = {}


def _const_tensor(values, dtype, device):
    """Small constant tensors are created once per (values, dtype, device): creating a tensor from host data is
    not allowed inside CUDA-graph capture."""
    key = (values, dtype, str(device))
    t = _CONST_TENSORS.get(key)
    if t is None:
        t = _CONST_TENSORS[key] = torch.tensor(values, dtype=dtype, device=device)
    return t


# rows: {B(w0), B'(w0), b2, 0, A(w0), A'(w0), a2, 0} for w0 = +1 then w0 = -1; columns: (b0, b1, b2, a0, a1, a2)
_PACK_MATRIX = (
    (1, 1, 1, 0, 0, 0), (0, 1, 2, 0, 0, 0), (0, 0, 1, 0, 0, 0), (0, 0, 0, 0, 0, 0),
    (0, 0, 0, 1, 1, 1), (0, 0, 0, 0, 1, 2), (0, 0, 0, 0, 0, 1), (0, 0, 0, 0, 0, 0),
    (1, -1, 1, 0, 0, 0), (0, 1, -2, 0, 0, 0), (0, 0, 1, 0, 0, 0), (0, 0, 0, 0, 0, 0),
    (0, 0, 0, 1, -1, 1), (0, 0, 0, 0, 1, -2), (0, 0, 0, 0, 0, 1), (0, 0, 0, 0, 0, 0),
)


def pack_sections(b: torch.Tensor, a: torch.Tensor, parallel: bool, real: torch.dtype) -> torch.Tensor:
    """(3, K, N_out, N_in) taps (or (3, K, N) for parallel) -> kernel layout [K][N_in][N_out][2][8]
    ([K][N][2][8]): per section the Taylor coefficients of B and A around w0 = +1 and w0 = -1
    (include/fsweep.h, FSWEEP_OP_SOS).  The packing is linear in the taps: ONE contraction with a constant 16 x 6
    matrix in the taps' own (float64) precision, then a cast — a handful of launches forward and backward instead of
    one per tap combination.  Differentiable, so autograd maps the kernel's packed gradient back onto (b, a)."""
    taps = torch.cat((b, a), dim=0)  # (6, K, ...)
    T = _const_tensor(_PACK_MATRIX, taps.dtype, taps.device)  # (16, 6)
    packed = torch.tensordot(T, taps, dims=([1], [0]))  # (16, K, ...)
    packed = packed.reshape(2, 8, *taps.shape[1:])
    if parallel:
        packed = packed.permute(2, 3, 0, 1)  # (K, N, 2, 8)
    else:
        packed = packed.permute(2, 4, 3, 0, 1)  # (2, 8, K, N_out, N_in) -> (K, N_in, N_out, 2, 8)
    return (packed if real is None else packed.to(real)).contiguous()
