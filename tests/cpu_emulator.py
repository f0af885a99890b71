"""TEST-ONLY: a float64 PyTorch emulator of the fsweep C ABI (include/fsweep.h).

It interprets the same flat op programs and the same kernel-layout coefficient buffers that
flamo_b200.sweep hands to libfsweep.so, so the host-side logic (lowering, coefficient packing,
autograd plumbing, series splitting, bin sharding, the distributed trainer) can be exercised on a
machine without a GPU.  It is installed by the `emulated_backend` fixture only; the product never
imports this file and has no CPU path.
"""
import math

import torch

from flamo_b200 import _lib, sweep
from flamo_b200._lib import (EPI_ABS, F_ISINT, OP_DELAY, OP_GAIN, OP_PDELAY, OP_PGAIN, OP_PSOS, OP_PTABLE,
                             OP_RECURSION, OP_SOS, OP_TABLE)


class _EmuPlan:
    def __init__(self, ops, nfft, alias_decay_db, dtype):
        # validate with the real library's host-side planner when it is built
        self.real = None
        try:
            self.real = _lib.Plan([_lib.Op(*o) for o in ops], nfft, alias_decay_db, dtype)  # (flags like DT_GRAD32 included)
        except _lib.Unsupported:
            pass  # e.g. loop width > 32: the emulator itself has no such limit
        except RuntimeError as e:
            if "not found" not in str(e):
                raise
        self.ops, self.nfft, self.alias, self.dtype = ops, nfft, alias_decay_db, dtype


def _response(op, coef, k, nfft, lng):
    """H for absolute bins k: (nb, n_out, n_in) complex128 (diagonal kinds: (nb, n))."""
    kind, n_out, n_in, K, flags = op[:5]
    om = 2 * math.pi * k.double() / nfft
    if kind in (OP_GAIN, OP_PGAIN):
        return torch.complex(coef.double(), torch.zeros_like(coef.double())).unsqueeze(0).expand(len(k), *coef.shape)
    if kind in (OP_TABLE, OP_PTABLE):
        return coef.to(torch.complex128)[k]
    if kind in (OP_DELAY, OP_PDELAY):
        d = coef.double()
        if flags & F_ISINT:
            d = d.round()
            idx = torch.remainder(k.view(-1, *[1] * d.dim()).to(torch.int64) * d.to(torch.int64).unsqueeze(0), nfft)
            ph = 2 * math.pi * idx.double() / nfft
        else:
            ph = om.view(-1, *[1] * d.dim()) * d.unsqueeze(0)
        return torch.exp(lng * d).unsqueeze(0) * torch.exp(-1j * ph)
    # SOS / PSOS: packed sections [K][n_in][n_out][2][8] / [K][n][2][8]
    c = coef.double()
    g = math.exp(lng)
    # z^-1 = exp(-j 2 pi k / nfft) with EXACT values at the quarter points, as the kernels' sincospi gives them
    # (torch.exp(-1j * pi) has a 1.2e-16 imaginary part, which matters where a section's zero and pole cancel at Nyquist)
    x = 2.0 * k.double() / nfft                                   # omega / pi in [0, 1]
    sinpi = torch.sin(math.pi * torch.minimum(x, 1.0 - x))
    cospi = torch.where(x <= 0.5, torch.sin(math.pi * (0.5 - x)), -torch.sin(math.pi * (x - 0.5)))
    w = g * torch.complex(cospi, -sinpi)
    plus = cospi >= 0
    v = torch.where(plus, w - 1, w + 1)
    shape = (-1,) + (1,) * (c.dim() - 2)
    v, plus = v.view(shape), plus.view(shape)
    cc = torch.where(plus.unsqueeze(-1), c[..., 0, :].unsqueeze(0), c[..., 1, :].unsqueeze(0))  # (nb, K, ..., 8)
    B = cc[..., 0] + cc[..., 1] * v + cc[..., 2] * v * v  # (nb, K, n_in, n_out) | (nb, K, n)
    A = cc[..., 4] + cc[..., 5] * v + cc[..., 6] * v * v
    num, den = B.prod(dim=1), A.prod(dim=1)
    H = torch.where(den.abs() != 0, num / den, torch.full_like(num, torch.finfo(torch.float64).eps))
    if kind == OP_SOS:
        H = H.transpose(1, 2)  # (nb, n_out, n_in)
    return H


def _apply(op, H, x):
    if op[0] in (OP_PGAIN, OP_PSOS, OP_PDELAY, OP_PTABLE):
        return H.unsqueeze(0).unsqueeze(-1) * x
    return torch.einsum("fmn,bfnc->bfmc", H, x)


def _run(ops, coefs, x, bin_begin, nfft, alias, epilogue):
    leaf = [o for o in ops if o[0] != OP_RECURSION]
    if any(o[7] for o in leaf):
        # per-item coefficient sets (fsweep_op_t::per_item): batch item b meets set b — one plain run per item
        plain = tuple(tuple(o[:7]) + (0,) for o in ops)
        return torch.cat([_run(plain, [c[b] if o[7] else c for o, c in zip(leaf, coefs)], x[b:b + 1], bin_begin, nfft,
                               alias, epilogue) for b in range(x.shape[0])])
    lng = -abs(alias) / nfft / 20.0 * math.log(10.0)
    k = torch.arange(bin_begin, bin_begin + x.shape[1])
    x = x.to(torch.complex128)
    ci, i = 0, 0
    while i < len(ops):
        op = ops[i]
        if op[0] == OP_RECURSION:
            n_ff, n_fb = op[5], op[6]
            Hs = [_response(ops[i + 1 + j], coefs[ci + j], k, nfft, lng) for j in range(n_ff + n_fb)]
            ff = list(zip(ops[i + 1:i + 1 + n_ff], Hs[:n_ff]))
            fb = list(zip(ops[i + 1 + n_ff:i + 1 + n_ff + n_fb], Hs[n_ff:]))
            N = op[1]
            b = x
            for o, H in ff:
                b = _apply(o, H, b)
            T = torch.eye(N, dtype=torch.complex128).expand(1, len(k), N, N)
            for o, H in fb + ff:
                T = _apply(o, H, T)
            A = torch.eye(N, dtype=torch.complex128) - T[0]
            x = torch.linalg.solve(A.unsqueeze(0).expand(b.shape[0], -1, -1, -1), b)
            ci += n_ff + n_fb
            i += 1 + n_ff + n_fb
        else:
            x = _apply(op, _response(op, coefs[ci], k, nfft, lng), x)
            ci += 1
            i += 1
    return torch.abs(x) if epilogue == EPI_ABS else x


class EmulatedBackend:
    name = "emulated-cpu"

    def plan(self, ops, nfft, alias_decay_db, dtype):
        return _EmuPlan(ops, nfft, alias_decay_db, dtype)

    def forward(self, plan, ops, coefs, x, y, cols, bin_begin, epilogue):
        with torch.no_grad():
            y.copy_(_run(ops, coefs, x, bin_begin, plan.nfft, plan.alias, epilogue).to(y.dtype))
        return 1

    def backward(self, plan, ops, coefs, x, gy, grads, gx, cols, bin_begin, epilogue):
        with torch.enable_grad():
            cs = [c.detach().clone().requires_grad_(g is not None) for c, g in zip(coefs, grads)]
            xs = x.detach().clone().requires_grad_(gx is not None)
            out = _run(ops, cs, xs, bin_begin, plan.nfft, plan.alias, epilogue)
            wanted = [c for c, g in zip(cs, grads) if g is not None] + ([xs] if gx is not None else [])
            got = torch.autograd.grad(out, wanted, gy.to(out.dtype), allow_unused=True)
        it = iter(got)
        for c, g in zip(cs, grads):
            if g is not None:
                v = next(it)
                g.copy_(torch.zeros_like(g) if v is None else v.to(g.dtype))
        if gx is not None:
            gx.copy_(next(it).to(gx.dtype))
        return 2  # the sweep kernel + the gradient finalize

    def loss(self, plan, ops, coefs, x, target, kind, scale, loss, grads, gx, bin_begin):
        want = grads is not None
        with torch.enable_grad():
            cs = [c.detach().clone().requires_grad_(want and grads[i] is not None) for i, c in enumerate(coefs)]
            xs = x.detach().clone().requires_grad_(want and gx is not None)
            mag = _run(ops, cs, xs, bin_begin, plan.nfft, plan.alias, EPI_ABS)[..., 0]
            e = (mag.sum(-1) if kind == _lib.CRIT_MSE_CHSUM else mag) - target.double()
            val = scale * (e * e).sum()
            loss.copy_(val.detach().to(loss.dtype))
            if not want:
                return 1
            wanted = [c for c, g in zip(cs, grads) if g is not None] + ([xs] if gx is not None else [])
            got = torch.autograd.grad(val, wanted, allow_unused=True) if wanted else []
        it = iter(got)
        for c, g in zip(cs, grads):
            if g is not None:
                v = next(it)
                g.copy_(torch.zeros_like(g) if v is None else v.to(g.dtype))
        if gx is not None:
            gx.copy_(next(it).to(gx.dtype))
        return 2


def install():
    prev = sweep._BACKEND
    sweep._BACKEND = EmulatedBackend()
    return prev


def uninstall(prev):
    sweep._BACKEND = prev
