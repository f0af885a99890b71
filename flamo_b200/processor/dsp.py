"""flamo.processor.dsp — same classes, constructor signatures, parameter shapes and attributes as the
reference (gdalsanto/flamo, flamo/processor/dsp.py), with every per-bin evaluation lowered to the
CUDA sweep (flamo_b200.sweep -> libfsweep.so).

What stays PyTorch: the O(#params) maps from raw `param` to coefficients (`map`, RBJ / SVF / GEQ
designers, softplus of delays, matrix exponential).  They are evaluated in float64 (free at this
size, removes the reference's float32 rounding of delays and section taps — SURVEY.md finding 4) and
handed to the kernel in its compute dtype.

`freq_response(param)` / `get_poly_coeff(...)` remain available as plain-PyTorch closed forms for
callers that want the response tensor itself (plots, probes, subclasses); they are not used by
`forward`.  A subclass or instance that overrides them is lowered as a TABLE op fed by its own
`freq_response(param)`.
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import sweep
from .._lib import OP_DELAY, OP_GAIN, OP_PDELAY, OP_PGAIN, OP_PSOS, OP_PTABLE, OP_SOS, OP_TABLE
from ..auxiliary.eq import eq_freqs, geq
from ..functional import (HadamardMatrix, RotationMatrix, bandpass_filter, expm_capturable, highpass_filter,
                          lowpass_filter, skew_matrix)
from ..utils import to_complex

# ======================================================================================= transforms


class Transform(nn.Module):
    """Callable wrapper used as Shell input / output layer (reference dsp.py:27-66)."""

    def __init__(self, transform: callable = lambda x: x, device: Optional[str] = None,
                 dtype: torch.dtype = torch.float32):
        super().__init__()
        self.transform = transform
        self.device = device
        self.dtype = dtype

    def forward(self, x):
        return self.transform(x)

    def probe(self, z):
        return None


class FFT(Transform):
    """rfft along dim 1 with n = nfft (reference dsp.py:69-93); cuFFT when x is on the GPU."""

    def __init__(self, nfft: int = 2 ** 11, norm: str = "backward", dtype: torch.dtype = torch.float32):
        self.nfft, self.norm = nfft, norm
        super().__init__(transform=self._transform, dtype=dtype)

    def _transform(self, x):
        return sweep.rfft(x, self.nfft, self.norm)  # libfsweep's two-launch FFT on the GPU, torch.fft elsewhere


class iFFT(Transform):
    """irfft along dim 1 with n = nfft (reference dsp.py:96-119)."""

    def __init__(self, nfft: int = 2 ** 11, norm: str = "backward", dtype: torch.dtype = torch.float32):
        self.nfft, self.norm = nfft, norm
        super().__init__(transform=self._transform, dtype=dtype)

    def _transform(self, x):
        x = sweep.gather_bins(x, self.nfft)  # inside a multi-GPU bin shard: the whole spectrum first
        return torch.fft.irfft(x, n=self.nfft, dim=1, norm=self.norm)


def _alias_envelope(nfft, alias_decay_db, device, dtype):
    gamma = 10 ** (-torch.abs(torch.tensor(alias_decay_db, device=device, dtype=dtype)) / nfft / 20)
    return gamma ** torch.arange(0, -nfft, -1, device=device, dtype=dtype)


class FFTAntiAlias(Transform):
    """rfft of x(n) * gamma^-n (reference dsp.py:122-163)."""

    def __init__(self, nfft: int = 2 ** 11, norm: str = "backward", alias_decay_db: float = 0.0,
                 device: Optional[str] = None, dtype: torch.dtype = torch.float32):
        self.nfft, self.norm = nfft, norm
        self.alias_envelope = _alias_envelope(nfft, alias_decay_db, device, dtype)
        super().__init__(transform=self._transform, device=device, dtype=dtype)

    def _transform(self, x):
        return sweep.rfft(x, self.nfft, self.norm, envelope=self.alias_envelope)


class iFFTAntiAlias(Transform):
    """irfft followed by the gamma^-n envelope (reference dsp.py:166-206)."""

    def __init__(self, nfft: int = 2 ** 11, norm: str = "backward", alias_decay_db: float = 0.0,
                 device: Optional[str] = None, dtype: torch.dtype = torch.float32):
        self.nfft, self.norm = nfft, norm
        self.alias_envelope = _alias_envelope(nfft, alias_decay_db, device, dtype)
        super().__init__(transform=self._transform, device=device, dtype=dtype)

    def _transform(self, x):
        x = sweep.gather_bins(x, self.nfft)  # inside a multi-GPU bin shard: the whole spectrum first
        return torch.fft.irfft(x, n=self.nfft, dim=1, norm=self.norm) * self.alias_envelope.view(1, -1, 1)


# ============================================================================================= core


def _identity(x):
    return x


class DSP(nn.Module):
    """Base of every LTI module: raw `param`, a `map` to a stable parameterisation, a closed-form
    response on the nfft/2+1 rFFT bins (reference dsp.py:212-352)."""

    _parallel = False  # 1-D (per-channel) variant

    def __init__(self, size: tuple, nfft: int = 2 ** 11, map: callable = _identity, requires_grad: bool = False,
                 alias_decay_db: float = 0.0, device: Optional[str] = None, dtype: torch.dtype = torch.float32):
        super().__init__()
        assert isinstance(size, tuple), "Size must be a tuple."
        self.size = size
        self.nfft = nfft
        self.map = map
        self.new_value = 0
        self.requires_grad = requires_grad
        self.device = device
        self.dtype = dtype
        self.param = nn.Parameter(torch.empty(self.size, device=device, dtype=dtype), requires_grad=requires_grad)
        self.fft = lambda x: torch.fft.rfft(x, n=self.nfft, dim=0)
        self.ifft = lambda x: torch.fft.irfft(x, n=self.nfft, dim=0)
        self.alias_decay_db = torch.tensor(alias_decay_db, device=device, dtype=dtype)
        self._alias_db = float(alias_decay_db)  # host copy: lowering never reads the device tensor back
        self._consts = {}
        self._coef_cache = None
        self.init_param()
        self.get_gamma()

    def _const(self, key, values, like):
        """Small constant tensor cached per (device, dtype): maps must not create device tensors from
        host data at call time (that would synchronise and is illegal under CUDA-graph capture)."""
        k = (key, like.device, like.dtype)
        t = self._consts.get(k)
        if t is None:
            t = torch.as_tensor(values).to(device=like.device, dtype=like.dtype)
            self._consts[k] = t
        return t

    # -- parameters ------------------------------------------------------------------------------
    def init_param(self):
        torch.nn.init.normal_(self.param)

    def get_gamma(self):
        self.gamma = 10 ** (-torch.abs(self.alias_decay_db) / self.nfft / 20)

    def assign_value(self, new_value: torch.Tensor, indx: tuple = tuple([slice(None)])):
        assert self.param[indx].shape == new_value.shape, (
            f"New values shape {new_value.shape} is not compatible with the parameter shape {self.param[indx].shape}.")
        with torch.no_grad():
            self.param[indx].copy_(new_value)
            self.new_value = 1

    # -- evaluation ------------------------------------------------------------------------------
    def forward(self, x, ext_param=None):
        """x: (B, M, N_in, ...) complex -> (B, M, N_out, ...).  `ext_param` replaces `param` on the
        gradient path after being logged into it (reference dsp.py:415-432)."""
        self.check_input_shape(x)
        return self.freq_convolve(x, self._select_param(ext_param))

    def _select_param(self, ext_param):
        if ext_param is None:
            return self.param
        with torch.no_grad():
            # one parameter set per batch item (see _emit_items): the reference's per-item loop leaves the last one behind
            self.assign_value(ext_param[-1] if self._per_item(ext_param) else ext_param)
        return ext_param

    def _per_item(self, param) -> bool:
        """`param` carries one parameter set PER BATCH ITEM: a leading batch axis in front of the module's own shape.
        The reference has no such call — its NN-in-the-loop examples run the module once per item
        (examples/e7_biquad_nn.py:149-156, e4_recursion_nn.py:243-250, "the only way to process batches larger than
        1") — here the whole batch is ONE launch."""
        return (torch.is_tensor(param) and param is not self.param and param.dim() == self.param.dim() + 1
                and tuple(param.shape[1:]) == tuple(self.param.shape))

    def _sweep(self, x, param):
        prog = sweep.Program(self.nfft, self._alias_db, x.dtype, x.device)
        if self._per_item(param):
            self._emit_items(prog, param)
        else:
            self._emit(prog, param)
        return prog.run(x)

    def _lower(self, prog, ext_param=None):
        param = self._select_param(ext_param)
        if self._per_item(param):
            self._emit_items(prog, param)
            return
        if param is self.param and not param.requires_grad:
            # frozen parameters: the mapped coefficient tensor is reused until the parameter is
            # written again (assign_value / load_state_dict bump `_version`); the optimizer never
            # touches it, so this also holds across CUDA-graph replays
            key = (param._version, param.data_ptr(), prog.real, self.map)
            if self._coef_cache is not None and self._coef_cache[0] == key:
                prog.items_target().append(self._coef_cache[1])
                return
            n0 = len(prog.items_target())
            with torch.no_grad():
                self._emit(prog, param)
            if len(prog.items_target()) == n0 + 1:
                self._coef_cache = (key, prog.items_target()[-1])
            return
        self._emit(prog, param)

    def _emit_items(self, prog, params):
        """Lower `params` (B, *param.shape) as ONE op whose coefficient slot holds B sets (fsweep_op_t::per_item): batch
        item b of the signal is filtered with set b.  The generic route maps / designs the sets one after the other
        (small parameter-sized launches) and stacks them; subclasses whose design treats its sections independently fold
        the batch into the section axis instead (_items_folded)."""
        if self._items_folded(prog, params):
            return
        tgt = prog.items_target()
        n0 = len(tgt)
        leaves = []
        for i in range(params.shape[0]):
            self._emit(prog, params[i])
            if len(tgt) != n0 + 1 or tgt[-1][0] != "leaf":
                raise sweep._lib.Unsupported(sweep._lib.E_UNSUPPORTED, f"{type(self).__name__} does not lower to a "
                                             "single op: per-item parameters are not supported for it")
            leaves.append(tgt.pop())
        op = leaves[0][1]
        if any(l[1][:4] != op[:4] for l in leaves):
            raise sweep._lib.Unsupported(sweep._lib.E_UNSUPPORTED, "per-item parameters lower to different ops")
        prog.leaf_items(op, torch.stack([l[2] for l in leaves]))

    def _items_folded(self, prog, params) -> bool:
        return False

    def _emit(self, prog, param):
        raise NotImplementedError

    def _up(self, param):
        """Raw parameter in float64 for the maps."""
        return param.to(torch.float64)

    def _omega(self, dtype=torch.float64, device=None):
        return 2 * math.pi * torch.arange(0, self.nfft // 2 + 1, dtype=dtype, device=device) / self.nfft

    def _bins_ok(self, x):
        """All nfft // 2 + 1 bins — or, inside sweep.bin_shard, a signal an earlier launch of the same Series already
        restricted to the shard (a Parallel node or any other eager module after the first sweep)."""
        if x.shape[1] == int(self.nfft / 2 + 1):
            return True
        shard = sweep.current_shard()
        return shard is not None and x.shape[1] == shard[1] - shard[0]

    def probe(self, z):
        raise NotImplementedError(f"probe() not implemented for {self.__class__.__name__}")

    def probe_w(self, w):
        return self.probe(1 / w)


# ============================================================================================ gains


class Gain(DSP):
    """y[b,f,m,...] = sum_n map(param)[m,n] x[b,f,n,...]   (reference dsp.py:357-496)."""

    def __init__(self, size: tuple = (1, 1), nfft: int = 2 ** 11, map: callable = _identity,
                 requires_grad: bool = False, alias_decay_db: float = 0.0, device: Optional[str] = None,
                 dtype: torch.dtype = torch.float32):
        super().__init__(size=size, nfft=nfft, map=map, requires_grad=requires_grad, alias_decay_db=alias_decay_db,
                         device=device, dtype=dtype)
        self.initialize_class()

    def initialize_class(self):
        self.check_param_shape()
        self.get_io()
        self.get_freq_convolve()

    def check_param_shape(self):
        assert len(self.size) == 2, "gains must be 2D. For 1D (parallel) gains use parallelGain module."

    def check_input_shape(self, x):
        if self.input_channels != x.shape[2]:
            raise ValueError(f"parameter shape = {self.size} not compatible with input signal of shape = ({x.shape}).")

    def get_io(self):
        self.input_channels = self.size[-1]
        self.output_channels = self.size[-2] if len(self.size) > 1 else self.size[-1]

    def get_freq_convolve(self):
        self.freq_convolve = lambda x, param: self._sweep(x, param)

    def _emit(self, prog, param):
        # the default identity map needs no float64 detour
        W = param if self.map is _identity else self.map(self._up(param))
        if self._parallel:
            prog.leaf(OP_PGAIN, self.output_channels, self.input_channels, W.reshape(-1))
        else:
            prog.leaf(OP_GAIN, self.output_channels, self.input_channels, W)

    def probe(self, z):
        h = to_complex(self.map(self.param))
        return torch.diag(h) if self._parallel else h


class parallelGain(Gain):
    """Per-channel gains, param (N,) (reference dsp.py:499-573)."""

    _parallel = True

    def __init__(self, size: tuple = (1,), **kwargs):
        super().__init__(size=size, **kwargs)

    def check_param_shape(self):
        assert len(self.size) == 1, "gains must be 1D, for 2D gains use Gain module."


class Matrix(Gain):
    """Gain whose map builds a structured matrix: random | orthogonal | hadamard | rotation
    (reference dsp.py:579-676)."""

    def __init__(self, size: tuple = (1, 1), nfft: int = 2 ** 11, map: callable = _identity,
                 matrix_type: str = "random", iter: int = 1, requires_grad: bool = False,
                 alias_decay_db: float = 0.0, device: Optional[str] = None, dtype: torch.dtype = torch.float32):
        self.matrix_type = matrix_type
        self.iter = iter
        super().__init__(size=size, nfft=nfft, map=map, requires_grad=requires_grad, alias_decay_db=alias_decay_db,
                         device=device, dtype=dtype)

    def matrix_gallery(self):
        N = self.size[0]
        t = self.matrix_type
        if t == "random":
            self.map = _identity
        elif t == "orthogonal":
            assert N == self.size[1], "Matrix must be square to be orthogonal"
            self.map = self._orthogonal_map
        elif t == "hadamard":
            assert N == self.size[1], "Matrix must be square to be Hadamard"
            assert N % 2 == 0, "Matrix must have even dimensions to be Hadamard"
            self.map = lambda x: HadamardMatrix(N, device=x.device, dtype=x.dtype)(x)
        elif t == "rotation":
            assert N == self.size[1], "Matrix must be square to be a rotation matrix"
            assert N % 2 == 0, "Matrix must have even dimensions to be a rotation matrix"
            # (the reference passes `iter` in RotationMatrix's SECOND positional slot, which is min_angle, dsp.py:665:
            #  the number of Kronecker squarings is therefore always log2(N) - 1 and `iter` bounds the angle from below)
            self.map = lambda x: RotationMatrix(N, self.iter, device=x.device, dtype=x.dtype)([x[0][0]])
        else:
            raise ValueError(f"unknown matrix_type {t}")

    def _orthogonal_map(self, x):
        """matrix_exp(skew(x)).  On CUDA the library's capture-safe kernel is used; the result is
        memoised per parameter version so that the sweep and parameter-only criteria (sparsity_loss
        calls map(param) again, reference loss.py:41-42) share one evaluation per step."""
        if not x.is_cuda:
            return torch.matrix_exp(skew_matrix(x))
        key = (x._version, torch.is_grad_enabled() and x.requires_grad)
        if x is self.param:
            hit = self._expm_cache
            if hit is not None and hit[0] == key:
                return hit[1]
        sp = None
        if sweep.OrthogonalMap.supported(x):
            out, sp = sweep.OrthogonalMap.apply(x)  # sp: sparsity_loss(out), for optimize.loss.sparsity_loss
        else:  # wider than the one-CTA kernel: capture-safe PyTorch scaling-and-squaring (float64 inside)
            out = expm_capturable(skew_matrix(x))
        if x is self.param:
            self._expm_cache = (key, out, sp)
        return out

    def _up(self, param):
        # the orthogonal map upcasts internally; handing it `param` itself lets the memo above hit
        if self.matrix_type == "orthogonal" and param.is_cuda:
            return param
        return super()._up(param)

    def invalidate_cache(self):
        """Drop the memoised map (parameters changed without a version bump, e.g. by a graph replay)."""
        self._expm_cache = None

    def initialize_class(self):
        self._expm_cache = None
        self.check_param_shape()
        self.get_io()
        self.matrix_gallery()
        self.get_freq_convolve()


class HouseholderMatrix(Gain):
    """U = I - 2 u u^T with u = param / ||param||, param (N, 1) (reference dsp.py:679-782).  The reference applies it
    as two vector products; here it is lowered as a dense GAIN whose N x N matrix is formed from u in the map's
    precision (O(N^2), differentiable), so it fuses with its neighbours like any other gain."""

    def __init__(self, size: tuple = (1, 1), nfft: int = 2 ** 11, requires_grad: bool = False,
                 alias_decay_db: float = 0.0, device: Optional[str] = None, dtype: torch.dtype = torch.float32):
        assert size[0] == size[1], "Matrix must be square"
        super().__init__(size=(size[0], 1), nfft=nfft, map=lambda x: to_complex(x) / torch.norm(x, dim=0, keepdim=True),
                         requires_grad=requires_grad, alias_decay_db=alias_decay_db, device=device, dtype=dtype)

    def check_input_shape(self, x):
        if self.size[0] != x.shape[2]:
            raise ValueError(f"parameter shape = {self.size} not compatible with input signal of shape = ({x.shape}).")

    def get_io(self):
        self.input_channels = self.size[0]
        self.output_channels = self.size[0]

    def _matrix(self, param):
        u = self.map(param)
        u = u.real if u.is_complex() else u
        return torch.eye(u.shape[0], dtype=u.dtype, device=u.device) - 2 * (u @ u.mT)

    def _emit(self, prog, param):
        prog.leaf(OP_GAIN, self.output_channels, self.input_channels, self._matrix(self._up(param)))

    def probe(self, z):
        return to_complex(self._matrix(self.param))


# ========================================================================================== filters


class Filter(DSP):
    """FIR filter bank: param (taps, N_out, N_in); H = rfft(map(param) * gamma^n) (reference
    dsp.py:788-962).  Lowered as a TABLE op: the response is built once per call by cuFFT and
    streamed by the sweep."""

    _native_response = True

    def __init__(self, size: tuple = (1, 1, 1), nfft: int = 2 ** 11, map: callable = _identity,
                 requires_grad: bool = False, alias_decay_db: float = 0.0, device: Optional[str] = None,
                 dtype: torch.dtype = torch.float32):
        super().__init__(size=size, nfft=nfft, map=map, requires_grad=requires_grad, alias_decay_db=alias_decay_db,
                         device=device, dtype=dtype)
        self.initialize_class()

    def initialize_class(self):
        self.check_param_shape()
        self.get_io()
        self.get_freq_response()
        self.get_freq_convolve()
        self._stock_freq_response = self.freq_response

    def check_param_shape(self):
        assert len(self.size) == 3, "Filter must be 3D, for 2D (parallel) filters use ParallelFilter module."

    def check_input_shape(self, x):
        if not self._bins_ok(x) or self.input_channels != x.shape[2]:
            raise ValueError(f"parameter shape not compatible with input signal of shape = ({x.shape}).")

    def get_io(self):
        self.input_channels = self.size[-1]
        self.output_channels = self.size[-1] if self._parallel else self.size[-2]

    def get_freq_response(self):
        self.ir = lambda x: self.map(x)
        self.freq_response = self._fir_response

    def _fir_response(self, param):
        ir = self.map(param)
        env = (10 ** (-abs(self._alias_db) / self.nfft / 20)) ** torch.arange(0, ir.shape[0], device=ir.device,
                                                                               dtype=ir.dtype)
        return torch.fft.rfft(ir * env.view(-1, *([1] * (ir.dim() - 1))), n=self.nfft, dim=0)

    def get_freq_convolve(self):
        self.freq_convolve = lambda x, param: self._sweep(x, param)

    def _emit_table(self, prog, param, upcast=True):
        H = self.freq_response(self._up(param) if upcast else param)
        prog.leaf(OP_PTABLE if self._parallel else OP_TABLE, self.output_channels, self.input_channels, H)

    def _emit(self, prog, param):
        self._emit_table(prog, param)

    def probe(self, z):
        coeff = self.map(self.param)
        k = torch.arange(coeff.shape[0], device=coeff.device, dtype=coeff.dtype)
        w = (self.gamma ** k) * z ** (-k)
        H = (to_complex(coeff) * w.view(-1, *([1] * (coeff.dim() - 1)))).sum(dim=0)
        return torch.diag(H) if self._parallel else H


class parallelFilter(Filter):
    """Per-channel FIR filters, param (taps, N) (reference dsp.py:965-1049)."""

    _parallel = True

    def __init__(self, size: tuple = (1, 1), **kwargs):
        super().__init__(size=size, **kwargs)

    def check_param_shape(self):
        assert len(self.size) == 2, "Filter must be 1D, for 2D filters use Filter module."


class _SectionFilter(Filter):
    """Shared machinery of Biquad / SVF / GEQ: a cascade of second-order sections per channel pair.
    Subclasses provide `_taps(mapped) -> (b, a)` of shape (3, K, ...)."""

    def initialize_class(self):
        self.check_param_shape()
        self.get_io()
        self.get_freq_response()
        self.get_freq_convolve()
        self._stock_freq_response = self.freq_response

    def _envelope(self, device, dtype):
        gamma = 10 ** (-abs(self._alias_db) / self.nfft / 20)
        return gamma ** torch.arange(0, 3, device=device, dtype=dtype)

    def get_freq_response(self):
        self.freq_response = lambda param: self.get_poly_coeff(self.map(param))[0]

    def get_poly_coeff(self, param):
        """(H, B, A) with B, A the per-section responses (M, K, ...), as the reference returns them
        (dsp.py:1464-1526); evaluated in closed form on the bin grid instead of zero-padded FFTs."""
        b, a = self._taps(param)
        env = self._envelope(b.device, b.dtype)
        shape = (3,) + (1,) * (b.dim() - 1)
        om = self._omega(b.dtype, b.device)
        zs = torch.exp(-1j * om.view(-1, 1) * torch.arange(3, device=b.device, dtype=b.dtype).view(1, 3))  # (M, 3)
        B = torch.einsum("fp,p...->f...", zs, to_complex(b * env.view(shape)))
        A = torch.einsum("fp,p...->f...", zs, to_complex(a * env.view(shape)))
        num, den = torch.prod(B, dim=1), torch.prod(A, dim=1)
        H = torch.where(torch.abs(den) != 0, num / den, torch.finfo(num.dtype).eps * torch.ones_like(num))
        return H, B, A

    def _fused_design(self, param):
        """Packed section coefficients straight from the raw parameter in one library launch, where one exists for this
        module (Biquad low-/high-pass with its stock map); None otherwise."""
        return None

    _fold_axis = None  # axis of the parameter that indexes the sections, for classes whose stock design treats them independently

    def _stock_design(self) -> bool:
        """The module still uses its stock map / tap formulas (subclasses that can fold per-item parameters say when)."""
        return False

    def _items_folded(self, prog, params) -> bool:
        """Per-item parameter sets of a section cascade whose design is elementwise over the sections (Biquad, SVF with
        their stock maps): the batch axis is folded into the section axis, so ONE design call yields the packed
        coefficients of every item, [B*K][...] = [B][K][...]."""
        ax = self._fold_axis
        if ax is None or self._overridden() or not self._stock_design():
            return False
        B = params.shape[0]
        moved = params.movedim(0, ax)  # (..., B, K, ...)
        flat = moved.reshape(*moved.shape[:ax], B * moved.shape[ax + 1], *moved.shape[ax + 2:])
        coef = self._fused_design(flat)
        if coef is None:
            b, a = self._taps(self.map(self._up(flat)))
            coef = sweep.pack_sections(b, a, self._parallel, None)
        K = coef.shape[0] // B
        coef = coef.view(B, K, *coef.shape[1:])
        prog.leaf_items((OP_PSOS if self._parallel else OP_SOS, int(self.output_channels), int(self.input_channels), K,
                         0, 0, 0, 0), coef)
        return True

    def _overridden(self):
        return (self.freq_response is not self._stock_freq_response
                or type(self).get_poly_coeff is not _SectionFilter.get_poly_coeff)

    def _emit(self, prog, param):
        if self._overridden():  # user / subclass supplies its own response: stream it as a table
            self._emit_table(prog, param, upcast=False)
            return
        coef = self._fused_design(param)
        if coef is None:
            b, a = self._taps(self.map(self._up(param)))
            coef = sweep.pack_sections(b, a, self._parallel, None)
        prog.leaf(OP_PSOS if self._parallel else OP_SOS, self.output_channels, self.input_channels, coef,
                  K=coef.shape[0])

    def probe(self, z):
        b, a = self._taps(self.map(self.param))
        env = self._envelope(b.device, b.dtype).view(3, *([1] * (b.dim() - 1)))
        zs = (z ** (-torch.arange(3, device=b.device, dtype=b.dtype))).view(3, *([1] * (b.dim() - 1)))
        H = torch.prod((to_complex(b * env) * zs).sum(0), dim=0) / torch.prod((to_complex(a * env) * zs).sum(0), dim=0)
        return torch.diag(H) if self._parallel else H


class Biquad(_SectionFilter):
    """Cascade of RBJ low/high/band-pass sections; param (K, 2|3, N_out, N_in) = normalised cut-off(s)
    and linear gain (reference dsp.py:1353-1603)."""

    def __init__(self, size: tuple = (1, 1), n_sections: int = 1, filter_type: str = "lowpass",
                 nfft: int = 2 ** 11, fs: int = 48000, requires_grad: bool = False, alias_decay_db: float = 0.0,
                 device: Optional[str] = None, dtype: torch.dtype = torch.float32):
        assert filter_type in ["lowpass", "highpass", "bandpass"], "Invalid filter type"
        self.n_sections, self.filter_type, self.fs = n_sections, filter_type, fs
        self.device, self.dtype = device, dtype
        self.get_map()
        self.alias_envelope_dcy = _alias_envelope(nfft, alias_decay_db, device, dtype)[:3].reciprocal()
        DSP.__init__(self, size=(n_sections, *self.get_size(), *size), nfft=nfft, map=self.map,
                     requires_grad=requires_grad, alias_decay_db=alias_decay_db, device=device, dtype=dtype)
        self.initialize_class()

    def get_size(self):
        return (3,) if self.filter_type == "bandpass" else (2,)

    def get_map(self):
        self.map = self._bounded_map

    def _bounded_map(self, x):
        """clamp([fc..., 20 log10 |g|]) to [0,1] / [-60,60] dB (reference dsp.py:1528-1563)."""
        cols = x.unbind(1)  # one UnbindBackward instead of a select_backward (zeros + copy + add) per column
        gain_db = 20 * torch.log10(torch.abs(cols[-1]))
        tail = (1,) * (x.dim() - 2)
        if self.filter_type == "bandpass":
            e = torch.finfo(self.dtype).eps
            v = torch.stack((cols[0], cols[1], gain_db), dim=1)
            lo, hi = [e, e, -60.0], [1 - e, 1 - e, 60.0]
        else:
            v = torch.stack((cols[0], gain_db), dim=1)
            lo, hi = [0.0, -60.0], [1.0, 60.0]
        lo = self._const("lo", lo, x).view(-1, *tail)
        hi = self._const("hi", hi, x).view(-1, *tail)
        return torch.clamp(v, min=lo, max=hi)

    def init_param(self):
        with torch.no_grad():
            torch.nn.init.uniform_(self.param[:, 0], a=0, b=0.5)
            if self.filter_type == "bandpass":
                torch.nn.init.uniform_(self.param[:, 1], a=self.param[:, 0].max().item(), b=1)
            torch.nn.init.uniform_(self.param[:, -1], a=-1, b=1)

    def check_param_shape(self):
        assert len(self.size) == 4, "Parameter size must be 4D, for 3D (parallel) biquads use parallelBiquad module."

    _fold_axis = 0

    def _stock_design(self):
        return getattr(self.map, "__func__", None) is Biquad._bounded_map and type(self)._taps is Biquad._taps

    def _fused_design(self, param):
        stock = self._stock_design() and self.filter_type in ("lowpass", "highpass")
        if not (stock and param.is_cuda and param.dtype in (torch.float32, torch.float64)
                and sweep._BACKEND.name == "cuda" and os.environ.get("FLAMO_B200_FUSED_DESIGN", "1") == "1"):
            return None
        return sweep.BiquadDesign.apply(param, self.output_channels, self.input_channels, self._parallel,
                                        self.filter_type == "highpass")

    def _taps(self, p):
        half_fs = self.fs / 2  # rad2hertz(param * pi)
        kw = dict(fs=self.fs, device=p.device, dtype=p.dtype)
        c = p.unbind(1)
        if self.filter_type == "lowpass":
            return lowpass_filter(fc=c[0] * half_fs, gain=c[1], **kw)
        if self.filter_type == "highpass":
            return highpass_filter(fc=c[0] * half_fs, gain=c[1], **kw)
        return bandpass_filter(fc1=c[0] * half_fs, fc2=c[1] * half_fs, gain=c[2], **kw)


class parallelBiquad(Biquad):
    """param (K, 2|3, N) (reference dsp.py:1607-1764)."""

    _parallel = True

    def __init__(self, size: tuple = (1,), **kwargs):
        super().__init__(size=size, **kwargs)

    def check_param_shape(self):
        assert len(self.size) == 3, "Parameter size must be 3D, for 3D sapce use Biquad module."


class SOSFilter(_SectionFilter):
    """Cascade of second-order sections given directly by their coefficients: param (K, 6, N_out, N_in) ordered
    [b0, b1, b2, a0, a1, a2]; the map optionally normalises every section by a0 (reference dsp.py:1767-1975).
    Lowered as the same SOS op as Biquad / SVF / GEQ."""

    def __init__(self, size: tuple = (1, 1), n_sections: int = 1, nfft: int = 2 ** 11, fs: int = 48000,
                 alias_decay_db: float = 0.0, device: Optional[str] = None, dtype: torch.dtype = torch.float32,
                 normalize_a0: bool = True):
        self.n_sections, self.fs, self.device, self.dtype, self.normalize_a0 = n_sections, fs, device, dtype, normalize_a0
        self.alias_envelope_dcy = _alias_envelope(nfft, alias_decay_db, device, dtype)[:3].reciprocal()
        self.get_map()
        DSP.__init__(self, size=(n_sections, *self.get_size(), *size), nfft=nfft, map=self.map, requires_grad=False,
                     alias_decay_db=alias_decay_db, device=device, dtype=dtype)
        self.initialize_class()

    def get_size(self):
        return (6,)

    def get_map(self):
        self.map = self._normalise

    def _normalise(self, x):
        if not self.normalize_a0:
            return x
        a0 = x[:, 3:4]
        eps = torch.finfo(x.dtype).eps
        safe = torch.where(torch.abs(a0) > eps, a0, torch.full_like(a0, eps))
        y = x / safe
        return torch.cat((y[:, :3], torch.ones_like(a0), y[:, 4:]), dim=1)

    def init_param(self):
        with torch.no_grad():
            self.param.zero_()
            self.param[:, 0] = 1.0
            self.param[:, 3] = 1.0

    def check_param_shape(self):
        assert len(self.size) == 4, "Parameter size must be 4D, expected (K, 6, N_out, N_in)."
        assert self.size[1] == 6, "Second dimension must be 6: [b0,b1,b2,a0,a1,a2]."

    def _taps(self, mapped):
        c = mapped.unbind(1)
        return torch.stack(c[:3], dim=0), torch.stack(c[3:], dim=0)


class parallelSOSFilter(SOSFilter):
    """param (K, 6, N) (reference dsp.py:1978-2073)."""

    _parallel = True

    def __init__(self, size: tuple = (1,), **kwargs):
        super().__init__(size=size, **kwargs)

    def check_param_shape(self):
        assert len(self.size) == 3, "Parameter size must be 3D, expected (K, 6, N)."
        assert self.size[1] == 6, "Second dimension must be 6: [b0,b1,b2,a0,a1,a2]."


class SVF(_SectionFilter):
    """State-variable filter sections; param (5, K, N_out, N_in) = (f, R, mLP, mBP, mHP) before their
    activations (reference dsp.py:2076-2373)."""

    _TYPES = ["lowpass", "highpass", "bandpass", "lowshelf", "highshelf", "peaking", "notch", None]

    def __init__(self, size: tuple = (1, 1), n_sections: int = 1, filter_type: str = None, nfft: int = 2 ** 11,
                 fs: int = 48000, requires_grad: bool = False, alias_decay_db: float = 0.0,
                 device: Optional[str] = None, dtype: torch.dtype = torch.float32):
        assert filter_type in self._TYPES, "Invalid filter type"
        self.fs, self.n_sections, self.filter_type = fs, n_sections, filter_type
        self.alias_envelope_dcy = _alias_envelope(nfft, alias_decay_db, device, dtype)[:3].reciprocal()
        DSP.__init__(self, size=(5, n_sections, *size), nfft=nfft, map=self.map_param2svf,
                     requires_grad=requires_grad, alias_decay_db=alias_decay_db, device=device, dtype=dtype)
        self.initialize_class()

    def check_param_shape(self):
        assert len(self.size) == 4, "Filter parameter space must be 4D, for 3D (parallel) filters use parallelSVF module."

    # activations (reference dsp.py:2234-2364)
    def param2freq(self, param):
        return torch.tan(math.pi * torch.sigmoid(param) * 0.5)

    def param2R(self, param):
        return F.softplus(param) / math.log(2.0)

    def param2mix(self, param, R=None):
        return torch.stack(self._mix(param.unbind(0), R), dim=0)

    def _mix(self, p, R):
        """(mLP, mBP, mHP) from the three raw mix parameters p = (p0, p1, p2)."""
        t = self.filter_type
        if t is None:
            return (p[0] + 1, p[1] + 2, p[2] + 1)
        G = 10 ** (-F.softplus(p[0]))
        one, zero = torch.ones_like(G), torch.zeros_like(G)
        if t == "lowpass":
            m = (one, zero, zero)
        elif t == "highpass":
            m = (zero, zero, one)
        elif t == "bandpass":
            m = (zero, one, zero)
        elif t == "lowshelf":
            m = (one, 2 * R * torch.sqrt(G), G)
        elif t == "highshelf":
            m = (G, 2 * R * torch.sqrt(G), one)
        elif t in ("peaking", "notch"):
            m = (one, 2 * R * torch.sqrt(G), one)
        return m

    def map_param2svf(self, param):
        p = param.unbind(0)  # one UnbindBackward (a stack) instead of five select_backward (zeros + copy + add)
        f = self.param2freq(p[0])
        r = self.param2R(p[1])
        R = 1 / r if self.filter_type == "peaking" else r
        m = self._mix(p[2:], r)
        return f, R, m[0], m[1], m[2]

    _fold_axis = 1

    def _stock_design(self):
        return (getattr(self.map, "__func__", None) is SVF.map_param2svf and type(self)._taps is SVF._taps
                and type(self)._mix is SVF._mix and type(self).param2freq is SVF.param2freq
                and type(self).param2R is SVF.param2R)

    def _fused_design(self, param):
        stock = self._stock_design() and self.filter_type is None
        if not (stock and param.is_cuda and param.dtype in (torch.float32, torch.float64)
                and sweep._BACKEND.name == "cuda" and os.environ.get("FLAMO_B200_FUSED_DESIGN", "1") == "1"):
            return None
        return sweep.SVFDesign.apply(param, self.output_channels, self.input_channels, self._parallel)

    def _taps(self, mapped):
        f, R, mLP, mBP, mHP = mapped
        f2 = f * f
        b = torch.stack((f2 * mLP + f * mBP + mHP, 2 * f2 * mLP - 2 * mHP, f2 * mLP - f * mBP + mHP))
        a = torch.stack((f2 + 2 * R * f + 1, 2 * f2 - 2, f2 - 2 * R * f + 1))
        return b, a


class parallelSVF(SVF):
    """param (5, K, N) (reference dsp.py:2377-2464)."""

    _parallel = True

    def __init__(self, size: tuple = (1,), **kwargs):
        super().__init__(size=size, **kwargs)

    def check_param_shape(self):
        assert len(self.size) == 3, "Filter parameter space must be 3D, for 4D filters use SVF module."


class GEQ(_SectionFilter):
    """Graphic equaliser: param (n_gains, N_out, N_in) linear command gains; `map` -> dB
    (reference dsp.py:2467-2611, auxiliary/eq.py:57-111)."""

    def __init__(self, size: tuple = (1, 1), octave_interval: int = 1, nfft: int = 2 ** 11, fs: int = 48000,
                 map: callable = lambda x: 20 * torch.log10(torch.abs(x)), requires_grad: bool = False,
                 alias_decay_db: float = 0.0, device: Optional[str] = None, dtype: torch.dtype = torch.float32):
        self.octave_interval, self.fs = octave_interval, fs
        self.center_freq, self.shelving_crossover = eq_freqs(interval=octave_interval)
        self.n_gains = len(self.center_freq) + 3
        self._R = float(torch.tensor(2.7))  # the reference's float32 constant (dsp.py:2575)
        self.alias_envelope_dcy = _alias_envelope(nfft, alias_decay_db, device, dtype)[:3].reciprocal()
        DSP.__init__(self, size=(self.n_gains, *size), nfft=nfft, map=map, requires_grad=requires_grad,
                     alias_decay_db=alias_decay_db, device=device, dtype=dtype)
        self.initialize_class()

    def init_param(self):
        torch.nn.init.uniform_(self.param, a=10 ** (-6 / 20), b=10 ** (6 / 20))

    def check_param_shape(self):
        assert len(self.size) == 3, "Filter must be 3D, for 2D (parallel) filters use ParallelGEQ module."

    def _taps(self, gain_db):
        cf = self._const("cf", self.center_freq, gain_db)
        sf = self._const("sf", self.shelving_crossover, gain_db)
        return geq(center_freq=cf, shelving_freq=sf, R=self._R, gain_db=gain_db, fs=self.fs,
                   device=gain_db.device, dtype=gain_db.dtype)


class parallelGEQ(GEQ):
    """param (n_gains, N) (reference dsp.py:2614-2692)."""

    _parallel = True

    def __init__(self, size: tuple = (1,), **kwargs):
        super().__init__(size=size, **kwargs)

    def check_param_shape(self):
        assert len(self.size) == 2, "Filter must be 2D, for 3D filters use GEQ module."


# =========================================================================================== delays


class Delay(DSP):
    """Delay lines: H = gamma^m exp(-j omega m), m = map(param) * fs / unit samples, rounded when
    `isint` (reference dsp.py:3226-3450).  Learnable delays use a softplus map."""

    def __init__(self, size: tuple = (1, 1), max_len: int = 2000, isint: bool = False, unit: int = 100,
                 nfft: int = 2 ** 11, fs: int = 48000, requires_grad: bool = False, alias_decay_db: float = 0.0,
                 device: Optional[str] = None, dtype: torch.dtype = torch.float32):
        self.fs, self.max_len, self.unit, self.isint = fs, max_len, unit, isint
        super().__init__(size=size, nfft=nfft, requires_grad=requires_grad, alias_decay_db=alias_decay_db,
                         device=device, dtype=dtype)
        self.initialize_class()

    def init_param(self):
        if self.isint:
            delay_len = torch.randint(1, int(self.max_len), self.size, device=self.device)
        else:
            delay_len = torch.rand(self.size, device=self.device) * self.max_len
        self.assign_value(self.sample2s(delay_len).to(self.param.dtype))
        self.order = delay_len.max() + 1

    def s2sample(self, delay):
        return delay * self.fs / self.unit

    def sample2s(self, delay):
        return delay / self.fs * self.unit

    def initialize_class(self):
        self.check_param_shape()
        self.get_io()
        if self.requires_grad:
            self.map = lambda x: F.softplus(x)
        self.omega = self._omega(self.dtype, self.device).unsqueeze(1)
        self.get_freq_response()
        self.get_freq_convolve()

    def check_param_shape(self):
        assert len(self.size) == 2, "delay must be 2D, for 1D (parallel) delay use parallelDelay module."

    def check_input_shape(self, x):
        if not self._bins_ok(x) or self.input_channels != x.shape[2]:
            raise ValueError(
                f"parameter shape = {self.param.shape} not compatible with input signal of shape = ({x.shape}).")

    def get_io(self):
        self.input_channels = self.size[-1]
        self.output_channels = self.size[-1] if self._parallel else self.size[-2]

    def get_delays(self):
        return lambda param: self.s2sample(self.map(param))

    def get_freq_response(self):
        self.freq_response = self._delay_response

    def _delay_response(self, param):
        m = self.s2sample(self.map(param))
        if self.isint:
            m = m.round()
        om = self._omega(m.dtype, m.device).view(-1, *([1] * m.dim()))
        return ((10 ** (-abs(self._alias_db) / self.nfft / 20)) ** m) * torch.exp(-1j * om * m.unsqueeze(0))

    def get_freq_convolve(self):
        self.freq_convolve = lambda x, param: self._sweep(x, param)

    def _emit(self, prog, param):
        m = self.s2sample(self.map(self._up(param)))
        if self._parallel:
            prog.leaf(OP_PDELAY, self.output_channels, self.input_channels, m.reshape(-1), isint=self.isint)
        else:
            prog.leaf(OP_DELAY, self.output_channels, self.input_channels, m, isint=self.isint)

    def probe(self, z):
        m = self.s2sample(self.map(self.param))
        if self.isint:
            m = m.round()
        H = (self.gamma ** m) * (1.0 / z) ** m
        return torch.diag_embed(H) if self._parallel else H


class parallelDelay(Delay):
    """param (N,) (reference dsp.py:3453-3551)."""

    _parallel = True

    def __init__(self, size: tuple = (1,), max_len: int = 2000, unit: int = 100, isint: bool = False, nfft=2 ** 11,
                 fs: int = 48000, requires_grad: bool = False, alias_decay_db: float = 0.0,
                 device: Optional[str] = None, dtype: torch.dtype = torch.float32):
        super().__init__(size=size, max_len=max_len, isint=isint, unit=unit, nfft=nfft, fs=fs,
                         requires_grad=requires_grad, alias_decay_db=alias_decay_db, device=device, dtype=dtype)

    def check_param_shape(self):
        assert len(self.size) == 1, "delays must be 1D, for 2D delays use Delay module."


class GainDelay(DSP):
    """Gain and delay per channel pair: H[m][n] = map_gain(g)[m][n] gamma^d exp(-j omega d), d = map_delay(param[1])
    * fs / unit samples; param (2, N_out, N_in) (reference dsp.py:3554-3724).  The dense variant is an elementwise
    (not a matrix) product of a gain and a delay matrix, so it is lowered as a streamed response table built in
    float64 with exact integer phases; the parallel variant is a PGAIN followed by a PDELAY."""

    def __init__(self, size: tuple = (1, 1), max_len: int = 2000, isint: bool = False, unit: int = 100,
                 nfft: int = 2 ** 11, fs: int = 48000, map_gain: Optional[callable] = None,
                 map_delay: Optional[callable] = None, requires_grad: bool = False, alias_decay_db: float = 0.0,
                 device: Optional[str] = None, dtype: torch.dtype = torch.float32):
        self.fs, self.max_len, self.unit, self.isint = fs, max_len, unit, isint
        self._custom_gain_map, self._custom_delay_map = map_gain is not None, map_delay is not None
        self.map_gain = map_gain if map_gain is not None else _identity
        self.map_delay = map_delay if map_delay is not None else _identity
        super().__init__(size=(2, *size), nfft=nfft, requires_grad=requires_grad, alias_decay_db=alias_decay_db,
                         device=device, dtype=dtype)
        self.initialize_class()

    def init_param(self):
        shape = self.size[1:]
        with torch.no_grad():
            nn.init.ones_(self.param[0])
            if self.isint:
                d = torch.randint(1, self.max_len, shape, device=self.device, dtype=torch.int64).to(self.param.dtype)
            else:
                d = torch.rand(shape, device=self.device, dtype=self.dtype) * self.max_len
            self.param[1].copy_(self.sample2s(d))
        self.order = int(torch.ceil(d).max().item()) + 1

    def s2sample(self, delay):
        return delay * self.fs / self.unit

    def sample2s(self, delay):
        return delay / self.fs * self.unit

    def check_input_shape(self, x):
        if not self._bins_ok(x) or self.input_channels != x.shape[2]:
            raise ValueError(
                f"parameter shape = {self.param.shape} not compatible with input signal of shape = ({x.shape}).")

    def check_param_shape(self):
        assert len(self.size) == 3 and self.size[0] == 2, "GainDelay parameters must have shape (2, N_out, N_in)."

    def get_io(self):
        self.input_channels = self.size[-1]
        self.output_channels = self.size[-1] if self._parallel else self.size[-2]

    def get_gains(self):
        return lambda param: to_complex(self.map_gain(param[0]))

    def get_delays(self):
        return lambda param: self.s2sample(self.map_delay(param[1]))

    def initialize_class(self):
        self.check_param_shape()
        self.get_io()
        if self.requires_grad and not self._custom_delay_map:
            self.map_delay = lambda x: F.softplus(x)
        self.omega = self._omega(self.dtype, self.device).unsqueeze(1)
        self.get_freq_response()
        self.get_freq_convolve()

    def get_freq_response(self):
        self.freq_response = self._response

    def _response(self, param):
        """(M, N_out, N_in) | (M, N) response; integer delays take their phase index k*d mod nfft in integers."""
        g = self.map_gain(param[0])
        d = self.s2sample(self.map_delay(param[1])).to(g.dtype)
        k = torch.arange(0, self.nfft // 2 + 1, device=d.device).view(-1, *([1] * d.dim()))
        if self.isint:
            d = d.round()
            idx = torch.remainder(k * d.detach().to(torch.int64).unsqueeze(0), self.nfft).to(d.dtype)
            ph = (2 * math.pi / self.nfft) * idx
        else:
            ph = (2 * math.pi / self.nfft) * k.to(d.dtype) * d.unsqueeze(0)
        mag = (g * (10 ** (-abs(self._alias_db) / self.nfft / 20)) ** d).unsqueeze(0)
        if mag.is_complex():  # a custom map_gain may return complex gains
            return mag * torch.exp(-1j * ph)
        return torch.complex(mag * torch.cos(ph), -mag * torch.sin(ph))

    def get_freq_convolve(self):
        self.freq_convolve = lambda x, param: self._sweep(x, param)

    def _emit(self, prog, param):
        p = self._up(param)
        if self._parallel:
            g = self.map_gain(p[0])
            prog.leaf(OP_PGAIN, self.output_channels, self.input_channels, g.reshape(-1))
            prog.leaf(OP_PDELAY, self.output_channels, self.input_channels, self.s2sample(self.map_delay(p[1])).reshape(-1),
                      isint=self.isint)
            return
        H = self._response(p)
        prog.leaf(OP_TABLE, self.output_channels, self.input_channels, H)

    def probe(self, z):
        g = to_complex(self.map_gain(self.param[0]))
        m = self.s2sample(self.map_delay(self.param[1]))
        if self.isint:
            m = m.round()
        H = g * (self.gamma ** m) * (1.0 / z) ** m
        return torch.diag_embed(H) if self._parallel else H


class parallelGainDelay(GainDelay):
    """param (2, N) (reference dsp.py:3727-3778)."""

    _parallel = True

    def __init__(self, size: tuple = (1,), **kwargs):
        super().__init__(size=size, **kwargs)

    def check_param_shape(self):
        assert len(self.size) == 2 and self.size[0] == 2, (
            "parallelGainDelay parameters must have shape (2, N), for MIMO use GainDelay module.")
