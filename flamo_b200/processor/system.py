"""flamo.processor.system — Series, Recursion and Shell with the reference's interface
(gdalsanto/flamo, flamo/processor/system.py), evaluated as ONE fused sweep launch per container.

A container does not call its children one by one: it asks each child to append its ops to a flat
sweep program (`_lower`) and runs the program through flamo_b200.sweep.  The closed loop of a
Recursion, y = (I - F Fb)^-1 F x (reference system.py:417-425), becomes a RECURSION op solved per bin
in registers; the (B, M, N, N) tensors the reference materialises never exist.
"""
from __future__ import annotations

import warnings
from collections import OrderedDict

import torch
import torch.nn as nn

from .. import sweep
from .._lib import EPI_ABS, EPI_NONE
from ..functional import signal_gallery
from .dsp import FFT, Transform, iFFT

# ============================================================================================ series


def _flatten(modules, taken):
    """Flatten nested Sequential / dict containers into an OrderedDict with the reference's key rules
    (system.py:127-209): integer-like keys are renumbered by position, custom keys must be unique."""
    out = OrderedDict()

    def position():
        return str(len(out) + len(taken))

    def visit(obj, key=None):
        if isinstance(obj, nn.Sequential):
            visit(obj._modules)
        elif isinstance(obj, (OrderedDict, dict)):
            for k, v in obj.items():
                if isinstance(v, (nn.Sequential, OrderedDict, dict)):
                    visit(v)
                else:
                    visit(v, k)
        elif isinstance(obj, nn.Module):
            if key is None:
                out[position()] = obj
                return
            try:
                int(key)
                numeric = True
            except (TypeError, ValueError):
                numeric = False
            if numeric:
                new_key = position()
                out[new_key] = obj
                if key != new_key:
                    warnings.warn(f"Key {key} is an integer, it will be overwritten.")
            else:
                if key in taken or key in out:
                    raise ValueError(f"Key {key} is already present in the Series.")
                out[key] = obj
        else:
            raise ValueError("Modules must be nn.Module, nn.Sequential, or OrderedDict.")

    for m in modules:
        visit(m)
    return out


def _common_attribute(modules, attr):
    """Value of `attr` shared by every module that has it (system.py:211-240)."""
    value, found = None, False
    for m in modules:
        if hasattr(m, attr):
            value, found = getattr(m, attr), True
            break
    if not found or value is None:
        warnings.warn(f"Attribute {attr} not found in any of the modules.")
        return None
    for i, m in enumerate(modules):
        if hasattr(m, attr) and getattr(m, attr) != value:
            raise ValueError(
                f"All modules must have the same {attr} value. Module {m.__class__.__name__} at index {i} is "
                "incoherent with the part of the Series preceding it.")
    return value


def _chain_io(modules):
    """Input channels of the first module that declares them and output channels of the last,
    asserting that consecutive declarations match (system.py:242-277)."""
    first_in, prev_out, prev_name, prev_idx = None, None, None, None
    for j, m in enumerate(modules):
        if not hasattr(m, "input_channels"):
            continue
        if first_in is None:
            first_in = m.input_channels
        else:
            assert m.input_channels == prev_out, (
                f"Module {prev_name} at index {prev_idx} has {prev_out} output channels, but module "
                f"{m.__class__.__name__} at index {j} has {m.input_channels} input_channels.")
        prev_out, prev_name, prev_idx = getattr(m, "output_channels", None), m.__class__.__name__, j
    return first_in, prev_out


def _entry_check(module, x):
    """check_input_shape of the FIRST leaf the signal meets (a Series hands it to its first module, a Recursion to its
    feedforward path): the reference raises there (dsp.py:441-444, 880-883), before anything is evaluated."""
    while module is not None:
        if hasattr(module, "check_input_shape"):
            module.check_input_shape(x)
            return
        if isinstance(module, Series):
            module = module[0] if len(module) else None
        elif isinstance(module, Recursion):
            module = module.feedforward
        else:
            return


def _alias_of(module) -> float:
    a = getattr(module, "_alias_db", None)
    if a is None:
        a = getattr(module, "alias_decay_db", 0.0)
        a = float(a) if a is not None else 0.0
    return a


class Series(nn.Sequential):
    """Modules applied one after the other (reference system.py:11-329)."""

    def __init__(self, *args):
        super().__init__(_flatten(args, []))
        self._refresh()

    def _refresh(self):
        mods = list(self)
        self.nfft = _common_attribute(mods, "nfft")
        self.alias_decay_db = _common_attribute(mods, "alias_decay_db")
        self.dtype = _common_attribute(mods, "dtype")
        self.input_channels, self.output_channels = _chain_io(mods)
        self._alias_db = _alias_of(self)

    def prepend(self, new_module):
        return self.insert(0, new_module)

    def append(self, new_module):
        for k, v in _flatten((new_module,), list(self._modules.keys())).items():
            self.add_module(k, v)
        self._refresh()
        return self

    def insert(self, index: int, new_module):
        n = len(self._modules)
        if not (-n <= index <= n):
            raise IndexError("Index out of range.")
        if index < 0:
            index += n
        items = list(self._modules.items())
        items[index:index] = list(_flatten((new_module,), list(self._modules.keys())).items())
        self._modules.clear()
        self._modules.update(items)
        self._refresh()
        return self

    # -- evaluation ------------------------------------------------------------------------------
    def _lower(self, prog, ext_param=None):
        for key, module in self._modules.items():
            ext = ext_param[key] if (ext_param is not None and key in ext_param) else None
            if hasattr(module, "_lower"):
                module._lower(prog, ext)
            elif prog._chain is not None and isinstance(module, Parallel):
                prog.table_of(module, ext)  # a Parallel inside a Recursion path: its response as a streamed table
            elif ext is not None:
                prog.eager(lambda t, m=module, e=ext: m(t, e))
            else:
                prog.eager(module)

    def forward(self, input, ext_param=None):
        if not (torch.is_tensor(input) and input.is_complex()) or self.nfft is None:
            # not a bin-domain signal (e.g. a Series used as a time-domain layer): plain sequential
            for key, module in self._modules.items():
                if ext_param is not None and key in ext_param:
                    input = module(input, ext_param[key])
                else:
                    input = module(input)
            return input
        _entry_check(self, input)
        prog = sweep.Program(self.nfft, self._alias_db, input.dtype, input.device)
        self._lower(prog, ext_param)
        return prog.run(input)

    def probe(self, z):
        H = None
        for m in self:
            Hi = m.probe(z)
            H = Hi if H is None else Hi @ H
        return H

    def probe_w(self, w):
        H = None
        for m in self:
            Hi = m.probe_w(w)
            H = Hi if H is None else Hi @ H
        return H


# ========================================================================================= recursion


class Recursion(nn.Module):
    """Closed loop with feedforward path fF and feedback path fB (reference system.py:335-565)."""

    def __init__(self, fF, fB):
        super().__init__()
        self.feedforward = self._as_path(fF, "Feedforward")
        self.feedback = self._as_path(fB, "Feedback")
        self.nfft = self._shared("nfft")
        self.alias_decay_db = self._shared("alias_decay_db")
        self.dtype = self._shared("dtype")
        self.input_channels, self.output_channels = self._check_io()
        self._alias_db = _alias_of(self.feedforward)
        self._eye = None

    @staticmethod
    def _as_path(p, name):
        if isinstance(p, (nn.Sequential, OrderedDict)) and not isinstance(p, Series):
            warnings.warn(f"{name} path has been converted to a Series class instance.")
            return Series(p)
        return p

    def _shared(self, attr):
        a, b = getattr(self.feedforward, attr, None), getattr(self.feedback, attr, None)
        if a is None:
            warnings.warn(f"The feedforward pass does not possess the attribute {attr}.")
        if b is None:
            warnings.warn(f"The feedback pass does not possess the attribute {attr}.")
        if a is not None and b is not None:
            assert a == b, (f"The feedforward pass has {attr} = {a} and feedback pass has {attr} = {b}. "
                            "They must have the same value.")
        return a if a is not None else b

    def _check_io(self):
        io = {}
        for name, path in (("feedforward", self.feedforward), ("feedback", self.feedback)):
            for side in ("input_channels", "output_channels"):
                v = getattr(path, side, None)
                if v is None:
                    raise ValueError(f"The {name} pass does not possess the attribute {side}.")
                io[name, side] = v
        ff_in, ff_out = io["feedforward", "input_channels"], io["feedforward", "output_channels"]
        fb_in, fb_out = io["feedback", "input_channels"], io["feedback", "output_channels"]
        assert ff_out == fb_in, (f"Feedforward pass has {ff_out} output channels, but feedback pass has {fb_in} "
                                 "input channels. They must be the same.")
        assert fb_out == ff_in, (f"Feedforward pass {ff_in} input channels, but the feedback pass has {fb_out} "
                                 "output channels. They must be the same.")
        return ff_in, ff_out

    @property
    def I(self):
        """(M, N, N) complex identity, built on first use (the reference allocates it eagerly,
        system.py:427-438; the sweep never needs it)."""
        if self._eye is None:
            n, M = self.output_channels, self.nfft // 2 + 1
            dev = self.alias_decay_db.device if torch.is_tensor(self.alias_decay_db) else None
            self._eye = torch.eye(n, dtype=self.dtype, device=dev).to(torch.complex64 if self.dtype == torch.float32
                                                                       else torch.complex128).expand(M, n, n)
        return self._eye

    def _lower(self, prog, ext_param=None):
        ext_ff = ext_fb = None
        if ext_param is not None:
            for key, p in ext_param.items():
                if "feedback" in key:
                    ext_fb = p
                elif "feedforward" in key:
                    ext_ff = p
        if prog._chain is not None:
            # a loop nested inside another loop's path: its closed-loop response enters the outer program as a table
            prog.table_of(self, ext_param)
            return

        def lower(path, ext):
            if hasattr(path, "_lower"):
                path._lower(prog, ext)
            elif isinstance(path, Parallel):
                prog.table_of(path, ext)
            else:
                raise sweep._lib.Unsupported(sweep._lib.E_UNSUPPORTED,
                                             f"{path.__class__.__name__} cannot be lowered inside a Recursion")

        prog.recursion(lambda: lower(self.feedforward, ext_ff), lambda: lower(self.feedback, ext_fb))

    def forward(self, X, ext_param=None):
        _entry_check(self, X)
        prog = sweep.Program(self.nfft, self._alias_db, X.dtype, X.device)
        self._lower(prog, ext_param)
        return prog.run(X)

    def probe(self, z):
        Fm, Bm = self.feedforward.probe(z), self.feedback.probe(z)
        A = torch.eye(Fm.shape[-2], dtype=Fm.dtype, device=Fm.device) - Fm @ Bm
        return torch.linalg.solve(A, Fm)

    def probe_recursion(self, z, include_shell_io: bool = False, **kwargs):
        Fm, Bm = self.feedforward.probe(z), self.feedback.probe(z)
        return torch.eye(Fm.shape[0], dtype=Fm.dtype, device=Fm.device) - Fm @ Bm

    def probe_recursion_w(self, w):
        Fm, Bm = self.feedforward.probe_w(w), self.feedback.probe_w(w)
        return torch.eye(Fm.shape[0], dtype=Fm.dtype, device=Fm.device) - Fm @ Bm


# ========================================================================================== parallel


class Parallel(nn.Module):
    """Two branches fed with the same input; outputs summed or concatenated along the channel axis (reference
    system.py:570-772).  Each branch is lowered and launched as its own fused sweep program; the sum / concatenation
    is the only PyTorch op.  Inside a Series it is therefore a launch boundary; inside a Recursion path it is
    not supported (FSWEEP_E_UNSUPPORTED)."""

    def __init__(self, brA, brB, sum_output: bool = True):
        super().__init__()
        self.branchA = self._as_branch(brA, "A")
        self.branchB = self._as_branch(brB, "B")
        self.sum_output = sum_output
        self.nfft = self._shared("nfft")
        self.alias_decay_db = self._shared("alias_decay_db")
        self.dtype = self._shared("dtype")
        self.input_channels, self.output_channels = self._check_io()

    @staticmethod
    def _as_branch(b, name):
        if isinstance(b, (nn.Sequential, OrderedDict)) and not isinstance(b, Series):
            warnings.warn(f"Branch {name} has been converted to a Series class instance.")
            return Series(b)
        return b

    def _shared(self, attr):
        a, b = getattr(self.branchA, attr, None), getattr(self.branchB, attr, None)
        if a is None:
            warnings.warn(f"The feedforward pass does not possess the attribute {attr}.")
        if b is None:
            warnings.warn(f"The feedback pass does not possess the attribute {attr}.")
        if a is not None and b is not None:
            assert a == b, (f"Branch A has {attr} = {a} and branch B has {attr} = {b}. They must have the same value.")
        return a if a is not None else b

    def _check_io(self):
        io = {}
        for name, br in (("A", self.branchA), ("B", self.branchB)):
            for side in ("input_channels", "output_channels"):
                v = getattr(br, side, None)
                if v is None:
                    raise ValueError(f"Branch {name} does not possess the attribute {side}.")
                io[name, side] = v
        assert io["A", "input_channels"] == io["B", "input_channels"], (
            f"Branch A has {io['A', 'input_channels']} input channels, but branch B has "
            f"{io['B', 'input_channels']} input channels. They must be the same.")
        if self.sum_output:
            assert io["A", "output_channels"] == io["B", "output_channels"], (
                f"Branch A has {io['A', 'output_channels']} output channels, but branch B has "
                f"{io['B', 'output_channels']} output channels. They must be the same if their output is being summed.")
            return io["A", "input_channels"], io["A", "output_channels"]
        return io["A", "input_channels"], io["A", "output_channels"] + io["B", "output_channels"]

    def forward(self, X, ext_param: dict = None):
        ext_a = ext_b = None
        if ext_param is not None:
            for key, p in ext_param.items():
                if "branchA" in key:
                    ext_a = p
                elif "branchB" in key:
                    ext_b = p
        YA = self.branchA(X, ext_a) if ext_a is not None else self.branchA(X)
        YB = self.branchB(X, ext_b) if ext_b is not None else self.branchB(X)
        return YA + YB if self.sum_output else torch.cat((YA, YB), dim=2)

    def probe(self, z):
        HA, HB = self.branchA.probe(z), self.branchB.probe(z)
        return HA + HB if self.sum_output else torch.cat([HA, HB], dim=0)

    def probe_w(self, w):
        HA, HB = self.branchA.probe_w(w), self.branchB.probe_w(w)
        return HA + HB if self.sum_output else torch.cat([HA, HB], dim=0)


# ============================================================================================= shell


def _is_abs_layer(layer) -> bool:
    """Cheap pre-filter: True if `layer` is a plain Transform that maps a tiny probe tensor to |x|.  Necessary, not
    sufficient (abs(x[:, :1000]), clamp(abs(x)), abs(x) + eps all pass it): `_abs_layer_verified` confirms the
    candidate on the real signal before anything is fused."""
    if not isinstance(layer, Transform) or type(layer) is not Transform:
        return False
    flag = getattr(layer, "_fsweep_is_abs", None)
    if flag is None:
        flag = False
        try:
            t = torch.tensor([[[3.0 + 4.0j], [-1.0 - 1.0j], [0.0 + 0.0j], [0.5j]]], dtype=torch.complex64)
            r = layer.transform(t)
            flag = bool(torch.is_tensor(r) and not r.is_complex() and r.shape == t.shape
                        and torch.allclose(r, torch.abs(t)))
        except Exception:
            flag = False
        layer._fsweep_is_abs = flag
    return flag


def _abs_layer_verified(layer, prog, X) -> bool:
    """The |.| epilogue replaces `layer` only after the layer has been RUN once on the unfused output of this very
    program for this signal shape and agreed with |Y| in shape, dtype and values (a transform that crops, clamps or
    offsets is then kept as it is).  One extra forward sweep per (layer, shape), outside any CUDA-graph capture."""
    if not _is_abs_layer(layer):
        return False
    seen = layer.__dict__.setdefault("_fsweep_abs_verified", {})
    key = (tuple(X.shape), X.dtype, current_shard_key())
    ok = seen.get(key)
    if ok is None:
        if X.is_cuda and torch.cuda.is_current_stream_capturing():
            return False  # cannot verify (host read) inside a capture: stay unfused
        try:
            with torch.no_grad():
                Y = prog.run(X, epilogue=EPI_NONE)
                ok = True
                for s in (1.0, 2.0 ** 20, 2.0 ** -20):  # other scales expose clamps and offsets this signal misses
                    r, ref = layer(Y * s), torch.abs(Y * s)
                    ok = ok and bool(torch.is_tensor(r) and r.shape == ref.shape and r.dtype == ref.dtype
                                     and torch.allclose(r, ref, rtol=1e-6, atol=0.0))
        except Exception:
            ok = False
        seen[key] = ok
    return ok


def current_shard_key():
    return sweep.current_shard()


class Shell(nn.Module):
    """input_layer -> core -> output_layer (reference system.py:776-1153)."""

    def __init__(self, core, input_layer=nn.Identity(), output_layer=nn.Identity()):
        super().__init__()
        self.__core = self._wrap(core, "Core")
        self.__input_layer = self._wrap(input_layer, "Input layer")
        self.__output_layer = self._wrap(output_layer, "Output layer")
        self.nfft = self._shared("nfft")
        self.alias_decay_db = self._shared("alias_decay_db")
        self.dtype = self._shared("dtype")
        self.input_channels, self.output_channels = self._check_io()
        self.fuse_output = True  # fold an |.| output layer into the sweep epilogue

    @staticmethod
    def _wrap(m, name):
        if isinstance(m, (nn.Sequential, OrderedDict)) and not isinstance(m, Series):
            warnings.warn(f"{name} has been converted to a Series class instance.")
            return Series(m)
        return m

    def _shared(self, attr):
        core = getattr(self.__core, attr, None)
        if core is None:
            raise ValueError(f"The core does not possess the attribute {attr}.")
        for name, layer in (("input layer", self.__input_layer), ("output layer", self.__output_layer)):
            v = getattr(layer, attr, None)
            if v is not None:
                assert core == v, (f"The {name} has {attr} = {v} and the core has {attr} = {core}. "
                                   "They must have the same value.")
        return core

    def _check_io(self):
        c_in, c_out = getattr(self.__core, "input_channels", None), getattr(self.__core, "output_channels", None)
        if c_in is None:
            raise ValueError("The core does not possess the attribute input_channels.")
        if c_out is None:
            raise ValueError("The core does not possess the attribute output_channels.")
        l_out = getattr(self.__input_layer, "output_channels", None)
        if l_out is not None:
            assert c_in == l_out, (f"The core should receive {c_in} input channels, but {l_out} channels arrive "
                                   "from the input layer.")
        o_in = getattr(self.__output_layer, "input_channels", None)
        if o_in is not None:
            assert c_out == o_in, (f"The core sends {c_out} output channels, but the output layer can only "
                                   f"receive {o_in} channels.")
        in_ch = getattr(self.__input_layer, "input_channels", None)
        out_ch = getattr(self.__output_layer, "output_channels", None)
        return (in_ch if in_ch is not None else c_in), (out_ch if out_ch is not None else c_out)

    # -- evaluation ------------------------------------------------------------------------------
    def _invalidate_caches(self):
        for m in self.modules():
            f = getattr(m, "invalidate_cache", None)
            if f is not None:
                f()

    def _input_and_program(self, x, ext_param):
        """Input layer, then lowering of the core into a sweep program (None if the core cannot be lowered or the
        input layer did not produce a bin-domain tensor).  Running the input-layer FFT on a side stream, concurrently
        with the parameter maps, was measured and dropped: the cross-stream edges cost a captured step 9 us
        (profiles/r01_notes.md)."""
        core = self.__core
        X = self.__input_layer(x)
        prog = None
        if hasattr(core, "_lower") and torch.is_tensor(X) and X.is_complex():
            _entry_check(core, X)
            prog = sweep.Program(self.nfft, _alias_of(core), X.dtype, X.device)
            core._lower(prog, ext_param)
        return X, prog

    def forward(self, x, ext_param=None, keep_caches: bool = False):
        if not keep_caches:
            self._invalidate_caches()  # memoised maps live for one forward (+ its criteria) only
        core, out = self.__core, self.__output_layer
        x, prog = self._input_and_program(x, ext_param)
        if prog is not None:
            if self.fuse_output and _abs_layer_verified(out, prog, x):
                return prog.run(x, epilogue=EPI_ABS)
            return out(prog.run(x, epilogue=EPI_NONE))
        x = core(x, ext_param) if ext_param is not None else core(x)
        return out(x)

    def forward_loss(self, x, target, kind: int, ext_param=None, keep_caches: bool = False):
        """criterion(self(x), target) with the |.| output layer and the criterion fused into the sweep kernel
        (sweep.SweepLossFunction; kind = _lib.CRIT_MSE | CRIT_MSE_CHSUM).  Returns None when this Shell cannot be
        fused (output layer is not |.|, core is not a single sweep launch, unexpected shapes): the caller then
        evaluates self(x) and the criterion separately."""
        core, out = self.__core, self.__output_layer
        if not (self.fuse_output and _is_abs_layer(out) and hasattr(core, "_lower")):
            return None
        if not keep_caches:
            self._invalidate_caches()
        x, prog = self._input_and_program(x, ext_param)
        if prog is None or not _abs_layer_verified(out, prog, x):
            return None
        return prog.run_loss(x, target, kind)

    # -- accessors -------------------------------------------------------------------------------
    def get_inputLayer(self):
        return self.__input_layer

    def set_inputLayer(self, input_layer=None):
        self.__input_layer = input_layer

    def get_outputLayer(self):
        return self.__output_layer

    def set_outputLayer(self, output_layer=None):
        self.__output_layer = output_layer

    def get_core(self):
        return self.__core

    def set_core(self, core):
        self.__core = core

    def probe(self, z, include_shell_io: bool = False):
        H = self.__core.probe(z)
        if include_shell_io:
            for layer, left in ((self.__input_layer, False), (self.__output_layer, True)):
                Hl = layer.probe(z) if hasattr(layer, "probe") else None
                if Hl is not None:
                    H = Hl if H is None else (Hl @ H if left else H @ Hl)
        return H

    # -- responses -------------------------------------------------------------------------------
    def _device(self):
        return self.alias_decay_db.device if torch.is_tensor(self.alias_decay_db) else None

    def _impulse(self, fs, identity):
        x = signal_gallery(batch_size=1, n_samples=self.nfft, n=self.input_channels, signal_type="impulse", fs=fs,
                           device=self._device(), dtype=self.dtype)
        if identity and self.input_channels > 1:
            x = x.diag_embed()
        return x

    def _rising_envelope(self, identity):
        a = _alias_of(self.__core)
        gamma = 10 ** (-abs(a) / self.nfft / 20)
        env = gamma ** torch.arange(0, -self.nfft, -1, device=self._device(), dtype=self.dtype)
        env = env.view(1, -1, 1)
        if identity and self.input_channels > 1:
            env = env.unsqueeze(-1)
        return env

    def _with_layers(self, input_layer, output_layer, x):
        saved = (self.__input_layer, self.__output_layer)
        self.__input_layer, self.__output_layer = input_layer, output_layer
        try:
            with torch.no_grad():
                return self.forward(x)
        finally:
            self.__input_layer, self.__output_layer = saved

    def get_time_response(self, fs: int = 48000, identity: bool = False):
        """Impulse response (1, nfft, N_out[, N_in]) with the anti-aliasing envelope undone
        (reference system.py:1012-1079)."""
        env = self._rising_envelope(identity)
        out = nn.Sequential(iFFT(self.nfft, dtype=self.dtype), Transform(lambda y: y * env))
        return self._with_layers(FFT(self.nfft, dtype=self.dtype), out, self._impulse(fs, identity))

    def get_freq_response(self, fs: int = 48000, identity: bool = False):
        """Frequency response on the unit circle (1, M, N_out[, N_in]): core response evaluated at
        radius gamma, brought back by iFFT -> rising envelope -> FFT (reference system.py:1081-1153)."""
        env = self._rising_envelope(identity)
        out = nn.Sequential(iFFT(self.nfft, dtype=self.dtype), Transform(lambda y: y * env),
                            FFT(self.nfft, dtype=self.dtype))
        return self._with_layers(FFT(self.nfft, dtype=self.dtype), out, self._impulse(fs, identity))
