"""CPU: the drop-in surface — class / attribute / state_dict compatibility with the reference, error
behaviour, Series key flattening, and the C ABI (exports + plan validation; no compute calls)."""
import ctypes
import os
import re
import warnings
from collections import OrderedDict

import numpy as np
import pytest
import torch

from flamo_b200 import _lib, sweep, workloads as W
from flamo_b200.processor import dsp, system

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ----------------------------------------------------------------------------------- C ABI
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "fsweep.h")).read()
    declared = set(re.findall(r"FSWEEP_API\s+[\w\s\*]+?\b(fsweep_\w+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert _lib.lib().fsweep_version() == 2


def _op(kind, n_out, n_in, K=0, flags=0, n_ff=0, n_fb=0):
    return (kind, n_out, n_in, K, flags, n_ff, n_fb, 0)


def test_plan_validation():
    P = lambda ops, nfft=1024: _lib.Plan([_lib.Op(*o) for o in ops], nfft, 30.0, _lib.C64)
    p = P([_op(_lib.OP_GAIN, 8, 1), _op(_lib.OP_RECURSION, 8, 8, n_ff=1, n_fb=1), _op(_lib.OP_PDELAY, 8, 8),
           _op(_lib.OP_GAIN, 8, 8), _op(_lib.OP_GAIN, 1, 8)])
    assert p.n_coeffs == 4
    assert p.coeff_numel(0, 513) == 8 and p.coeff_numel(2, 513) == 64
    assert P([_op(_lib.OP_SOS, 3, 2, K=5)]).coeff_numel(0, 513) == 5 * 2 * 3 * 16
    assert P([_op(_lib.OP_TABLE, 3, 2)]).coeff_numel(0, 513) == 513 * 6
    with pytest.raises(_lib.SweepError, match="inputs"):  # channel mismatch in a series
        P([_op(_lib.OP_GAIN, 4, 2), _op(_lib.OP_GAIN, 2, 3)])
    with pytest.raises(_lib.Unsupported, match="width"):
        P([_op(_lib.OP_GAIN, 65, 65)])
    assert P([_op(_lib.OP_GAIN, 64, 64)]).n_coeffs == 1  # two warps per bin
    with pytest.raises(_lib.Unsupported, match="FDN shape only"):
        _lib.Plan([_lib.Op(*_op(_lib.OP_GAIN, 64, 64))], 1024, 30.0, _lib.C128)
    with pytest.raises(_lib.Unsupported, match="more than one"):
        P([_op(_lib.OP_RECURSION, 2, 2, n_ff=1, n_fb=1), _op(_lib.OP_PGAIN, 2, 2), _op(_lib.OP_PGAIN, 2, 2),
           _op(_lib.OP_RECURSION, 2, 2, n_ff=1, n_fb=1), _op(_lib.OP_PGAIN, 2, 2), _op(_lib.OP_PGAIN, 2, 2)])
    with pytest.raises(_lib.SweepError, match="feedback"):
        P([_op(_lib.OP_RECURSION, 3, 2, n_ff=1, n_fb=1), _op(_lib.OP_GAIN, 3, 2), _op(_lib.OP_GAIN, 3, 3)])
    with pytest.raises(_lib.SweepError):
        P([_op(_lib.OP_PGAIN, 3, 2)])
    with pytest.raises(_lib.SweepError):
        P([_op(_lib.OP_SOS, 2, 2, K=0)])


def test_no_cpu_fallback():
    m = dsp.Gain(size=(2, 2), nfft=64)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.ones(1, 33, 2, dtype=torch.complex64))


# ------------------------------------------------------------------------- module interface
def test_state_dict_matches_reference_checkpoints():
    ck = np.load(os.path.join(ROOT, "tests", "golden", "reference_checkpoints.npz"))
    # FDN checkpoints of notebooks/output/ex_fdn (N = 6)
    core = W.build(W.fdn(6), dsp, system, 2048, 30.0, dtype=torch.float64)
    model = system.Shell(core, dsp.FFT(2048, dtype=torch.float64),
                         dsp.Transform(lambda x: torch.abs(x), dtype=torch.float64))
    sd = {k.split("|")[2]: torch.tensor(ck[k]) for k in ck.files if k.startswith("fdn|e7|")}
    assert set(sd) == set(model.state_dict().keys())
    model.load_state_dict(sd)
    assert torch.equal(model.get_core().feedback_loop.feedback.param, sd["_Shell__core.feedback_loop.feedback.param"])
    # biquad checkpoint of notebooks/output/ex_biquad: param (2, 3, 2, 1) = 2 bandpass sections, 2 outputs, 1 input
    sdb = {k.split("|")[2]: torch.tensor(ck[k]) for k in ck.files if k.startswith("biquad|e7|")}
    shape = tuple(sdb["_Shell__core.param"].shape)
    filt = dsp.Biquad(size=shape[2:], n_sections=shape[0], filter_type="bandpass" if shape[1] == 3 else "lowpass",
                      nfft=2048, dtype=sdb["_Shell__core.param"].dtype)
    mb = system.Shell(filt, dsp.FFT(2048))
    mb.load_state_dict(sdb)


def test_parameter_shapes_and_attributes():
    kw = dict(nfft=256, alias_decay_db=30.0)
    cases = [
        (dsp.Gain(size=(3, 2), **kw), (3, 2), 2, 3),
        (dsp.parallelGain(size=(4,), **kw), (4,), 4, 4),
        (dsp.Matrix(size=(4, 4), matrix_type="orthogonal", **kw), (4, 4), 4, 4),
        (dsp.Biquad(size=(2, 3), n_sections=4, filter_type="bandpass", **kw), (4, 3, 2, 3), 3, 2),
        (dsp.parallelBiquad(size=(5,), n_sections=2, **kw), (2, 2, 5), 5, 5),
        (dsp.SVF(size=(2, 3), n_sections=3, **kw), (5, 3, 2, 3), 3, 2),
        (dsp.parallelSVF(size=(6,), n_sections=1, **kw), (5, 1, 6), 6, 6),
        (dsp.GEQ(size=(2, 3), octave_interval=1, **kw), (12, 2, 3), 3, 2),
        (dsp.GEQ(size=(1, 1), octave_interval=3, **kw), (30, 1, 1), 1, 1),
        (dsp.parallelGEQ(size=(4,), octave_interval=1, **kw), (12, 4), 4, 4),
        (dsp.Delay(size=(2, 3), **kw), (2, 3), 3, 2),
        (dsp.parallelDelay(size=(7,), **kw), (7,), 7, 7),
        (dsp.Filter(size=(9, 2, 3), **kw), (9, 2, 3), 3, 2),
        (dsp.parallelFilter(size=(9, 4), **kw), (9, 4), 4, 4),
    ]
    for m, shape, n_in, n_out in cases:
        assert tuple(m.param.shape) == shape, type(m)
        assert (m.input_channels, m.output_channels) == (n_in, n_out), type(m)
        assert m.nfft == 256 and float(m.alias_decay_db) == 30.0
        assert abs(float(m.gamma) - 10 ** (-30 / 256 / 20)) < 1e-7
        assert list(m.state_dict().keys()) == ["param"]
        for attr in ("map", "freq_convolve", "new_value", "size", "dtype"):
            assert hasattr(m, attr)


def test_seeded_initialisation_matches_reference_distributions():
    torch.manual_seed(0)
    b = dsp.Biquad(size=(2, 2), n_sections=3, filter_type="highpass", nfft=64)
    assert 0 <= float(b.param[:, 0].min()) and float(b.param[:, 0].max()) <= 0.5
    assert -1 <= float(b.param[:, 1].min()) and float(b.param[:, 1].max()) <= 1
    g = dsp.GEQ(size=(2, 2), nfft=64)
    assert 10 ** (-0.3) <= float(g.param.min()) and float(g.param.max()) <= 10 ** 0.3
    d = dsp.parallelDelay(size=(4,), max_len=1000, isint=True, nfft=64)
    samples = d.s2sample(d.param)
    assert torch.allclose(samples, samples.round(), atol=1e-3)
    d.assign_value(d.sample2s(torch.tensor([3.0, 5.0, 7.0, 11.0])))
    assert d.new_value == 1


def test_error_behaviour():
    g = dsp.Gain(size=(3, 2), nfft=64)
    with pytest.raises(ValueError):
        g(torch.ones(1, 33, 5, dtype=torch.complex64))
    f = dsp.Biquad(size=(1, 1), nfft=64)
    with pytest.raises(ValueError):
        f(torch.ones(1, 10, 1, dtype=torch.complex64))  # wrong number of bins
    with pytest.raises(AssertionError):
        dsp.Gain(size=(3,), nfft=64)
    with pytest.raises(AssertionError):
        dsp.parallelGain(size=(3, 2), nfft=64)
    with pytest.raises(AssertionError):
        dsp.Delay(size=(3,), nfft=64)
    with pytest.raises(AssertionError):
        dsp.Biquad(size=(2, 2), filter_type="allpass", nfft=64)
    with pytest.raises(AssertionError):
        g.assign_value(torch.zeros(2, 2))
    with pytest.raises(AssertionError):  # channel chaining
        system.Series(dsp.Gain(size=(3, 2), nfft=64), dsp.Gain(size=(2, 4), nfft=64))
    with pytest.raises(ValueError):  # nfft mismatch
        system.Series(dsp.Gain(size=(3, 2), nfft=64), dsp.Gain(size=(2, 3), nfft=128))
    with pytest.raises(AssertionError):  # recursion io
        system.Recursion(fF=dsp.parallelDelay(size=(3,), nfft=64), fB=dsp.Gain(size=(2, 3), nfft=64))


def test_series_key_flattening():
    a, b, c, d = (dsp.parallelGain(size=(2,), nfft=64) for _ in range(4))
    s = system.Series(OrderedDict({"first": a, "inner": system.Series(b, c)}), d)
    assert list(s._modules.keys()) == ["first", "1", "2", "3"]
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        s2 = system.Series(OrderedDict({"5": a, "x": b}))
    assert list(s2._modules.keys()) == ["0", "x"] and any("overwritten" in str(x.message) for x in w)
    with pytest.raises(ValueError):
        system.Series(OrderedDict({"k": a}), OrderedDict({"k": b}))
    s.append(dsp.parallelGain(size=(2,), nfft=64))
    s.prepend(OrderedDict({"head": dsp.parallelGain(size=(2,), nfft=64)}))
    assert list(s._modules.keys())[0] == "head" and len(s) == 6
    assert (s.input_channels, s.output_channels) == (2, 2)


def test_lowering_produces_one_fused_program():
    core = W.build(W.active_acoustics(), dsp, system, 512, 30.0)
    prog = sweep.Program(512, 30.0, torch.complex64, "cpu")
    core._lower(prog, None)
    segs = list(prog._segments())
    assert len(segs) == 1
    ops, coefs, n_out = prog.flatten_segment(segs[0][1], torch.complex64)
    kinds = [o[0] for o in ops]
    assert kinds == [_lib.OP_GAIN, _lib.OP_RECURSION, _lib.OP_SOS, _lib.OP_PDELAY, _lib.OP_PGAIN, _lib.OP_DELAY,
                     _lib.OP_PGAIN, _lib.OP_GAIN]
    assert ops[1][5:7] == (3, 2) and n_out == 1
    assert coefs[1].shape == (2, 4, 13, 2, 8) and coefs[1].dtype == torch.float32
    assert coefs[2].dtype == torch.float64 and coefs[4].dtype == torch.float64  # delays stay float64
    _lib.Plan([_lib.Op(*o) for o in ops], 512, 30.0, _lib.C64)  # accepted by the C planner


def test_series_with_two_recursions_splits_into_two_launches():
    rec = lambda: system.Recursion(fF=dsp.parallelDelay(size=(2,), nfft=64), fB=dsp.Gain(size=(2, 2), nfft=64))
    s = system.Series(rec(), rec())
    prog = sweep.Program(64, 0.0, torch.complex64, "cpu")
    s._lower(prog, None)
    assert [t for t, _ in prog._segments()] == ["sweep", "sweep"]


def test_kernel_family_selection():
    """Host-side dispatch (no GPU needed: plans are host objects): which kernel family a plan launches."""
    import torch

    from flamo_b200 import sweep, workloads as W
    from flamo_b200.processor import dsp, system

    def plan_of(desc, nfft=4096, dtype=torch.float32):
        torch.manual_seed(0)
        core = W.build(desc, dsp, system, nfft, 30.0, dtype=dtype, device="cpu")
        cdt = torch.complex64 if dtype == torch.float32 else torch.complex128
        with torch.enable_grad():
            prog = sweep.Program(nfft, 30.0, cdt, "cpu")
            core._lower(prog, None)
            (tag, payload), = list(prog._segments())
            ops, coefs, n_out = prog.flatten_segment(payload)
        return _lib.Plan([_lib.Op(*o) for o in ops], nfft, 30.0, _lib.C64 if dtype == torch.float32 else _lib.C128)

    p = plan_of(W.fdn(8))
    assert "tpc" in p.kernel_family(48001, True) and "tpc" in p.kernel_family(48001, False)  # one thread per bin
    assert "loop" in p.kernel_family(2049, True)                                             # few bins: 8 lanes per bin
    assert "loop" in plan_of(W.fdn(8), dtype=torch.float64).kernel_family(48001, True)       # float64: no tpc
    assert "loop" in plan_of(W.fdn(16, delays=list(range(601, 601 + 16 * 37, 37)))).kernel_family(48001, True)
    assert "cta" in plan_of(W.fdn(64)).kernel_family(192001, True)                           # CTA per bin
    assert "cta" in plan_of(W.fdn(40, delays=list(range(601, 601 + 40 * 37, 37)))).kernel_family(1000, False)
    fir = ("Series", [("Filter", dict(size=(20, 5, 3), requires_grad=True)),
                      ("parallelGain", dict(size=(5,), requires_grad=True)),
                      ("Filter", dict(size=(12, 2, 5), requires_grad=True))])
    assert "stream" in plan_of(fir).kernel_family(2049, True)                                # TABLE streaming
    assert p.kernel_family(48001, True) != plan_of(W.geq(4, 4, 3)).kernel_family(48001, True)
    assert plan_of(W.geq(4, 4, 3)).kernel_family(48001, True) == "fsweep_bwd_kernel"         # generic interpreter
    assert plan_of(W.active_acoustics()).kernel_family(48001, True) == "fsweep_bwd_kernel"


def test_header_is_plain_c_and_a_c_program_links_against_the_library(tmp_path):
    """include/fsweep.h is the reference-facing seam: it must compile as strict C99 (no C++, no torch types) and a C
    program must link and call into libfsweep.so with nothing but that header (what a cgo / ctypes / FFI binding does)."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "t.c"
    src.write_text('#include "fsweep.h"\n'
                   "int main(void) {\n"
                   "  fsweep_op_t op = {FSWEEP_OP_GAIN, 2, 2, 0, FSWEEP_F_GRAD, 0, 0, 0};\n"
                   "  fsweep_plan_t* plan = 0;\n"
                   "  if (fsweep_version() != FSWEEP_VERSION) return 1;\n"
                   "  if (fsweep_plan_create(&op, 1, 64, 0.0, FSWEEP_C64, &plan) != FSWEEP_OK) return 2;\n"
                   "  if (fsweep_plan_num_coeffs(plan) != 1 || fsweep_plan_coeff_numel(plan, 0, 33) != 4) return 3;\n"
                   "  if (fsweep_forward(plan, 0, 0, 0, 0, 0, 1, 1, 0, 33, FSWEEP_EPI_NONE, 0) != FSWEEP_E_BADARG) return 4;\n"
                   "  if (!fsweep_last_error()[0]) return 5;\n"
                   "  return fsweep_plan_destroy(plan) == FSWEEP_OK ? 0 : 6;\n"
                   "}\n")
    exe = tmp_path / "t"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    str(src), "-o", str(exe), "-L", libdir, "-l:" + os.path.basename(_lib.LIB_PATH),
                    "-Wl,-rpath," + libdir], check=True)
    assert subprocess.run([str(exe)]).returncode == 0


def test_integration_md_stub_matches_the_real_binding():
    """The ctypes stub INTEGRATION.md shows a flamo maintainer is executable and declares the same argument lists as
    the binding this package ships (flamo_b200/_lib.py)."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = next(b.split("```")[0] for b in text.split("```python")[1:] if "flamo/_fsweep.py" in b)
    assert "flamo/_fsweep.py" in block
    stub = block.split("class Sweep")[0].replace('C.CDLL("libfsweep.so")', f'C.CDLL({_lib.LIB_PATH!r})')
    ns = {}
    exec(compile(stub, "INTEGRATION.md", "exec"), ns)
    real = _lib.lib()
    for name in ("fsweep_plan_create", "fsweep_forward", "fsweep_backward"):
        shown, shipped = getattr(ns["L"], name).argtypes, getattr(real, name).argtypes
        assert len(shown) == len(shipped), name
        for a, b in zip(shown, shipped):
            assert ctypes.sizeof(a) == ctypes.sizeof(b), (name, a, b)
    assert ctypes.sizeof(ns["Op"]) == ctypes.sizeof(_lib.Op) == 32
