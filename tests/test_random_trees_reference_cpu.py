"""CPU, build container only (skipped where /root/reference is absent — e.g. on the GPU box): randomly composed module
trees evaluated by the UNMODIFIED reference (imported from /root/reference with the stubs of
tests/golden/make_golden.py) and by this package's host side through the ABI emulator, same description, same seed,
same raw parameters.  Checks that constructors draw the same parameters from the same RNG stream (seeded
initialisation parity), then forward response and gradients in float64.  The reference keeps SVF / GEQ internals in
float32 (SURVEY.md §8c caveat), so trees containing them are compared at that noise level."""
import os
import sys
import types

import pytest
import torch
from hypothesis import HealthCheck, assume, given, settings, strategies as st

import cases as C
from flamo_b200 import workloads as W
from flamo_b200.processor import dsp, system
from helpers import rel_err
from test_random_trees_cpu import NFFT, tree

REF = "/root/reference"
pytestmark = [pytest.mark.usefixtures("emulated_backend"),
              pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "flamo")), reason="reference checkout not present")]


def reference_modules():
    for n in ["soundfile", "nnAudio", "nnAudio.features", "pyfar", "matplotlib", "matplotlib.pyplot"]:
        if n not in sys.modules:
            try:
                __import__(n)
            except Exception:
                sys.modules[n] = types.ModuleType(n)
    if not hasattr(sys.modules["nnAudio"], "features"):
        sys.modules["nnAudio"].features = sys.modules["nnAudio.features"]
    if REF not in sys.path:
        sys.path.append(REF)
    from flamo.processor import dsp as rdsp, system as rsystem

    return rdsp, rsystem


def kinds_of(desc, out):
    if desc[0] == "Series":
        for d in desc[1]:
            kinds_of(d, out)
    elif desc[0] in ("Recursion", "Parallel"):
        kinds_of(desc[1], out)
        kinds_of(desc[2], out)
    else:
        out.add(desc[0])
    return out


def kink(Y):
    """|Y| = 0 somewhere (a section's zero on a bin, paths cancelling exactly at alias_decay_db = 0, an all-zero
    channel): the gradient of |.| there is the direction of Y's rounding noise — any subgradient is right and no two
    evaluation routes agree, so gradient comparisons skip such draws (responses are still compared)."""
    a = Y.detach().abs()
    return float(a.min()) <= 1e-9 * float(a.max())


def assert_matches_reference(desc, X, model, Y, Yr, alias, tol_full=1e-8, nfft=NFFT):
    """Response parity with the reference.  Trees without SVF / GEQ: directly, 1e-8.  Trees with them: the reference
    keeps those modules' tap buffers (and eq.geq's frequency terms) in float32 whatever the module dtype (SURVEY.md §8c
    caveat), which costs it up to 1.4e-1 — so parity is shown as an exact chain instead of a loose tolerance:
    reference == oracle with REF_FP32_INTERNALS (1e-10, the oracle reproduces that dtype flow bit for bit) and
    oracle at full precision == this package (1e-8)."""
    from oracle import flamo_oracle as O

    kinds = kinds_of(desc, set())
    if not kinds & {"SVF", "parallelSVF", "GEQ", "parallelGEQ"}:
        assert rel_err(Y.detach().numpy(), Yr.detach().numpy()) <= 1e-8, desc
        return
    ps = [p.detach().clone() for p in model.parameters()]
    node = O.from_desc(desc)
    with torch.no_grad():
        Yo = O.forward(node, X, ps, nfft, alias)
        O.REF_FP32_INTERNALS = True
        try:
            Yo32 = O.forward(node, X, ps, nfft, alias)
        finally:
            O.REF_FP32_INTERNALS = False
    assert rel_err(Yo32.numpy(), Yr.detach().numpy()) <= 1e-10, desc
    assert rel_err(Y.detach().numpy(), Yo.numpy()) <= tol_full, desc


@settings(max_examples=150, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(tree(extra=True))  # + SOSFilter / parallelSOSFilter, Matrix "hadamard" / "rotation" (no oracle restatement)
def test_random_tree_matches_the_reference_itself(t):
    rdsp, rsystem = reference_modules()
    desc, n_in, B, cols, seed, alias = t
    X = C.make_input(B, NFFT // 2 + 1, n_in, cols)
    torch.manual_seed(seed)
    try:  # trees the reference itself cannot build or run (e.g. parallelBiquad "bandpass": IndexError in its
        # init_param, dsp.py:1572) say nothing about parity
        ref = W.build(desc, rdsp, rsystem, NFFT, alias, dtype=torch.float64)
        Yr = ref(X)
    except Exception:
        assume(False)
    torch.manual_seed(seed)
    model = W.build(desc, dsp, system, NFFT, alias, dtype=torch.float64, device="cpu")
    rp, mp = list(ref.parameters()), list(model.parameters())
    assert [tuple(p.shape) for p in rp] == [tuple(p.shape) for p in mp]
    assert [p.requires_grad for p in rp] == [p.requires_grad for p in mp]
    for a, b in zip(rp, mp):  # same constructor draws from the same RNG stream
        assert torch.equal(a.detach(), b.detach()), desc
    assert list(ref.state_dict().keys()) == list(model.state_dict().keys())
    Y = model(X)
    assert Yr.shape == Y.shape
    fp32_internals = bool(kinds_of(desc, set()) & {"SVF", "parallelSVF", "GEQ", "parallelGEQ"})
    assert_matches_reference(desc, X, model, Y, Yr, alias)
    gr = [p for p in rp if p.requires_grad]
    if gr and not fp32_internals and Yr.requires_grad and not kink(Yr):  # (a Hadamard matrix "requires grad" but nothing depends on it)
        assert Y.requires_grad, desc
        C.golden_loss(Y).backward()
        go = torch.autograd.grad(C.golden_loss(Yr), gr, allow_unused=True)
        scale = max([float(g.abs().max()) for g in go if g is not None] + [1e-6])  # (floor: where every true gradient is zero only rounding noise is left)
        k = 0
        for a, b in zip(rp, mp):
            if not a.requires_grad:
                continue
            g = go[k]
            k += 1
            if g is None:
                continue
            assert b.grad is not None, desc
            assert float((b.grad - g).abs().max()) <= 1e-7 * scale, desc


def _make_ext(ref_mod, rsystem, gen, p_use=0.6):
    """Nested ext_param structure for a REFERENCE module tree (hyper-conditioning, reference dsp.py:415-432,
    system.py:279-301, 397-411): tensors for leaves, {key: ...} for a Series, {"feedforward" / "feedback": ...} for a
    Recursion, {"branchA" / "branchB": ...} for a Parallel.  New values = stored parameter + a small perturbation."""
    if isinstance(ref_mod, rsystem.Series):
        d = {}
        for key, child in ref_mod._modules.items():
            e = _make_ext(child, rsystem, gen, p_use)
            if e is not None:
                d[key] = e
        return d or None
    if isinstance(ref_mod, rsystem.Recursion):
        d = {}
        for name in ("feedforward", "feedback"):
            e = _make_ext(getattr(ref_mod, name), rsystem, gen, p_use)
            if e is not None:
                d[name] = e
        return d or None
    if isinstance(ref_mod, rsystem.Parallel):
        d = {}
        for name in ("branchA", "branchB"):
            e = _make_ext(getattr(ref_mod, name), rsystem, gen, p_use)
            if e is not None:
                d[name] = e
        return d or None
    if torch.rand((), generator=gen).item() > p_use:
        return None
    p = ref_mod.param.detach()
    return (p + 0.01 * torch.randn(p.shape, generator=gen, dtype=p.dtype)).requires_grad_(True)


def _ext_tensors(e, out):
    if torch.is_tensor(e):
        out.append(e)
    elif e is not None:
        for v in e.values():
            _ext_tensors(v, out)
    return out


def _clone_ext(e):
    if torch.is_tensor(e):
        return e.detach().clone().requires_grad_(True)
    return None if e is None else {k: _clone_ext(v) for k, v in e.items()}


@settings(max_examples=80, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(tree())
def test_random_tree_ext_param_routing_matches_the_reference(t):
    rdsp, rsystem = reference_modules()
    desc, n_in, B, cols, seed, alias = t
    kinds = kinds_of(desc, set())
    assume(not kinds & {"SVF", "parallelSVF", "GEQ", "parallelGEQ"})  # float32 internals: covered above
    X = C.make_input(B, NFFT // 2 + 1, n_in, cols)
    torch.manual_seed(seed)
    try:
        ref = W.build(desc, rdsp, rsystem, NFFT, alias, dtype=torch.float64)
        ext_r = _make_ext(ref, rsystem, torch.Generator().manual_seed(seed))
        assume(ext_r is not None)
        Yr = ref(X, ext_r)
    except Exception as e:
        if isinstance(e, (AssertionError,)) and "Unsatisfied" in type(e).__name__:
            raise
        assume(False)
    torch.manual_seed(seed)
    model = W.build(desc, dsp, system, NFFT, alias, dtype=torch.float64, device="cpu")
    ext_m = _clone_ext(ext_r)
    Y = model(X, ext_m)
    assert rel_err(Y.detach().numpy(), Yr.detach().numpy()) <= 1e-8, desc
    for a, b in zip(ref.parameters(), model.parameters()):  # the external values were logged into the modules
        assert torch.equal(a.detach(), b.detach()), desc
    if kink(Yr):
        return
    tr, tm = _ext_tensors(ext_r, []), _ext_tensors(ext_m, [])
    try:  # (the reference's a0-normalising SOSFilter map writes in place and cannot be differentiated)
        gr = torch.autograd.grad(C.golden_loss(Yr), tr, allow_unused=True)
    except RuntimeError:
        assume(False)
    gm = torch.autograd.grad(C.golden_loss(Y), tm, allow_unused=True) if Y.requires_grad else [None] * len(tm)
    assume(all(g is None or bool(torch.isfinite(g).all()) for g in gr))  # the reference's own gradient is NaN for some
    # draws (clamped Biquad maps): nothing to compare with
    scale = max([float(g.abs().max()) for g in gr if g is not None] + [1e-6])  # (floor: where every true gradient is zero only rounding noise is left)
    for a, b in zip(gr, gm):
        if a is None or float(a.abs().max()) == 0.0:
            assert b is None or float(b.abs().max()) <= 1e-12 * scale, desc
        else:
            assert b is not None and float((a - b).abs().max()) <= 1e-7 * scale, desc


@settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(tree())
def test_random_tree_train_steps_match_the_reference_trainer(t):
    """Two Trainer.train_step calls (reference optimize/trainer.py:155-192: zero_grad, forward, weighted criteria,
    backward, Adam) of the reference Trainer on the reference model vs this package's Trainer on its model."""
    import numpy as np

    rdsp, rsystem = reference_modules()
    from flamo.optimize.loss import mse_loss as r_mse_loss
    from flamo.optimize.trainer import Trainer as RTrainer

    from flamo_b200.optimize.loss import mse_loss
    from flamo_b200.optimize.trainer import Trainer

    desc, n_in, B, cols, seed, alias = t
    assume(not kinds_of(desc, set()) & {"SVF", "parallelSVF", "GEQ", "parallelGEQ"})
    M = NFFT // 2 + 1
    x = torch.zeros(B, NFFT, n_in, dtype=torch.float64)
    x[:, 0] = 1
    x[:, 7] = 0.5
    tgt = torch.full((B, M, 1), 0.6, dtype=torch.float64)

    def run(dsp_, system_, trainer_cls, crit):
        torch.manual_seed(seed)
        core = W.build(desc, dsp_, system_, NFFT, alias, dtype=torch.float64)
        model = system_.Shell(core, dsp_.FFT(NFFT, dtype=torch.float64),
                              dsp_.Transform(lambda v: torch.abs(v), dtype=torch.float64))
        if not any(p.requires_grad for p in model.parameters()):
            return None
        with torch.no_grad():
            est = model(x)
        if float(est.min()) <= 1e-9 * float(est.max()):
            return "kink"  # |Y| = 0 at some bin (e.g. z^-5 + z^-3 at omega = pi / 2): the gradient of |.| there is
            # whatever direction the rounding noise of Y has — any subgradient is right, none is comparable
        tr = trainer_cls(model, max_epochs=1, lr=1e-2, log=False, device="cpu")
        tr.register_criterion(crit, 1)
        tr.train_loss_log = {crit.__class__.__name__: []}
        losses = [tr.train_step((x, tgt)) for _ in range(2)]
        return losses, [p.detach().clone() for p in model.parameters()]

    try:
        r = run(rdsp, rsystem, RTrainer, r_mse_loss(nfft=NFFT, device="cpu"))
    except Exception:
        assume(False)
    assume(r is not None and r != "kink")
    assume(all(np.isfinite(r[0])) and all(bool(torch.isfinite(p).all()) for p in r[1]))
    m = run(dsp, system, Trainer, mse_loss(nfft=NFFT))
    assume(m != "kink")
    assert np.allclose(m[0], r[0], rtol=1e-7, atol=1e-13), desc
    for a, b in zip(m[1], r[1]):
        # atol: where the true gradient is zero (a pure delay in front of |.|) autograd returns ~1e-15 of rounding
        # noise and Adam turns it into lr * g / (|g| + eps) ~ 1e-9 ... 1e-6 of parameter movement, different on either
        # side (more where the gradient is a small remainder of large terms); a real update is lr = 1e-2 per step, so
        # parameters are pinned to 1e-2 of an update and the second step's loss is the sharp check
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-4), desc


def _has(desc, name):
    if desc[0] == name:
        return True
    if desc[0] == "Series":
        return any(_has(d, name) for d in desc[1])
    if desc[0] in ("Recursion", "Parallel"):
        return _has(desc[1], name) or _has(desc[2], name)
    return False


@settings(max_examples=80, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(tree(), st.sampled_from([-3.0, 0.0, 0.05, 4.0, 40.0]))
def test_random_tree_far_from_the_initial_parameters(t, scale):
    """Raw parameters far from where the constructors draw them (x -3, 0, 0.05, 4, 40: clamps of the Biquad map,
    saturated sigmoids / softplus of the SVF map, negative and zero gains, long delays): response against the
    reference itself.  Loop-free trees only (a scaled loop gain says more about conditioning than about parity)."""
    rdsp, rsystem = reference_modules()
    desc, n_in, B, cols, seed, alias = t
    kinds = kinds_of(desc, set())
    assume(not _has(desc, "Recursion"))
    X = C.make_input(B, NFFT // 2 + 1, n_in, cols)
    torch.manual_seed(seed)
    try:
        ref = W.build(desc, rdsp, rsystem, NFFT, alias, dtype=torch.float64)
        with torch.no_grad():
            for p in ref.parameters():
                p.mul_(scale)
            Yr = ref(X)
    except Exception:
        assume(False)
    assume(bool(torch.isfinite(torch.view_as_real(Yr)).all()) and float(Yr.abs().max()) > 1e-9)  # not an all-stop filter
    torch.manual_seed(seed)
    model = W.build(desc, dsp, system, NFFT, alias, dtype=torch.float64, device="cpu")
    with torch.no_grad():
        for p in model.parameters():
            p.mul_(scale)
        Y = model(X)
    # saturated SVF maps put zeros and poles within 1e-6 of z = 1 (f = tan(pi/2 sigmoid(-15)) ~ 6e-7): A(1) = 4 f^2 is
    # then a 1e-12 remainder of O(1) taps and ANY float64 evaluation from the taps (the reference's rfft, the
    # oracle's, this package's Taylor blocks) is only good to ~1e-4 there: the full-precision step is held to 1e-3
    assert_matches_reference(desc, X, model, Y, Yr, alias, tol_full=1e-3 if scale in (-3.0, 4.0, 40.0) else 1e-8)


@settings(max_examples=60, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(tree(), st.booleans())
def test_random_tree_shell_responses_match_the_reference(t, identity):
    """Shell.get_time_response / get_freq_response (reference system.py:1012-1153), input-to-output and input-free
    (`identity=True`: a diagonal of impulses, one column per input channel) against the reference's own."""
    rdsp, rsystem = reference_modules()
    desc, n_in, B, cols, seed, alias = t
    assume(not kinds_of(desc, set()) & {"SVF", "parallelSVF", "GEQ", "parallelGEQ"})

    def shell(dsp_, system_):
        torch.manual_seed(seed)
        core = W.build(desc, dsp_, system_, NFFT, alias, dtype=torch.float64)
        return system_.Shell(core, dsp_.FFT(NFFT, dtype=torch.float64),
                             dsp_.Transform(lambda v: torch.abs(v), dtype=torch.float64))

    try:
        ref = shell(rdsp, rsystem)
        hr = ref.get_time_response(identity=identity)
        Hr = ref.get_freq_response(identity=identity)
    except Exception:
        assume(False)
    assume(bool(torch.isfinite(hr).all()) and float(hr.abs().max()) > 1e-9)
    model = shell(dsp, system)
    h, H = model.get_time_response(identity=identity), model.get_freq_response(identity=identity)
    assert h.shape == hr.shape and H.shape == Hr.shape
    assert float((h - hr).abs().max()) <= 1e-9 * float(hr.abs().max()), desc
    assert float((H - Hr).abs().max()) <= 1e-9 * float(Hr.abs().max()), desc


def test_full_training_loop_matches_the_reference(tmp_path):
    """Trainer.train (reference optimize/trainer.py:106-153: epochs over a DataLoader, StepLR, validation, checkpoint
    per epoch, early stopping) with the colourless-FDN criteria of examples/e8_colorless_fdn.py on a 4-line FDN: same
    train / validation loss curves, same per-criterion logs, same checkpoint files as the reference Trainer."""
    import numpy as np

    rdsp, rsystem = reference_modules()
    from flamo.optimize import dataset as rdataset, loss as rloss, trainer as rtrainer

    from flamo_b200.optimize import dataset, loss, trainer

    nfft = 512
    desc = W.fdn(4, delays=[11, 17, 23, 31])

    def run(dsp_, system_, ds_, loss_, tr_, out):
        torch.manual_seed(11)
        core = W.build(desc, dsp_, system_, nfft, 30.0, dtype=torch.float64)
        model = system_.Shell(core, dsp_.FFT(nfft, dtype=torch.float64),
                              dsp_.Transform(lambda v: torch.abs(v), dtype=torch.float64))
        data = ds_.DatasetColorless(input_shape=(1, nfft // 2 + 1, 1), target_shape=(1, nfft // 2 + 1, 1), expand=10,
                                    device="cpu", dtype=torch.float64)
        train_loader, valid_loader = ds_.load_dataset(data, batch_size=2, split=0.8, shuffle=False, device="cpu")
        os.makedirs(out)
        tr = tr_.Trainer(model, max_epochs=3, lr=1e-2, step_size=2, train_dir=str(out), device="cpu")
        tr.register_criterion(loss_.mse_loss(nfft=nfft, device="cpu"), 1)
        tr.register_criterion(loss_.sparsity_loss(), 0.2, requires_model=True)
        tr.train(train_loader, valid_loader)
        return tr, model

    rt, rm = run(rdsp, rsystem, rdataset, rloss, rtrainer, tmp_path / "ref")
    mt, mm = run(dsp, system, dataset, loss, trainer, tmp_path / "mine")
    assert np.allclose(mt.train_loss, rt.train_loss, rtol=1e-9) and np.allclose(mt.valid_loss, rt.valid_loss, rtol=1e-9)
    assert mt.train_loss_log.keys() == rt.train_loss_log.keys()
    for k in rt.train_loss_log:
        assert np.allclose(mt.train_loss_log[k], rt.train_loss_log[k], rtol=1e-9), k
        assert np.allclose(mt.valid_loss_log[k], rt.valid_loss_log[k], rtol=1e-9), k
    for a, b in zip(mm.parameters(), rm.parameters()):
        assert torch.allclose(a, b, rtol=1e-8, atol=1e-11)
    files = sorted(os.listdir(tmp_path / "ref" / "checkpoints"))
    assert files == sorted(os.listdir(tmp_path / "mine" / "checkpoints")) and len(files) == 3
    sd_r = torch.load(tmp_path / "ref" / "checkpoints" / files[-1], weights_only=True)
    sd_m = torch.load(tmp_path / "mine" / "checkpoints" / files[-1], weights_only=True)
    assert sd_r.keys() == sd_m.keys()
    for k in sd_r:
        assert torch.allclose(sd_r[k], sd_m[k], rtol=1e-8, atol=1e-11), k


def _outcome(fn):
    try:
        fn()
        return "ok"
    except Exception as e:  # noqa: BLE001 - the exception TYPE is the thing compared
        return type(e).__name__


@settings(max_examples=120, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(st.integers(1, 4), st.integers(1, 4), st.integers(1, 4), st.integers(1, 4), st.integers(1, 4),
       st.sampled_from(["series", "recursion", "input", "param_rank", "assign", "nfft"]))
def test_error_behaviour_matches_the_reference(a, b, c, d, n_x, what):
    """error types are part of the interface — mismatched channel chains (AssertionError), wrong input
    widths (ValueError), wrong parameter ranks (AssertionError), wrong assign_value shapes, mixed nfft (ValueError):
    whatever the reference does for a construction (including accepting it), this package does too."""
    rdsp, rsystem = reference_modules()
    kw = dict(nfft=64, dtype=torch.float64)

    def attempt(dsp_, system_):
        if what == "series":  # Gain (a <- b) then Gain (c <- d): valid iff a == d
            return _outcome(lambda: system_.Series(dsp_.Gain(size=(a, b), **kw), dsp_.Gain(size=(c, d), **kw)))
        if what == "recursion":  # fF: b -> a, fB: d -> c: valid iff a == d and c == b
            return _outcome(lambda: system_.Recursion(fF=dsp_.Gain(size=(a, b), **kw), fB=dsp_.Gain(size=(c, d), **kw)))
        if what == "input":
            m = dsp_.Delay(size=(a, b), max_len=20, **kw) if c % 2 else dsp_.Gain(size=(a, b), **kw)
            x = C.make_input(1, 33, n_x, None)
            return _outcome(lambda: m(x))
        if what == "param_rank":
            cls = [dsp_.Gain, dsp_.parallelGain, dsp_.Delay, dsp_.parallelDelay][c % 4]
            size = (a, b) if d % 2 else (a,)
            return _outcome(lambda: cls(size=size, **kw))
        if what == "assign":
            m = dsp_.Gain(size=(a, b), **kw)
            return _outcome(lambda: m.assign_value(torch.zeros(c, d, dtype=torch.float64)))
        return _outcome(lambda: system_.Series(dsp_.Gain(size=(a, b), nfft=64), dsp_.Gain(size=(c, a), nfft=64 * d)))

    assert attempt(dsp, system) == attempt(rdsp, rsystem), (what, a, b, c, d, n_x)


CLASSES = ["Gain", "parallelGain", "Matrix", "HouseholderMatrix", "Filter", "parallelFilter", "Biquad", "parallelBiquad",
           "SVF", "parallelSVF", "GEQ", "parallelGEQ", "Delay", "parallelDelay", "GainDelay", "parallelGainDelay",
           "SOSFilter", "parallelSOSFilter"]


@settings(max_examples=200, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(st.sampled_from(CLASSES), st.lists(st.integers(1, 4), min_size=1, max_size=3))
def test_constructor_size_checks_match_the_reference(cls, size):
    """Every module class with every rank of `size` (1 to 3 entries): accepted or rejected exactly like the
    reference, with the same exception type; when accepted, the same parameter shape and channel counts."""
    rdsp, rsystem = reference_modules()
    kw = dict(nfft=64, dtype=torch.float64)

    def build(dsp_):
        torch.manual_seed(0)
        return getattr(dsp_, cls)(size=tuple(size), **kw)

    mine, ref = _outcome(lambda: build(dsp)), _outcome(lambda: build(rdsp))
    assert mine == ref, (cls, size)
    if ref == "ok":
        m, r = build(dsp), build(rdsp)
        assert tuple(m.param.shape) == tuple(r.param.shape), (cls, size)
        assert (m.input_channels, m.output_channels) == (r.input_channels, r.output_channels), (cls, size)


@pytest.mark.parametrize("name,kwargs", [
    ("Delay", dict(size=(2, 3), max_len=500, isint=False, unit=1, fs=44100)),
    ("Delay", dict(size=(2, 3), max_len=500, isint=True, unit=1000, fs=48000)),
    ("parallelDelay", dict(size=(3,), max_len=123, isint=True, unit=10, fs=8000)),
    ("parallelDelay", dict(size=(3,), max_len=50, isint=False, unit=100, fs=48000, requires_grad=True)),
    ("GainDelay", dict(size=(2, 2), max_len=77, isint=False, unit=1, fs=16000)),
    ("parallelGainDelay", dict(size=(3,), max_len=77, isint=True, unit=100, fs=48000)),
    ("GEQ", dict(size=(1, 2), octave_interval=3, fs=48000)),
    ("parallelGEQ", dict(size=(2,), octave_interval=3, fs=44100)),
    ("Biquad", dict(size=(2, 1), n_sections=3, filter_type="highpass", fs=22050)),
    ("parallelSVF", dict(size=(2,), n_sections=2, filter_type="peaking", fs=96000)),
    ("Matrix", dict(size=(4, 4), matrix_type="rotation", iter=2)),   # (the reference hands `iter` to RotationMatrix's
    ("Matrix", dict(size=(4, 4), matrix_type="rotation")),           #  min_angle slot, dsp.py:665: with the default
    ("Matrix", dict(size=(2, 2), matrix_type="rotation", iter=0)),   #  iter = 1 > pi/4 the angle is ALWAYS pi/4;
    ("Matrix", dict(size=(4, 4), matrix_type="rotation", iter=None)),  # iter = 0 / None let the parameter through)
    ("Matrix", dict(size=(4, 4), matrix_type="hadamard")),
    ("Filter", dict(size=(7, 2, 2))),
    ("Filter", dict(size=(5, 1, 3), requires_grad=True)),
    ("parallelFilter", dict(size=(9, 3))),
    ("Delay", dict(size=(2, 2), max_len=300, isint=False, requires_grad=True)),          # learnable: softplus map
    ("Delay", dict(size=(2, 2), max_len=300, isint=True, requires_grad=True)),
    ("parallelDelay", dict(size=(4,), max_len=2000, isint=True)),
    ("GainDelay", dict(size=(3, 1), max_len=40, isint=True, requires_grad=True)),
    ("parallelGainDelay", dict(size=(2,), max_len=40, isint=False, requires_grad=True)),
    ("Gain", dict(size=(2, 3), map=lambda x: torch.tanh(x) * 2, requires_grad=True)),
    ("parallelGain", dict(size=(3,), map=lambda x: x ** 2)),
    ("HouseholderMatrix", dict(size=(4, 4), requires_grad=True)),
    ("HouseholderMatrix", dict(size=(2, 2))),
    ("Matrix", dict(size=(3, 3), matrix_type="orthogonal", requires_grad=True)),
    ("Matrix", dict(size=(2, 5), matrix_type="random")),
    ("Biquad", dict(size=(1, 2), n_sections=2, filter_type="bandpass", fs=48000, requires_grad=True)),
    ("parallelBiquad", dict(size=(3,), n_sections=1, filter_type="highpass", fs=32000)),
    ("SVF", dict(size=(2, 2), n_sections=1, filter_type="notch", fs=48000)),
    ("SVF", dict(size=(1, 1), n_sections=3, filter_type="lowshelf", fs=44100, requires_grad=True)),
    ("parallelSVF", dict(size=(3,), n_sections=2, filter_type="highshelf", fs=48000)),
    ("GEQ", dict(size=(2, 1), octave_interval=1, fs=44100, requires_grad=True)),
    ("SOSFilter", dict(size=(2, 2), n_sections=2)),
    ("parallelSOSFilter", dict(size=(3,), n_sections=1)),
])
def test_constructor_keywords_match_the_reference(name, kwargs):
    """Less common constructor keywords (delay `unit` / `fs` / `max_len`, third-octave GEQ, sampling rates, matrix
    types): same parameter draw from the same seed and the same response as the reference."""
    rdsp, rsystem = reference_modules()
    nfft = 128

    def build(dsp_):
        torch.manual_seed(5)
        return getattr(dsp_, name)(nfft=nfft, alias_decay_db=20.0, dtype=torch.float64, **kwargs)

    try:
        r = build(rdsp)
        X = C.make_input(1, nfft // 2 + 1, r.input_channels, None)
        Yr = r(X)
    except Exception as e:
        pytest.skip(f"the reference itself fails here: {type(e).__name__}: {e}")
    m = build(dsp)
    assert torch.equal(m.param.detach(), r.param.detach())
    for attr in ("fs", "unit", "max_len", "n_sections", "n_gains"):
        if hasattr(r, attr):
            assert getattr(m, attr) == getattr(r, attr), attr
    assert_matches_reference((name, dict(kwargs)), X, m, m(X), Yr, 20.0, nfft=nfft)


def _io(desc):
    """(input channels, output channels) of a description."""
    if desc[0] == "Series":
        return _io(desc[1][0])[0], _io(desc[1][-1])[1]
    if desc[0] == "Recursion":
        return _io(desc[1])
    if desc[0] == "Parallel":
        a, b = _io(desc[1]), _io(desc[2])
        return a[0], (a[1] if (len(desc) < 4 or desc[3]) else a[1] + b[1])
    size = desc[1]["size"]
    return (size[-1], size[-1]) if desc[0].startswith("parallel") else (size[-1], size[-2])


def _rect_loop(desc):
    if desc[0] == "Series":
        return any(_rect_loop(d) for d in desc[1])
    if desc[0] == "Recursion":
        a, b = _io(desc[1])
        return a != b
    if desc[0] == "Parallel":
        return _rect_loop(desc[1]) or _rect_loop(desc[2])
    return False


@settings(max_examples=80, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(tree(), st.floats(0.05, 3.0), st.floats(0.8, 1.2))
def test_random_tree_probe_matches_the_reference(t, angle, radius):
    """probe(z): the transfer matrix at an arbitrary point of the z plane (reference examples/e10_probe.py; the
    per-module probe methods of dsp.py / system.py) on random trees."""
    rdsp, rsystem = reference_modules()
    desc, n_in, B, cols, seed, alias = t
    # kinds whose probe the reference defines (dsp.py:487, 563, 945, 1032, 3436, 3539): the section filters and
    # GainDelay inherit the FIR probe, which reads their parameter as taps, and Recursion.probe sizes its identity
    # with F.shape[-1] (system.py:530), which is only right for square loops — neither says anything about parity
    assume(kinds_of(desc, set()) <= {"Gain", "parallelGain", "Matrix", "HouseholderMatrix", "Filter", "parallelFilter",
                                     "Delay", "parallelDelay"})
    assume(not _rect_loop(desc) and not _has(desc, "Parallel"))
    z = torch.polar(torch.tensor(radius, dtype=torch.float64), torch.tensor(angle, dtype=torch.float64))
    torch.manual_seed(seed)
    try:
        ref = W.build(desc, rdsp, rsystem, NFFT, alias, dtype=torch.float64)
        with torch.no_grad():
            Hr = ref.probe(z)
    except Exception:
        assume(False)
    assume(Hr is not None and bool(torch.isfinite(torch.view_as_real(Hr)).all()))
    torch.manual_seed(seed)
    model = W.build(desc, dsp, system, NFFT, alias, dtype=torch.float64, device="cpu")
    with torch.no_grad():
        H = model.probe(z)
    assert H.shape == Hr.shape, desc
    assert float((H - Hr).abs().max()) <= 1e-9 * float(Hr.abs().max() + 1e-300), desc


@settings(max_examples=100, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(st.sampled_from(["lowpass", "highpass", "bandpass", "shelving_low", "shelving_high", "peak"]),
       st.lists(st.floats(20.0, 20000.0), min_size=1, max_size=4), st.floats(-24.0, 24.0), st.floats(0.3, 8.0),
       st.sampled_from([8000, 44100, 48000, 96000]), st.sampled_from([torch.float32, torch.float64]))
def test_filter_designers_match_the_reference(kind, fcs, gain, Q, fs, dtype):
    """functional.{lowpass, highpass, bandpass, shelving, peak}_filter (reference functional.py:376-675): same taps,
    same shapes, same dtypes for random cutoffs / gains / Q / sampling rates."""
    reference_modules()
    import flamo.functional as RF

    from flamo_b200 import functional as MF

    fc = torch.tensor(fcs, dtype=dtype).clamp(max=0.45 * fs)
    if kind.startswith("shelving") or kind == "peak":
        fc = fc[0]  # the reference's shelving / peak designers are scalar (eq.geq loops over bands and channel pairs in
        # Python, eq.py:57-111); this package's accept tensors — compared on what the reference can do
    g = torch.full_like(fc, gain)

    def call(F):
        if kind == "lowpass":
            return F.lowpass_filter(fc=fc, gain=g, fs=fs, dtype=dtype)
        if kind == "highpass":
            return F.highpass_filter(fc=fc, gain=g, fs=fs, dtype=dtype)
        if kind == "bandpass":
            return F.bandpass_filter(fc1=fc, fc2=(fc * 1.7).clamp(max=0.49 * fs), gain=g, fs=fs, dtype=dtype)
        if kind.startswith("shelving"):
            return F.shelving_filter(fc=fc, gain=10 ** (g / 20), type=kind.split("_")[1], fs=fs, dtype=dtype)
        return F.peak_filter(fc=fc, gain=10 ** (g / 20), Q=torch.full_like(fc, Q), fs=fs, dtype=dtype)

    try:
        br, ar = call(RF)
    except Exception:
        assume(False)
    b, a = call(MF)
    tol = 1e-12 if dtype == torch.float64 else 2e-6
    for mine, ref in ((b, br), (a, ar)):
        assert mine.shape == ref.shape and mine.dtype == ref.dtype, kind
        assert float((mine - ref).abs().max()) <= tol * float(ref.abs().max() + 1e-30), kind


@pytest.mark.parametrize("signal_type", ["impulse", "sine", "sweep", "wgn", "exp", "noise", "reference", "foo"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_signal_gallery_matches_the_reference(signal_type, dtype):
    """functional.signal_gallery (reference functional.py:164-270): same signals (seeded for the noise types), same
    errors for an unknown type / a missing reference."""
    reference_modules()
    import flamo.functional as RF

    from flamo_b200 import functional as MF

    def call(F):
        torch.manual_seed(3)
        return F.signal_gallery(2, 480, 3, signal_type=signal_type, fs=48000, rate=5.0, dtype=dtype)

    ref, mine = _outcome(lambda: call(RF)), _outcome(lambda: call(MF))
    assert ref == mine
    if ref == "ok":
        r, m = call(RF), call(MF)
        assert r.shape == m.shape and r.dtype == m.dtype
        assert float((r - m).abs().max()) <= (1e-12 if dtype == torch.float64 else 1e-5)


def test_small_helpers_match_the_reference():
    """functional.{hertz2rad, rad2hertz, db2mag, mag2db, get_magnitude, skew_matrix}, utils.to_complex, eq.eq_freqs:
    same values, shapes and dtypes as the reference (functional.py:42-56, 306-353; utils.py:12-22; eq.py:8-54)."""
    reference_modules()
    import flamo.functional as RF
    import importlib

    RU = importlib.import_module("flamo.utils")  # (the package attribute `flamo.utils` is shadowed by optimize.utils)
    REQ = importlib.import_module("flamo.auxiliary.eq")

    from flamo_b200 import functional as MF, utils as MU
    from flamo_b200.auxiliary import eq as MEQ

    g = torch.Generator().manual_seed(0)
    for dtype in (torch.float32, torch.float64):
        x = torch.rand(4, 3, generator=g, dtype=dtype) * 100 + 0.1
        for name, args in (("hertz2rad", (x, 48000)), ("rad2hertz", (x / 100, 44100)), ("db2mag", (x - 50,)),
                           ("mag2db", (x,)), ("skew_matrix", (torch.randn(5, 5, generator=g, dtype=dtype),))):
            r, m = getattr(RF, name)(*args), getattr(MF, name)(*args)
            assert r.shape == m.shape and r.dtype == m.dtype, name
            assert torch.allclose(r, m, rtol=1e-6 if dtype == torch.float32 else 1e-13, atol=0), name
        z = torch.complex(x, x.flip(0))
        assert torch.equal(RF.get_magnitude(z), MF.get_magnitude(z))
        rc, mc = RU.to_complex(x), MU.to_complex(x)
        assert rc.dtype == mc.dtype and torch.equal(rc, mc)
    for interval in (1, 3):
        rf, mf = REQ.eq_freqs(interval=interval), MEQ.eq_freqs(interval=interval)
        rf, mf = (rf[0] if isinstance(rf, tuple) else rf), (mf[0] if isinstance(mf, tuple) else mf)
        assert torch.allclose(torch.as_tensor(rf).double(), torch.as_tensor(mf).double(), rtol=1e-6)


@settings(max_examples=100, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(st.lists(st.tuples(st.sampled_from(["append", "prepend", "insert"]), st.integers(-4, 4),
                          st.sampled_from([None, "a", "b", "7", "gain"]), st.booleans()), min_size=1, max_size=5))
def test_series_editing_matches_the_reference(edits):
    """Series.append / prepend / insert with plain modules, keyed OrderedDicts and nested Series (reference
    system.py:33-125): the same keys in the same order, the same channel counts, the same exceptions."""
    from collections import OrderedDict

    rdsp, rsystem = reference_modules()
    kw = dict(size=(2,), nfft=64, dtype=torch.float64)

    def play(dsp_, system_):
        s = system_.Series(OrderedDict({"first": dsp_.parallelGain(**kw)}), dsp_.parallelGain(**kw))
        log = []
        for op, idx, key, nested in edits:
            new = dsp_.parallelGain(**kw)
            if nested:
                new = system_.Series(new, dsp_.parallelGain(**kw))
            if key is not None:
                new = OrderedDict({key: new})
            if op == "append":
                log.append(_outcome(lambda: s.append(new)))
            elif op == "prepend":
                log.append(_outcome(lambda: s.prepend(new)))
            else:
                log.append(_outcome(lambda: s.insert(idx, new)))
            log.append(tuple(s._modules.keys()))
        return log, (s.input_channels, s.output_channels), len(s)

    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert play(dsp, system) == play(rdsp, rsystem)


@settings(max_examples=150, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(st.tuples(st.integers(1, 3), st.integers(1, 3)), st.one_of(st.none(), st.tuples(st.integers(1, 3), st.integers(1, 3))),
       st.one_of(st.none(), st.tuples(st.integers(1, 3), st.integers(1, 3))), st.sampled_from([64, 128]),
       st.sampled_from([64, 128]), st.sampled_from(["plain", "sequential", "ordered"]))
def test_shell_construction_checks_match_the_reference(core, inp, out, nfft_core, nfft_layer, wrap):
    """Shell.__init__ (reference system.py:800-976): channel / nfft consistency checks between the core and the input
    / output layers, conversion of nn.Sequential / OrderedDict cores to Series, resulting channel counts."""
    from collections import OrderedDict

    import torch.nn as nn

    rdsp, rsystem = reference_modules()

    def build(dsp_, system_):
        c = dsp_.Gain(size=core, nfft=nfft_core, dtype=torch.float64)
        if wrap == "sequential":
            c = nn.Sequential(c)
        elif wrap == "ordered":
            c = OrderedDict({"g": c})
        i = dsp_.FFT(nfft_layer, dtype=torch.float64) if inp is None else dsp_.Gain(size=inp, nfft=nfft_layer,
                                                                                      dtype=torch.float64)
        o = dsp_.iFFT(nfft_layer, dtype=torch.float64) if out is None else dsp_.Gain(size=out, nfft=nfft_layer,
                                                                                       dtype=torch.float64)
        s = system_.Shell(core=c, input_layer=i, output_layer=o)
        return (s.input_channels, s.output_channels, s.nfft, type(s.get_core()).__name__,
                tuple(s.state_dict().keys()))

    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mine, ref = _outcome(lambda: build(dsp, system)), _outcome(lambda: build(rdsp, rsystem))
        assert mine == ref, (core, inp, out, nfft_core, nfft_layer, wrap)
        if ref == "ok":
            assert build(dsp, system) == build(rdsp, rsystem)


@settings(max_examples=150, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(st.sampled_from(["Recursion", "Parallel_sum", "Parallel_cat"]), st.tuples(st.integers(1, 3), st.integers(1, 3)),
       st.tuples(st.integers(1, 3), st.integers(1, 3)), st.sampled_from([64, 128]), st.sampled_from([64, 128]),
       st.sampled_from(["plain", "sequential", "ordered"]))
def test_recursion_and_parallel_construction_match_the_reference(kind, sa, sb, nfft_a, nfft_b, wrap):
    """Recursion.__init__ / Parallel.__init__ (reference system.py:363-395, 473-515, 600-738): path conversion to
    Series, nfft / channel consistency assertions, resulting channel counts and state_dict keys."""
    from collections import OrderedDict

    import torch.nn as nn

    rdsp, rsystem = reference_modules()

    def build(dsp_, system_):
        a = dsp_.Gain(size=sa, nfft=nfft_a, dtype=torch.float64)
        b = dsp_.Gain(size=sb, nfft=nfft_b, dtype=torch.float64)
        if wrap == "sequential":
            a = nn.Sequential(a)
        elif wrap == "ordered":
            b = OrderedDict({"g": b})
        if kind == "Recursion":
            m = system_.Recursion(fF=a, fB=b)
        else:
            m = system_.Parallel(a, b, sum_output=kind.endswith("sum"))
        return m.input_channels, m.output_channels, m.nfft, tuple(m.state_dict().keys())

    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mine, ref = _outcome(lambda: build(dsp, system)), _outcome(lambda: build(rdsp, rsystem))
        assert mine == ref, (kind, sa, sb, nfft_a, nfft_b, wrap)
        if ref == "ok":
            assert build(dsp, system) == build(rdsp, rsystem)


@settings(max_examples=80, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(st.sampled_from(["FFT", "iFFT", "FFTAntiAlias", "iFFTAntiAlias"]), st.sampled_from(["backward", "ortho", "forward"]),
       st.sampled_from([64, 96]), st.sampled_from([0.0, 20.0, -35.0]), st.integers(1, 3), st.integers(1, 3),
       st.sampled_from([torch.float32, torch.float64]), st.booleans())
def test_transform_layers_match_the_reference(cls, norm, nfft, alias, B, N, dtype, short):
    """dsp.FFT / iFFT / FFTAntiAlias / iFFTAntiAlias (reference dsp.py:45-160): same values, shapes and dtypes for
    inputs shorter than or equal to nfft, every norm, with and without the anti-aliasing envelope."""
    rdsp, rsystem = reference_modules()
    g = torch.Generator().manual_seed(nfft + B + 10 * N)
    if cls.startswith("i"):
        x = torch.randn(B, nfft // 2 + 1, N, generator=g, dtype=dtype) + 1j * torch.randn(B, nfft // 2 + 1, N, generator=g,
                                                                                         dtype=dtype)
    else:
        x = torch.randn(B, nfft // 2 + 3 if short else nfft, N, generator=g, dtype=dtype)

    def run(dsp_):
        kw = dict(nfft=nfft, norm=norm, dtype=dtype)
        if "AntiAlias" in cls:
            kw["alias_decay_db"] = alias
        return getattr(dsp_, cls)(**kw)(x)

    ref = _outcome(lambda: run(rdsp))
    assert _outcome(lambda: run(dsp)) == ref
    if ref == "ok":
        r, m = run(rdsp), run(dsp)
        assert r.shape == m.shape and r.dtype == m.dtype
        assert float((r - m).abs().max()) <= (1e-4 if dtype == torch.float32 else 1e-11) * float(r.abs().max() + 1e-30)


@pytest.mark.parametrize("name,kwargs,tol", [
    ("Biquad", dict(size=(2, 3), n_sections=2, filter_type="lowpass"), 1e-10),
    ("Biquad", dict(size=(1, 2), n_sections=3, filter_type="bandpass"), 1e-10),
    ("parallelBiquad", dict(size=(3,), n_sections=2, filter_type="highpass"), 1e-10),
    ("SVF", dict(size=(2, 2), n_sections=2, filter_type=None), 5e-3),            # float32 tap buffers in the reference
    ("parallelSVF", dict(size=(3,), n_sections=1, filter_type="peaking"), 5e-3),
    ("GEQ", dict(size=(1, 2), octave_interval=1), 0.2),                          # float32 eq.geq in the reference
    ("parallelGEQ", dict(size=(2,), octave_interval=1), 0.2),
    ("SOSFilter", dict(size=(2, 2), n_sections=2), 1e-10),
])
def test_public_module_methods_match_the_reference(name, kwargs, tol):
    """The methods other flamo code calls on a section filter: get_poly_coeff(map(param)) -> (H, B, A),
    freq_response(param), assign_value with an index, s2sample / sample2s of the delays."""
    rdsp, rsystem = reference_modules()
    nfft = 128

    def build(dsp_):
        torch.manual_seed(9)
        return getattr(dsp_, name)(nfft=nfft, alias_decay_db=10.0, dtype=torch.float64, fs=48000, **kwargs)

    r, m = build(rdsp), build(dsp)
    with torch.no_grad():
        outs_r, outs_m = r.get_poly_coeff(r.map(r.param)), m.get_poly_coeff(m.map(m.param))
        assert len(outs_r) == len(outs_m) == 3
        for a, b in zip(outs_r, outs_m):
            assert a.shape == b.shape and a.dtype == b.dtype, name
            assert float((a - b).abs().max()) <= tol * float(a.abs().max()), name
        Hr, Hm = r.freq_response(r.param), m.freq_response(m.param)
        assert Hr.shape == Hm.shape and float((Hr - Hm).abs().max()) <= tol * float(Hr.abs().max())
    new = torch.full_like(r.param[0], 0.25)
    r.assign_value(new, (0,))
    m.assign_value(new, (0,))
    assert torch.equal(r.param.detach(), m.param.detach()) and m.new_value == r.new_value
    d_r = rdsp.parallelDelay(size=(3,), max_len=500, unit=10, fs=44100, nfft=nfft, dtype=torch.float64)
    d_m = dsp.parallelDelay(size=(3,), max_len=500, unit=10, fs=44100, nfft=nfft, dtype=torch.float64)
    v = torch.tensor([1.0, 12.5, 499.0], dtype=torch.float64)
    assert torch.equal(d_r.sample2s(v), d_m.sample2s(v)) and torch.equal(d_r.s2sample(v), d_m.s2sample(v))


def test_datasets_and_criteria_match_the_reference():
    """optimize.dataset.{Dataset, DatasetColorless, load_dataset} and optimize.loss.{mse_loss, sparsity_loss}
    (reference dataset.py:9-174, loss.py:12-103): same items, same loader lengths, same loss values and gradients."""
    rdsp, rsystem = reference_modules()
    from flamo.optimize import dataset as RD, loss as RL

    from flamo_b200.optimize import dataset as MD, loss as ML

    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(1, 33, 2, generator=g), torch.rand(1, 33, 1, generator=g)
    for D in (RD, MD):
        ds = D.Dataset(input=x, target=y, expand=7, device="cpu", dtype=torch.float64)
        assert len(ds) == 7 and ds[3][0].dtype == torch.float64 and torch.equal(ds[3][0], x[0].double())
        tr, va = D.load_dataset(ds, batch_size=2, split=0.7, shuffle=False, device="cpu")
        assert (len(tr), len(va)) == (2, 1)  # int(7 * 0.7) = 4 -> 2 batches; 3 -> 1 batch (drop_last)
        dc = D.DatasetColorless(input_shape=(1, 17, 3), target_shape=(1, 17, 1), expand=4, device="cpu")
        assert dc.input.shape == (4, 17, 3) and float(dc.input[:, 0].min()) == 1 and float(dc.input[:, 1:].abs().max()) == 0
        assert dc.target.shape == (4, 17, 1) and float(dc.target.min()) == 1
    pred = torch.rand(3, 33, 4, generator=g, dtype=torch.float64, requires_grad=True)
    tgt = torch.rand(3, 33, 1, generator=g, dtype=torch.float64)
    lr, lm = RL.mse_loss(nfft=64, device="cpu")(pred, tgt), ML.mse_loss(nfft=64, device="cpu")(pred, tgt)
    assert torch.equal(lr, lm)
    assert torch.equal(torch.autograd.grad(lr, pred)[0], torch.autograd.grad(lm, pred)[0])
    for N in (4, 6):
        def fdn(dsp_, system_):
            torch.manual_seed(N)
            core = W.build(W.fdn(N, delays=list(range(11, 11 + 2 * N, 2))), dsp_, system_, 64, 30.0, dtype=torch.float64)
            return system_.Shell(core, dsp_.FFT(64, dtype=torch.float64))
        mr, mm = fdn(rdsp, rsystem), fdn(dsp, system)
        sr, sm = RL.sparsity_loss()(None, None, mr), ML.sparsity_loss()(None, None, mm)
        assert abs(float(sr) - float(sm)) <= 1e-12
        gr = torch.autograd.grad(sr, mr.get_core().feedback_loop.feedback.param)[0]
        gm = torch.autograd.grad(sm, mm.get_core().feedback_loop.feedback.param)[0]
        assert torch.allclose(gr, gm, rtol=1e-9, atol=1e-12)
