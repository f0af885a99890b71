"""CPU: randomly composed module trees (hypothesis) through this package's host side (constructors, maps, lowering,
coefficient packing, autograd seams; float64, ABI emulator) against the oracle on the same raw parameters — forward
response and every parameter gradient.  Complements the fixed case list of tests/cases.py with compositions nobody
wrote down: nested Series inside Recursion paths, rectangular loops, trailing columns, mixed filter kinds."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

import cases as C
from flamo_b200 import workloads as W
from flamo_b200.processor import dsp, system
from helpers import rel_err
from oracle import flamo_oracle as O

pytestmark = pytest.mark.usefixtures("emulated_backend")
FS = W.FS
NFFT = 256


EXTRA = False  # set by tree(extra=True)


@st.composite
def leaf(draw, n_in, n_out=None, allow_delay=True):
    """One dsp module description with n_in inputs (and n_out outputs if given, else drawn)."""
    square = n_out is not None and n_out == n_in
    kinds = ["Gain", "Biquad", "SVF", "GEQ", "Filter", "GainDelay"]
    if EXTRA:  # (kept as a switch: the reference-differential file asks for a denser mix of the rarer kinds)
        kinds += ["SOSFilter", "SOSFilter"]
    kinds.append("SOSFilter")
    if n_out is None or square:
        kinds.append("parallelSOSFilter")
        if n_in in (2, 4):
            kinds += ["Matrix:hadamard", "Matrix:rotation"]
    if allow_delay:
        kinds.append("Delay")
    if n_out is None or square:
        kinds += ["parallelGain", "parallelBiquad", "parallelSVF", "parallelFilter", "parallelGainDelay"]
        if allow_delay:
            kinds.append("parallelDelay")
        if n_in > 1:
            kinds += ["Matrix", "HouseholderMatrix"]
    kind = draw(st.sampled_from(kinds))
    rg = draw(st.booleans()) or kind in ("Gain", "parallelGain")
    if kind.startswith("Matrix:"):
        return ("Matrix", dict(size=(n_in, n_in), matrix_type=kind.split(":")[1], requires_grad=rg)), n_in
    if kind.startswith("parallel") or kind in ("Matrix", "HouseholderMatrix"):
        m = n_in
    else:
        m = n_out if n_out is not None else draw(st.integers(1, 4))
    size = (m,) if kind.startswith("parallel") else (m, n_in)
    kw = dict(size=size, requires_grad=rg)
    if kind in ("Biquad", "parallelBiquad"):
        kw.update(n_sections=draw(st.integers(1, 3)), filter_type=draw(st.sampled_from(["lowpass", "highpass", "bandpass"])),
                  fs=FS)
    elif kind in ("SVF", "parallelSVF"):
        kw.update(n_sections=draw(st.integers(1, 2)), fs=FS,
                  filter_type=draw(st.sampled_from([None, "lowpass", "highpass", "bandpass", "lowshelf", "highshelf",
                                                    "peaking", "notch"])))
    elif kind == "GEQ":
        kw.update(octave_interval=1, fs=FS)
    elif kind in ("Filter", "parallelFilter"):
        kw["size"] = (draw(st.integers(1, 9)),) + size
    elif kind in ("Delay", "parallelDelay", "GainDelay", "parallelGainDelay"):
        kw.update(max_len=draw(st.integers(5, 60)), isint=draw(st.booleans()), fs=FS)
        if kw["isint"] and kind in ("Delay", "parallelDelay"):
            kw["requires_grad"] = False
    elif kind in ("SOSFilter", "parallelSOSFilter"):
        K = draw(st.integers(1, 3))
        kw = dict(size=size, n_sections=K, fs=FS)  # (no requires_grad argument; the a0-normalising map of the reference
        # writes in place and cannot be differentiated, dsp.py:1857-1862)
        return (kind, kw, {"assign": C._sos_coeffs(K, size, draw(st.integers(0, 999)))}), m
    elif kind == "Matrix":
        kw["matrix_type"] = draw(st.sampled_from(["orthogonal", "random"]))
    elif kind == "HouseholderMatrix":
        pass
    return (kind, kw), m


@st.composite
def chain(draw, n_in, n_out=None, max_len=3, allow_delay=True):
    n = draw(st.integers(1, max_len))
    descs, cur = [], n_in
    for j in range(n):
        last = j == n - 1
        d, cur = draw(leaf(cur, n_out if last else None, allow_delay))
        descs.append(d)
    if n_out is not None and cur != n_out:  # a diagonal / square kind came last: close with a Gain
        descs.append(("Gain", dict(size=(n_out, cur), requires_grad=True)))
        cur = n_out
    return (descs[0] if len(descs) == 1 and draw(st.booleans()) else ("Series", descs)), cur


@st.composite
def tree(draw, extra=False):
    global EXTRA
    EXTRA = extra
    try:
        return draw(_tree())
    finally:
        EXTRA = False


@st.composite
def _tree(draw):
    n_in = draw(st.integers(1, 3))
    parts, cur = [], n_in
    if draw(st.booleans()):
        d, cur = draw(chain(cur, None, 2))
        parts.append(d)
    if draw(st.booleans()):
        N = draw(st.integers(1, 5))
        ff, _ = draw(chain(cur, N, 2))
        fb, _ = draw(chain(N, cur, 2))
        # keep the loop gain small so that I - F Fb is well conditioned whatever the random parameters are
        damp = ("parallelGain", dict(size=(cur,), requires_grad=True), {"assign": [0.02] * cur})
        fb = ("Series", (list(fb[1]) if fb[0] == "Series" else [fb]) + [damp])
        parts.append(("Recursion", ff, fb))
        cur = N
    if draw(st.integers(0, 3)) == 0:  # a Parallel node: two branches on the same input, summed or concatenated
        total = draw(st.booleans())
        a, na = draw(chain(cur, None, 2))
        b, nb = draw(chain(cur, na if total else None, 2))
        parts.append(("Parallel", a, b, total))
        cur = na if total else na + nb
    if draw(st.integers(0, 4)) == 0:  # a second loop: the series splits into two launches
        N = draw(st.integers(1, 3))
        damp = ("parallelGain", dict(size=(cur,), requires_grad=True), {"assign": [0.03] * cur})
        parts.append(("Recursion", ("Gain", dict(size=(N, cur), requires_grad=True)),
                      ("Series", [("Gain", dict(size=(cur, N), requires_grad=True)), damp])))
        cur = N
    if draw(st.booleans()) or not parts:
        d, cur = draw(chain(cur, None, 2))
        parts.append(d)
    desc = parts[0] if len(parts) == 1 and parts[0][0] not in ("Recursion", "Parallel") and draw(st.booleans()) \
        else ("Series", parts)
    return desc, n_in, draw(st.integers(1, 2)), draw(st.sampled_from([None, None, 2])), draw(st.integers(0, 10 ** 6)), \
        draw(st.sampled_from([0.0, 30.0]))


@settings(max_examples=150, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(tree())
def test_random_tree_matches_oracle(t):
    desc, n_in, B, cols, seed, alias = t
    torch.manual_seed(seed)
    model = W.build(desc, dsp, system, NFFT, alias, dtype=torch.float64, device="cpu")
    assert model.input_channels == n_in
    M = NFFT // 2 + 1
    X = C.make_input(B, M, n_in, cols).requires_grad_(True)  # the gradient w.r.t. the input signal is checked too
    Xo = X.detach().clone().requires_grad_(True)
    params = list(model.parameters())
    Y = model(X)
    ps = [p.detach().clone().requires_grad_(p.requires_grad) for p in params]
    Yo = O.forward(O.from_desc(desc), Xo, ps, NFFT, alias)
    assert Y.shape == Yo.shape
    assert rel_err(Y.detach().numpy(), Yo.detach().numpy()) <= 1e-8, desc
    gp = [p for p in ps if p.requires_grad]
    a = Yo.detach().abs()
    if float(a.min()) <= 1e-9 * float(a.max()):
        return  # |Y| = 0 somewhere: the gradient of |.| is not defined there (see kink() in the reference file)
    C.golden_loss(Y).backward()
    *go, gxo = torch.autograd.grad(C.golden_loss(Yo), gp + [Xo], allow_unused=True)
    assert X.grad is not None and float((X.grad - gxo).abs().max()) <= 1e-8 * float(gxo.abs().max() + 1e-300), desc
    if gp:
        k = 0
        for p, q in zip(params, ps):
            if not q.requires_grad:
                continue
            ref = go[k]
            k += 1
            if ref is None or float(ref.abs().max()) == 0.0:
                assert p.grad is None or float(p.grad.abs().max()) <= 1e-12, desc
                continue
            assert p.grad is not None, desc
            scale = max([float(g.abs().max()) for g in go if g is not None] + [1e-6])  # (floor: all-zero true gradients leave rounding noise)
            assert float((p.grad - ref).abs().max()) <= 1e-7 * scale, desc


@settings(max_examples=150, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(tree(), st.integers(1, NFFT // 2 - 1))
def test_random_tree_is_bin_shard_invariant(t, cut):
    """Evaluating two bin ranges separately (what the bin-sharded multi-GPU step does) and concatenating
    equals the full sweep, whatever the tree: launch boundaries (Parallel nodes, second loops) included."""
    from flamo_b200 import sweep

    desc, n_in, B, cols, seed, alias = t
    torch.manual_seed(seed)
    model = W.build(desc, dsp, system, NFFT, alias, dtype=torch.float64, device="cpu")
    M = NFFT // 2 + 1
    X = C.make_input(B, M, n_in, cols)
    with torch.no_grad():
        full = model(X)
        with sweep.bin_shard(0, cut):
            lo = model(X)
        with sweep.bin_shard(cut, M):
            hi = model(X)
    assert lo.shape[1] == cut and hi.shape[1] == M - cut
    assert torch.allclose(torch.cat((lo, hi), dim=1), full, rtol=1e-12, atol=1e-12 * float(full.abs().max()))


@settings(max_examples=60, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(tree(), st.sampled_from(["mse_loss", "MSELoss"]))
def test_random_tree_trainer_fused_equals_unfused(t, crit_kind):
    """Trainer.train_step on a random Shell(FFT, tree, |.|): the fused-criterion route (one backward_loss call, or its
    fallback when the tree is more than one launch) takes exactly the steps of the unfused route."""
    from flamo_b200.optimize.loss import mse_loss
    from flamo_b200.optimize.trainer import Trainer

    desc, n_in, B, cols, seed, alias = t
    M = NFFT // 2 + 1

    def run(fuse):
        torch.manual_seed(seed)
        core = W.build(desc, dsp, system, NFFT, alias, dtype=torch.float64, device="cpu")
        model = system.Shell(core, dsp.FFT(NFFT, dtype=torch.float64),
                             dsp.Transform(lambda x: torch.abs(x), dtype=torch.float64))
        tr = Trainer(model, max_epochs=1, lr=1e-2, log=False, device="cpu", fuse_criterion=fuse)
        n_out = core.output_channels
        if crit_kind == "mse_loss":
            tr.register_criterion(mse_loss(nfft=NFFT), 1)
            tgt = torch.full((B, M, 1), 0.7, dtype=torch.float64)
        else:
            tr.register_criterion(torch.nn.MSELoss(), 0.5)
            tgt = torch.full((B, M, n_out), 0.3, dtype=torch.float64)
        x = torch.zeros(B, NFFT, n_in, dtype=torch.float64)
        x[:, 0] = 1
        x[:, 5] = -0.25
        with torch.no_grad():
            est = model(x)
        if float(est.min()) <= 1e-9 * float(est.max()):
            return None  # |Y| = 0 at some bin (e.g. a low-pass section's zero at Nyquist with alias_decay_db = 0): the
            # gradient of |.| there is the direction of Y's rounding noise, different on every evaluation route
        losses = [tr.train_step((x, tgt)) for _ in range(2)]
        return losses, [p.detach().clone() for p in model.parameters()]

    if not any(p.requires_grad for p in W.build(desc, dsp, system, NFFT, alias, dtype=torch.float64).parameters()):
        return
    def attempt(fuse):
        try:
            return run(fuse)
        except RuntimeError as e:  # e.g. the only trainable parameter is a Hadamard matrix's: nothing to differentiate
            return f"RuntimeError: {str(e)[:40]}"

    rf, ru = attempt(True), attempt(False)
    if isinstance(rf, str) or isinstance(ru, str):
        assert rf == ru, (rf, ru, desc)
        return
    if rf is None or ru is None:
        return
    (lf, pf), (lu, pu) = rf, ru
    assert np.allclose(lf, lu, rtol=1e-7, atol=1e-14), desc
    for a, b in zip(pf, pu):
        # Adam normalises: a parameter whose true gradient is (near) zero moves by lr * g / (|g| + eps) with g the
        # rounding noise of the route taken, so parameters are only pinned to 1e-2 of a real update (lr = 1e-2);
        # the second step's loss above is the sharp check of the parameters that matter
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-4), desc


@settings(max_examples=60, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(tree())
def test_random_tree_responses_match_oracle(t):
    """Shell.get_time_response / get_freq_response (layer swapping, rising envelope; reference system.py:1012-1153) on
    random trees against the oracle's restatement (itself pinned on the shipped checkpoints)."""
    desc, n_in, B, cols, seed, alias = t
    torch.manual_seed(seed)
    core = W.build(desc, dsp, system, NFFT, alias, dtype=torch.float64, device="cpu")
    out_layer = dsp.Transform(lambda x: torch.abs(x), dtype=torch.float64)
    model = system.Shell(core, dsp.FFT(NFFT, dtype=torch.float64), out_layer)
    ps = [p.detach().clone() for p in model.parameters()]
    node = O.from_desc(desc)
    with torch.no_grad():
        h, H = model.get_time_response(), model.get_freq_response()
        ho = O.time_response(node, ps, NFFT, alias, n_in)
        Ho = O.freq_response(node, ps, NFFT, alias, n_in)
    assert h.shape == ho.shape and H.shape == Ho.shape
    assert float((h - ho).abs().max()) <= 1e-9 * float(ho.abs().max() + 1e-300)
    assert float((H - Ho).abs().max()) <= 1e-9 * float(Ho.abs().max() + 1e-300)
    assert model.get_outputLayer() is out_layer  # layers restored


@settings(max_examples=100, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(tree())
def test_random_tree_float32_modules_stay_within_the_float32_bar(t):
    """float32 modules (the bench dtype): raw parameters, maps evaluated in float64 and cast once, coefficient
    packing (Taylor blocks of the sections, float64 delays) — everything the host hands the float32 kernels — against
    the float64 oracle on the same (float32-representable) parameters, at BASELINE.md's bar: 1e-4 relative on the
    response, denominator floored at 1e-3 of the peak.  (The emulator computes in float64 from these float32
    coefficients: what is measured is the error the HOST side contributes to the bar.)"""
    desc, n_in, B, cols, seed, alias = t
    torch.manual_seed(seed)
    model = W.build(desc, dsp, system, NFFT, alias, dtype=torch.float32, device="cpu")
    M = NFFT // 2 + 1
    X = C.make_input(B, M, n_in, cols).to(torch.complex64)
    with torch.no_grad():
        Y = model(X)
        ps = [p.detach().double() for p in model.parameters()]
        Yo = O.forward(O.from_desc(desc), X.to(torch.complex128), ps, NFFT, alias)
    assert Y.dtype == torch.complex64
    assert rel_err(Y.numpy().astype(np.complex128), Yo.numpy()) <= 1e-4, desc
