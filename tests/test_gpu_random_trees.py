"""GPU: (1) the coalesced gradient finalize kernel (fsweep_finalize_v2_kernel, the default since round 2) against the
round-1 kernel (FSWEEP_FINALIZE_V2=0) on every parity case; (2) the random module trees of
tests/test_random_trees_cpu.py on the REAL kernels: float64 kernels at 1e-8 / gradients 1e-6, float32 kernels at the
1e-4 bar; (3) captured vs eager training steps on random trees.  First run green on a B200 in round 2
(gpurun_out/r02_experimental_all.log: 89 passed; the two failures were test tolerances, see below)."""
import os

import pytest
import torch

import cases as C
from helpers import build_case

pytestmark = [pytest.mark.gpu]


def _grads(name, dtype, variant):
    os.environ["FSWEEP_FINALIZE_V2"] = variant  # read by libfsweep at every backward call ("0": the round-1 kernel)
    try:
        case, g, model = build_case(name, dtype, "cuda")
        M = case["nfft"] // 2 + 1
        cdt = torch.complex64 if dtype == torch.float32 else torch.complex128
        X = C.make_input(case["B"], M, model.input_channels, case["C"]).to(cdt).cuda()
        params = [p for p in model.parameters() if p.requires_grad]
        if not params:
            return None
        C.golden_loss(model(X)).backward()
        torch.cuda.synchronize()
        return [p.grad.detach().clone() for p in params]
    finally:
        os.environ.pop("FSWEEP_FINALIZE_V2", None)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("name", list(C.CASES))
def test_finalize_v2_equals_default(name, dtype):
    a = _grads(name, dtype, "0")
    if a is None:
        return  # (cases without a trainable parameter have no gradient to finalize)
    b = _grads(name, dtype, "1")
    # both kernels sum the per-block partials in float64; programs whose accumulators overflow shared memory add
    # float atomics in a flat buffer whose order differs from run to run (cfg4: 1.4e-6 between two runs of the SAME
    # kernel), hence 5e-6 rather than one ulp
    eps = 5e-6 if dtype == torch.float32 else 1e-13
    for ga, gb in zip(a, b):
        assert torch.isfinite(gb).all()
        assert float((ga - gb).abs().max()) <= eps * float(ga.abs().max() + 1e-30)


# ---------------------------------------------------------------------------------------------------------------------
# The random module trees of tests/test_random_trees_cpu.py on the REAL kernels: float64 kernels at 1e-8 / gradients
# 1e-6, float32 kernels at the 1e-4 bar (loops are damped by the generator, so conditioning is benign).
from hypothesis import HealthCheck, given, settings  # noqa: E402

from test_random_trees_cpu import NFFT, tree  # noqa: E402


@settings(max_examples=150, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(tree())
def test_random_tree_on_the_kernels(t):
    import numpy as np

    from flamo_b200 import workloads as W
    from flamo_b200.processor import dsp, system
    from helpers import rel_err
    from oracle import flamo_oracle as O

    desc, n_in, B, cols, seed, alias = t
    M = NFFT // 2 + 1
    X = C.make_input(B, M, n_in, cols)
    torch.manual_seed(seed)
    m64 = W.build(desc, dsp, system, NFFT, alias, dtype=torch.float64, device="cuda")
    torch.manual_seed(seed)
    m32 = W.build(desc, dsp, system, NFFT, alias, dtype=torch.float32, device="cuda")
    W.set_params(m64, [p.detach() for p in m32.parameters()])  # the same float32-representable parameters
    ps = [p.detach().cpu().double().requires_grad_(p.requires_grad) for p in m64.parameters()]
    Yo = O.forward(O.from_desc(desc), X, ps, NFFT, alias)
    Y64 = m64(X.cuda())
    with torch.no_grad():
        Y32 = m32(X.to(torch.complex64).cuda())
    assert rel_err(Y64.detach().cpu().numpy(), Yo.detach().numpy()) <= 1e-8, desc
    assert rel_err(Y32.cpu().numpy().astype(np.complex128), Yo.detach().numpy()) <= 1e-4, desc
    gp = [p for p in ps if p.requires_grad]
    if gp:
        C.golden_loss(Y64).backward()
        go = torch.autograd.grad(C.golden_loss(Yo), gp, allow_unused=True)
        scale = max([float(g.abs().max()) for g in go if g is not None] + [1e-6])  # (floor: where every true gradient is zero only rounding noise is left)
        k = 0
        for p, q in zip(m64.parameters(), ps):
            if not q.requires_grad:
                continue
            ref = go[k]
            k += 1
            if ref is None or float(ref.abs().max()) == 0.0:
                continue
            assert p.grad is not None, desc
            assert float((p.grad.cpu() - ref).abs().max()) <= 1e-6 * scale, desc


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("ft", ["lowpass", "highpass", "bandpass"])
def test_biquad_cutoff_clamped_at_nyquist(ft, dtype):
    """Raw cutoff parameters beyond 1 are clamped to Nyquist by the Biquad map (reference dsp.py:1528-1563): with
    alias_decay_db = 0 the low-pass section's double zero and its poles then cancel AT the Nyquist bin (B = 0 and
    A = -1.1e-16 there; the reference returns 0).  The kernels form z^-1 with sincospi, i.e. exactly -1 at that bin, and
    must return ~0 too — the float64 emulator did not until it was given exact quarter-point phases (found by
    tests/test_random_trees_reference_cpu.py::test_random_tree_far_from_the_initial_parameters)."""
    import numpy as np

    from flamo_b200 import workloads as W
    from flamo_b200.processor import dsp, system
    from helpers import rel_err
    from oracle import flamo_oracle as O

    nfft = 256
    desc = ("Biquad", dict(size=(2, 1), n_sections=1, filter_type=ft, fs=48000, requires_grad=False))
    torch.manual_seed(0)
    model = W.build(desc, dsp, system, nfft, 0.0, dtype=dtype, device="cuda")
    with torch.no_grad():
        for p in model.parameters():
            p.mul_(4.0)
        cdt = torch.complex64 if dtype == torch.float32 else torch.complex128
        X = C.make_input(1, nfft // 2 + 1, 1, None).to(cdt).cuda()
        Y = model(X).cpu().numpy().astype(np.complex128)
        ps = [p.detach().cpu().double() for p in model.parameters()]
        Yo = O.forward(O.from_desc(desc), X.cpu().to(torch.complex128), ps, nfft, 0.0).numpy()
    peak = np.abs(Yo).max()
    if peak < 1e-9:  # high-pass / band-pass at Nyquist: an all-stop filter
        assert np.abs(Y).max() < 1e-6
    else:
        assert rel_err(Y, Yo) <= (1e-4 if dtype == torch.float32 else 1e-8)


@settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(tree())
def test_random_tree_captured_step_equals_eager(t):
    """Trainer.train_step replayed from a CUDA graph vs stepped eagerly, on random trees (float32): same losses, same
    parameters, and the capture itself must not have fallen back (maps that build tensors from host data, eager
    Parallel segments, multi-launch series are all legal inside a capture or the Trainer has a bug)."""
    import warnings

    import numpy as np

    from flamo_b200 import workloads as W
    from flamo_b200.optimize.loss import mse_loss
    from flamo_b200.optimize.trainer import Trainer
    from flamo_b200.processor import dsp, system

    desc, n_in, B, cols, seed, alias = t
    M = NFFT // 2 + 1
    x = torch.zeros(B, NFFT, n_in, device="cuda")
    x[:, 0] = 1
    x[:, 9] = 0.5
    tgt = torch.full((B, M, 1), 0.6, device="cuda")

    def run(graph):
        torch.manual_seed(seed)
        core = W.build(desc, dsp, system, NFFT, alias, dtype=torch.float32, device="cuda")
        model = system.Shell(core, dsp.FFT(NFFT), dsp.Transform(lambda v: torch.abs(v)))
        if not any(p.requires_grad for p in model.parameters()):
            return None
        tr = Trainer(model, max_epochs=1, lr=1e-3, log=False, device="cuda", graph=graph)
        tr.register_criterion(mse_loss(nfft=NFFT, device="cuda"), 1)
        with warnings.catch_warnings(record=True) as caught:
            warnings.simplefilter("always")
            losses = [tr.train_step((x, tgt)) for _ in range(7)]  # 3 eager warm-ups, capture, 3 replays
        assert not [w for w in caught if "capture" in str(w.message)], [str(w.message) for w in caught]
        assert tr.use_graph == graph
        return losses, [p.detach().clone() for p in model.parameters()]

    e = run(False)
    if e is None:
        return
    g = run(True)
    assert np.allclose(g[0], e[0], rtol=2e-4, atol=1e-7), desc
    for a, b in zip(g[1], e[1]):
        # (Adam moves a parameter whose true gradient is zero — e.g. a delay in front of a |.| output — by
        #  lr * noise / (|noise| + eps) = +-lr per step in either run: 7 steps * 2 * lr is the resolution of this
        #  comparison; the losses above are the sharp check)
        assert torch.allclose(a, b, rtol=1e-3, atol=7 * 2 * 1e-3), desc
