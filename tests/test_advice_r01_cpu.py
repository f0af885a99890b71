"""CPU regression tests for the round-1 advisor findings (host logic through the ABI emulator):
eager modules that change the bin axis inside a Series / a Sequential output layer, sparsity_loss on a
HouseholderMatrix, output transforms that merely resemble |x|, frozen parameters and the captured-step key."""
import numpy as np
import pytest
import torch

from oracle import flamo_oracle as O

pytestmark = pytest.mark.usefixtures("emulated_backend")

NFFT, M = 64, 33


def _mods():
    from flamo_b200.processor import dsp, system

    return dsp, system


def test_series_with_ifft_keeps_the_modules_own_output_shape():
    """Series(Gain, iFFT): the eager module's output (B, nfft, N) is returned as it is (reference system.py:279-301
    runs the modules in sequence); round 1 reshaped it back to the bin count and crashed / scrambled."""
    dsp, system = _mods()
    torch.manual_seed(0)
    g = dsp.Gain(size=(2, 2), nfft=NFFT, dtype=torch.float64)
    ser = system.Series(g, dsp.iFFT(NFFT, dtype=torch.float64))
    X = torch.randn(1, M, 2, dtype=torch.complex128)
    y = ser(X)
    W = g.param.detach().to(torch.complex128)
    ref = torch.fft.irfft(torch.einsum("mn,bfn->bfm", W, X), n=NFFT, dim=1)
    assert y.shape == (1, NFFT, 2) and not y.is_complex()
    assert torch.allclose(y, ref, atol=1e-12)


def test_shell_with_sequential_output_layer():
    """Shell(core, FFT, nn.Sequential(iFFT, Transform)): _wrap turns the output layer into a Series whose modules
    are all eager and change dim 1."""
    dsp, system = _mods()
    torch.manual_seed(1)
    core = system.Series(dsp.Gain(size=(1, 2), nfft=NFFT, dtype=torch.float64))
    out = torch.nn.Sequential(dsp.iFFT(NFFT, dtype=torch.float64), dsp.Transform(lambda v: 2.0 * v, dtype=torch.float64))
    with pytest.warns(UserWarning):
        model = system.Shell(core, dsp.FFT(NFFT, dtype=torch.float64), out)
    x = torch.zeros(1, NFFT, 2, dtype=torch.float64)
    x[0, 0, :] = 1.0
    x[0, 5, 1] = -0.25
    y = model(x)
    W = core[0].param.detach()
    ref = 2.0 * torch.einsum("mn,btn->btm", W, x)
    assert y.shape == (1, NFFT, 1)
    assert torch.allclose(y, ref, atol=1e-12)


def test_sweep_after_a_non_bin_module_raises_clearly():
    dsp, system = _mods()
    ser = system.Series(dsp.Gain(size=(1, 1), nfft=NFFT, dtype=torch.float64), dsp.iFFT(NFFT, dtype=torch.float64),
                        dsp.Gain(size=(1, 1), nfft=NFFT, dtype=torch.float64))
    with pytest.raises((TypeError, ValueError)):
        ser(torch.randn(1, M, 1, dtype=torch.complex128))


def test_sparsity_loss_householder_matches_reference_formula():
    """reference loss.py:53-55: for a HouseholderMatrix the loss is taken on A = I - 2 u u^T, not on the (N, 1)
    parameter (round 1: N = 1 in the denominator -> -inf)."""
    from flamo_b200.optimize.loss import sparsity_loss

    dsp, system = _mods()
    torch.manual_seed(2)
    N = 4
    fb = system.Series()
    mm = dsp.HouseholderMatrix(size=(N, N), nfft=NFFT, requires_grad=True, dtype=torch.float64)
    fb.add_module("mixing_matrix", mm)
    fb._refresh()
    ff = dsp.parallelDelay(size=(N,), max_len=20, nfft=NFFT, isint=True, dtype=torch.float64)
    core = system.Series()
    core.add_module("feedback_loop", system.Recursion(ff, fb))
    core._refresh()
    model = system.Shell(core, dsp.FFT(NFFT, dtype=torch.float64))
    loss = sparsity_loss()(None, None, model)
    u = (mm.param / torch.norm(mm.param, dim=0, keepdim=True)).detach()
    A = torch.eye(N, dtype=torch.float64) - 2 * u @ u.T
    ref = -(A.abs().sum() - N * np.sqrt(N)) / (N * (np.sqrt(N) - 1))
    assert torch.isfinite(loss)
    assert abs(float(loss) - float(ref)) < 1e-12
    loss.backward()
    assert mm.param.grad is not None and torch.isfinite(mm.param.grad).all()


@pytest.mark.parametrize("transform,shape", [
    (lambda v: torch.abs(v[:, :10]), (1, 10, 2)),
    (lambda v: torch.clamp(torch.abs(v), max=0.5), (1, M, 2)),
    (lambda v: torch.abs(v) + 0.25, (1, M, 2)),
])
def test_output_transforms_that_only_resemble_abs_are_run_as_written(transform, shape):
    """The |.| epilogue must not replace a transform that crops, clamps or offsets (round 1 decided on a 4-bin probe
    and silently dropped the real transform)."""
    dsp, system = _mods()
    torch.manual_seed(3)
    core = system.Series(dsp.Gain(size=(2, 1), nfft=NFFT, dtype=torch.float64))
    model = system.Shell(core, dsp.FFT(NFFT, dtype=torch.float64), dsp.Transform(transform, dtype=torch.float64))
    x = torch.zeros(1, NFFT, 1, dtype=torch.float64)
    x[0, 0] = 1.0
    x[0, 3] = 0.7
    y = model(x)
    Y = torch.einsum("mn,bfn->bfm", core[0].param.detach().to(torch.complex128), torch.fft.rfft(x, n=NFFT, dim=1))
    assert y.shape == shape
    assert torch.allclose(y, transform(Y), atol=1e-12)
    # ... and the Trainer's fused-criterion route refuses it too
    assert model.forward_loss(x, torch.ones(1, M, 2, dtype=torch.float64), 1) is None


def test_plain_abs_transform_is_still_fused():
    from flamo_b200 import sweep

    dsp, system = _mods()
    core = system.Series(dsp.Gain(size=(2, 1), nfft=NFFT, dtype=torch.float64))
    model = system.Shell(core, dsp.FFT(NFFT, dtype=torch.float64),
                         dsp.Transform(lambda v: torch.abs(v), dtype=torch.float64))
    x = torch.zeros(1, NFFT, 1, dtype=torch.float64)
    x[0, 0] = 1.0
    model(x)  # verification pass
    n0 = sweep.launch_count
    y = model(x)
    assert not y.is_complex() and y.shape == (1, M, 2)
    assert model.forward_loss(x, torch.ones(1, M, 2, dtype=torch.float64), 1) is not None
    assert sweep.launch_count > n0


def test_graph_key_tracks_frozen_parameter_versions():
    """assign_value on a frozen module must invalidate a captured step (its mapped coefficients are baked in)."""
    from flamo_b200.optimize.trainer import Trainer

    dsp, system = _mods()
    d = dsp.parallelDelay(size=(2,), max_len=20, nfft=NFFT, isint=True, dtype=torch.float64)
    core = system.Series(dsp.Gain(size=(2, 1), nfft=NFFT, requires_grad=True, dtype=torch.float64), d)
    model = system.Shell(core, dsp.FFT(NFFT, dtype=torch.float64))
    tr = Trainer(model, max_epochs=1, lr=1e-3, log=False, device="cpu")
    x, t = torch.zeros(1, NFFT, 1, dtype=torch.float64), torch.zeros(1, M, 2, dtype=torch.float64)
    k0 = tr._graph_key(x, t)
    d.assign_value(torch.tensor([0.0003, 0.0004], dtype=torch.float64))
    assert tr._graph_key(x, t) != k0
