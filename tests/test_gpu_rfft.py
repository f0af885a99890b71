"""GPU: libfsweep's two-launch four-step FFT (fsweep_rfft, the input transform of a step) against torch.fft.rfft in
float64 — dsp.FFT / dsp.FFTAntiAlias semantics (reference flamo/processor/dsp.py:69-93, :122-163): zero padding and
cropping to nfft, the three norms, batch and channel axes, the anti-alias envelope."""
import pytest
import torch

from flamo_b200 import _lib, sweep
from flamo_b200.processor import dsp

pytestmark = pytest.mark.gpu

TOL = 2e-6  # of the largest bin magnitude: float32 transform, float64 reference


def _check(x, nfft, norm="backward", envelope=None):
    n0 = sweep.launch_count
    X = sweep.rfft(x, nfft, norm, envelope, force=True)
    assert sweep.launch_count - n0 == 2, "the libfsweep FFT did not run"
    xr = x.double()
    if envelope is not None:
        xr = xr * envelope.double().view(1, -1, 1)
    R = torch.fft.rfft(xr, n=nfft, dim=1, norm=norm)
    assert X.shape == R.shape and X.dtype == torch.complex64
    err = float((X.to(torch.complex128) - R).abs().max() / R.abs().max())
    assert err < TOL, err


@pytest.mark.parametrize("nfft", [512, 1000, 2048, 4096, 6000, 48000, 96000, 192000, 384000])
def test_sizes(nfft):
    torch.manual_seed(nfft)
    _check(torch.randn(1, nfft, 1, device="cuda"), nfft)


@pytest.mark.parametrize("norm", ["backward", "forward", "ortho"])
def test_norms_batch_channels_padding_cropping(norm):
    torch.manual_seed(3)
    _check(torch.randn(3, 4001, 5, device="cuda"), 6000, norm)   # zero padded, odd length
    _check(torch.randn(2, 7000, 3, device="cuda"), 6000, norm)   # cropped
    _check(torch.randn(2, 6000, 3, device="cuda")[:, :, 1:], 6000, norm)  # non-contiguous view


def test_impulse_and_nyquist():
    x = torch.zeros(1, 96000, 1, device="cuda")
    x[0, 0, 0] = 1.0
    X = sweep.rfft(x, 96000)
    assert float((X - 1).abs().max()) < 1e-6
    x = torch.ones(1, 96000, 1, device="cuda")
    x[0, 1::2, 0] = -1.0  # all the energy in the Nyquist bin
    X = sweep.rfft(x, 96000)
    assert abs(float(X[0, -1, 0].real) - 96000) < 0.1 and float(X[0, :-1, 0].abs().max()) < 0.05


def test_modules_use_it_and_refuse_what_cufft_must_do():
    torch.manual_seed(0)
    x = torch.randn(2, 32768, 2, device="cuda")
    n0 = sweep.launch_count
    X = dsp.FFT(32768)(x)
    assert sweep.launch_count - n0 == 2
    assert torch.allclose(X, torch.fft.rfft(x, n=32768, dim=1), atol=1e-3)
    aa = dsp.FFTAntiAlias(32768, alias_decay_db=30.0, device="cuda")
    n0 = sweep.launch_count
    Xa = aa(x)
    assert sweep.launch_count - n0 == 2
    Ra = torch.fft.rfft(x.double() * aa.alias_envelope.double().view(1, -1, 1), n=32768, dim=1)
    assert float((Xa.to(torch.complex128) - Ra).abs().max() / Ra.abs().max()) < TOL
    n0 = sweep.launch_count  # small sizes (one cuFFT launch) and big batches stay with cuFFT
    dsp.FFT(4096)(x)
    dsp.FFT(32768)(torch.randn(16, 32768, 4, device="cuda"))
    assert sweep.launch_count == n0
    # not supported by the kernels: cuFFT, same result type
    assert not _lib.lib().fsweep_rfft_supported(4098 * 2 + 1)
    for nfft in (257 * 2, 2 * 1031 * 4, 300):  # a prime above the radix bound, a small size
        n0 = sweep.launch_count
        Y = dsp.FFT(nfft)(x)
        assert sweep.launch_count == n0 and Y.shape == (2, nfft // 2 + 1, 2)
    # float64 signals and signals that need a gradient stay on torch.fft
    n0 = sweep.launch_count
    xg = x.clone().requires_grad_(True)
    dsp.FFT(32768)(xg).abs().sum().backward()
    assert sweep.launch_count == n0 and xg.grad is not None
    assert dsp.FFT(32768)(x.double()).dtype == torch.complex128


def test_capture():
    x = torch.randn(1, 96000, 1, device="cuda")
    ref = sweep.rfft(x, 96000).clone()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        sweep.rfft(x, 96000)
        with torch.cuda.graph(g):
            X = sweep.rfft(x, 96000)
    torch.cuda.current_stream().wait_stream(s)
    x.mul_(2.0)
    g.replay()
    torch.cuda.synchronize()
    assert torch.allclose(X, 2 * ref, atol=1e-3)
