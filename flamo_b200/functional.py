"""Coefficient designers and helpers with the signatures of flamo.functional.

Only what the sweep path needs: parameter -> (b, a) second-order
taps, matrix helpers, unit conversions and the impulse of `signal_gallery`.  All designers are
elementwise over the shape of their frequency / gain arguments, follow the dtype of their inputs
(so they run in float64 when the sweep asks for it) and are differentiable.

Math references: RBJ cookbook low/high/band-pass (reference: flamo/functional.py:376-552), first
order-matched shelving and peaking sections of the graphic EQ (functional.py:555-675).
"""
from __future__ import annotations

import math

import numpy as np
import torch


# ------------------------------------------------------------------------------ conversions
def hertz2rad(hertz, fs):
    return hertz / fs * (2 * math.pi)


def rad2hertz(rad, fs):
    return rad * fs / (2 * math.pi)


def db2mag(dB):
    return 10 ** (dB / 20)


def mag2db(mag):
    return 20 * torch.log10(torch.abs(mag))


def get_magnitude(x: torch.Tensor):
    return torch.abs(x)


def skew_matrix(X: torch.Tensor) -> torch.Tensor:
    """Skew-symmetric matrix from the strict upper triangle of X (functional.py:42-56)."""
    U = torch.triu(X, diagonal=1)
    return U - U.mT


def expm_capturable(S: torch.Tensor, max_squarings: int = 20) -> torch.Tensor:
    """exp(S) for square matrices too large for the one-CTA device kernel (libfsweep fsweep_expm_*, n <= 28),
    written so that it never reads anything back to the host and can therefore be captured in a CUDA graph
    (torch.matrix_exp copies the matrix norm to the host to choose its Pade degree).  Scaling and squaring in
    float64: X = S / 2^s with ||X||_1 <= 1/4 (s computed on the device), degree-12 Taylor polynomial evaluated
    Paterson-Stockmeyer style (7 products, remainder 0.25^13/13! ~ 2e-18), then `max_squarings` squarings of which
    the first s are kept (the rest are masked out).  Plain differentiable PyTorch: autograd gives the exact
    derivative of the approximant."""
    dt = S.dtype
    X = S.to(torch.float64)
    n = X.shape[-1]
    norm = X.abs().sum(dim=-2).amax(dim=-1)
    s = torch.clamp(torch.ceil(torch.log2(torch.clamp(norm, min=1e-300) * 4.0)), min=0.0, max=float(max_squarings))
    X = X * torch.exp2(-s)[..., None, None]
    I = torch.eye(n, dtype=torch.float64, device=S.device)
    X2 = X @ X
    X3 = X2 @ X
    X4 = X2 @ X2
    X5 = X3 @ X2
    X6 = X3 @ X3
    c = [1.0]
    for k in range(1, 13):
        c.append(c[-1] / k)
    B0 = c[0] * I + c[1] * X + c[2] * X2 + c[3] * X3 + c[4] * X4 + c[5] * X5
    B1 = c[6] * I + c[7] * X + c[8] * X2 + c[9] * X3 + c[10] * X4 + c[11] * X5
    E = B0 + X6 @ (B1 + c[12] * X6)
    for i in range(max_squarings):
        E = torch.where((s > i)[..., None, None], E @ E, E)
    return E.to(dt)


def _as_tensor(v, like=None, dtype=None, device=None):
    if isinstance(v, torch.Tensor):
        return v
    return torch.as_tensor(v, dtype=dtype if dtype is not None else (like.dtype if like is not None else None),
                           device=device if device is not None else (like.device if like is not None else None))


def _taps(rows):
    return torch.stack(rows, dim=0)


# --------------------------------------------------------------------------------- RBJ biquads
def _rbj_common(fc, fs, dtype, device):
    fc = _as_tensor(fc, dtype=dtype, device=device)
    w = hertz2rad(fc, fs)
    c = torch.cos(w)
    alpha = torch.sin(w) * (math.sqrt(2.0) / 2.0)  # Q = 1/sqrt(2)
    return c, alpha


def lowpass_filter(fc=500.0, gain=0.0, fs: int = 48000, device=None, dtype=torch.float32):
    """Second-order RBJ low-pass, Q = 1/sqrt(2); `gain` in dB scales the numerator.  -> (b, a): (3, *fc.shape)"""
    c, alpha = _rbj_common(fc, fs, dtype, device)
    g = db2mag(_as_tensor(gain, like=c))
    h = (1 - c) / 2
    return g * _taps([h, 1 - c, h]), _taps([1 + alpha, -2 * c, 1 - alpha])


def highpass_filter(fc=10000.0, gain=0.0, fs: int = 48000, device=None, dtype=torch.float32):
    """Second-order RBJ high-pass, Q = 1/sqrt(2)."""
    c, alpha = _rbj_common(fc, fs, dtype, device)
    g = db2mag(_as_tensor(gain, like=c))
    h = (1 + c) / 2
    return g * _taps([h, -(1 + c), h]), _taps([1 + alpha, -2 * c, 1 - alpha])


def bandpass_filter(fc1, fc2, gain=0.0, fs: int = 48000, device=None, dtype=torch.float32):
    """RBJ band-pass (constant 0 dB peak) between fc1 < fc2; bandwidth in octaves log2(fc2/fc1)."""
    fc1 = _as_tensor(fc1, dtype=dtype, device=device)
    fc2 = _as_tensor(fc2, like=fc1)
    w = (hertz2rad(fc1, fs) + hertz2rad(fc2, fs)) / 2
    bw = torch.log2(fc2 / fc1)
    sw = torch.sin(w)
    alpha = sw * torch.sinh(math.log(2.0) / 2 * bw * (w / sw))
    c = torch.cos(w)
    g = db2mag(_as_tensor(gain, like=c))
    return g * _taps([alpha, torch.zeros_like(alpha), -alpha]), _taps([1 + alpha, -2 * c, 1 - alpha])


# ------------------------------------------------------------------- graphic-EQ building blocks
def shelving_filter(fc, gain, type: str = "low", fs: int = 48000, device=None, dtype=torch.float32):
    """Second-order shelving section with linear gain `gain` and crossover fc (functional.py:555-622)."""
    gain = _as_tensor(gain, dtype=dtype, device=device)
    fc = _as_tensor(fc, like=gain).to(gain.dtype)
    t = torch.tan(hertz2rad(fc, fs) / 2)
    t2, g2, g4 = t * t, torch.sqrt(gain), gain ** 0.25
    r2 = math.sqrt(2.0)
    num = _taps([g2 * t2 + r2 * t * g4 + 1, 2 * g2 * t2 - 2, g2 * t2 - r2 * t * g4 + 1]) * g2
    den = _taps([g2 + r2 * t * g4 + t2, 2 * t2 - 2 * g2, g2 - r2 * t * g4 + t2])
    if type == "high":
        return den * gain, num
    return num, den


def peak_filter(fc, gain, Q, fs: int = 48000, device=None, dtype=torch.float32):
    """Second-order peaking section with linear gain `gain` (functional.py:625-675)."""
    gain = _as_tensor(gain, dtype=dtype, device=device)
    fc = _as_tensor(fc, like=gain).to(gain.dtype)
    if isinstance(Q, torch.Tensor):
        Q = Q.to(gain.dtype)
    w = hertz2rad(fc, fs)
    t = torch.tan(w / Q / 2)
    sg = torch.sqrt(gain)
    mid = -2 * sg * torch.cos(w) + torch.zeros_like(gain)
    return _taps([sg + gain * t, mid, sg - gain * t]), _taps([sg + t, mid, sg - t])


# ---------------------------------------------------------------------------- fixed matrices
class HadamardMatrix:
    """Orthonormal Sylvester-Hadamard matrix of size N (power of two) applied as a constant map."""

    def __init__(self, N, device=None, dtype=torch.float32):
        n = int(N)
        if n & (n - 1):
            raise ValueError("Hadamard matrix size must be a power of two")
        H = torch.ones(1, 1, dtype=dtype, device=device)
        while H.shape[0] < n:
            H = torch.cat((torch.cat((H, H), 1), torch.cat((H, -H), 1)), 0)
        self.H = H / math.sqrt(n)

    def __call__(self, x):
        return self.H.to(x.dtype if isinstance(x, torch.Tensor) else self.H.dtype)


class RotationMatrix:
    """Rotation matrix built from ONE angle by Kronecker squaring, with the reference's semantics
    (flamo/functional.py:96-158): the angle is clamped to [min_angle, max_angle] (default [0, pi/4]) and the 2 x 2
    rotation is squared `iter` times with torch.kron (iter = log2(N) - 1 when None), i.e. the result is N x N only for
    N = 2 and N = 4.  Unlike the reference's (fill_diagonal_ with a tensor) this one is differentiable."""

    def __init__(self, N, min_angle=0, max_angle=math.pi / 4, iter=None, device=None, dtype=torch.float32):
        self.N, self.min_angle, self.max_angle, self.iter = int(N), min_angle, max_angle, iter
        self.device, self.dtype = device, dtype

    def __call__(self, theta):
        th = theta[0] if isinstance(theta, (list, tuple)) else theta
        if self.min_angle is not None and self.min_angle > self.max_angle:
            th = torch.full_like(th, self.max_angle)  # torch.clamp with min > max returns max
        else:
            th = torch.clamp(th, self.min_angle, self.max_angle)
        c, s = torch.cos(th), torch.sin(th)
        out = torch.stack((torch.stack((c, s)), torch.stack((-s, c))))
        iters = self.iter if self.iter is not None else int(math.log2(self.N)) - 1
        for _ in range(iters):
            out = torch.kron(out, out)
        return out


# -------------------------------------------------------------------------------- test signals
def signal_gallery(batch_size: int, n_samples: int, n: int, signal_type: str = "impulse", fs: int = 48000,
                   rate: float = 1.0, reference=None, device=None, dtype=torch.float32):
    """Subset of flamo.functional.signal_gallery (functional.py:164-270) used on the sweep path."""
    if signal_type == "impulse":
        x = torch.zeros(batch_size, n_samples, n, dtype=dtype, device=device)
        x[:, 0, :] = 1
        return x
    if signal_type in ("wgn", "noise"):
        return torch.randn((batch_size, n_samples, n), device=device, dtype=dtype)
    if signal_type == "sine":
        t = torch.linspace(0, n_samples / fs, n_samples, dtype=dtype, device=device)
        return torch.sin(2 * np.pi * rate / fs * t).unsqueeze(-1).expand(batch_size, n_samples, n)
    if signal_type == "exp":
        t = torch.arange(n_samples, dtype=dtype, device=device)
        return torch.exp(-rate * t / fs).unsqueeze(-1).expand(batch_size, n_samples, n)
    if signal_type == "sweep":
        # linear chirp 20 Hz -> 20 kHz over the whole signal: scipy.signal.chirp(t, 20, t[-1], 20000) with phi = 0,
        # evaluated in `dtype` in scipy's operation order (the reference hands it a `dtype` time axis)
        t = torch.linspace(0, n_samples / fs - 1 / fs, n_samples, dtype=dtype)
        beta = (20000.0 - 20.0) / t[-1]
        x = torch.cos(2 * np.pi * (20.0 * t + 0.5 * beta * t * t))
        return x.to(device).unsqueeze(-1).expand(batch_size, n_samples, n)
    if signal_type == "reference":
        return torch.as_tensor(reference, dtype=dtype, device=device).expand(batch_size, n_samples, n)
    raise ValueError(f"Signal type {signal_type} not recognized.")
