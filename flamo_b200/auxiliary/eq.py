"""Graphic-EQ designer with the interface of flamo.auxiliary.eq (reference: eq.py:8-111).

`geq` is vectorised over any trailing shape of `gain_db` — the reference calls it once per
(output, input) channel pair from a Python double loop (processor/dsp.py:2576-2585); here one call
produces the taps of every pair.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from ..functional import db2mag, peak_filter, shelving_filter


def octave_bands(interval: int = 1, start_freq: float = 31.25, end_freq: float = 16000.0):
    """Centre frequencies start*2^(i/interval), i = 1, 2, ... up to and including the first >= end."""
    out, f = [], start_freq
    while f < end_freq:
        f = f * np.power(2, 1 / interval)
        out.append(f)
    return out


def eq_freqs(interval: int = 1, start_freq: float = 31.25, end_freq: float = 16000.0, device="cpu",
             dtype=torch.float32):
    centre = torch.tensor(octave_bands(interval, start_freq, end_freq), device=device, dtype=dtype)
    half = np.power(2, 1 / interval / 2)
    cross = torch.tensor([centre[0] / half, centre[-1] * half], device=device, dtype=dtype)
    return centre, cross


def geq(center_freq, shelving_freq, R, gain_db, fs: int = 48000, device="cpu", dtype=None):
    """Cascade taps of the graphic EQ.  gain_db: (n_bands, ...) command gains in dB with
    n_bands = len(center_freq) + 3: [broadband, low shelf, peaks..., high shelf].
    Returns (b, a), each (3, n_bands, ...)."""
    dtype = dtype or gain_db.dtype
    gain_db = gain_db.to(dtype)
    n = len(center_freq) + len(shelving_freq) + 1
    if gain_db.shape[0] != n:
        raise AssertionError("The number of gains must be equal to the number of frequencies.")
    g = db2mag(gain_db)
    cf = torch.as_tensor(center_freq).to(dtype=dtype, device=g.device)  # no-op when already there
    sf = torch.as_tensor(shelving_freq).to(dtype=dtype, device=g.device)
    Rf = float(R)  # host scalar: keeps the designer free of host->device copies
    Q = math.sqrt(Rf) / (Rf - 1)
    tail = (1,) * (g.dim() - 1)
    zero, one = torch.zeros_like(g[0]), torch.ones_like(g[0])
    b0 = torch.stack((g[0], zero, zero))
    a0 = torch.stack((one, zero, zero))
    bl, al = shelving_filter(sf[0], g[1], "low", fs=fs)
    bh, ah = shelving_filter(sf[1], g[-1], "high", fs=fs)
    bp, ap = peak_filter(cf.view(-1, *tail), g[2:-1], Q, fs=fs)  # (3, n-3, ...)
    b = torch.cat((b0.unsqueeze(1), bl.unsqueeze(1), bp, bh.unsqueeze(1)), dim=1)
    a = torch.cat((a0.unsqueeze(1), al.unsqueeze(1), ap, ah.unsqueeze(1)), dim=1)
    return b, a
