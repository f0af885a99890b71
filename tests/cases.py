"""Parity cases shared by the golden generator, the oracle tests and the GPU tests.

Each case = a model description (flamo_b200.workloads format) + nfft, alias decay, batch,
trailing columns and the seed under which the REFERENCE constructors drew the raw parameters
(the drawn values are stored in the golden file; nothing here depends on RNG streams).
"""
import math

import numpy as np
import torch

from flamo_b200 import workloads as W

FS = W.FS


def make_input(B, M, N, C=None, dtype=torch.complex128):
    """Deterministic closed-form bin-domain input (no RNG): unit-ish modulus, all bins distinct."""
    k = torch.arange(M, dtype=torch.float64).view(1, M, 1)
    n = torch.arange(N, dtype=torch.float64).view(1, 1, N)
    b = torch.arange(B, dtype=torch.float64).view(B, 1, 1)
    ph = 0.37 * k + 1.3 * n + 0.7 * b
    mag = 1.0 + 0.5 * torch.cos(0.011 * k * (n + 1) + 0.3 * b)
    X = torch.polar(mag, ph)
    if C is not None:
        c = torch.arange(C, dtype=torch.float64).view(1, 1, 1, C)
        X = X.unsqueeze(-1) * torch.polar(1.0 / (1.0 + c), 0.9 * c)
    return X.to(dtype)


def golden_loss(Y):
    """mean((sum_ch |Y| - 1)^2): the colorless-FDN criterion (optimize/loss.py:90-103 with target 1)."""
    a = torch.abs(Y)
    s = a.sum(dim=2)
    return torch.mean((s - 1.0) ** 2)


def select_bins(M):
    stride = max(1, M // 509)
    idx = np.unique(np.concatenate([np.arange(min(M, 24)), np.arange(0, M, stride), [M - 1]]))
    return idx.astype(np.int64)


def probe_fdn_desc():
    """examples/e10_probe.py:16-84: 4x4 FDN, m=[101,157,211,263], A = Q diag(g^m), identity map."""
    g = torch.Generator().manual_seed(130709)
    N = 4
    m = torch.tensor([101, 157, 211, 263])
    R = torch.randn(N, N, generator=g, dtype=torch.float64)
    U = R.triu(1)
    A = torch.matrix_exp(U - U.T) @ torch.diag(0.999 ** m.double())
    b = torch.randn(N, 1, generator=g, dtype=torch.float64)
    c = torch.randn(1, N, generator=g, dtype=torch.float64)
    return (
        "Series",
        [
            ("Gain", dict(size=(N, 1), requires_grad=True), {"assign": b.tolist()}),
            ("Recursion",
             ("parallelDelay", dict(size=(N,), max_len=263, isint=True, requires_grad=False),
              {"delay_samples": m.tolist()}),
             ("Matrix", dict(size=(N, N), matrix_type="random", requires_grad=True), {"assign": A.tolist()})),
            ("Gain", dict(size=(1, N), requires_grad=True), {"assign": c.tolist()}),
        ],
        ["input_gain", "feedback_loop", "output_gain"],
    )


def _case(desc, nfft, alias=30.0, B=1, C=None, seed=0, grads=True):
    return dict(desc=desc, nfft=nfft, alias=alias, B=B, C=C, seed=seed, grads=grads)


def _svf(ft, size=(3, 2), K=2):
    return ("SVF", dict(size=size, n_sections=K, filter_type=ft, fs=FS, requires_grad=True))


def _sos_coeffs(K, shape, seed):
    """Deterministic stable second-order sections [b0,b1,b2,a0,a1,a2] of shape (K, 6, *shape)."""
    g = torch.Generator().manual_seed(seed)
    b = torch.rand(K, 3, *shape, generator=g, dtype=torch.float64) * 2 - 1
    a0 = 0.8 + 0.7 * torch.rand(K, 1, *shape, generator=g, dtype=torch.float64)
    r = 0.3 + 0.6 * torch.rand(K, 1, *shape, generator=g, dtype=torch.float64)  # pole radius
    th = math.pi * torch.rand(K, 1, *shape, generator=g, dtype=torch.float64)
    a = torch.cat((a0, -2 * r * torch.cos(th) * a0, r * r * a0), dim=1)
    return torch.cat((b, a), dim=1).tolist()


CASES = {
    # --- the five BASELINE configs (full nfft where the reference fits in RAM/time) ---------
    "cfg1_biquad_full": _case(W.biquad(), 96000, seed=130709),
    "cfg2_fdn8_full": _case(W.fdn(8), 96000, seed=130709),
    "cfg3_geq16_small": _case(W.geq(16, 16, 3), 2048, seed=130710),
    "cfg4_active_full": _case(W.active_acoustics(), 96000, seed=130297),
    "cfg5_fdn64_small": _case(W.fdn(64), 2048, B=2, seed=0),
    # --- variants -----------------------------------------------------------------------
    "fdn6_example": _case(W.fdn(6), 8192, seed=130709),
    "fdn8_lossless": _case(W.fdn(8), 4096, alias=0.0, seed=1),
    "fdn8_batch3": _case(W.fdn(8), 4096, B=3, seed=2),
    "fdn16": _case(W.fdn(16, delays=[593, 641, 701, 769, 839, 907, 977, 1051, 1117, 1201, 1279, 1361, 1447, 1531,
                                      1613, 1699]), 4096, seed=3),
    "fdn32": _case(W.fdn(32, delays=list(range(601, 601 + 32 * 37, 37))), 2048, seed=4),
    "fdn8_fracdelay": _case(W.fdn(8, isint=False), 4096, seed=5),
    "biquad_lowpass": _case(W.biquad(2, 3, 3, "lowpass"), 4096, seed=6),
    "biquad_bandpass": _case(W.biquad(2, 2, 2, "bandpass"), 4096, seed=7),
    "biquad_highpass_noalias": _case(W.biquad(2, 1, 2, "highpass"), 4096, alias=0.0, seed=8),
    "pbiquad_lowpass": _case(("parallelBiquad", dict(size=(3,), n_sections=2, filter_type="lowpass", fs=FS,
                                                     requires_grad=True)), 4096, seed=9),
    "svf_general": _case(_svf(None), 4096, seed=10),
    "svf_lowpass": _case(_svf("lowpass"), 2048, seed=11),
    "svf_highpass": _case(_svf("highpass"), 2048, seed=12),
    "svf_bandpass": _case(_svf("bandpass"), 2048, seed=13),
    "svf_lowshelf": _case(_svf("lowshelf"), 2048, seed=14),
    "svf_highshelf": _case(_svf("highshelf"), 2048, seed=15),
    "svf_peaking": _case(_svf("peaking"), 2048, seed=16),
    "svf_notch": _case(_svf("notch"), 2048, seed=17),
    "psvf_general": _case(("parallelSVF", dict(size=(4,), n_sections=3, filter_type=None, fs=FS,
                                               requires_grad=True)), 2048, seed=18),
    "geq_oct1": _case(("GEQ", dict(size=(2, 3), octave_interval=1, fs=FS, requires_grad=True)), 4096, seed=19),
    "geq_oct3": _case(("GEQ", dict(size=(3, 2), octave_interval=3, fs=FS, requires_grad=True)), 8192, seed=20),
    "pgeq_oct1": _case(("parallelGEQ", dict(size=(5,), octave_interval=1, fs=FS, requires_grad=True)), 4096, seed=21),
    "delay_mimo_frac": _case(("Delay", dict(size=(3, 2), max_len=1500, isint=False, fs=FS, requires_grad=True)),
                             4096, seed=22),
    "delay_mimo_int": _case(("Delay", dict(size=(2, 4), max_len=1500, isint=True, fs=FS, requires_grad=False)),
                            4096, seed=23, grads=False),
    "pdelay_frac": _case(("parallelDelay", dict(size=(5,), max_len=900, isint=False, fs=FS, requires_grad=True)),
                         4096, seed=24),
    "series_mixed": _case(("Series", [
        ("Gain", dict(size=(4, 2), requires_grad=True)),
        ("parallelBiquad", dict(size=(4,), n_sections=2, filter_type="lowpass", fs=FS, requires_grad=True)),
        ("parallelDelay", dict(size=(4,), max_len=300, isint=True, fs=FS, requires_grad=False)),
        ("Matrix", dict(size=(4, 4), matrix_type="orthogonal", requires_grad=True)),
        ("parallelGain", dict(size=(4,), requires_grad=True)),
        ("Gain", dict(size=(3, 4), requires_grad=True)),
    ]), 4096, B=2, seed=25),
    "series_trailing_cols": _case(("Series", [
        ("Gain", dict(size=(3, 3), requires_grad=True)),
        ("parallelDelay", dict(size=(3,), max_len=200, isint=True, fs=FS, requires_grad=False)),
        ("Biquad", dict(size=(2, 3), n_sections=1, filter_type="lowpass", fs=FS, requires_grad=True)),
    ]), 2048, B=2, C=3, seed=26),
    "recursion_rect": _case(("Series", [
        ("Gain", dict(size=(3, 2), requires_grad=True)),
        ("Recursion",
         ("Series", [("Gain", dict(size=(5, 3), requires_grad=True), ),
                     ("parallelDelay", dict(size=(5,), max_len=700, isint=True, fs=FS, requires_grad=False))]),
         ("Series", [("Delay", dict(size=(3, 5), max_len=400, isint=True, fs=FS, requires_grad=False)),
                     ("Gain", dict(size=(3, 3), requires_grad=True),
                      {"assign": (0.12 * np.eye(3) + 0.05).tolist()})])),
        ("Gain", dict(size=(2, 5), requires_grad=True)),
    ]), 4096, B=2, seed=27),
    "recursion_filters": _case(("Recursion",
                                ("Series", [("parallelDelay", dict(size=(4,), max_len=800, isint=True, fs=FS,
                                                                  requires_grad=False)),
                                            ("parallelBiquad", dict(size=(4,), n_sections=1, filter_type="lowpass",
                                                                    fs=FS, requires_grad=True)),
                                            ("parallelGain", dict(size=(4,), requires_grad=True),
                                             {"assign": [0.7, 0.65, 0.6, 0.55]})]),
                                ("Matrix", dict(size=(4, 4), matrix_type="orthogonal", requires_grad=True),
                                 )), 4096, seed=28),
    # --- SOSFilter, HouseholderMatrix, GainDelay, Parallel ---------------
    "sosfilter": _case(("SOSFilter", dict(size=(2, 3), n_sections=2, fs=FS),
                        {"assign": _sos_coeffs(2, (2, 3), 31)}), 4096, seed=31, grads=False),  # the reference's a0
    # normalisation map writes in place: its own autograd cannot differentiate it (dsp.py:1857-1862)
    "psosfilter_raw_a0": _case(("parallelSOSFilter", dict(size=(4,), n_sections=3, fs=FS, normalize_a0=False),
                                {"assign": _sos_coeffs(3, (4,), 32), "requires_grad": True}), 2048, B=2, seed=32),
    "householder_series": _case(("Series", [
        ("Gain", dict(size=(4, 2), requires_grad=True)),
        ("HouseholderMatrix", dict(size=(4, 4), requires_grad=True)),
        ("Gain", dict(size=(3, 4), requires_grad=True)),
    ]), 2048, B=2, seed=33),
    "fdn4_householder": _case(("Series", [
        ("Gain", dict(size=(4, 1), requires_grad=True)),
        ("Recursion",
         ("Series", [("parallelDelay", dict(size=(4,), max_len=700, isint=True, fs=FS, requires_grad=False),
                      {"delay_samples": [241, 331, 419, 547]}),
                     ("parallelGain", dict(size=(4,), requires_grad=True), {"assign": [0.9, 0.88, 0.86, 0.84]})]),
         ("HouseholderMatrix", dict(size=(4, 4), requires_grad=True))),
        ("Gain", dict(size=(1, 4), requires_grad=True)),
    ], ["input_gain", "feedback_loop", "output_gain"]), 4096, seed=34),
    "gaindelay_frac": _case(("GainDelay", dict(size=(3, 2), max_len=900, isint=False, fs=FS, requires_grad=True)),
                            4096, seed=35),
    "gaindelay_int": _case(("GainDelay", dict(size=(2, 3), max_len=900, isint=True, fs=FS, requires_grad=False)),
                           2048, seed=36, grads=False),
    "pgaindelay_frac": _case(("parallelGainDelay", dict(size=(4,), max_len=600, isint=False, fs=FS,
                                                       requires_grad=True)), 2048, B=2, seed=37),
    "parallel_sum": _case(("Parallel",
                           ("Series", [("Gain", dict(size=(3, 2), requires_grad=True)),
                                       ("parallelDelay", dict(size=(3,), max_len=300, isint=True, fs=FS,
                                                              requires_grad=False))]),
                           ("Biquad", dict(size=(3, 2), n_sections=1, filter_type="lowpass", fs=FS, requires_grad=True)),
                           True), 2048, B=2, seed=38),
    "parallel_cat_in_series": _case(("Series", [
        ("Gain", dict(size=(2, 2), requires_grad=True)),
        ("Parallel",
         ("Gain", dict(size=(2, 2), requires_grad=True)),
         ("Series", [("Gain", dict(size=(3, 2), requires_grad=True)), ("parallelGain", dict(size=(3,), requires_grad=True))]),
         False),
        ("Gain", dict(size=(1, 5), requires_grad=True)),
    ]), 2048, seed=39),
    "fir_filter": _case(("Filter", dict(size=(24, 2, 3), requires_grad=True)), 2048, seed=29),
    "pfir_filter": _case(("parallelFilter", dict(size=(40, 3), requires_grad=True)), 2048, seed=30),
}

