"""CPU: the float32 numerical choices of the kernels restated in numpy float32 next to the
naive float32 formulas the reference's own float32 path amounts to, both against float64 truth.  Not a test of the
kernels (tests/test_gpu_parity.py is) — a check that the FORMULAS the kernels use are the reason float32 meets the
1e-4 bar, and by what margin, so that nobody "simplifies" them away."""
import numpy as np
import pytest
import torch

from flamo_b200 import sweep
from flamo_b200.functional import peak_filter, shelving_filter

f32, f64 = np.float32, np.float64


def _taylor_eval_f32(packed, k, nfft, gamma):
    """B(w) and A(w) from the packed Taylor blocks exactly as fsweep_kernels.cuh make_ctx / the section op do it:
    v = w -/+ 1 with w - 1 = (g-1) - 2 g sin^2(omega/2) - j g sin(omega) (and the mirrored form around -1)."""
    fr = (2.0 * k / nfft)  # omega / pi, float64 like the kernel
    s, co = np.sin(np.pi * fr).astype(f32), np.cos(np.pi * fr).astype(f32)
    sh, ch = np.sin(np.pi * fr / 2).astype(f32), np.cos(np.pi * fr / 2).astype(f32)
    g, gm1 = f32(gamma), f32(gamma - 1.0)
    plus = co >= 0
    vr = np.where(plus, gm1 - f32(2) * g * sh * sh, f32(2) * g * ch * ch - gm1).astype(f32)
    vi = (-g * s).astype(f32)
    v = (vr + 1j * vi).astype(np.complex64)
    blk = np.where(plus[:, None], packed[0][None, :], packed[1][None, :]).astype(f32)  # (nb, 8)
    B = blk[:, 0] + blk[:, 1] * v + blk[:, 2] * v * v
    A = blk[:, 4] + blk[:, 5] * v + blk[:, 6] * v * v
    return B.astype(np.complex64), A.astype(np.complex64)


@pytest.mark.parametrize("kind", ["lowshelf", "peak"])
def test_taylor_blocks_remove_the_dc_cancellation(kind):
    """A 31 Hz section at fs = 48 kHz, nfft = 192000 (config 3's lowest band): around DC, b0 + b1 w + b2 w^2 is a
    ~1e-5 remainder of O(1) terms — float32 loses it (error ~1e-2 of |H|), the Taylor form around w = 1 does not."""
    fs, nfft, alias = 48000, 192000, 30.0
    gamma = 10 ** (-alias / nfft / 20)
    fc, gain_db = torch.tensor(31.25, dtype=torch.float64), torch.tensor(4.0, dtype=torch.float64)
    if kind == "lowshelf":
        b, a = shelving_filter(fc, 10 ** (gain_db / 20), "low", fs=fs, dtype=torch.float64)
    else:
        b, a = peak_filter(fc, 10 ** (gain_db / 20), torch.tensor(1.4, dtype=torch.float64), fs=fs, dtype=torch.float64)
    b, a = b.reshape(3).numpy(), a.reshape(3).numpy()
    k = np.arange(0, 400, dtype=f64)  # DC ... 100 Hz
    w = gamma * np.exp(-2j * np.pi * k / nfft)
    H = (b[0] + b[1] * w + b[2] * w * w) / (a[0] + a[1] * w + a[2] * w * w)  # float64 truth
    # naive float32: taps and w rounded to float32, polynomial in float32 (what a float32 rfft of the taps amounts to)
    w32 = w.astype(np.complex64)
    b32, a32 = b.astype(f32), a.astype(f32)
    Hn = (b32[0] + b32[1] * w32 + b32[2] * w32 * w32) / (a32[0] + a32[1] * w32 + a32[2] * w32 * w32)
    # kernel formula: blocks formed in float64 by the host (sweep.pack_sections), rounded once, evaluated in float32
    tb = torch.tensor(b).view(3, 1, 1)
    ta = torch.tensor(a).view(3, 1, 1)
    packed = sweep.pack_sections(tb, ta, True, torch.float32)[0, 0].numpy()  # (2, 8)
    B, A = _taylor_eval_f32(packed, k, nfft, gamma)
    Ht = B / A
    err_naive = np.abs(Hn - H).max() / np.abs(H).max()
    err_taylor = np.abs(Ht - H).max() / np.abs(H).max()
    assert err_taylor < 2e-6, err_taylor
    assert err_naive > 100 * err_taylor, (err_naive, err_taylor)


def test_integer_phase_is_exact_where_float32_omega_times_delay_is_not():
    """parallelDelay(2287) at nfft = 96000 (config 2): exp(-j omega_k m) with omega_k * m formed in float32 is off by
    ~1e-3 at the high bins (SURVEY.md finding 4); the kernels reduce (k * m) mod nfft in integers first."""
    nfft, m = 96000, 2287
    k = np.arange(0, nfft // 2 + 1, dtype=np.int64)
    truth = np.exp(-2j * np.pi * ((k * m) % nfft) / nfft)
    omega32 = (f32(2) * f32(np.pi) * k.astype(f32) / f32(nfft)).astype(f32)
    naive = np.exp(-1j * (omega32 * f32(m)).astype(f32).astype(f64))
    idx = (k * m) % nfft  # exact integers
    fr = (2.0 * idx / nfft).astype(f32)  # in [0, 2): one rounding of a reduced argument
    exact = (np.cos(np.pi * fr.astype(f64)) - 1j * np.sin(np.pi * fr.astype(f64)))
    assert np.abs(naive - truth).max() > 3e-4
    assert np.abs(exact - truth).max() < 5e-7


def test_product_of_ratios_survives_where_separate_products_underflow():
    """30 third-octave peak sections (config 3's GEQ): near DC every B_s and A_s is a ~1e-5 ... 1e-1 remainder, so
    prod B and prod A taken separately (the reference's formula, dsp.py:2587-2593) underflow float32 and give 0 / 0,
    while the product of the per-section ratios B_s / A_s — what the kernels accumulate — stays O(1)."""
    fs, nfft, alias = 48000, 192000, 30.0
    gamma = 10 ** (-alias / nfft / 20)
    fcs = 31.25 * 2.0 ** (np.arange(30) / 3.0)
    fcs = fcs[fcs < 0.45 * fs]
    k = np.arange(0, 40, dtype=f64)
    B64 = np.ones((len(k),), dtype=np.complex128)
    A64 = np.ones_like(B64)
    num32 = np.ones((len(k),), dtype=np.complex64)
    den32 = np.ones_like(num32)
    ratio32 = np.ones_like(num32)
    for i, fc in enumerate(fcs):
        g = 10 ** ((3.0 if i % 2 else -3.0) / 20)
        b, a = peak_filter(torch.tensor(fc, dtype=torch.float64), torch.tensor(g, dtype=torch.float64),
                           torch.tensor(4.3, dtype=torch.float64), fs=fs, dtype=torch.float64)
        b, a = b.reshape(3).numpy(), a.reshape(3).numpy()
        w = gamma * np.exp(-2j * np.pi * k / nfft)
        B64 *= b[0] + b[1] * w + b[2] * w * w
        A64 *= a[0] + a[1] * w + a[2] * w * w
        packed = sweep.pack_sections(torch.tensor(b).view(3, 1, 1), torch.tensor(a).view(3, 1, 1), True,
                                     torch.float32)[0, 0].numpy()
        Bs, As = _taylor_eval_f32(packed, k, nfft, gamma)
        num32 = (num32 * Bs).astype(np.complex64)
        den32 = (den32 * As).astype(np.complex64)
        ratio32 = (ratio32 * (Bs / As)).astype(np.complex64)
    H = B64 / A64
    assert np.abs(A64).min() < 1e-45  # float64 holds it; float32 (tiny = 1.2e-38, denormals to 1.4e-45) cannot
    with np.errstate(all="ignore"):
        separate = num32 / den32
    assert not np.isfinite(separate).all() or np.abs(separate - H).max() > 0.1 * np.abs(H).max()
    assert np.abs(ratio32 - H).max() <= 2e-5 * np.abs(H).max()


def test_float32_loop_solve_meets_the_bar_only_with_exact_phases():
    """Config 2 (8 x 8 FDN, nfft = 96000, alias 30 dB) modelled in numpy: H = c^T (I - D W)^-1 D b per bin with a
    float32 LU (LAPACK complex64, partial pivoting — the arithmetic of the loop kernels).  With the delay phases formed
    exactly (integer k * m mod nfft) the magnitude response is within the 1e-4 bar of the float64 truth; with the
    reference's float32 omega * m it is not — the phase error, not the float32 LU, is what breaks float32 parity."""
    from flamo_b200 import workloads as W

    nfft, N, alias = 96000, 8, 30.0
    m = np.array(W.fdn_delays(N), dtype=np.int64)
    rng = np.random.default_rng(130709)
    P = rng.standard_normal((N, N))
    S = np.triu(P, 1) - np.triu(P, 1).T
    Wm = torch.matrix_exp(torch.tensor(S)).numpy()  # orthogonal feedback matrix
    b, c = rng.standard_normal((N, 1)), rng.standard_normal((1, N))
    gamma = 10 ** (-alias / nfft / 20)
    k = np.arange(0, nfft // 2 + 1, dtype=np.int64)[::7]  # every 7th bin: 6858 systems
    gm = gamma ** m.astype(f64)

    def response(D, dtype):
        A = np.eye(N, dtype=dtype)[None] - D[:, :, None].astype(dtype) * Wm.astype(dtype)[None]
        rhs = (D.astype(dtype) * b[:, 0].astype(dtype)[None])[:, :, None]
        y = np.linalg.solve(A, rhs)[:, :, 0]
        return np.abs((y * c[0].astype(dtype)[None]).sum(-1))

    D_true = gm[None] * np.exp(-2j * np.pi * ((k[:, None] * m[None]) % nfft) / nfft)
    truth = response(D_true, np.complex128)
    fr = (2.0 * ((k[:, None] * m[None]) % nfft) / nfft).astype(f32).astype(f64)
    D_exact = (gm.astype(f32)[None] * (np.cos(np.pi * fr) - 1j * np.sin(np.pi * fr))).astype(np.complex64)
    omega32 = (f32(2) * f32(np.pi) * k.astype(f32) / f32(nfft)).astype(f32)
    ph32 = (omega32[:, None] * m.astype(f32)[None]).astype(f32).astype(f64)
    D_naive = (gm.astype(f32)[None] * (np.cos(ph32) - 1j * np.sin(ph32))).astype(np.complex64)

    def rel(a):
        return float((np.abs(a - truth) / np.maximum(truth, 1e-3 * truth.max())).max())

    err_exact, err_naive = rel(response(D_exact, np.complex64)), rel(response(D_naive, np.complex64))
    assert err_exact < 1e-4, err_exact
    assert err_naive > 10 * err_exact and err_naive > 1e-4, (err_naive, err_exact)
