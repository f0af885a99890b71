"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

import cases as C
from flamo_b200 import workloads as W
from flamo_b200.processor import dsp, system

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def golden_params(g):
    out, i = [], 0
    while f"param_{i}" in g:
        out.append(g[f"param_{i}"])
        i += 1
    return out


def build_case(name, dtype, device):
    """This package's module tree for a parity case, with the reference's raw parameters."""
    case, g = C.CASES[name], load_golden(name)
    model = W.build(case["desc"], dsp, system, case["nfft"], case["alias"], dtype=dtype, device=device)
    W.set_params(model, golden_params(g))
    return case, g, model


def rel_err(Y, Yref, floor=1e-3):
    """|Y - Yref| / max(|Yref|, floor * max|Yref|)  (BASELINE.md §2 metric)."""
    Y, Yref = np.asarray(Y), np.asarray(Yref)
    den = np.maximum(np.abs(Yref), floor * np.abs(Yref).max())
    return float((np.abs(Y - Yref) / den).max())


def grad_err(g, gref):
    g, gref = np.asarray(g, dtype=np.float64), np.asarray(gref, dtype=np.float64)
    return float(np.abs(g - gref).max() / (np.abs(gref).max() + 1e-300))


# ------------------------------------------------------------------ the reference's shipped checkpoints
CHECKPOINTS = [("fdn", e) for e in range(20)] + [("biquad", e) for e in range(8)]


def checkpoint_shell(tag, dtype, device):
    """The notebook model a shipped checkpoint belongs to (notebooks/e8_colorless_fdn.ipynb: 6-line FDN, nfft 2**16,
    alias 30 dB; notebooks/e7_biquad.ipynb: 1 -> 2 bandpass Biquad, 2 sections, alias 0 dB), built from this package."""
    nfft = 2 ** 16
    if tag == "fdn":
        core = W.build(W.fdn(6), dsp, system, nfft, 30, dtype=dtype, device=device)
    else:
        core = W.build(W.biquad(2, 1, 2, "bandpass"), dsp, system, nfft, 0, dtype=dtype, device=device)
    return system.Shell(core, dsp.FFT(nfft, dtype=dtype), dsp.Transform(lambda x: torch.abs(x), dtype=dtype))


def check_checkpoint(tag, epoch, dtype, device, tol_mag, tol_resp):
    """Load checkpoint `epoch` through load_state_dict (the reference's key names) and compare forward |.|,
    get_freq_response and get_time_response with what the reference returned for it
    (tests/golden/reference_checkpoint_responses.npz).  tol_mag: floored relative error of the magnitude (BASELINE.md
    §2); tol_resp: error of the complex / time responses relative to their peak."""
    g = load_golden("reference_checkpoint_responses")
    model = checkpoint_shell(tag, dtype, device)
    keys = list(model.state_dict().keys())
    sd = {k: torch.tensor(g[f"{tag}|e{epoch}|param_{i}"]) for i, k in enumerate(keys)}
    model.load_state_dict(sd)
    nfft, bins, taps = model.nfft, g["bins"], g["taps"]
    x = torch.zeros(1, nfft, 1, dtype=dtype, device=device)
    x[:, 0, :] = 1
    with torch.no_grad():
        mag = model(x)[0, bins].cpu().numpy()
        assert rel_err(mag, g[f"{tag}|e{epoch}|mag"]) <= tol_mag
        H = model.get_freq_response(identity=False)
        ref = g[f"{tag}|e{epoch}|H"]
        assert H.shape == (1, nfft // 2 + 1, ref.shape[-1]) and H.is_complex()
        assert np.abs(H[0, bins].cpu().numpy() - ref).max() <= tol_resp * np.abs(ref).max()
        if f"{tag}|e{epoch}|h" in g.files:
            h = model.get_time_response(identity=False)
            ref = g[f"{tag}|e{epoch}|h"]
            assert h.shape == (1, nfft, ref.shape[-1]) and not h.is_complex()
            assert np.abs(h[0, taps].cpu().numpy() - ref).max() <= tol_resp * np.abs(ref).max()
