"""CPU: pin oracle/flamo_oracle.py against the golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  Bit-level agreement is expected when the oracle
reproduces the reference's float32 internals; the full-precision oracle must stay within
the reference's own float32 noise recorded in each fixture."""
import os

import numpy as np
import pytest
import torch

import cases as C
from oracle import flamo_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def params_of(g, dtype=torch.float64):
    out, i = [], 0
    while f"param_{i}" in g:
        out.append(torch.tensor(g[f"param_{i}"], dtype=dtype))
        i += 1
    return out


def rel_err(Y, Yref):
    den = np.maximum(np.abs(Yref), 1e-3 * np.abs(Yref).max())
    return (np.abs(Y - Yref) / den).max()


@pytest.mark.parametrize("name", list(C.CASES))
def test_oracle_matches_reference(name):
    case, g = C.CASES[name], load(name)
    node = O.from_desc(case["desc"])
    M = case["nfft"] // 2 + 1
    n_in = g["Y"].shape  # noqa
    params = params_of(g)
    # requires_grad flags come from which grads the reference produced
    for i, p in enumerate(params):
        p.requires_grad_(f"grad_{i}" in g)
    X = C.make_input(case["B"], M, _n_in(case["desc"]), case["C"])
    O.REF_FP32_INTERNALS = True
    try:
        Y = O.forward(node, X, params, case["nfft"], case["alias"])
        assert rel_err(Y.detach()[:, g["bins"]].numpy(), g["Y"]) < 1e-11
        gp = [p for p in params if p.requires_grad]
        if gp:
            loss = C.golden_loss(Y)
            assert abs(loss.item() - float(g["loss"])) <= 1e-11 * max(1.0, abs(float(g["loss"])))
            grads = torch.autograd.grad(loss, gp)
            k = 0
            for i, p in enumerate(params):
                if p.requires_grad:
                    ref = g[f"grad_{i}"]
                    assert np.abs(grads[k].numpy() - ref).max() <= 1e-10 * (np.abs(ref).max() + 1e-30), (name, i)
                    k += 1
    finally:
        O.REF_FP32_INTERNALS = False
    # full-precision oracle: within the reference's own float32 noise
    with torch.no_grad():
        Yt = O.forward(node, X, [p.detach() for p in params], case["nfft"], case["alias"])
    assert rel_err(Yt[:, g["bins"]].numpy(), g["Y"]) <= max(1e-11, 1.01 * float(g["ref_fp32_noise"]))


def _n_in(desc):
    name = desc[0]
    if name == "Series":
        return _n_in(desc[1][0])
    if name in ("Recursion", "Parallel"):
        return _n_in(desc[1])
    size = desc[1]["size"]
    return size[-1]


def test_known_answer_probe():
    """examples/e10_probe.py:149-157: z-plane probe == bin sweep (< 5e-3) for the 4x4 FDN."""
    g = load("kat_probe_fdn4")
    node = O.from_desc(C.probe_fdn_desc())
    nfft = 2**15
    M = nfft // 2 + 1
    X = torch.ones(1, M, 1, dtype=torch.complex128)
    Y = O.forward(node, X, params_of(g), nfft, 0.0)[0, g["bins"]].numpy()
    assert np.abs(Y - g["Y"][0]).max() < 1e-11
    assert np.abs(Y - g["probe"][:, :, 0]).max() < 5e-3


def test_train_trace():
    """Three reference Trainer.train_step calls (Adam, mse + 0.2*sparsity) reproduced by the oracle."""
    g = np.load(os.path.join(GOLD, "train_trace_fdn8.npz"))
    from flamo_b200 import workloads as W

    node = O.from_desc(W.fdn(8))
    nfft = 4096
    params, i = [], 0
    while f"param0_{i}" in g:
        params.append(torch.tensor(g[f"param0_{i}"]))
        i += 1
    for p, flag in zip(params, (True, False, True, True)):
        p.requires_grad_(flag)
    fb_node = node.children[1].children[1]
    crit = [
        (1, lambda est, tgt, ps: O.mse_loss(est, tgt)),
        (0.2, lambda est, tgt, ps: O.sparsity_loss(O.mapped_matrix(fb_node, ps[2]))),
    ]
    tr = O.OracleTrainer(node, params, nfft, W.ALIAS_DECAY_DB, crit, lr=1e-3)
    x = torch.zeros(1, nfft // 2 + 1, 1, dtype=torch.float64)
    x[:, 0, :] = 1
    y = torch.ones(1, nfft // 2 + 1, 1, dtype=torch.float64)
    losses = [tr.train_step(x, y) for _ in range(3)]
    assert np.allclose(losses, g["losses"], rtol=1e-10, atol=0)
    for i, p in enumerate(params):
        assert np.allclose(p.detach().numpy(), g[f"param3_{i}"], rtol=1e-9, atol=1e-12)


CHECKPOINTS = [("fdn", e) for e in range(20)] + [("biquad", e) for e in range(8)]


def checkpoint_model(tag):
    """(description, nfft, alias, n_in) of the two notebook models whose checkpoints the reference ships."""
    from flamo_b200 import workloads as W

    return (W.fdn(6), 2 ** 16, 30.0, 1) if tag == "fdn" else (W.biquad(2, 1, 2, "bandpass"), 2 ** 16, 0.0, 1)


@pytest.mark.parametrize("tag,epoch", CHECKPOINTS)
def test_shipped_checkpoint_responses(tag, epoch):
    """SURVEY §8c fixtures: the reference's get_freq_response / get_time_response / forward of every checkpoint under
    notebooks/output (recorded by make_golden.checkpoint_responses) reproduced by the oracle."""
    g = load("reference_checkpoint_responses")
    desc, nfft, alias, n_in = checkpoint_model(tag)
    node = O.from_desc(desc)
    params, i = [], 0
    while f"{tag}|e{epoch}|param_{i}" in g:
        params.append(torch.tensor(g[f"{tag}|e{epoch}|param_{i}"]))
        i += 1
    with torch.no_grad():
        H = O.freq_response(node, params, nfft, alias, n_in)[0, g["bins"]].numpy()
        ref = g[f"{tag}|e{epoch}|H"]
        assert np.abs(H - ref).max() <= 1e-10 * np.abs(ref).max()
        x = torch.zeros(1, nfft, n_in, dtype=torch.float64)
        x[:, 0, :] = 1
        mag = O.shell_forward(node, x, params, nfft, alias)[0, g["bins"]].numpy()
        assert rel_err(mag, g[f"{tag}|e{epoch}|mag"]) < 1e-11
        if f"{tag}|e{epoch}|h" in g:
            h = O.time_response(node, params, nfft, alias, n_in)[0, g["taps"]].numpy()
            ref = g[f"{tag}|e{epoch}|h"]
            assert np.abs(h - ref).max() <= 1e-11 * np.abs(ref).max()


@pytest.mark.parametrize("name", ["cfg1_biquad_full", "cfg2_fdn8_full", "cfg3_geq16_small", "cfg4_active_full",
                                  "cfg5_fdn64_small", "svf_general", "geq_oct3", "delay_mimo_frac", "fir_filter",
                                  "gaindelay_frac", "recursion_rect"])
def test_bin_subset_mode_equals_full_sweep(name):
    """oracle.forward(..., bins=idx) — the per-bin closed form used to check the BASELINE-size configs on a few
    hundred bins — is the full-M oracle restricted to those bins (response and parameter gradients)."""
    if name not in C.CASES:
        pytest.skip("case not in this build")
    case, g = C.CASES[name], load(name)
    node = O.from_desc(case["desc"])
    M = case["nfft"] // 2 + 1
    X = C.make_input(case["B"], M, _n_in(case["desc"]), case["C"])
    idx = torch.as_tensor(C.select_bins(M)[::3].copy())

    def run(bins):
        ps = [p.requires_grad_(f"grad_{i}" in g) for i, p in enumerate(params_of(g))]
        if bins is None:
            Y = O.forward(node, X, ps, case["nfft"], case["alias"])[:, idx]
        else:
            Y = O.forward(node, X[:, idx], ps, case["nfft"], case["alias"], bins=idx)
        gp = [p for p in ps if p.requires_grad]
        gr = torch.autograd.grad(C.golden_loss(Y), gp) if gp else ()
        return Y.detach().numpy(), [t.numpy() for t in gr]

    Yf, gf = run(None)
    Ys, gs = run(idx)
    assert rel_err(Ys, Yf) < 1e-10
    for a, b in zip(gs, gf):
        assert np.abs(a - b).max() <= 1e-9 * (np.abs(b).max() + 1e-30)
