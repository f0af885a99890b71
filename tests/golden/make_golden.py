"""Generate tests/golden/*.npz from the UNMODIFIED reference (build container only).

Run:  python tests/golden/make_golden.py            (needs /root/reference; ~2-3 min on 8 cores)

The reference (gdalsanto/flamo @ 6ab8e6b) is imported from /root/reference with empty stub
modules for the optional packages it imports at module scope but never touches on the hot
path (soundfile, nnAudio, pyfar, matplotlib — SURVEY.md §8c).  Every case is run in float64
(the examples' default dtype, e.g. examples/e8_colorless_fdn.py:197).  For each case we store

  params      raw nn.Parameters of the reference model (parameters() order)
  bins, Y     complex core output at a subset of bins for the deterministic input of
              tests/cases.py:make_input
  loss, grads reference autograd gradients of  mean((sum_ch |Y| - 1)^2)  w.r.t. every
              parameter that requires grad (all bins)

The same script cross-checks oracle/flamo_oracle.py against the reference on the full
tensors and prints the worst deviation: ~1e-12 with REF_FP32_INTERNALS=True (the reference's
float32 SVF/GEQ tap buffers reproduced, SURVEY.md §8c caveat), and the deviation of the
full-precision oracle, which shows how far the reference's own float32 internals are from
exact arithmetic (up to ~7e-3 for GEQ at DC).
"""
import os
import sys
import types

for _n in ["soundfile", "nnAudio", "nnAudio.features", "pyfar", "matplotlib", "matplotlib.pyplot"]:
    sys.modules[_n] = types.ModuleType(_n)
sys.modules["nnAudio"].features = sys.modules["nnAudio.features"]
sys.path.insert(0, "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

from flamo.processor import dsp as rdsp, system as rsystem  # the reference
from flamo_b200 import workloads as W
import cases as C
from oracle import flamo_oracle as O

torch.set_num_threads(os.cpu_count())


def run_case(name, case):
    desc, nfft, alias, B, Ccols, seed = case["desc"], case["nfft"], case["alias"], case["B"], case["C"], case["seed"]
    torch.manual_seed(seed)
    model = W.build(desc, rdsp, rsystem, nfft, alias, dtype=torch.float64)
    params = [p.detach().clone() for p in model.parameters()]
    n_in = model.input_channels
    M = nfft // 2 + 1
    X = C.make_input(B, M, n_in, Ccols)
    Y = model(X)
    loss = C.golden_loss(Y)
    gp = [p for p in model.parameters() if p.requires_grad]
    grads = torch.autograd.grad(loss, gp) if (gp and case.get("grads", True)) else []
    bins = C.select_bins(M)
    out = {"bins": bins, "Y": Y.detach()[:, bins].numpy(), "loss": float(loss)}
    for i, p in enumerate(params):
        out[f"param_{i}"] = p.numpy()
    gi = 0
    for i, p in enumerate(model.parameters()):
        if p.requires_grad and len(grads):
            out[f"grad_{i}"] = grads[gi].numpy()
            gi += 1
    # cross-check the oracle on the full tensor (reference dtype flow reproduced)
    O.REF_FP32_INTERNALS = True
    node = O.from_desc(desc)
    op = [p.clone().requires_grad_(q.requires_grad) for p, q in zip(params, model.parameters())]
    Yo = O.forward(node, X, op, nfft, alias)
    den = torch.clamp(Y.detach().abs(), min=1e-3 * Y.detach().abs().max())
    err = ((Yo.detach() - Y.detach()).abs() / den).max().item()
    gerr = 0.0
    if len(grads):
        lo = C.golden_loss(Yo)
        go = torch.autograd.grad(lo, [p for p in op if p.requires_grad])
        for a, b in zip(go, grads):
            gerr = max(gerr, ((a - b).abs().max() / (b.abs().max() + 1e-30)).item())
    O.REF_FP32_INTERNALS = False
    with torch.no_grad():
        Yt = O.forward(node, X, params, nfft, alias)
    terr = ((Yt - Y.detach()).abs() / den).max().item()
    print(f"{name:28s} M={M:6d} loss={float(loss):.6e}  oracle-vs-ref  Y:{err:.2e} grad:{gerr:.2e}"
          f"   (full-precision oracle vs ref: {terr:.2e})")
    out["ref_fp32_noise"] = terr  # deviation of the reference from exact arithmetic (its float32 internals)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    return err, gerr


def probe_kat():
    """examples/e10_probe.py:30-157 — the reference's only known-answer check:
    4x4 FDN, fixed delays/gains, nfft=2**15: |probe(z_k) - core(ones)| < 5e-3."""
    torch.manual_seed(130709)
    nfft, N, alias = 2**15, 4, 0.0
    desc = C.probe_fdn_desc()
    model = W.build(desc, rdsp, rsystem, nfft, alias, dtype=torch.float64)
    M = nfft // 2 + 1
    X = torch.ones(1, M, 1, dtype=torch.complex128)
    Y = model(X)
    bins = np.arange(0, M, 97)
    probes = []
    for k in bins:
        z = torch.exp(torch.tensor(1j * 2 * np.pi * k / nfft, dtype=torch.complex128))
        probes.append(model.probe(z))
    P = torch.stack(probes).detach()  # (nb, 1, 1)
    err = (P[:, :, 0] - Y.detach()[0, bins]).abs().max().item()
    print(f"probe KAT: max|probe - sweep| = {err:.2e} (reference bound 5e-3)")
    out = {"bins": bins, "probe": P.numpy(), "Y": Y.detach()[:, bins].numpy()}
    for i, p in enumerate(model.parameters()):
        out[f"param_{i}"] = p.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "kat_probe_fdn4.npz"), **out)


def checkpoints():
    """notebooks/output/ex_fdn/checkpoints/model_e*.pt — state_dict naming/shape pin."""
    base = "/root/reference/notebooks/output"
    out = {}
    for sub, tag in (("ex_fdn", "fdn"), ("ex_biquad", "biquad")):
        d = os.path.join(base, sub, "checkpoints")
        for e in (0, 7):
            sd = torch.load(os.path.join(d, f"model_e{e}.pt"), map_location="cpu", weights_only=True)
            for k, v in sd.items():
                out[f"{tag}|e{e}|{k}"] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "reference_checkpoints.npz"), **out)
    print("checkpoint keys:", sorted({k.split('|')[2] for k in out}))


def checkpoint_responses():
    """SURVEY §8c fixtures: every shipped checkpoint (notebooks/output/ex_fdn: 20 epochs of the 6-line colourless FDN,
    notebooks/e8_colorless_fdn.ipynb, nfft = 2**16, alias 30 dB; notebooks/output/ex_biquad: 8 epochs of the 2-section
    1 -> 2 bandpass Biquad, notebooks/e7_biquad.ipynb, alias 0 dB) is loaded into the reference model (float64 modules)
    and the reference's Shell.get_freq_response / get_time_response (system.py:1012-1153) are recorded."""
    base = "/root/reference/notebooks/output"
    nfft = 2 ** 16
    M = nfft // 2 + 1
    bins = C.select_bins(M)
    taps = np.unique(np.concatenate([np.arange(0, 2200), np.arange(2200, nfft, 499)]))
    out = {"bins": bins, "taps": taps}

    def shell(core):
        return rsystem.Shell(core=core, input_layer=rdsp.FFT(nfft, dtype=torch.float64),
                             output_layer=rdsp.Transform(lambda x: torch.abs(x), dtype=torch.float64))

    fdn = shell(W.build(W.fdn(6), rdsp, rsystem, nfft, 30, dtype=torch.float64))
    bq = shell(W.build(W.biquad(2, 1, 2, "bandpass"), rdsp, rsystem, nfft, 0, dtype=torch.float64))
    for tag, model, sub, n in (("fdn", fdn, "ex_fdn", 20), ("biquad", bq, "ex_biquad", 8)):
        for e in range(n):
            sd = torch.load(os.path.join(base, sub, "checkpoints", f"model_e{e}.pt"), map_location="cpu",
                            weights_only=True)
            model.load_state_dict(sd)
            for i, q in enumerate(model.parameters()):
                out[f"{tag}|e{e}|param_{i}"] = q.detach().numpy().copy()  # copy: load_state_dict writes in place
            H = model.get_freq_response(identity=False)
            assert H.shape[:2] == (1, M)
            out[f"{tag}|e{e}|H"] = H[0, bins].numpy()
            if e in (0, n - 1):  # impulse responses of the first and the last epoch only (size)
                h = model.get_time_response(identity=False)
                assert h.shape[:2] == (1, nfft)
                out[f"{tag}|e{e}|h"] = h[0, taps].numpy()
            out[f"{tag}|e{e}|mag"] = model(signal_impulse(nfft, model.input_channels))[0, bins].detach().numpy()
        print(f"checkpoint responses: {tag} x {n}")
    np.savez_compressed(os.path.join(HERE, "reference_checkpoint_responses.npz"), **out)


def signal_impulse(nfft, n):
    x = torch.zeros(1, nfft, n, dtype=torch.float64)
    x[:, 0, :] = 1
    return x


def train_trace():
    """Three Trainer.train_step calls of the reference on a reduced config 2 (N=8, nfft=4096),
    float64, Adam lr=1e-3, mse_loss + 0.2*sparsity_loss (examples/e8_colorless_fdn.py:128-138)."""
    sys.modules["nnAudio"].features = sys.modules["nnAudio.features"]
    from flamo.optimize.trainer import Trainer
    from flamo.optimize.loss import mse_loss, sparsity_loss
    from flamo.optimize.dataset import DatasetColorless

    nfft, alias = 4096, W.ALIAS_DECAY_DB
    torch.manual_seed(130709)
    core = W.build(W.fdn(8), rdsp, rsystem, nfft, alias, dtype=torch.float64)
    model = rsystem.Shell(core=core, input_layer=rdsp.FFT(nfft, dtype=torch.float64),
                          output_layer=rdsp.Transform(lambda x: torch.abs(x), dtype=torch.float64))
    params0 = [p.detach().clone().numpy() for p in model.parameters()]
    ds = DatasetColorless(input_shape=(1, nfft // 2 + 1, 1), target_shape=(1, nfft // 2 + 1, 1), expand=4,
                          device="cpu", dtype=torch.float64)
    tr = Trainer(model, max_epochs=1, lr=1e-3, log=False, device="cpu")
    tr.register_criterion(mse_loss(nfft=nfft, device="cpu"), 1)
    tr.register_criterion(sparsity_loss(), 0.2, requires_model=True)
    tr.train_loss_log = {"mse_loss": [], "sparsity_loss": []}
    losses = []
    for i in range(3):
        x, y = ds[i]
        losses.append(tr.train_step((x.unsqueeze(0), y.unsqueeze(0))))
    out = {"losses": np.array(losses), "mse": np.array(tr.train_loss_log["mse_loss"]),
           "sparsity": np.array(tr.train_loss_log["sparsity_loss"])}
    for i, p in enumerate(params0):
        out[f"param0_{i}"] = p
    for i, p in enumerate(model.parameters()):
        out[f"param3_{i}"] = p.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "train_trace_fdn8.npz"), **out)
    print("train trace losses:", losses)


if __name__ == "__main__":
    only = sys.argv[1:]
    if only == ["checkpoint_responses"]:
        checkpoint_responses()
        sys.exit(0)
    worst = 0.0
    for name, case in C.CASES.items():
        if only and name not in only:
            continue
        e, g = run_case(name, case)
    if not only:
        probe_kat()
        checkpoints()
        checkpoint_responses()
        train_trace()
