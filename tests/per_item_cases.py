"""Per-item parameter sets ("batched ext_param"): one parameter set per batch item inside ONE
launch, where the reference's NN-in-the-loop examples call the model once per item (examples/e7_biquad_nn.py:149-156,
e4_recursion_nn.py:243-250).  Shared by the CPU tests (ABI emulator) and the GPU tests (the real kernels).

Each case: a model description with keyed Series modules, the key that receives the external parameters, and how the
B parameter sets are drawn.  Checks (run by `check_case`):
  * model(x, {key: P}) with P of shape (B, *param.shape) == the reference's per-item loop over model(x[b:b+1], {key: P[b]}),
    outputs and gradients (w.r.t. P, the model's own trainable parameters, and x);
  * every item == the float64 oracle evaluated with that item's parameters;
  * afterwards the module's `param` holds the LAST set, as the reference's loop leaves it (dsp.py:415-432).
"""
import numpy as np
import torch

from flamo_b200 import workloads as W

NFFT = 512
FS = W.FS


def _pdelay(n, first, isint=False, grad=False):
    return ("parallelDelay", dict(size=(n,), max_len=2000, isint=isint, requires_grad=grad),
            {"delay_samples": [first + 37.25 * i for i in range(n)]})


CASES = {
    # e7_biquad_nn: a Biquad conditioned by a network, 1 -> 2 channels, 2 sections (folded design: batch -> sections)
    "biquad_highpass": dict(
        desc=("Series", [("Biquad", dict(size=(2, 1), n_sections=2, filter_type="highpass", fs=FS, requires_grad=False))],
              ["biquad"]),
        key="biquad", n_in=1, alias=30.0, draw="biquad"),
    "biquad_bandpass_after_gain": dict(
        desc=("Series", [("Gain", dict(size=(2, 1), requires_grad=True)),
                         ("parallelBiquad", dict(size=(2,), n_sections=3, filter_type="bandpass", fs=FS, requires_grad=False))],
              ["in", "biquad"]),
        key="biquad", n_in=1, alias=0.0, draw="biquad"),
    "svf_general": dict(
        desc=("Series", [("SVF", dict(size=(2, 2), n_sections=2, fs=FS, requires_grad=False)),
                         ("Gain", dict(size=(1, 2), requires_grad=True))], ["svf", "out"]),
        key="svf", n_in=2, alias=30.0, draw="normal"),
    "svf_peaking": dict(
        desc=("Series", [("parallelSVF", dict(size=(3,), n_sections=2, filter_type="peaking", fs=FS, requires_grad=False))],
              ["svf"]),
        key="svf", n_in=3, alias=10.0, draw="normal"),
    "geq": dict(
        desc=("Series", [("GEQ", dict(size=(2, 2), octave_interval=1, fs=FS, requires_grad=False))], ["geq"]),
        key="geq", n_in=2, alias=30.0, draw="positive"),
    "gain_matrix": dict(
        desc=("Series", [("Gain", dict(size=(3, 2), requires_grad=False)), _pdelay(3, 5.0, grad=True)], ["mix", "delay"]),
        key="mix", n_in=2, alias=30.0, draw="normal"),
    "orthogonal_matrix": dict(
        desc=("Series", [("Matrix", dict(size=(4, 4), matrix_type="orthogonal", requires_grad=False)),
                         ("parallelGain", dict(size=(4,), requires_grad=True))], ["mix", "g"]),
        key="mix", n_in=4, alias=0.0, draw="normal"),
    "delays": dict(
        desc=("Series", [("Gain", dict(size=(3, 1), requires_grad=True)),
                         ("parallelDelay", dict(size=(3,), max_len=2000, isint=False, requires_grad=False))], ["in", "delay"]),
        key="delay", n_in=1, alias=30.0, draw="delay"),
    "fir_table": dict(
        desc=("Series", [("Filter", dict(size=(12, 2, 2), requires_grad=False)),
                         ("parallelGain", dict(size=(2,), requires_grad=True))], ["fir", "g"]),
        key="fir", n_in=2, alias=30.0, draw="normal"),
    # e4_recursion_nn: a comb whose FEEDBACK filter is conditioned by a network (key "feedback" routes into the loop)
    "comb_feedback_biquad": dict(
        desc=("Recursion", _pdelay(2, 7.0, isint=True),
              ("Biquad", dict(size=(2, 2), n_sections=1, filter_type="lowpass", fs=FS, requires_grad=False))),
        key="feedback", n_in=2, alias=30.0, draw="biquad_small"),
    "fdn_feedback_gain": dict(
        desc=("Series", [("Gain", dict(size=(4, 1), requires_grad=True)),
                         ("Recursion", _pdelay(4, 11.0, isint=True), ("Gain", dict(size=(4, 4), requires_grad=False))),
                         ("Gain", dict(size=(1, 4), requires_grad=True))], ["in", "loop", "out"]),
        key="loop", sub="feedback", n_in=1, alias=30.0, draw="small"),
}


def draw(kind, shape, seed):
    g = torch.Generator().manual_seed(seed)
    if kind == "normal":
        return torch.randn(shape, generator=g, dtype=torch.float64)
    if kind == "small":
        return 0.2 * torch.randn(shape, generator=g, dtype=torch.float64)
    if kind == "positive":
        return 0.5 + torch.rand(shape, generator=g, dtype=torch.float64)
    if kind == "delay":  # seconds
        return (3.0 + 40.0 * torch.rand(shape, generator=g, dtype=torch.float64)) / FS
    if kind in ("biquad", "biquad_small"):  # (B, K, 2|3, ...): cut-offs in (0.05, 0.9), gains around 1 (or small)
        p = 0.05 + 0.85 * torch.rand(shape, generator=g, dtype=torch.float64)
        if shape[2] == 3:
            p[:, :, 1] = p[:, :, 0] + (0.95 - p[:, :, 0]) * torch.rand(p[:, :, 0].shape, generator=g, dtype=torch.float64)
        p[:, :, -1] = (0.3 if kind == "biquad_small" else 1.0) * (0.5 + torch.rand(p[:, :, -1].shape, generator=g, dtype=torch.float64))
        return p
    raise KeyError(kind)


def target_module(model, case):
    from flamo_b200.processor import system

    m = model._modules[case["key"]] if isinstance(model, system.Series) else getattr(model, case["key"])
    if case.get("sub"):
        m = getattr(m, case["sub"])
    return m


def ext_dict(case, P):
    return {case["key"]: ({case["sub"]: P} if case.get("sub") else P)}


def check_case(name, dtype, device, B=3, cols=None, tol_item=1e-9, tol_oracle=1e-8, tol_grad=1e-7, shard=None):
    from flamo_b200 import sweep
    from flamo_b200.processor import dsp, system
    from helpers import rel_err
    from oracle import flamo_oracle as O

    case = CASES[name]
    torch.manual_seed(7)
    model = W.build(case["desc"], dsp, system, NFFT, case["alias"], dtype=dtype, device=device)
    tm = target_module(model, case)
    P = draw(case["draw"], (B,) + tuple(tm.param.shape), 11).to(dtype=dtype, device=device).requires_grad_(True)
    M = NFFT // 2 + 1
    cdt = torch.complex64 if dtype == torch.float32 else torch.complex128
    g = torch.Generator().manual_seed(3)
    shape = (B, M, case["n_in"]) + ((cols,) if cols else ())
    X = torch.complex(torch.randn(shape, generator=g, dtype=torch.float64), torch.randn(shape, generator=g, dtype=torch.float64))
    X = X.to(cdt).to(device).requires_grad_(True)
    own = [p for p in model.parameters() if p.requires_grad]

    def run_all():
        return model(X, ext_dict(case, P))

    def run_loop():  # the reference's way: one call per item
        return torch.cat([model(X[b:b + 1], ext_dict(case, P[b])) for b in range(B)])

    def grads_of(fn):
        for p in own:
            p.grad = None
        Y = fn()
        w = torch.linspace(0.5, 1.5, Y.numel(), dtype=torch.float64, device=device).reshape(Y.shape)
        loss = (w * (Y.real.double() ** 2 + 0.5 * Y.imag.double() ** 2)).sum()
        gs = torch.autograd.grad(loss, [P, X] + own, allow_unused=True)
        return Y.detach(), [None if t is None else t.detach() for t in gs]

    ctx = sweep.bin_shard(*shard) if shard else None
    if ctx:
        ctx.__enter__()
    try:
        Ya, ga = grads_of(run_all)
        with torch.no_grad():
            assert torch.equal(tm.param.detach(), P[-1].detach().to(tm.param.dtype))  # the loop's leftover
        Yl, gl = grads_of(run_loop)
    finally:
        if ctx:
            ctx.__exit__(None, None, None)
    assert Ya.shape == Yl.shape
    scale = float(Yl.abs().max())
    assert float((Ya - Yl).abs().max()) <= tol_item * scale, name
    for a, l in zip(ga, gl):
        assert (a is None) == (l is None)
        if a is not None:
            assert a.shape == l.shape
            assert float((a - l).abs().max()) <= tol_grad * float(l.abs().max() + 1e-30), name
    if shard:
        return
    # every item against the oracle with ITS parameters
    node = O.from_desc(case["desc"])
    for b in range(B):
        with torch.no_grad():
            tm.param.copy_(P[b])
        ps = [p.detach().cpu().double() for p in model.parameters()]
        Yo = O.forward(node, X[b:b + 1].detach().cpu().to(torch.complex128), ps, NFFT, case["alias"]).numpy()
        assert rel_err(Ya[b:b + 1].cpu().numpy().astype(np.complex128), Yo) <= tol_oracle, (name, b)
