"""GPU: Shell / Trainer level behaviour of the CUDA path — the reference's train_step trace, CUDA-graph
replay vs eager, the fused |.| epilogue, bin sharding, input gradients, ext_param routing, and the
size-independent properties at BASELINE.json's full sizes."""
import os

import numpy as np
import pytest
import torch

import cases as C
from flamo_b200 import sweep, workloads as W
from flamo_b200.optimize.dataset import DatasetColorless
from flamo_b200.optimize.loss import mse_loss, sparsity_loss
from flamo_b200.optimize.trainer import Trainer
from flamo_b200.processor import dsp, system
from helpers import rel_err
from oracle import flamo_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda"


def fdn_shell(N, nfft, dtype, alias=30.0, seed=130709):
    torch.manual_seed(seed)
    core = W.build(W.fdn(N), dsp, system, nfft, alias, dtype=dtype, device=DEV)
    return system.Shell(core, dsp.FFT(nfft, dtype=dtype), dsp.Transform(lambda x: torch.abs(x), dtype=dtype))


def colorless(nfft, dtype, B=1):
    M = nfft // 2 + 1
    ds = DatasetColorless(input_shape=(1, M, 1), target_shape=(1, M, 1), expand=B, device=DEV, dtype=dtype)
    return ds.input[:B], ds.target[:B]


def make_trainer(model, nfft, graph):
    tr = Trainer(model, max_epochs=1, lr=1e-3, log=False, device=DEV, graph=graph)
    tr.register_criterion(mse_loss(nfft=nfft, device=DEV), 1)
    tr.register_criterion(sparsity_loss(), 0.2, requires_model=True)
    return tr


def test_train_trace_matches_reference():
    """Three Trainer.train_step calls reproduce the reference's losses and parameters (float64)."""
    g = np.load(os.path.join(GOLD, "train_trace_fdn8.npz"))
    nfft = 4096
    model = fdn_shell(8, nfft, torch.float64)
    W.set_params(model, [g[f"param0_{i}"] for i in range(4)])
    tr = make_trainer(model, nfft, graph=False)
    x, y = colorless(nfft, torch.float64)
    losses = [tr.train_step((x, y)) for _ in range(3)]
    assert np.allclose(losses, g["losses"], rtol=1e-9)
    assert np.allclose(tr.train_loss_log["mse_loss"], g["mse"], rtol=1e-9)
    assert np.allclose(tr.train_loss_log["sparsity_loss"], g["sparsity"], rtol=1e-9)
    for i, p in enumerate(model.parameters()):
        assert np.allclose(p.detach().cpu().numpy(), g[f"param3_{i}"], rtol=1e-8, atol=1e-11)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_graph_replay_equals_eager(dtype):
    nfft = 8192
    ma, mb = fdn_shell(8, nfft, dtype), fdn_shell(8, nfft, dtype)
    ta, tb = make_trainer(ma, nfft, graph=False), make_trainer(mb, nfft, graph=True)
    x, y = colorless(nfft, dtype, B=2)
    la = [ta.train_step((x, y)) for _ in range(10)]
    lb = [tb.train_step((x, y)) for _ in range(10)]
    assert tb.use_graph and len(tb._graphs) == 1, "the step was not captured"
    # the captured step uses fused capturable Adam with a float32 learning-rate tensor (1e-3 rounds by 5e-8)
    tol = 1e-5 if dtype == torch.float32 else 1e-6
    assert np.allclose(la, lb, rtol=tol)
    for pa, pb in zip(ma.parameters(), mb.parameters()):
        assert torch.allclose(pa, pb, rtol=tol * 10, atol=tol)
    assert la[-1] < la[0]  # it actually trains


@pytest.mark.parametrize("env", [{}, {"FLAMO_B200_NOTIFY": "0"}, {"FLAMO_B200_DEFER_TOTAL": "1"},
                                 {"FLAMO_B200_TOTAL_BRANCH": "1"}, {"FLAMO_B200_RFFT": "0"}],
                         ids=["default", "copy-node", "rider", "branch", "cufft"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_captured_step_loss_delivery(env, dtype, monkeypatch):
    """How the captured step hands its losses to the host — written into mapped pinned memory by the criteria-total
    kernel and polled (default), as a rider block of the adjoint-map launch, on a side branch of the graph, or through a
    copy node and a stream synchronize — and which input transform it runs (libfsweep's two-launch FFT or cuFFT) must not
    change a single loss: every variant against the eager trainer, 12 steps, time-domain input (the FFT is inside)."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    nfft = 32768
    ma, mb = fdn_shell(8, nfft, dtype), fdn_shell(8, nfft, dtype)
    mb.load_state_dict(ma.state_dict())
    ta, tb = make_trainer(ma, nfft, graph=False), make_trainer(mb, nfft, graph=True)
    x, y = colorless(nfft, dtype, B=1)
    la = [ta.train_step((x, y)) for _ in range(12)]
    lb = [tb.train_step((x, y)) for _ in range(12)]
    assert tb.use_graph and len(tb._graphs) == 1, "the step was not captured"
    (g,) = tb._graphs.values()
    assert (g[6] is not None) == (env.get("FLAMO_B200_NOTIFY") != "0")  # the notification slot is in use
    tol = 1e-5 if dtype == torch.float32 else 1e-6
    assert np.allclose(la, lb, rtol=tol), (la, lb)
    assert np.allclose(ta.train_loss_log["sparsity_loss"], tb.train_loss_log["sparsity_loss"], rtol=tol)
    for pa, pb in zip(ma.parameters(), mb.parameters()):
        assert torch.allclose(pa, pb, rtol=tol * 10, atol=tol)


@pytest.mark.parametrize("kind", ["pinned", "pageable", "pinned-no-kernel"])
def test_host_batches_reach_the_captured_step(kind, monkeypatch):
    """Batches that live on the host (reference trainer.py:176 move_to_device): pinned ones are uploaded by ONE kernel that
    reads them over PCIe (fsweep_upload), pageable ones by ordinary copies — every step must see the CURRENT content of
    the host tensors (they change between steps here), odd byte counts and both tensors included."""
    if kind == "pinned-no-kernel":
        monkeypatch.setenv("FLAMO_B200_UPLOAD_KERNEL", "0")
    nfft = 16384
    ma, mb = fdn_shell(8, nfft, torch.float32), fdn_shell(8, nfft, torch.float32)
    mb.load_state_dict(ma.state_dict())
    ta, tb = make_trainer(ma, nfft, graph=True), make_trainer(mb, nfft, graph=True)
    x, y = colorless(nfft, torch.float32, B=1)  # (1, 8193, 1): 32772 bytes, not a multiple of 16
    xh, yh = x.cpu(), y.cpu()
    if kind != "pageable":
        xh, yh = xh.pin_memory(), yh.pin_memory()
    n0 = sweep.launch_count
    la, lb = [], []
    for i in range(10):
        scale = 1.0 + 0.1 * (i % 3)
        xd, yd = x * scale, y * (2.0 - scale)
        xh.copy_(xd)
        yh.copy_(yd)
        torch.cuda.synchronize()
        la.append(ta.train_step((xd, yd)))
        lb.append(tb.train_step((xh, yh)))
        torch.cuda.synchronize()  # the host tensors are rewritten in the next iteration
    assert tb.use_graph and len(tb._graphs) == 1
    assert np.allclose(la, lb, rtol=1e-6), (la, lb)
    for pa, pb in zip(ma.parameters(), mb.parameters()):
        assert torch.allclose(pa, pb, rtol=1e-5, atol=1e-6)


def test_fused_abs_epilogue_equals_unfused():
    case = C.CASES["cfg4_active_full"]
    torch.manual_seed(case["seed"])
    core = W.build(case["desc"], dsp, system, 4096, 30.0, device=DEV)
    model = system.Shell(core, dsp.FFT(4096), dsp.Transform(lambda x: torch.abs(x)))
    x, _ = colorless(4096, torch.float32, B=3)
    a = model(x)
    ga = torch.autograd.grad(a.square().mean(), [p for p in model.parameters() if p.requires_grad])
    model.fuse_output = False
    b = model(x)
    gb = torch.autograd.grad(b.square().mean(), [p for p in model.parameters() if p.requires_grad])
    assert not a.is_complex() and torch.allclose(a, b, rtol=1e-6, atol=1e-7)
    for u, v in zip(ga, gb):
        assert torch.allclose(u, v, rtol=1e-4, atol=1e-7 * float(v.abs().max()))


def test_bin_sharding_is_exact():
    """Any partition of the bins gives bit-identical outputs and gradients that sum to the whole."""
    nfft = 4096
    M = nfft // 2 + 1
    model = fdn_shell(8, nfft, torch.float32)
    x, _ = colorless(nfft, torch.float32, B=2)
    ps = [p for p in model.parameters() if p.requires_grad]
    full = model(x)
    gfull = torch.autograd.grad(full.sum(), ps)
    cuts = [0, 700, 701, 1500, M]
    outs, gsum = [], [torch.zeros_like(g) for g in gfull]
    for b0, b1 in zip(cuts[:-1], cuts[1:]):
        with sweep.bin_shard(b0, b1):
            part = model(x)
        assert part.shape[1] == b1 - b0
        outs.append(part)
        for acc, g in zip(gsum, torch.autograd.grad(part.sum(), ps)):
            acc += g
    assert torch.equal(torch.cat(outs, dim=1), full)
    for a, b in zip(gsum, gfull):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6 * float(b.abs().max()))


def test_input_gradient_and_complex_output():
    case = C.CASES["recursion_rect"]
    torch.manual_seed(3)
    model = W.build(case["desc"], dsp, system, 1024, 30.0, dtype=torch.float64, device=DEV)
    M = 513
    X = C.make_input(2, M, model.input_channels, None).to(DEV).requires_grad_(True)
    Y = model(X)
    wgt = C.make_input(2, M, model.output_channels, None).to(DEV)
    (Y * wgt).real.sum().backward()
    params64 = [p.detach().cpu() for p in model.parameters()]
    Xo = X.detach().cpu().requires_grad_(True)
    Yo = O.forward(O.from_desc(case["desc"]), Xo, params64, 1024, 30.0)
    (Yo * wgt.cpu()).real.sum().backward()
    assert rel_err(Y.detach().cpu().numpy(), Yo.detach().numpy()) < 1e-9
    assert rel_err(X.grad.cpu().numpy(), Xo.grad.numpy()) < 1e-8


def test_ext_param_routing():
    nfft = 1024
    core = W.build(W.fdn(4, delays=[101, 157, 211, 263]), dsp, system, nfft, 30.0, device=DEV)
    M = nfft // 2 + 1
    X = torch.ones(1, M, 1, dtype=torch.complex64, device=DEV)
    new_gain = torch.full((1, 4), 0.25, device=DEV, requires_grad=True)
    y = core(X, {"output_gain": new_gain})
    assert torch.equal(core.output_gain.param.detach(), new_gain.detach())  # logged into the module
    y.abs().sum().backward()
    assert new_gain.grad is not None and new_gain.grad.abs().sum() > 0
    assert torch.allclose(y, core(X))


def test_responses_are_consistent():
    """get_freq_response == FFT(get_time_response) (reference system.py:1012-1153)."""
    nfft = 4096
    model = fdn_shell(6, nfft, torch.float64)
    H = model.get_freq_response()
    h = model.get_time_response()
    assert H.shape == (1, nfft // 2 + 1, 1) and h.shape == (1, nfft, 1)
    assert torch.allclose(torch.fft.rfft(h, n=nfft, dim=1), H, rtol=1e-9, atol=1e-9)
    Hi = model.get_freq_response(identity=False)
    assert torch.equal(H, Hi)
    assert isinstance(model.get_outputLayer(), dsp.Transform)  # layers restored


@pytest.mark.parametrize("name", ["cfg1_biquad", "cfg2_fdn8", "cfg3_geq16", "cfg4_active"])
def test_full_size_properties(name):
    """BASELINE.json sizes: the sweep is linear in its input, bin-sharding invariant, and the float32
    kernels agree with the float64 kernels (which the smaller cases pin to the oracle at 1e-9)."""
    desc, nfft, B, seed, _ = W.CONFIGS[name]
    M = nfft // 2 + 1
    torch.manual_seed(seed)
    m32 = W.build(desc, dsp, system, nfft, W.ALIAS_DECAY_DB, dtype=torch.float32, device=DEV)
    m64 = W.build(desc, dsp, system, nfft, W.ALIAS_DECAY_DB, dtype=torch.float64, device=DEV)
    W.set_params(m64, [p.detach() for p in m32.parameters()])
    X = C.make_input(B, M, m32.input_channels, None).to(DEV)
    with torch.no_grad():
        y32 = m32(X.to(torch.complex64))
        y64 = m64(X)
        a32, a64 = np.abs(y32.cpu().numpy()), np.abs(y64.cpu().numpy())
        # floored relative metric over ALL bins (configs 3 and 4 run float64 arithmetic inside: sweep._wants_f64)
        assert rel_err(a32, a64) < 1e-4
        assert np.abs(a32 - a64).max() / a64.max() < 2e-6
        # linearity: f(a x1 + b x2) = a f(x1) + b f(x2)
        X2 = torch.flip(X, dims=[1]) * (0.3 - 0.8j)
        lhs = m64((1.7 + 0.2j) * X + X2)
        rhs = (1.7 + 0.2j) * y64 + m64(X2)
        assert rel_err(lhs.cpu().numpy(), rhs.cpu().numpy()) < 1e-10
        half = M // 2 + 13
        with sweep.bin_shard(0, half):
            a = m32(X.to(torch.complex64))
        with sweep.bin_shard(half, M):
            b = m32(X.to(torch.complex64))
        # (the kernel family is chosen from the number of bins in the launch, so the two halves may run on a
        #  different family than the whole: equal to float32 rounding, not bit for bit -- bit-identity within
        #  one family is asserted in test_bin_sharding_is_exact)
        ab = torch.cat((a, b), dim=1)
        assert float((ab - y32).abs().max() / y32.abs().max()) < 2e-6


def test_config5_full_size_properties():
    """BASELINE config 5 at full size (64x64 FDN, nfft = 384000 -> 192001 bins, batch 32), float32, wide
    kernels: linear in the input, batch items independent, bin-shard invariant.  (The reference needs 201 GB
    per intermediate for this size; parity itself is pinned on the nfft = 2048 version of the same model.)"""
    desc, nfft, B, seed, _ = W.CONFIGS["cfg5_fdn64"]
    M = nfft // 2 + 1
    torch.manual_seed(seed)
    model = W.build(desc, dsp, system, nfft, W.ALIAS_DECAY_DB, dtype=torch.float32, device=DEV)
    X = C.make_input(B, M, 1, None).to(torch.complex64).to(DEV)
    with torch.no_grad():
        y = model(X)
        assert y.shape == (B, M, 1) and torch.isfinite(torch.view_as_real(y)).all()
        # batch items are independent: item 5 alone gives the same bits
        assert torch.equal(model(X[5:6]), y[5:6])
        # linearity (float32: relative to the peak)
        a = 1.5 - 0.5j
        lhs = model(a * X[:2] + X[2:4])
        rhs = a * y[:2] + y[2:4]
        assert float((lhs - rhs).abs().max() / rhs.abs().max()) < 2e-5
        half = M // 3
        with sweep.bin_shard(0, half):
            p0 = model(X[:2])
        with sweep.bin_shard(half, M):
            p1 = model(X[:2])
        assert torch.equal(torch.cat((p0, p1), dim=1), y[:2])


def _subset(M, n):
    stride = max(1, M // n)
    idx = np.unique(np.concatenate([np.arange(min(M, 16)), np.arange(0, M, stride), [M - 1]]))
    return torch.as_tensor(idx.astype(np.int64))


FULL_SIZE_REPORT = {}


@pytest.mark.parametrize("name,n_sub", [("cfg1_biquad", 509), ("cfg2_fdn8", 509), ("cfg3_geq16", 509),
                                         ("cfg4_active", 509), ("cfg5_fdn64", 127)])
def test_full_size_vs_oracle_bin_subset(name, n_sub):
    """All five BASELINE.json configs AT FULL SIZE (config 3: nfft 192000, config 5: nfft 384000 / batch 32 / 64 x 64),
    float32 kernels, against the float64 oracle evaluated in its per-bin closed form on a subset of the bins
    (oracle.forward(..., bins=idx); the full-M oracle needs 11.8 GB / 201 GB per intermediate for configs 3 / 5):
    magnitude response within 1e-4 on the floored metric, parameter gradients of
    mean((sum_ch |Y| - 1)^2) over the subset bins within 1e-3."""
    desc, nfft, B, seed, _ = W.CONFIGS[name]
    M = nfft // 2 + 1
    torch.manual_seed(seed)
    model = W.build(desc, dsp, system, nfft, W.ALIAS_DECAY_DB, dtype=torch.float32, device=DEV)
    X = C.make_input(B, M, model.input_channels, None).to(torch.complex64).to(DEV)
    idx = _subset(M, n_sub)
    Y = model(X)
    Ys = Y[:, idx.to(DEV)]
    params = list(model.parameters())
    C.golden_loss(Ys).backward()
    torch.cuda.synchronize()
    ps = [p.detach().cpu().double().requires_grad_(p.requires_grad) for p in params]
    Yo = O.forward(O.from_desc(desc), X[:, idx.to(DEV)].cpu().to(torch.complex128), ps, nfft, W.ALIAS_DECAY_DB, bins=idx)
    go = torch.autograd.grad(C.golden_loss(Yo), [p for p in ps if p.requires_grad])
    a, ao = np.abs(Ys.detach().cpu().numpy()), np.abs(Yo.detach().numpy())
    ferr = rel_err(a, ao)
    gerr, k = 0.0, 0
    for p in params:
        if p.requires_grad:
            assert p.grad is not None
            ref = go[k].numpy()
            gerr = max(gerr, float(np.abs(p.grad.cpu().numpy() - ref).max() / (np.abs(ref).max() + 1e-300)))
            k += 1
    FULL_SIZE_REPORT[name] = (ferr, gerr)
    print(f"[full-size parity] {name}: |Y| floored rel err {ferr:.3e}, grad err {gerr:.3e}, bins {len(idx)}")
    assert ferr <= 1e-4, f"{name}: magnitude rel err {ferr:.3e}"
    assert gerr <= 1e-3, f"{name}: gradient rel err {gerr:.3e}"


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("name", ["loop_in_feedforward", "loop_as_feedback", "parallel_in_feedback",
                                  "parallel_cat_as_feedforward"])
def test_nested_recursion_and_parallel_inside_a_loop(name, dtype):
    """A Recursion or a Parallel INSIDE a Recursion path (round 1 raised): the nested system's response enters the outer
    program as a streamed table (sweep.Program.table_of); response and gradients against the oracle on the kernels."""
    from nested_cases import NESTED

    desc = NESTED[name]
    nfft, alias = 512, 20.0
    M = nfft // 2 + 1
    torch.manual_seed(3)
    model = W.build(desc, dsp, system, nfft, alias, dtype=dtype, device=DEV)
    cdt = torch.complex64 if dtype == torch.float32 else torch.complex128
    X = C.make_input(2, M, model.input_channels, None).to(cdt).to(DEV)
    Y = model(X)
    C.golden_loss(Y).backward()
    ps = [p.detach().cpu().double().requires_grad_(p.requires_grad) for p in model.parameters()]
    Yo = O.forward(O.from_desc(desc), X.cpu().to(torch.complex128), ps, nfft, alias)
    go = torch.autograd.grad(C.golden_loss(Yo), [p for p in ps if p.requires_grad])
    ftol, gtol = (1e-4, 1e-3) if dtype == torch.float32 else (1e-9, 1e-6)
    assert rel_err(np.abs(Y.detach().cpu().numpy()), np.abs(Yo.detach().numpy())) <= ftol
    k = 0
    for p in model.parameters():
        if p.requires_grad:
            assert p.grad is not None
            assert float((p.grad.cpu().double() - go[k]).abs().max()) <= gtol * float(go[k].abs().max() + 1e-30)
            k += 1


@pytest.mark.parametrize("isint", [True, False])
def test_register_matrix_kernel_with_several_bins_per_thread(isint):
    """fsweep_tpr.cuh runs one block of 352 threads per SM: beyond 52 096 bins a thread owns several bins and ADDS into the
    accumulators its first bin wrote.  8 x 8 FDN at nfft = 400 000 (200 001 bins: up to four per thread), batch 2, with
    integer delays (the 32-bit phase-index path) and fractional ones (the generic chain): magnitudes against the
    oracle's per-bin closed form on a bin subset, gradients of a subset loss likewise."""
    nfft, B = 400000, 2
    M = nfft // 2 + 1
    desc = W.fdn(8, isint=isint)
    torch.manual_seed(11)
    model = W.build(desc, dsp, system, nfft, W.ALIAS_DECAY_DB, dtype=torch.float32, device=DEV)
    if not isint:
        with torch.no_grad():
            for m in model.modules():
                if isinstance(m, dsp.parallelDelay):
                    m.assign_value(m.sample2s(m.s2sample(m.param) + 0.37))
    X = C.make_input(B, M, model.input_channels, None).to(torch.complex64).to(DEV)
    idx = _subset(M, 401)
    Y = model(X)
    with torch.no_grad(), sweep.bin_shard(70000, 130001):  # a bin shard of the same launch family: bit-identical bins
        assert torch.equal(model(X), Y[:, 70000:130001])
    Ys = Y[:, idx.to(DEV)]
    params = list(model.parameters())
    C.golden_loss(Ys).backward()
    torch.cuda.synchronize()
    ps = [p.detach().cpu().double().requires_grad_(p.requires_grad) for p in params]
    Yo = O.forward(O.from_desc(desc), X[:, idx.to(DEV)].cpu().to(torch.complex128), ps, nfft, W.ALIAS_DECAY_DB, bins=idx)
    go = torch.autograd.grad(C.golden_loss(Yo), [p for p in ps if p.requires_grad], allow_unused=True)
    assert rel_err(np.abs(Ys.detach().cpu().numpy()), np.abs(Yo.detach().numpy())) <= 1e-4
    k = 0
    for p in params:
        if p.requires_grad:
            ref = go[k]
            k += 1
            if ref is None:
                continue
            assert p.grad is not None
            assert float(np.abs(p.grad.cpu().numpy() - ref.numpy()).max() / (np.abs(ref.numpy()).max() + 1e-300)) <= 1e-3
