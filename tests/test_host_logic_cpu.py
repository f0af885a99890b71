"""CPU: the host side of the product (module mirror, lowering to sweep programs, coefficient
packing, autograd plumbing) driven through the float64 emulator of the C ABI, against the golden
vectors of the unmodified reference."""
import numpy as np
import pytest
import torch

import cases as C
from helpers import build_case, grad_err, rel_err

pytestmark = pytest.mark.usefixtures("emulated_backend")


@pytest.mark.parametrize("name", list(C.CASES))
def test_case_matches_reference(name):
    case, g, model = build_case(name, torch.float64, "cpu")
    M = case["nfft"] // 2 + 1
    X = C.make_input(case["B"], M, model.input_channels, case["C"])
    Y = model(X)
    # cond(I - F Fb) ~ 5e5 for the lossless loop: float64 LU orderings differ at ~1e-9
    tol = max(1e-7 if case["alias"] == 0.0 else 1e-9, 1.05 * float(g["ref_fp32_noise"]))
    assert rel_err(Y.detach()[:, g["bins"]].numpy(), g["Y"]) <= tol
    params = list(model.parameters())
    if any(p.requires_grad for p in params) and "loss" in g and any(k.startswith("grad_") for k in g.files):
        loss = C.golden_loss(Y)
        loss.backward()
        for i, p in enumerate(params):
            if f"grad_{i}" in g.files:
                assert p.grad is not None, (name, i)
                assert grad_err(p.grad.numpy(), g[f"grad_{i}"]) <= max(1e-8, 20 * float(g["ref_fp32_noise"])), (name, i)


@pytest.mark.parametrize("crit_kind", ["mse_loss", "MSELoss"])
def test_trainer_fused_criterion_equals_unfused(crit_kind):
    """Trainer.train_step with the criterion fused into the sweep (one backward_loss call) takes the same steps
    as the unfused path (forward, criterion in PyTorch, backward) — host logic, through the ABI emulator."""
    from flamo_b200 import sweep, workloads as W
    from flamo_b200.optimize.loss import mse_loss, sparsity_loss
    from flamo_b200.optimize.trainer import Trainer
    from flamo_b200.processor import dsp, system

    nfft, M = 512, 257

    def run(fuse):
        torch.manual_seed(5)
        core = W.build(W.fdn(4, delays=[11, 17, 23, 31]), dsp, system, nfft, 30.0, dtype=torch.float64, device="cpu")
        model = system.Shell(core, dsp.FFT(nfft, dtype=torch.float64),
                             dsp.Transform(lambda x: torch.abs(x), dtype=torch.float64))
        tr = Trainer(model, max_epochs=1, lr=1e-2, log=False, device="cpu", fuse_criterion=fuse)
        if crit_kind == "mse_loss":
            tr.register_criterion(mse_loss(nfft=nfft), 1)
            tgt = torch.ones(2, M, 1, dtype=torch.float64)
        else:
            tr.register_criterion(torch.nn.MSELoss(), 0.7)
            tgt = torch.full((2, M, 1), 0.8, dtype=torch.float64)
        tr.register_criterion(sparsity_loss(), 0.2, requires_model=True)
        x = torch.zeros(2, nfft, 1, dtype=torch.float64)
        x[:, 0] = 1
        x[1, 3] = 0.5
        n0 = sweep.launch_count
        losses = [tr.train_step((x, tgt)) for _ in range(3)]
        return losses, [p.detach().clone() for p in model.parameters()], sweep.launch_count - n0, tr

    lf, pf, nf, trf = run(True)
    lu, pu, nu, tru = run(False)
    assert np.allclose(lf, lu, rtol=1e-10)
    for a, b in zip(pf, pu):
        assert torch.allclose(a, b, rtol=1e-9, atol=1e-12)
    assert nf < nu  # the fused step makes fewer sweep launches (no forward sweep)
    assert np.allclose(list(trf.train_loss_log.values())[0], list(tru.train_loss_log.values())[0], rtol=1e-10)


@pytest.mark.parametrize("n,scale", [(6, 1.0), (40, 1.0), (64, 4.0), (33, 1e-3)])
def test_capture_safe_expm_matches_matrix_exp(n, scale):
    """functional.expm_capturable (the orthogonal map of matrices wider than the one-CTA device kernel) against
    torch.matrix_exp and its autograd, float64 and float32 inputs."""
    from flamo_b200.functional import expm_capturable, skew_matrix

    torch.manual_seed(n)
    P = (torch.randn(n, n, dtype=torch.float64) * scale).requires_grad_(True)
    w = torch.randn(n, n, dtype=torch.float64)
    E, R = expm_capturable(skew_matrix(P)), torch.matrix_exp(skew_matrix(P))
    assert torch.allclose(E, R, rtol=1e-12, atol=1e-12)
    (g1,) = torch.autograd.grad((E * w).sum(), P)
    (g2,) = torch.autograd.grad((R * w).sum(), P)
    assert float((g1 - g2).abs().max()) <= 1e-9 * float(g2.abs().max())
    assert torch.allclose(E.T @ E, torch.eye(n, dtype=torch.float64), atol=1e-11)  # orthogonal
    E32 = expm_capturable(skew_matrix(P.detach().float()))
    assert E32.dtype == torch.float32 and torch.allclose(E32.double(), R.detach(), atol=5e-6)


def test_pack_sections_matches_definition():
    """sweep.pack_sections (one contraction with a constant 16 x 6 matrix) against the layout include/fsweep.h
    defines: per section {B(w0), B'(w0), b2, 0, A(w0), A'(w0), a2, 0} for w0 = +1 and w0 = -1."""
    from flamo_b200 import sweep

    torch.manual_seed(0)
    b, a = torch.randn(3, 2, 4, 3, dtype=torch.float64), torch.randn(3, 2, 4, 3, dtype=torch.float64)
    p = sweep.pack_sections(b, a, False, torch.float64)  # (K, N_in, N_out, 2, 8)
    assert p.shape == (2, 3, 4, 2, 8)
    for s in range(2):
        for m in range(4):
            for n in range(3):
                for blk, sg in ((0, 1.0), (1, -1.0)):
                    for off, t in ((0, b), (4, a)):
                        c0, c1, c2 = (float(t[i, s, m, n]) for i in range(3))
                        want = [c0 + sg * c1 + c2, c1 + 2 * sg * c2, c2, 0.0]
                        got = p[s, n, m, blk, off:off + 4].tolist()
                        assert np.allclose(got, want, rtol=1e-13, atol=1e-13)
    pp = sweep.pack_sections(b[..., 0], a[..., 0], True, torch.float32)  # parallel: (K, N, 2, 8)
    assert pp.shape == (2, 4, 2, 8) and pp.dtype == torch.float32
    assert np.allclose(pp.double().numpy(), p[:, 0].numpy(), rtol=1e-6, atol=1e-6)


def test_trainer_chooses_fused_path_only_when_it_is_safe():
    """The fused criterion is taken iff exactly ONE criterion consumes the prediction and it is an MSE; anything else
    (two consumers, a custom criterion, fuse_criterion=False, a Shell without |.| output) evaluates the prediction."""
    from flamo_b200 import workloads as W
    from flamo_b200.optimize.loss import mse_loss, sparsity_loss
    from flamo_b200.optimize.trainer import Trainer
    from flamo_b200.processor import dsp, system

    nfft, M = 256, 129

    def shell(abs_out=True):
        torch.manual_seed(3)
        core = W.build(W.fdn(4, delays=[7, 11, 13, 17]), dsp, system, nfft, 30.0, dtype=torch.float64, device="cpu")
        fn = (lambda x: torch.abs(x)) if abs_out else (lambda x: torch.abs(x) ** 2)
        out = dsp.Transform(fn, dtype=torch.float64)
        return system.Shell(core, dsp.FFT(nfft, dtype=torch.float64), out)

    x = torch.zeros(1, nfft, 1, dtype=torch.float64)
    x[:, 0] = 1
    tgt = torch.ones(1, M, 1, dtype=torch.float64)

    class Custom(torch.nn.Module):
        def forward(self, y_pred, y_true):
            return (y_pred - y_true).abs().mean()

    def fused(tr):
        est, done = tr._predict(x, tgt)
        return est is None

    t = Trainer(shell(), log=False, device="cpu")
    t.register_criterion(mse_loss(nfft=nfft), 1)
    t.register_criterion(sparsity_loss(), 0.2, requires_model=True)
    assert fused(t)
    t2 = Trainer(shell(), log=False, device="cpu")
    t2.register_criterion(mse_loss(nfft=nfft), 1)
    t2.register_criterion(Custom(), 1)  # a second consumer of the prediction
    assert not fused(t2)
    t3 = Trainer(shell(), log=False, device="cpu")
    t3.register_criterion(Custom(), 1)  # not an MSE the kernels know
    assert not fused(t3)
    t4 = Trainer(shell(), log=False, device="cpu", fuse_criterion=False)
    t4.register_criterion(mse_loss(nfft=nfft), 1)
    assert not fused(t4)
    t5 = Trainer(shell(abs_out=False), log=False, device="cpu")  # output layer is not |.|: declined, remembered
    t5.register_criterion(mse_loss(nfft=nfft), 1)
    assert not fused(t5) and len(t5._unfusable) == 1
    # every variant still trains
    for tr in (t, t2, t3, t4, t5):
        a = tr.train_step((x, tgt))
        b = tr.train_step((x, tgt))
        assert np.isfinite(a) and np.isfinite(b)


@pytest.mark.parametrize("tag,epoch", [("fdn", 0), ("fdn", 19), ("biquad", 0), ("biquad", 7)])
def test_shipped_checkpoints_load_and_respond(tag, epoch):
    """SURVEY §8c fixtures through the host side (load_state_dict with the reference's keys, Shell.get_freq_response /
    get_time_response layer swapping) and the ABI emulator; all 28 checkpoints run on the GPU
    (tests/test_gpu_zz_checkpoints.py)."""
    from helpers import check_checkpoint

    check_checkpoint(tag, epoch, torch.float64, "cpu", tol_mag=1e-9, tol_resp=1e-9)


def test_entry_shape_check_follows_the_first_leaf():
    """Found by tests/test_random_trees_cpu.py: a Series that STARTS with a Recursion checked the input against the
    next module that owns check_input_shape (a later one, with other widths).  The check belongs to the first leaf the
    signal meets; wrong widths are rejected before any launch, also for a bare Recursion."""
    from flamo_b200.processor import dsp, system

    kw = dict(nfft=64, dtype=torch.float64)
    rec = system.Recursion(fF=dsp.Gain(size=(2, 1), **kw),
                           fB=system.Series(dsp.Gain(size=(1, 2), **kw), dsp.parallelGain(size=(1,), **kw)))
    rec.feedback[1].assign_value(torch.tensor([0.02], dtype=torch.float64))
    s = system.Series(rec, dsp.Delay(size=(1, 2), max_len=5, **kw))
    X = C.make_input(1, 33, 1, None)
    assert s(X).shape == (1, 33, 1)
    with pytest.raises(ValueError):
        s(C.make_input(1, 33, 2, None))      # the Recursion's feedforward Gain reads ONE channel
    with pytest.raises(ValueError):
        rec(C.make_input(1, 33, 3, None))
    with pytest.raises(ValueError):
        rec(C.make_input(1, 40, 1, None))    # more bins than nfft // 2 + 1
    shell = system.Shell(s)
    with pytest.raises(ValueError):
        shell(C.make_input(1, 33, 2, None))


def test_long_series_splits_into_launches_of_at_most_24_ops():
    """A Series of 30 leaves exceeds the kernels' step table (24 ops per launch, include/fsweep.h): the lowering
    splits it into two launches; response and gradients equal the oracle's."""
    from flamo_b200 import sweep, workloads as W
    from flamo_b200.processor import dsp, system
    from oracle import flamo_oracle as O

    nfft = 128
    leaves = []
    for i in range(30):
        if i % 3 == 0:
            leaves.append(("Gain", dict(size=(2, 2), requires_grad=True)))
        elif i % 3 == 1:
            leaves.append(("parallelDelay", dict(size=(2,), max_len=20, isint=bool(i % 2), fs=48000,
                                                 requires_grad=not bool(i % 2))))
        else:
            leaves.append(("parallelGain", dict(size=(2,), requires_grad=True)))
    desc = ("Series", leaves)
    torch.manual_seed(3)
    model = W.build(desc, dsp, system, nfft, 30.0, dtype=torch.float64, device="cpu")
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() == 2:
                p.copy_(torch.eye(2, dtype=torch.float64) + 0.1 * p)  # keep the 10 dense gains well scaled
    prog = sweep.Program(nfft, 30.0, torch.complex128, "cpu")
    model._lower(prog, None)
    segs = list(prog._segments())
    assert [t for t, _ in segs] == ["sweep", "sweep"] and all(len(p) <= sweep.MAX_OPS_PER_LAUNCH for _, p in segs)
    X = C.make_input(2, nfft // 2 + 1, 2, None)
    Y = model(X)
    ps = [p.detach().clone().requires_grad_(p.requires_grad) for p in model.parameters()]
    Yo = O.forward(O.from_desc(desc), X, ps, nfft, 30.0)
    assert rel_err(Y.detach().numpy(), Yo.detach().numpy()) <= 1e-9
    C.golden_loss(Y).backward()
    gp = [p for p in ps if p.requires_grad]
    go = torch.autograd.grad(C.golden_loss(Yo), gp)
    mine = [p.grad for p in model.parameters() if p.requires_grad]
    scale = max(float(g.abs().max()) for g in go)
    for a, b in zip(mine, go):
        assert float((a - b).abs().max()) <= 1e-8 * scale
