"""GPU: the criterion fused into the sweep (fsweep_forward_loss / fsweep_backward_loss, include/fsweep.h) against
the unfused path — Shell output |Y| from fsweep_forward, the criterion in PyTorch, gradients from fsweep_backward —
which test_gpu_parity.py pins to the oracle.  Both criterion kinds (nn.MSELoss on the magnitudes; flamo's mse_loss
with its channel sum, reference optimize/loss.py:90-103), every kernel family, batch 1 and batch > 1."""
import numpy as np
import pytest
import torch

import cases as C
from flamo_b200 import _lib, sweep
from flamo_b200.optimize.loss import mse_loss
from flamo_b200.processor import dsp, system
from helpers import build_case, grad_err

pytestmark = pytest.mark.gpu

# (a Parallel node is a launch boundary: such programs decline the fused path, see test_unfusable_programs_decline)
NAMES = [n for n, c in C.CASES.items() if c["C"] is None and c["grads"] and not n.startswith("parallel_")]


def target_for(kind, B, M, n_out, dtype):
    k = torch.arange(M, dtype=torch.float64).view(1, M, 1)
    b = torch.arange(B, dtype=torch.float64).view(B, 1, 1)
    r = torch.arange(n_out if kind == _lib.CRIT_MSE else 1, dtype=torch.float64).view(1, 1, -1)
    return (1.0 + 0.4 * torch.sin(0.05 * k + 0.9 * b + 0.3 * r)).to(dtype).cuda()


def both_paths(name, dtype, kind, B=None):
    case, g, core = build_case(name, dtype, "cuda")
    nfft = case["nfft"]
    M = nfft // 2 + 1
    B = B or case["B"]
    shell = system.Shell(core, output_layer=dsp.Transform(lambda x: torch.abs(x), dtype=dtype))
    cdt = torch.complex64 if dtype == torch.float32 else torch.complex128
    X = C.make_input(B, M, core.input_channels, None).to(cdt).cuda()
    tgt = target_for(kind, B, M, core.output_channels, dtype)
    crit = mse_loss() if kind == _lib.CRIT_MSE_CHSUM else torch.nn.MSELoss()
    ps = [p for p in shell.parameters() if p.requires_grad]
    lu = crit(shell(X), tgt)
    gu = torch.autograd.grad(lu, ps, allow_unused=True)
    n0 = sweep.launch_count
    lf = shell.forward_loss(X, tgt, kind)
    assert lf is not None, "the fused path declined this program"
    gf = torch.autograd.grad(2.5 * lf, ps, allow_unused=True)  # upstream gradient != 1 on purpose
    # one sweep launch + finalize (+ the device expm forward / backward of an orthogonal Matrix map, or the
    # accumulator memset, response-table and deferred-gradient launches of a large section cascade)
    assert sweep.launch_count - n0 <= 5 + 2 * sum(1 for _ in ps), "fused loss + gradients must be ONE sweep launch"
    with torch.no_grad():
        lv = shell.forward_loss(X, tgt, kind)  # loss-only entry point (validation)
    return float(lu.detach()), float(lf.detach()), float(lv), gu, gf


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("kind", [_lib.CRIT_MSE, _lib.CRIT_MSE_CHSUM])
def test_fused_equals_unfused_c64(name, kind):
    lu, lf, lv, gu, gf = both_paths(name, torch.float32, kind)
    assert abs(lf - lu) <= 2e-5 * abs(lu) and abs(lv - lu) <= 2e-5 * abs(lu)
    for a, b in zip(gu, gf):
        assert (a is None) == (b is None)
        if a is not None:
            assert grad_err(b.cpu().numpy() / 2.5, a.cpu().numpy()) <= 2e-4


@pytest.mark.parametrize("name", [n for n in NAMES if n != "cfg5_fdn64_small"])
@pytest.mark.parametrize("kind", [_lib.CRIT_MSE, _lib.CRIT_MSE_CHSUM])
def test_fused_equals_unfused_c128(name, kind):
    lu, lf, lv, gu, gf = both_paths(name, torch.float64, kind)
    assert abs(lf - lu) <= 1e-12 * abs(lu) and abs(lv - lu) <= 1e-12 * abs(lu)
    # cond(I - F Fb) ~ 5e5 for the lossless loop: |.| formed from a recomputed Y differs at ~1e-9
    gtol = 1e-7 if C.CASES[name]["alias"] == 0.0 else 1e-10
    for a, b in zip(gu, gf):
        if a is not None:
            assert grad_err(b.cpu().numpy() / 2.5, a.cpu().numpy()) <= gtol


@pytest.mark.parametrize("path", ["generic", "loop", "tpb", "tpc", "tpc_v1"])
@pytest.mark.parametrize("kind", [_lib.CRIT_MSE, _lib.CRIT_MSE_CHSUM])
@pytest.mark.parametrize("B", [1, 5])
def test_fused_on_every_kernel_family(path, kind, B, monkeypatch):
    if path in ("tpb", "tpc", "tpc_v1"):
        monkeypatch.setenv("FSWEEP_FORCE_TPB" if path == "tpb" else "FSWEEP_FORCE_TPC", "1")
        if path == "tpc_v1":
            monkeypatch.setenv("FSWEEP_TPC_V1", "1")
    else:
        monkeypatch.setenv("FSWEEP_DISABLE_TPB", "1")
    if path == "generic":
        monkeypatch.setenv("FSWEEP_DISABLE_LOOP_KERNEL", "1")
    saved = dict(sweep._PLANS)
    sweep._PLANS.clear()
    try:
        lu, lf, lv, gu, gf = both_paths("fdn8_batch3", torch.float32, kind, B=B)
    finally:
        sweep._PLANS.clear()
        sweep._PLANS.update(saved)
    assert abs(lf - lu) <= 2e-5 * abs(lu) and abs(lv - lu) <= 2e-5 * abs(lu)
    for a, b in zip(gu, gf):
        if a is not None:
            assert grad_err(b.cpu().numpy() / 2.5, a.cpu().numpy()) <= 2e-4


def test_fused_respects_bin_shards():
    """Sum over shards of (shard loss * shard share) equals the whole, and so do the gradients."""
    case, g, core = build_case("fdn8_batch3", torch.float32, "cuda")
    M = case["nfft"] // 2 + 1
    shell = system.Shell(core, output_layer=dsp.Transform(lambda x: torch.abs(x)))
    X = C.make_input(3, M, 1, None).to(torch.complex64).cuda()
    tgt = target_for(_lib.CRIT_MSE_CHSUM, 3, M, 1, torch.float32)
    ps = [p for p in shell.parameters() if p.requires_grad]
    whole = shell.forward_loss(X, tgt, _lib.CRIT_MSE_CHSUM)
    gw = torch.autograd.grad(whole, ps)
    cuts = [0, 333, 334, 1200, M]
    tot, gs = 0.0, [torch.zeros_like(t) for t in gw]
    for b0, b1 in zip(cuts[:-1], cuts[1:]):
        with sweep.bin_shard(b0, b1):
            part = shell.forward_loss(X, tgt, _lib.CRIT_MSE_CHSUM) * ((b1 - b0) / M)
        tot += float(part.detach())
        for acc, t in zip(gs, torch.autograd.grad(part, ps)):
            acc += t
    assert abs(tot - float(whole)) <= 1e-5 * abs(float(whole))
    for a, b in zip(gs, gw):
        assert grad_err(a.cpu().numpy(), b.cpu().numpy()) <= 1e-4


def test_unfusable_programs_decline():
    case, g, core = build_case("series_trailing_cols", torch.float32, "cuda")
    M = case["nfft"] // 2 + 1
    shell = system.Shell(core, output_layer=dsp.Transform(lambda x: torch.abs(x)))
    X = C.make_input(2, M, core.input_channels, 3).to(torch.complex64).cuda()
    assert shell.forward_loss(X, torch.ones(2, M, 2, 3, device="cuda"), _lib.CRIT_MSE) is None  # trailing columns
    X3 = C.make_input(2, M, core.input_channels, None).to(torch.complex64).cuda()
    assert shell.forward_loss(X3, torch.ones(2, M, 5, device="cuda"), _lib.CRIT_MSE) is None  # wrong target shape
    case, g, pcore = build_case("parallel_sum", torch.float32, "cuda")
    pshell = system.Shell(pcore, output_layer=dsp.Transform(lambda x: torch.abs(x)))
    Xp = C.make_input(2, case["nfft"] // 2 + 1, 2, None).to(torch.complex64).cuda()
    assert pshell.forward_loss(Xp, torch.ones(2, case["nfft"] // 2 + 1, 3, device="cuda"), _lib.CRIT_MSE) is None
    shell2 = system.Shell(core, output_layer=dsp.Transform(lambda x: x))
    assert shell2.forward_loss(X3, torch.ones(2, M, 2, device="cuda"), _lib.CRIT_MSE) is None  # no |.| output layer
