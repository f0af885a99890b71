"""GPU: the CUDA sweep (through the module API -> C ABI -> libfsweep.so) against the float64
oracle on the same parameters, and against the golden vectors of the unmodified reference.

Tolerances: forward 1e-4 relative on the response with the denominator floored
at 1e-3 * max|H|; parameter gradients 1e-3 relative to the largest gradient entry.  The float64
instantiation of the kernels must agree to ~1e-9.
"""
import numpy as np
import pytest
import torch

import cases as C
from flamo_b200 import sweep
from helpers import build_case, golden_params, grad_err, load_golden, rel_err
from oracle import flamo_oracle as O

pytestmark = pytest.mark.gpu



def oracle_on(case, params64, X64):
    node = O.from_desc(case["desc"])
    ps = [p.clone().requires_grad_(r) for p, r in params64]
    Y = O.forward(node, X64, ps, case["nfft"], case["alias"])
    gp = [p for p in ps if p.requires_grad]
    grads = {}
    if gp:
        loss = C.golden_loss(Y)
        gs = torch.autograd.grad(loss, gp)
        k = 0
        for i, p in enumerate(ps):
            if p.requires_grad:
                grads[i] = gs[k].numpy()
                k += 1
    return Y.detach().numpy(), grads


def run_case(name, dtype):
    assert sweep._BACKEND.name == "cuda"
    case, g, model = build_case(name, dtype, "cuda")
    M = case["nfft"] // 2 + 1
    cdt = torch.complex64 if dtype == torch.float32 else torch.complex128
    X64 = C.make_input(case["B"], M, model.input_channels, case["C"])
    X = X64.to(cdt).cuda()
    params = list(model.parameters())
    Y = model(X)
    loss = C.golden_loss(Y)
    if any(p.requires_grad for p in params):
        loss.backward()
    torch.cuda.synchronize()
    # oracle on exactly the parameters / input the device saw (after rounding to `dtype`)
    p64 = [(p.detach().cpu().double(), p.requires_grad) for p in params]
    Yo, go = oracle_on(case, p64, X.cpu().to(torch.complex128))
    Yd = Y.detach().cpu().numpy()
    ferr = rel_err(Yd, Yo)
    run_case.mag_err = rel_err(np.abs(Yd), np.abs(Yo))  # the north-star metric: error of the magnitude response
    run_case.peak_err = float(np.abs(np.abs(Yd) - np.abs(Yo)).max() / np.abs(Yo).max())  # relative to the peak
    gerrs = {i: grad_err(params[i].grad.cpu().numpy(), go[i]) for i in go if params[i].grad is not None}
    missing = [i for i in go if params[i].grad is None]
    return case, g, Y, ferr, gerrs, missing


@pytest.mark.parametrize("name", [n for n in C.CASES])
def test_c64_vs_oracle(name):
    case, g, Y, ferr, gerrs, missing = run_case(name, torch.float32)
    ftol = 5e-3 if case["alias"] == 0.0 else 1e-4  # lossless loop: cond ~5e5 (SURVEY §7 "Conditioning")
    assert run_case.peak_err <= (5e-4 if case["alias"] == 0.0 else 2e-6), f"peak-relative err {run_case.peak_err:.3e}"
    assert not missing, f"no gradient for params {missing}"
    # BASELINE.json: "float32 match ... within 1e-4 rel on magnitude response"; the complex-valued
    # error (which also counts phase) is held to 2x that
    assert run_case.mag_err <= ftol, f"magnitude rel err {run_case.mag_err:.3e}"
    assert ferr <= 2 * ftol, f"complex rel err {ferr:.3e}"
    gtol = 5e-2 if case["alias"] == 0.0 else 1e-3
    for i, e in gerrs.items():
        assert e <= gtol, f"grad of param {i}: rel err {e:.3e}"


@pytest.mark.parametrize("name", ["cfg4_active_full", "cfg3_geq16_small", "fdn16", "fdn32"])
def test_pure_float32_arithmetic_on_promoted_shapes(name, monkeypatch):
    """Loops wider than 8 channels and long dense cascades are swept in float64 arithmetic by default
    (sweep._wants_f64: float32 arithmetic reads 1.3e-4 on config 4, above the 1e-4 bar).  The float32 kernels for these
    shapes stay reachable (FLAMO_B200_PRECISION=float32) and are held to float32 resolution of the PEAK here."""
    monkeypatch.setenv("FLAMO_B200_PRECISION", "float32")
    case, g, Y, ferr, gerrs, missing = run_case(name, torch.float32)
    assert not missing
    assert run_case.peak_err <= 2e-6 and run_case.mag_err <= 2.5e-4
    assert all(e <= 1e-3 for e in gerrs.values())


@pytest.mark.parametrize("name", [n for n in C.CASES])
def test_c128_vs_oracle_and_golden(name):
    case, g, Y, ferr, gerrs, missing = run_case(name, torch.float64)
    tol = 1e-6 if case["alias"] == 0.0 else 1e-9
    assert not missing
    assert ferr <= tol, f"forward rel err {ferr:.3e}"
    for i, e in gerrs.items():
        assert e <= 1e3 * tol, f"grad of param {i}: rel err {e:.3e}"
    # and directly against the reference's stored outputs (its own float32 internals allowed for)
    ref_tol = max(10 * tol, 1.05 * float(g["ref_fp32_noise"]))
    assert rel_err(Y.detach().cpu().numpy()[:, g["bins"]], g["Y"]) <= ref_tol


FDN_CASES = ["cfg2_fdn8_full", "fdn6_example", "fdn8_batch3", "fdn16", "fdn32", "fdn8_fracdelay", "recursion_filters"]


def _family_combos():
    """(case, dtype, kernel family) combinations that exist: the thread-per-bin families are float32, width <= 8, and
    the compact one needs the N x 1 / 1 x N gains around the loop."""
    out = []
    for name in FDN_CASES:
        for dtype in (torch.float32, torch.float64):
            for path in ("generic", "loop", "tpb", "tpc", "tpc_v1"):
                if path in ("tpb", "tpc", "tpc_v1") and (dtype != torch.float32 or name in ("fdn16", "fdn32")):
                    continue
                if path in ("tpc", "tpc_v1") and name == "recursion_filters":
                    continue
                out.append(pytest.param(name, dtype, path, id=f"{path}-{str(dtype).split('.')[-1]}-{name}"))
    return out


@pytest.mark.parametrize("name,dtype,path", _family_combos())
def test_alternative_kernel_paths_on_fdn_cases(name, dtype, path, monkeypatch):
    """FDN-shaped programs can run on four kernel families: the generic step-table interpreter
    (fsweep_kernels.cuh), the row-distributed loop kernels (fsweep_loop.cuh) and, for widths <= 8 in float32 and
    enough bins, the unrolled (fsweep_tpb.cuh) and the compact thread-per-bin kernels — "tpc": the default of that
    family, the matrix in registers for widths 5..8 (fsweep_tpr.cuh); "tpc_v1": the matrix in shared memory
    (fsweep_tpc.cuh, FSWEEP_TPC_V1=1).  Each family is forced here in turn; all must agree with the oracle."""
    monkeypatch.setenv("FLAMO_B200_PRECISION", "float32")  # the float32 kernels themselves, width 16 / 32 included
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
        case, g, Y, ferr, gerrs, missing = run_case(name, dtype)
    finally:
        sweep._PLANS.clear()
        sweep._PLANS.update(saved)
    assert not missing
    if dtype == torch.float32:
        assert run_case.mag_err <= 1e-4 and all(e <= 1e-3 for e in gerrs.values())
    else:
        assert ferr <= 1e-9 and all(e <= 1e-6 for e in gerrs.values())


@pytest.mark.parametrize("name", ["cfg3_geq16_small", "geq_oct3", "pgeq_oct1", "cfg4_active_full"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_deferred_sos_gradients_match_in_kernel_atomics(name, dtype, monkeypatch):
    """Section cascades whose accumulators exceed shared memory get their coefficient gradient from
    fsweep_sos_defer_kernel (one block per channel pair, registers across bins).  It must reproduce the in-kernel
    global-atomic path (FSWEEP_DISABLE_DEFER=1) — on the whole spectrum, on bin shards that straddle the
    cos(w) = 0 boundary between the two Taylor blocks, and with several batch items."""
    monkeypatch.setenv("FLAMO_B200_PRECISION", "float32")  # float32 models on the float32 kernels

    def grads(disable, shard=None, B=None):
        if disable:
            monkeypatch.setenv("FSWEEP_DISABLE_DEFER", "1")
        else:
            monkeypatch.delenv("FSWEEP_DISABLE_DEFER", raising=False)
        saved = dict(sweep._PLANS)
        sweep._PLANS.clear()
        try:
            case, g, model = build_case(name, dtype, "cuda")
            M = case["nfft"] // 2 + 1
            cdt = torch.complex64 if dtype == torch.float32 else torch.complex128
            X = C.make_input(B or case["B"], M, model.input_channels, None).to(cdt).cuda()
            ps = [p for p in model.parameters() if p.requires_grad]
            if shard is None:
                out = torch.autograd.grad(C.golden_loss(model(X)), ps)
            else:
                with sweep.bin_shard(*shard):
                    out = torch.autograd.grad(C.golden_loss(model(X)), ps)
            return [t.cpu().numpy() for t in out]
        finally:
            sweep._PLANS.clear()
            sweep._PLANS.update(saved)

    M = C.CASES[name]["nfft"] // 2 + 1
    tol = 2e-4 if dtype == torch.float32 else 1e-10
    for shard, B in ((None, None), ((M // 4 - 37, M // 2 + 501), 3), ((M // 2 + 3, M), None)):
        a, b = grads(False, shard, B), grads(True, shard, B)
        for u, v in zip(a, b):
            assert grad_err(u, v) <= tol, (shard, B)


@pytest.mark.parametrize("N,B", [(40, 1), (64, 3), (48, 35)])
def test_wide_fdn_in_float64(N, B):
    """float64 models with a 33..64-wide FDN loop (the reference's examples default to float64) run on the float64
    instantiation of the CTA-per-bin kernels: response and gradients against the oracle at float64 resolution."""
    from flamo_b200 import workloads as W
    from flamo_b200.processor import dsp, system

    nfft, alias = 1024, 30.0
    M = nfft // 2 + 1
    desc = W.fdn(N, delays=[601 + 37 * i for i in range(N)])
    torch.manual_seed(11)
    model = W.build(desc, dsp, system, nfft, alias, dtype=torch.float64, device="cuda")
    X = C.make_input(B, M, 1, None).cuda().requires_grad_(True)
    Y = model(X)
    ps = [p for p in model.parameters() if p.requires_grad]
    gs = torch.autograd.grad(C.golden_loss(Y), ps + [X])
    fam = next(iter(pl.kernel_family(M, True) for pl in sweep._PLANS.values() if pl.dtype == 1 and "cta" in pl.kernel_family(M, True)), None)
    assert fam is not None and "tcgen05" not in fam
    params64 = [p.detach().cpu().double().requires_grad_(p.requires_grad) for p in model.parameters()]
    Xo = C.make_input(B, M, 1, None).requires_grad_(True)
    Yo = O.forward(O.from_desc(desc), Xo, params64, nfft, alias)
    go = torch.autograd.grad(C.golden_loss(Yo), [p for p in params64 if p.requires_grad] + [Xo])
    assert rel_err(Y.detach().cpu().numpy(), Yo.detach().numpy()) <= 1e-9
    for u, v in zip(gs, go):
        assert grad_err(u.detach().cpu().numpy().view(np.float64) if u.is_complex() else u.cpu().numpy(),
                        v.detach().numpy().view(np.float64) if v.is_complex() else v.numpy()) <= 1e-6


@pytest.mark.parametrize("tc", [False, True], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("N,B", [(40, 1), (64, 3), (48, 35)])
def test_cta_per_bin_kernels_on_wide_fdn(N, B, tc, monkeypatch):
    """Wide FDN loops (32 < N <= 64, float32) run on the CTA-per-bin kernels (fsweep_cta.cuh): against the oracle,
    and against the row-distributed two-warps-per-bin path (FSWEEP_DISABLE_CTA=1) — forward, input gradient and
    parameter gradients, with a padded width (N < 64) and more than one pass of 32 right-hand sides (B = 35)."""
    from flamo_b200 import workloads as W
    from flamo_b200.processor import dsp, system

    nfft, alias = 1024, 30.0
    M = nfft // 2 + 1
    delays = [601 + 37 * i for i in range(N)]
    desc = W.fdn(N, delays=delays)

    monkeypatch.setenv("FSWEEP_CTA_TC", "1" if tc else "0")  # tensor-core (fsweep_tc.cuh) or SIMT elimination

    def run(disable):
        if disable:
            monkeypatch.setenv("FSWEEP_DISABLE_CTA", "1")
        else:
            monkeypatch.delenv("FSWEEP_DISABLE_CTA", raising=False)
        saved = dict(sweep._PLANS)
        sweep._PLANS.clear()
        try:
            torch.manual_seed(11)
            model = W.build(desc, dsp, system, nfft, alias, dtype=torch.float32, device="cuda")
            X = C.make_input(B, M, 1, None).to(torch.complex64).cuda().requires_grad_(True)
            Y = model(X)
            ps = [p for p in model.parameters() if p.requires_grad]
            gs = torch.autograd.grad(C.golden_loss(Y), ps + [X])
            plan = next(iter(sweep._PLANS.values()))
            fam = plan.kernel_family(M, True)
            return model, Y.detach(), [g.detach() for g in gs], fam
        finally:
            sweep._PLANS.clear()
            sweep._PLANS.update(saved)

    model, Ya, ga, fam_a = run(False)
    _, Yb, gb, fam_b = run(True)
    assert "cta" in fam_a and "cta" not in fam_b and ("tcgen05" in fam_a) == tc
    e_ab = rel_err(Ya.cpu().numpy(), Yb.cpu().numpy())
    assert e_ab <= 5e-5, f"cta vs row-distributed: {e_ab:.3e}"  # two float32 paths, each held to 1e-4 below
    for u, v in zip(ga[:-1], gb[:-1]):
        assert grad_err(u.cpu().numpy(), v.cpu().numpy()) <= 2e-4
    gxa, gxb = ga[-1].cpu().numpy(), gb[-1].cpu().numpy()  # complex input gradient
    assert np.abs(gxa - gxb).max() <= 2e-4 * np.abs(gxb).max()
    # oracle (float64) on the same rounded parameters
    params64 = [p.detach().cpu().double().requires_grad_(p.requires_grad) for p in model.parameters()]
    X64 = C.make_input(B, M, 1, None)
    Yo = O.forward(O.from_desc(desc), X64.to(torch.complex64).to(torch.complex128), params64, nfft, alias)
    go = torch.autograd.grad(C.golden_loss(Yo), [p for p in params64 if p.requires_grad])
    assert rel_err(np.abs(Ya.cpu().numpy()), np.abs(Yo.detach().numpy())) <= 1e-4
    for u, v in zip(ga[:-1], go):
        assert grad_err(u.cpu().numpy(), v.numpy()) <= 1e-3


@pytest.mark.parametrize("B,cols", [(1, None), (1, 4), (2, 2), (3, None)])
@pytest.mark.parametrize("tma", [False, True, None], ids=["cp.async", "tma", "reg"])
def test_streaming_table_kernels(B, cols, tma, monkeypatch):
    """TABLE-heavy programs without recursion (FIR filter banks + gains, the shape of the real
    examples/e8_active_acoustics.py path) run on the streaming kernels (fsweep_stream.cuh) when batch*cols is a power of
    two <= 16: against the generic interpreter (FSWEEP_DISABLE_STREAM=1) and the oracle — outputs, table / gain
    gradients, input gradient, whole spectrum and a bin shard."""
    from flamo_b200 import workloads as W
    from flamo_b200.processor import dsp, system

    monkeypatch.setenv("FSWEEP_STREAM_TMA", "1" if tma else "0")  # bulk-copy ring (UBLKCP + mbarrier) or cp.async tiles
    # tma None: the forward pass on the register-state kernel (fsweep_streamr.cuh, the default)
    monkeypatch.setenv("FSWEEP_STREAM_FWD", "reg" if tma is None else "smem")
    nfft, alias = 2048, 30.0
    M = nfft // 2 + 1
    desc = ("Series", [
        ("Filter", dict(size=(20, 5, 3), requires_grad=True)),
        ("parallelFilter", dict(size=(33, 5), requires_grad=True)),
        ("parallelGain", dict(size=(5,), requires_grad=True)),
        ("Gain", dict(size=(4, 5), requires_grad=False)),
        ("Filter", dict(size=(12, 2, 4), requires_grad=True)),
    ])

    def run(disable, shard=None):
        if disable:
            monkeypatch.setenv("FSWEEP_DISABLE_STREAM", "1")
        else:
            monkeypatch.delenv("FSWEEP_DISABLE_STREAM", raising=False)
        saved = dict(sweep._PLANS)
        sweep._PLANS.clear()
        try:
            torch.manual_seed(21)
            model = W.build(desc, dsp, system, nfft, alias, dtype=torch.float32, device="cuda")
            X = C.make_input(B, M, 3, cols).to(torch.complex64).cuda().requires_grad_(True)
            ps = [p for p in model.parameters() if p.requires_grad]
            if shard is None:
                Y = model(X)
            else:
                with sweep.bin_shard(*shard):
                    Y = model(X)
            gs = torch.autograd.grad((Y.abs() ** 2).mean(), ps + [X])
            fam = [pl.kernel_family(M, True) for pl in sweep._PLANS.values()]
            return model, Y.detach(), [g.detach() for g in gs], fam
        finally:
            sweep._PLANS.clear()
            sweep._PLANS.update(saved)

    q = B * (cols or 1)
    streamable = q & (q - 1) == 0
    for shard in (None, (301, 777)):
        model, Ya, ga, fam_a = run(False, shard)
        _, Yb, gb, fam_b = run(True, shard)
        assert any("stream" in f for f in fam_a) and not any("stream" in f for f in fam_b)
        assert np.abs(Ya.cpu().numpy() - Yb.cpu().numpy()).max() <= 2e-6 * np.abs(Yb.cpu().numpy()).max()
        for u, v in zip(ga, gb):
            u, v = u.cpu().numpy(), v.cpu().numpy()
            assert np.abs(u - v).max() <= 2e-5 * np.abs(v).max(), (shard, streamable)
    # oracle on the whole spectrum
    model, Ya, ga, _ = run(False)
    params64 = [p.detach().cpu().double().requires_grad_(p.requires_grad) for p in model.parameters()]
    X64 = C.make_input(B, M, 3, cols).to(torch.complex64).to(torch.complex128).requires_grad_(True)
    Yo = O.forward(O.from_desc(desc), X64, params64, nfft, alias)
    go = torch.autograd.grad((Yo.abs() ** 2).mean(), [p for p in params64 if p.requires_grad] + [X64])
    assert rel_err(Ya.cpu().numpy(), Yo.detach().numpy()) <= 1e-4
    for u, v in zip(ga, go):
        u, v = u.cpu().numpy(), v.numpy()
        assert np.abs(u - v).max() <= 1e-3 * np.abs(v).max()


@pytest.mark.parametrize("widths", [(4, 13), (2, 4), (7, 8), (16, 16), (1, 3)])
@pytest.mark.parametrize("B,cols", [(1, 4), (1, None), (2, 8)])
def test_forward_streaming_kernels_widths(widths, B, cols, monkeypatch):
    """The forward streaming kernels — thread per (bin, column) with register arrays of 4, 8 or 16 entries
    (fsweep_streamr.cuh, default; tiles by cp.async and by bulk copies) and one warp per bin with the signal distributed
    over the lanes (fsweep_streamw.cuh, opt-in) — on every width class, ragged widths, bin shards that are not tile
    aligned and a partial tile, against the shared-memory-state kernel and the generic interpreter."""
    from flamo_b200 import workloads as W
    from flamo_b200.processor import dsp, system

    n_m, n_l = widths
    nfft, alias = 4096, 30.0
    M = nfft // 2 + 1
    desc = ("Series", [
        ("Filter", dict(size=(10, n_l, n_m), requires_grad=False)),
        ("parallelFilter", dict(size=(30, n_l), requires_grad=False)),
        ("parallelGain", dict(size=(n_l,), requires_grad=False)),
        ("Gain", dict(size=(n_l, n_l), requires_grad=False)),
        ("Filter", dict(size=(15, n_m, n_l), requires_grad=False)),
    ])

    def run(env, shard):
        for k in ("FSWEEP_DISABLE_STREAM", "FSWEEP_STREAM_FWD", "FSWEEP_STREAM_REG_TMA"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        saved = dict(sweep._PLANS)
        sweep._PLANS.clear()
        try:
            torch.manual_seed(5)
            model = W.build(desc, dsp, system, nfft, alias, dtype=torch.float32, device="cuda")
            X = C.make_input(B, M, n_m, cols).to(torch.complex64).cuda()
            with torch.no_grad():
                if shard is None:
                    Y = model(X)
                else:
                    with sweep.bin_shard(*shard):
                        Y = model(X)
            fam = [pl.kernel_family(M, False) for pl in sweep._PLANS.values()]
            return Y, fam
        finally:
            sweep._PLANS.clear()
            sweep._PLANS.update(saved)

    for shard in (None, (3, 1999), (1030, 1041)):
        Yw, fam_w = run({"FSWEEP_STREAM_FWD": "warp"}, shard)
        Yr, fam_r = run({"FSWEEP_STREAM_FWD": "reg"}, shard)
        Yt, _ = run({"FSWEEP_STREAM_FWD": "reg", "FSWEEP_STREAM_REG_TMA": "1"}, shard)
        Ys, _ = run({"FSWEEP_STREAM_FWD": "smem"}, shard)
        Yg, fam_g = run({"FSWEEP_DISABLE_STREAM": "1"}, shard)
        assert any("streamr" in f for f in fam_r) and not any("stream" in f for f in fam_g)
        ref = Yg.abs().max()
        for Y in (Yw, Yr, Yt, Ys):
            assert float((Y - Yg).abs().max()) <= 2e-6 * float(ref), (shard, widths)


@pytest.mark.parametrize("scale", [0.3, 2.0, 6.0])
def test_register_matrix_kernel_on_loops_that_need_row_interchanges(scale, monkeypatch):
    """fsweep_tpr.cuh eliminates without row interchanges as long as every diagonal pivot is at least a quarter of the
    largest candidate below it — always the case for an FDN (orthogonal feedback, |D| < 1) — and finishes a bin that
    violates this with partial pivoting in a slow path.  A general feedback matrix with entries of size `scale` makes
    that path run (numpy on these matrices: an interchange is wanted at > 30 % of the bins for scale >= 2, at none for
    0.3): the results must agree with the float64 oracle as well as those of the shared-memory kernel
    (FSWEEP_TPC_V1=1), which pivots at every step."""
    from flamo_b200 import workloads as W
    from flamo_b200.processor import dsp, system
    from oracle import flamo_oracle as O

    nfft, N = 65536, 8
    g = torch.Generator().manual_seed(5)
    Wfb = (scale * torch.randn(N, N, generator=g, dtype=torch.float64)).tolist()
    desc = ("Series", [
        ("Gain", dict(size=(N, 1), requires_grad=True)),
        ("Recursion", ("parallelDelay", dict(size=(N,), max_len=3000, isint=True, requires_grad=False),
                       {"delay_samples": [float(d) for d in W.fdn_delays(N)]}),
         ("Gain", dict(size=(N, N), requires_grad=True), {"assign": Wfb})),
        ("Gain", dict(size=(1, N), requires_grad=True)),
    ])
    monkeypatch.setenv("FSWEEP_FORCE_TPC", "1")
    M = nfft // 2 + 1
    X = C.make_input(2, M, 1, None)
    errs = {}
    for v1 in ("0", "1"):
        monkeypatch.setenv("FSWEEP_TPC_V1", v1)
        saved = dict(sweep._PLANS)
        sweep._PLANS.clear()
        try:
            torch.manual_seed(3)
            model = W.build(desc, dsp, system, nfft, 30.0, dtype=torch.float32, device="cuda")
            ps = [p for p in model.parameters() if p.requires_grad]
            Y = model(X.to(torch.complex64).cuda())
            C.golden_loss(Y).backward()
            torch.cuda.synchronize()
        finally:
            sweep._PLANS.clear()
            sweep._PLANS.update(saved)
        p64 = [p.detach().cpu().double().requires_grad_(p.requires_grad) for p in model.parameters()]
        Yo = O.forward(O.from_desc(desc), X, p64, nfft, 30.0)
        go = torch.autograd.grad(C.golden_loss(Yo), [p for p in p64 if p.requires_grad])
        errs[v1] = (rel_err(np.abs(Y.detach().cpu().numpy()), np.abs(Yo.detach().numpy())),
                    max(grad_err(p.grad.cpu().numpy(), r.numpy()) for p, r in zip(ps, go)))
    # (an arbitrary loop matrix is not well conditioned at every bin: the bar is the pivot-every-step kernel's own error)
    assert errs["0"][0] <= max(1e-4, 3 * errs["1"][0]), errs
    assert errs["0"][1] <= max(1e-3, 3 * errs["1"][1]), errs
