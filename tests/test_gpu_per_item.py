"""GPU: per-item parameter sets (batched ext_param, include/fsweep.h fsweep_op_t::per_item) on the real kernels — the
generic kernels with one grid slice per batch item — against the reference's per-item loop (which runs whatever kernel
family the plain program gets), against the float64 oracle item by item, and the fused criterion through the same
launch."""
import pytest
import torch

import per_item_cases as PC

pytestmark = [pytest.mark.gpu]

TOL = {torch.float32: dict(tol_item=5e-5, tol_oracle=1e-4, tol_grad=1e-3),
       torch.float64: dict(tol_item=1e-10, tol_oracle=1e-8, tol_grad=1e-7)}


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("name", list(PC.CASES))
def test_per_item_equals_the_per_item_loop_and_the_oracle(name, dtype):
    PC.check_case(name, dtype, "cuda", **TOL[dtype])


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("name", ["biquad_highpass", "gain_matrix", "fdn_feedback_gain", "fir_table"])
def test_per_item_with_trailing_columns_bin_shards_and_many_items(name, dtype):
    PC.check_case(name, dtype, "cuda", B=2, cols=3, **TOL[dtype])
    PC.check_case(name, dtype, "cuda", B=4, shard=(40, 101), **TOL[dtype])
    PC.check_case(name, dtype, "cuda", B=37, **TOL[dtype])


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_fused_criterion_through_a_per_item_launch(dtype):
    """fsweep_backward_loss on a plan with a per-item op: loss and gradients equal |.| + MSE formed by PyTorch on the
    per-item sweep output (the Python layer itself routes per-item programs through the unfused path)."""
    from flamo_b200 import _lib, sweep
    from flamo_b200.processor import dsp

    nfft, B = 1024, 5
    M = nfft // 2 + 1
    cdt = torch.complex64 if dtype == torch.float32 else torch.complex128
    filt = dsp.Biquad(size=(2, 2), n_sections=2, filter_type="lowpass", nfft=nfft, fs=PC.FS, alias_decay_db=30,
                      device="cuda", dtype=dtype)
    gain = dsp.Gain(size=(2, 2), nfft=nfft, requires_grad=True, alias_decay_db=30, device="cuda", dtype=dtype)
    P = PC.draw("biquad", (B,) + tuple(filt.param.shape), 5).to(dtype=dtype, device="cuda").requires_grad_(True)
    g = torch.Generator().manual_seed(1)
    X = torch.complex(torch.randn(B, M, 2, generator=g, dtype=torch.float64),
                      torch.randn(B, M, 2, generator=g, dtype=torch.float64)).to(cdt).cuda()
    tgt = torch.rand(B, M, 2, generator=g, dtype=torch.float64).to(dtype).cuda()

    def program():
        prog = sweep.Program(nfft, 30.0, cdt, X.device)
        gain._lower(prog)
        filt._lower(prog, P)
        return prog

    Y = program().run(X, _lib.EPI_ABS)
    ref = ((Y - tgt) ** 2).mean()
    gref = torch.autograd.grad(ref, [P, gain.param])
    ops, coefs, _ = sweep.Program.flatten_segment(program().items, cdt)
    assert [o[7] for o in ops] == [0, 1]
    plan = sweep._get_plan(ops, nfft, 30.0, _lib.C64 if dtype == torch.float32 else _lib.C128)
    loss = sweep.SweepLossFunction.apply(X.unsqueeze(-1), tgt, plan, ops, _lib.CRIT_MSE, 1.0 / tgt.numel(), 0, *coefs)
    got = torch.autograd.grad(loss, [P, gain.param])
    tol = 1e-4 if dtype == torch.float32 else 1e-10
    assert abs(float(loss) - float(ref)) <= tol * abs(float(ref))
    for a, b in zip(got, gref):
        assert float((a - b).abs().max()) <= 10 * tol * float(b.abs().max())
