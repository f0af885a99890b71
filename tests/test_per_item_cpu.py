"""Per-item parameter sets (batched ext_param) — host logic on the ABI emulator: lowering (one op with per_item set, folded
section designs), the autograd plumbing of per-item gradients, the batch checks, and what the library's planner makes of
such a program.  The kernels themselves: tests/test_gpu_per_item.py."""
from collections import OrderedDict

import pytest
import torch

import per_item_cases as PC
from flamo_b200 import _lib, sweep
from flamo_b200.processor import dsp, system


@pytest.mark.parametrize("name", list(PC.CASES))
def test_per_item_equals_the_per_item_loop_and_the_oracle(name, emulated_backend):
    PC.check_case(name, torch.float64, "cpu")


@pytest.mark.parametrize("name", ["biquad_highpass", "gain_matrix", "fdn_feedback_gain"])
def test_per_item_with_trailing_columns_and_bin_shards(name, emulated_backend):
    PC.check_case(name, torch.float64, "cpu", B=2, cols=3)
    PC.check_case(name, torch.float64, "cpu", B=4, shard=(40, 101))


def test_one_signal_shared_by_all_sets(emulated_backend):
    """The reference examples feed the SAME excitation to every item (`z[0].unsqueeze(0)`): a single-item signal is
    broadcast over the parameter sets."""
    nfft = PC.NFFT
    filt = dsp.Biquad(size=(2, 1), n_sections=2, filter_type="highpass", nfft=nfft, fs=PC.FS, alias_decay_db=30,
                      dtype=torch.float64)
    model = system.Shell(core=OrderedDict({"biquad": filt}), input_layer=dsp.FFT(nfft, dtype=torch.float64),
                         output_layer=dsp.Transform(lambda x: torch.abs(x), dtype=torch.float64))
    P = PC.draw("biquad", (5,) + tuple(filt.param.shape), 1).requires_grad_(True)
    z = torch.zeros(1, nfft, 1, dtype=torch.float64)
    z[:, 0] = 1
    Y = model(z, {"biquad": P})
    assert Y.shape == (5, nfft // 2 + 1, 2)
    ref = torch.vstack([model(z, {"biquad": P[i]}) for i in range(5)])  # e7_biquad_nn.py:149-156
    assert torch.allclose(Y, ref, rtol=1e-10, atol=1e-12)
    g, = torch.autograd.grad(Y.square().sum(), P)
    gr, = torch.autograd.grad(ref.square().sum(), P)
    assert torch.allclose(g, gr, rtol=1e-8, atol=1e-12)


def test_lowering_is_one_per_item_op(emulated_backend):
    nfft = PC.NFFT
    filt = dsp.Biquad(size=(2, 1), n_sections=2, filter_type="highpass", nfft=nfft, fs=PC.FS, dtype=torch.float64)
    prog = sweep.Program(nfft, 0.0, torch.complex128, "cpu")
    filt._lower(prog, PC.draw("biquad", (3,) + tuple(filt.param.shape), 2))
    (tag, op, coef), = prog.items
    assert tag == "leaf" and op[0] == _lib.OP_SOS and op[3] == 2 and op[7] == 1
    assert tuple(coef.shape) == (3, 2, 1, 2, 2, 8)  # [item][section][n_in][n_out][Taylor block][8]


def test_batch_mismatch_raises(emulated_backend):
    nfft = PC.NFFT
    g = dsp.Gain(size=(2, 2), nfft=nfft, dtype=torch.float64)
    x = torch.ones(3, nfft // 2 + 1, 2, dtype=torch.complex128)
    with pytest.raises(ValueError, match="per-item parameter sets"):
        g(x, torch.randn(4, 2, 2, dtype=torch.float64))
    with pytest.raises(AssertionError):  # neither the module's shape nor a batch of it: the reference's own check
        g(x, torch.randn(2, 3, dtype=torch.float64))


def test_planner_runs_per_item_programs_on_the_generic_kernels():
    """include/fsweep.h: a plan with a per_item op uses the generic kernels (grid slice per item), whatever faster
    family the same program would get otherwise, and its workspace grows with the item count."""
    try:
        _lib.lib()
    except RuntimeError:
        pytest.skip("libfsweep.so not built")
    F = _lib.F_GRAD
    fdn = lambda per: [(_lib.OP_GAIN, 8, 1, 0, F, 0, 0, 0), (_lib.OP_RECURSION, 8, 8, 0, 0, 1, 1, 0),
                       (_lib.OP_PDELAY, 8, 8, 0, _lib.F_ISINT, 0, 0, 0), (_lib.OP_GAIN, 8, 8, 0, F, 0, 0, per),
                       (_lib.OP_GAIN, 1, 8, 0, F, 0, 0, 0)]
    plain = _lib.Plan([_lib.Op(*o) for o in fdn(0)], 96000, 30.0, _lib.C64)
    items = _lib.Plan([_lib.Op(*o) for o in fdn(1)], 96000, 30.0, _lib.C64)
    assert "tpc" in plain.kernel_family(48001, True)
    assert items.kernel_family(48001, True) == "fsweep_bwd_kernel" and items.kernel_family(48001, False) == "fsweep_fwd_kernel"
    assert items.workspace_bytes(8, 1, 48001) > 4 * items.workspace_bytes(1, 1, 48001) > 0
    assert plain.workspace_bytes(8, 1, 48001) == plain.workspace_bytes(1, 1, 48001)
