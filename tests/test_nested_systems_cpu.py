"""CPU: Recursion / Parallel nested inside a Recursion path against the oracle, through the ABI emulator (the host side:
table fallback of sweep.Program.table_of, autograd through the inner launch, bin sharding)."""
import numpy as np
import pytest
import torch

import cases as C
from flamo_b200 import sweep, workloads as W
from flamo_b200.processor import dsp, system
from helpers import rel_err
from nested_cases import NESTED
from oracle import flamo_oracle as O

pytestmark = pytest.mark.usefixtures("emulated_backend")
NFFT, ALIAS = 256, 20.0


@pytest.mark.parametrize("name", list(NESTED))
def test_nested_system_matches_oracle(name):
    desc = NESTED[name]
    torch.manual_seed(3)
    model = W.build(desc, dsp, system, NFFT, ALIAS, dtype=torch.float64, device="cpu")
    M = NFFT // 2 + 1
    X = C.make_input(2, M, model.input_channels, None)
    Y = model(X)
    ps = [p.detach().clone().requires_grad_(p.requires_grad) for p in model.parameters()]
    Yo = O.forward(O.from_desc(desc), X, ps, NFFT, ALIAS)
    assert rel_err(Y.detach().numpy(), Yo.detach().numpy()) < 1e-10
    C.golden_loss(Y).backward()
    go = torch.autograd.grad(C.golden_loss(Yo), [p for p in ps if p.requires_grad])
    k = 0
    for p in model.parameters():
        if p.requires_grad:
            assert p.grad is not None
            assert float((p.grad - go[k]).abs().max()) <= 1e-9 * float(go[k].abs().max() + 1e-30)
            k += 1
    # a bin shard of the outer sweep: the nested tables are still built for all bins
    with torch.no_grad(), sweep.bin_shard(40, 97):
        Ys = model(X)
    assert rel_err(Ys.numpy(), Yo.detach().numpy()[:, 40:97]) < 1e-10


@pytest.mark.parametrize("name", list(NESTED))
def test_oracle_equals_the_reference_on_nested_systems(name):
    """The oracle is only a checker if it restates the reference here too: the UNMODIFIED reference (build container
    only) evaluates the same nested trees, same parameters — identical to the oracle."""
    import os

    if not os.path.isdir("/root/reference/flamo"):
        pytest.skip("reference checkout not present")
    from test_random_trees_reference_cpu import reference_modules

    rdsp, rsystem = reference_modules()
    desc = NESTED[name]
    torch.manual_seed(3)
    ref = W.build(desc, rdsp, rsystem, NFFT, ALIAS, dtype=torch.float64, device="cpu")
    M = NFFT // 2 + 1
    X = C.make_input(2, M, ref.input_channels, None)
    with torch.no_grad():
        Yr = ref(X)
        Yo = O.forward(O.from_desc(desc), X, [p.detach() for p in ref.parameters()], NFFT, ALIAS)
    assert rel_err(Yo.numpy(), Yr.numpy()) < 1e-12
