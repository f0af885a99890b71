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
