"""GPU, OPT-IN (FLAMO_B200_EXPERIMENTAL=1): kernels written after round 1's GPU budget was spent and therefore never
run on a B200.  They are compiled into libfsweep.so but not dispatched unless their own environment switch is set; this
file is what to run first in the next round:

    FLAMO_B200_EXPERIMENTAL=1 python -m pytest tests/test_gpu_zz_experimental.py -q

* FSWEEP_FINALIZE_V2=1 — fsweep_finalize_v2_kernel (coalesced per-block partial sums): every gradient of every parity
  case must equal the default finalize kernel's to float rounding (both sum in float64; only the order differs)."""
import os

import pytest
import torch

import cases as C
from helpers import build_case

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("FLAMO_B200_EXPERIMENTAL", "0") != "1",
                                 reason="experimental kernels are opt-in (FLAMO_B200_EXPERIMENTAL=1)")]


def _grads(name, dtype, variant):
    os.environ["FSWEEP_FINALIZE_V2"] = variant  # read by libfsweep at every backward call
    try:
        case, g, model = build_case(name, dtype, "cuda")
        M = case["nfft"] // 2 + 1
        cdt = torch.complex64 if dtype == torch.float32 else torch.complex128
        X = C.make_input(case["B"], M, model.input_channels, case["C"]).to(cdt).cuda()
        params = [p for p in model.parameters() if p.requires_grad]
        if not params:
            return None
        C.golden_loss(model(X)).backward()
        torch.cuda.synchronize()
        return [p.grad.detach().clone() for p in params]
    finally:
        os.environ.pop("FSWEEP_FINALIZE_V2", None)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("name", list(C.CASES))
def test_finalize_v2_equals_default(name, dtype):
    if dtype == torch.float64 and name == "cfg5_fdn64_small":
        pytest.skip("no float64 kernels at loop width 64")
    a = _grads(name, dtype, "0")
    if a is None:
        pytest.skip("no trainable parameter")
    b = _grads(name, dtype, "1")
    eps = 1e-6 if dtype == torch.float32 else 1e-14
    for ga, gb in zip(a, b):
        assert torch.isfinite(gb).all()
        assert float((ga - gb).abs().max()) <= eps * float(ga.abs().max() + 1e-30)
