"""GPU: the capture-safe expm(skew(P)) kernel against torch.matrix_exp and its autograd."""
import numpy as np
import pytest
import torch

from flamo_b200 import sweep
from flamo_b200.functional import skew_matrix

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 2, 6, 8, 13, 16, 28])
@pytest.mark.parametrize("scale", [0.05, 1.0, 7.0])
def test_expm_matches_torch(n, scale):
    torch.manual_seed(n)
    P = (scale * torch.randn(n, n, dtype=torch.float64, device="cuda")).requires_grad_(True)
    Q = P.detach().clone().requires_grad_(True)
    E = sweep.OrthogonalMap.apply(P)[0]
    Er = torch.matrix_exp(skew_matrix(Q))
    import scipy.linalg

    Es = torch.tensor(scipy.linalg.expm(skew_matrix(Q.detach()).cpu().numpy()), device="cuda")
    assert torch.allclose(E, Es, rtol=0, atol=2e-13 * max(1.0, scale * n))
    assert torch.allclose(E, Er, rtol=0, atol=1e-9)  # sanity only: torch picks a low Taylor degree for small norms
    assert torch.allclose(E @ E.T, torch.eye(n, dtype=torch.float64, device="cuda"), atol=1e-11)
    G = torch.randn(n, n, dtype=torch.float64, device="cuda")
    (E * G).sum().backward()
    (Er * G).sum().backward()
    assert torch.allclose(P.grad, Q.grad, rtol=1e-10, atol=1e-11 * float(Q.grad.abs().max() + 1))
    assert float(P.grad.tril().abs().max()) == 0.0  # only the strict upper triangle is a free parameter


def test_expm_float32_parameter_and_capture():
    P = torch.randn(8, 8, device="cuda", requires_grad=True)
    E = sweep.OrthogonalMap.apply(P)[0]
    assert E.dtype == torch.float32
    assert torch.allclose(E, torch.matrix_exp(skew_matrix(P.detach().double())).float(), atol=1e-6)
    del E  # an autograd graph kept alive from before the capture would pin P's AccumulateGrad to the default stream
    for _ in range(2):  # warm-up on the current stream, like Trainer._graph_for
        sweep.OrthogonalMap.apply(P)[0].sum().backward()
    P.grad = None
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = sweep.OrthogonalMap.apply(P)[0]
        out.square().sum().backward()
    with torch.no_grad():
        P.add_(0.1)
    g.replay()
    torch.cuda.synchronize()
    assert torch.allclose(out, torch.matrix_exp(skew_matrix(P.detach().double())).float(), atol=1e-6)
    assert P.grad is not None and torch.isfinite(P.grad).all()


@pytest.mark.parametrize("shape", [(8, 8), (3, 6, 6), (64, 64)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_sparsity_kernels_match_reference_formula(shape, dtype):
    """fsweep_sparsity_* against the reference's expression (optimize/loss.py:36-63) and its autograd."""
    from flamo_b200 import sweep

    torch.manual_seed(1)
    A = torch.randn(*shape, device="cuda", dtype=dtype, requires_grad=True)
    N = shape[-1]
    if A.dim() == 3:
        ref = torch.mean((torch.sum(torch.abs(A), dim=(-2, -1)) - N * np.sqrt(N)) / (N * (1 - np.sqrt(N))))
    else:
        ref = -(torch.sum(torch.abs(A)) - N * np.sqrt(N)) / (N * (np.sqrt(N) - 1))
    (gr,) = torch.autograd.grad(1.7 * ref, A)
    assert sweep.SparsityFunction.supported(A)
    out = sweep.SparsityFunction.apply(A)
    (go,) = torch.autograd.grad(1.7 * out, A)
    tol = 1e-6 if dtype == torch.float32 else 1e-13
    assert abs(float(out.detach()) - float(ref.detach())) <= tol * max(1.0, abs(float(ref.detach())))
    assert torch.allclose(go, gr, rtol=tol * 10, atol=tol)


@pytest.mark.parametrize("n", [2, 8, 13])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_sparsity_rides_along_with_the_orthogonal_map(n, dtype):
    """OrthogonalMap returns (E, sparsity_loss(E)) from one launch and takes both gradients back in one launch:
    against torch.matrix_exp + the reference formula (optimize/loss.py:55-63), separately and together."""
    torch.manual_seed(n)
    P = torch.randn(n, n, device="cuda", dtype=dtype, requires_grad=True)
    Wt = torch.randn(n, n, device="cuda", dtype=dtype)

    def ref(p):
        u = p.triu(1)
        E = torch.matrix_exp((u - u.T).double())
        sp = -(E.abs().sum() - n * n ** 0.5) / (n * (n ** 0.5 - 1))
        return E, sp

    tol = 2e-5 if dtype == torch.float32 else 1e-11
    for use_e, use_sp in ((True, True), (False, True), (True, False)):
        P.grad = None
        E, sp = sweep.OrthogonalMap.apply(P)
        loss = (E * Wt).sum() * (1.0 if use_e else 0.0) * (1 if use_e else 0) if use_e else 0.0
        loss = loss + (0.7 * sp if use_sp else 0.0)
        loss.backward()
        Pr = P.detach().clone().requires_grad_(True)
        Er, spr = ref(Pr)
        lr_ = ((Er * Wt.double()).sum() if use_e else 0.0) + (0.7 * spr if use_sp else 0.0)
        lr_.backward()
        assert abs(float(sp) - float(spr)) <= tol * max(1.0, abs(float(spr)))
        assert float((E.double() - Er).abs().max()) <= tol
        g, gr = P.grad.double(), Pr.grad.double()
        assert float((g - gr).abs().max()) <= 5 * tol * max(1.0, float(gr.abs().max()))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("parallel", [False, True])
@pytest.mark.parametrize("ft", ["lowpass", "highpass"])
def test_fused_biquad_designer_equals_the_pytorch_designer(ft, parallel, dtype, monkeypatch):
    """fsweep_biquad_design (bounded map -> RBJ taps -> Taylor packing in one launch, and its adjoint) against the
    PyTorch map / designer / pack_sections chain it replaces: packed coefficients and parameter gradients, including
    parameters outside the clamps (cut-off > 1, gain beyond +-60 dB) where the gradient must vanish."""
    from flamo_b200.processor import dsp

    torch.manual_seed(5)
    kw = dict(n_sections=3, filter_type=ft, nfft=256, fs=48000, requires_grad=True, alias_decay_db=10.0, device="cuda", dtype=dtype)
    mod = dsp.parallelBiquad(size=(4,), **kw) if parallel else dsp.Biquad(size=(3, 2), **kw)
    with torch.no_grad():
        mod.param[0, 0].mul_(3.0)        # some cut-offs beyond the [0, 1] clamp
        mod.param[1, 1].mul_(5000.0)     # some gains beyond +60 dB
        mod.param[2, 1].mul_(1e-5)       # ... and below -60 dB
    Wt = torch.randn((3, 4, 2, 8) if parallel else (3, 2, 3, 2, 8), device="cuda", dtype=torch.float64)

    def run(fused):
        monkeypatch.setenv("FLAMO_B200_FUSED_DESIGN", "1" if fused else "0")
        mod.param.grad = None
        coef = mod._fused_design(mod.param)
        assert (coef is not None) == fused
        if coef is None:
            b, a = mod._taps(mod.map(mod._up(mod.param)))
            coef = sweep.pack_sections(b, a, parallel, None)
        (coef * Wt).sum().backward()
        return coef.detach().clone(), mod.param.grad.detach().clone()

    c1, g1 = run(True)
    c0, g0 = run(False)
    assert float((c1 - c0).abs().max()) <= 1e-12 * float(c0.abs().max())
    tol = 1e-5 if dtype == torch.float32 else 1e-11
    assert float((g1 - g0).abs().max()) <= tol * float(g0.abs().max())
    assert float(g0.abs().max()) > 0


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("parallel", [False, True])
def test_fused_svf_designer_equals_the_pytorch_designer(parallel, dtype, monkeypatch):
    """fsweep_svf_design (activations -> taps -> Taylor packing in one launch, and its adjoint; general mixing) against
    the PyTorch chain it replaces: packed coefficients and parameter gradients."""
    from flamo_b200.processor import dsp

    torch.manual_seed(6)
    kw = dict(n_sections=2, filter_type=None, nfft=256, fs=48000, requires_grad=True, alias_decay_db=10.0, device="cuda", dtype=dtype)
    mod = dsp.parallelSVF(size=(4,), **kw) if parallel else dsp.SVF(size=(3, 2), **kw)
    with torch.no_grad():
        mod.param.mul_(1.7)
        mod.param[1, 0].add_(25.0)  # softplus beyond its linear threshold
    Wt = torch.randn((2, 4, 2, 8) if parallel else (2, 2, 3, 2, 8), device="cuda", dtype=torch.float64)

    def run(fused):
        monkeypatch.setenv("FLAMO_B200_FUSED_DESIGN", "1" if fused else "0")
        mod.param.grad = None
        coef = mod._fused_design(mod.param)
        assert (coef is not None) == fused
        if coef is None:
            b, a = mod._taps(mod.map(mod._up(mod.param)))
            coef = sweep.pack_sections(b, a, parallel, None)
        (coef * Wt).sum().backward()
        return coef.detach().clone(), mod.param.grad.detach().clone()

    c1, g1 = run(True)
    c0, g0 = run(False)
    assert float((c1 - c0).abs().max()) <= 1e-12 * float(c0.abs().max())
    tol = 1e-5 if dtype == torch.float32 else 1e-11
    assert float((g1 - g0).abs().max()) <= tol * float(g0.abs().max())
    # the typed variants (lowpass ... notch) keep the PyTorch designer
    typed = dsp.SVF(size=(2, 2), n_sections=1, filter_type="peaking", nfft=256, device="cuda", dtype=dtype)
    assert typed._fused_design(typed.param) is None


@pytest.mark.parametrize("n", [2, 6, 8, 13, 16, 28])
@pytest.mark.parametrize("scale", [0.05, 1.0, 7.0])
def test_expm_float32_parameter_against_float64_autograd(n, scale):
    """The orthogonal map and its adjoint for a float32 parameter (float64 arithmetic inside, the scaling threshold of
    float32 results: theta = 2 forward for the skew map, 1 backward) against float64 autograd of torch.matrix_exp on the
    same (float32-representable) parameter.  (float32 ARITHMETIC for the adjoint was measured: same 10.2 us inside the
    step — the kernel is a serial instruction stream, not bound by the float64 pipe — so it stays float64.)"""
    torch.manual_seed(n)
    P = (scale * torch.randn(n, n, device="cuda")).requires_grad_(True)
    Q = P.detach().double().requires_grad_(True)
    G = torch.randn(n, n, device="cuda")
    E = sweep.OrthogonalMap.apply(P)[0]
    (E * G).sum().backward()
    Er = torch.matrix_exp(skew_matrix(Q))
    (Er * G.double()).sum().backward()
    assert float((E.double() - Er).abs().max()) <= 2e-7
    assert float((P.grad.double() - Q.grad).abs().max()) <= 2e-6 * float(Q.grad.abs().max())
