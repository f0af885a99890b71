"""GPU: the one-launch Adam (fsweep_adam_step, flamo_b200/optimize/adam.py) against torch.optim.Adam, the optimizer the
reference Trainer builds (flamo/optimize/trainer.py:42), eager and replayed from a CUDA graph."""
import pytest
import torch

from flamo_b200.optimize.adam import SweepAdam

pytestmark = pytest.mark.gpu

SHAPES = [(8, 8), (8,), (1, 3), (257,), (30, 16, 16)]


def _params(dtype, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return [torch.randn(*s, generator=g, dtype=dtype).cuda().requires_grad_(True) for s in SHAPES]


def _grads(step, dtype):
    g = torch.Generator(device="cpu").manual_seed(100 + step)
    return [torch.randn(*s, generator=g, dtype=dtype).cuda() * (10.0 ** (step % 3 - 1)) for s in SHAPES]


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.float64, 1e-13)])
def test_matches_torch_adam(dtype, tol):
    pa, pb = _params(dtype), _params(dtype)
    assert SweepAdam.supported(pa)
    lr = 2.0 ** -7  # exact in float32: SweepAdam keeps the learning rate as a float32 device scalar
    a = SweepAdam(pa, lr=lr)
    b = torch.optim.Adam(pb, lr=lr)
    for step in range(7):
        for p, q, g in zip(pa, pb, _grads(step, dtype)):
            p.grad, q.grad = g.clone(), g.clone()
        a.step()
        b.step()
        for p, q in zip(pa, pb):
            # relative to the parameter scale: one update is lr-sized, the two differ by rounding of the update
            assert float((p - q).detach().abs().max()) <= tol * float(q.detach().abs().max() + 1.0)
    for p in pa:
        assert float(a.state[p]["step"]) == 7.0


def test_lr_tensor_is_read_on_the_device_and_capture_replays():
    pa, pb = _params(torch.float32), _params(torch.float32)
    lr = torch.tensor(1e-2, device="cuda")
    a = SweepAdam(pa, lr=lr)
    b = torch.optim.Adam(pb, lr=torch.tensor(1e-2, device="cuda"), capturable=True)
    static = [torch.zeros_like(p) for p in pa]
    for p, s in zip(pa, static):
        p.grad = s
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        a.step()  # creates the state outside the capture
    torch.cuda.current_stream().wait_stream(s)
    for q in pb:
        q.grad = torch.zeros_like(q)
    b.step()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        a.step()
    for step in range(4):
        if step == 2:  # a scheduler fills the learning rate in place
            lr.fill_(3e-3)
            b.param_groups[0]["lr"].fill_(3e-3)
        for sg, q, g in zip(static, pb, _grads(step, torch.float32)):
            sg.copy_(g)
            q.grad = g.clone()
        graph.replay()
        b.step()
    torch.cuda.synchronize()
    for p, q in zip(pa, pb):
        assert float((p - q).detach().abs().max()) <= 2e-6 * float(q.detach().abs().max() + 1.0)
    assert float(a.state[pa[0]]["step"]) == 5.0
