"""CPU, world_size = 2, gloo: the multi-GPU host logic (bin-range / batch sharding, one all-reduce of
the flat gradient buffer with the losses in its tail) gives the same training trajectory as a single
process.  The sweep itself runs through the float64 ABI emulator (tests/cpu_emulator.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NFFT, N, STEPS = 1024, 4, 3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


FS = 48000
# a tree with launch boundaries inside the Series: a loop, a Parallel node (eager sum of two launches), a second loop
BOUNDARIES = ("Series", [
    ("Gain", dict(size=(3, 1), requires_grad=True)),
    ("Recursion", ("parallelDelay", dict(size=(3,), max_len=300, isint=True, requires_grad=False),
                   {"delay_samples": [101, 157, 211]}),
     ("Series", [("Gain", dict(size=(3, 3), requires_grad=True)),
                 ("parallelGain", dict(size=(3,), requires_grad=True), {"assign": [0.1, 0.1, 0.1]})])),
    ("Parallel",
     ("Series", [("Biquad", dict(size=(2, 3), n_sections=1, filter_type="lowpass", fs=FS, requires_grad=True))]),
     ("Series", [("Delay", dict(size=(2, 3), max_len=40, isint=False, fs=FS, requires_grad=True))]), True),
    ("Recursion", ("Gain", dict(size=(2, 2), requires_grad=True)),
     ("Series", [("parallelDelay", dict(size=(2,), max_len=50, isint=False, fs=FS, requires_grad=True)),
                 ("parallelGain", dict(size=(2,), requires_grad=True), {"assign": [0.05, 0.05]})])),
    ("Gain", dict(size=(1, 2), requires_grad=True)),
])


def _build(batch, kind="fdn"):
    from flamo_b200 import workloads as W
    from flamo_b200.optimize.loss import mse_loss, sparsity_loss
    from flamo_b200.processor import dsp, system

    torch.manual_seed(7)
    desc = W.fdn(N, delays=[101, 157, 211, 263]) if kind == "fdn" else BOUNDARIES
    if kind != "fdn":
        sparsity_loss = None
    core = W.build(desc, dsp, system, NFFT, 30.0, dtype=torch.float64)
    model = system.Shell(core, dsp.FFT(NFFT, dtype=torch.float64),
                         dsp.Transform(lambda x: torch.abs(x), dtype=torch.float64))
    M = NFFT // 2 + 1
    g = torch.Generator().manual_seed(3)
    x = torch.zeros(batch, M, 1, dtype=torch.float64)
    x[:, 0, :] = 1.0 + 0.1 * torch.arange(batch, dtype=torch.float64).view(-1, 1)  # items differ
    y = 1.0 + 0.05 * torch.rand(batch, M, 1, generator=g, dtype=torch.float64)
    return model, x, y, mse_loss, sparsity_loss


def _train(trainer_cls, model, x, y, mse_loss, sparsity_loss, **kw):
    tr = trainer_cls(model, max_epochs=1, lr=1e-2, log=False, device="cpu", **kw)
    tr.register_criterion(mse_loss(nfft=NFFT), 1)
    if sparsity_loss is not None:
        tr.register_criterion(sparsity_loss(), 0.2, requires_model=True)
    losses = [tr.train_step((x, y)) for _ in range(STEPS)]
    return losses, [p.detach().clone() for p in model.parameters()]


def _worker(rank, world, port, shard, out, kind="fdn"):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cpu_emulator
    from flamo_b200.parallel import DataParallelTrainer

    cpu_emulator.install()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        model, x, y, mse_loss, sparsity_loss = _build(2, kind)
        if shard == "batch":
            x, y = x[rank:rank + 1], y[rank:rank + 1]  # each rank trains its own item
        losses, params = _train(DataParallelTrainer, model, x, y, mse_loss, sparsity_loss, shard=shard)
        if rank == 0:
            torch.save({"losses": losses, "params": params}, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shard", ["bins", "batch"])
def test_two_ranks_match_single_process(shard, tmp_path, emulated_backend):
    from flamo_b200.optimize.trainer import Trainer

    model, x, y, mse_loss, sparsity_loss = _build(batch=2)
    ref_losses, ref_params = _train(Trainer, model, x, y, mse_loss, sparsity_loss)
    out = str(tmp_path / "rank0.pt")
    mp.spawn(_worker, args=(2, _free_port(), shard, out), nprocs=2, join=True)
    got = torch.load(out)
    assert np.allclose(got["losses"], ref_losses, rtol=1e-9), (got["losses"], ref_losses)
    for a, b in zip(got["params"], ref_params):
        assert torch.allclose(a, b, rtol=1e-8, atol=1e-10)


def test_two_ranks_bin_shards_across_launch_boundaries(tmp_path, emulated_backend):
    """Bin-sharded step of a Series with launch boundaries (two loops, a Parallel node): the modules behind the first
    launch see a signal already restricted to the rank's bins (found by tests/test_random_trees_cpu.py)."""
    from flamo_b200.optimize.trainer import Trainer

    model, x, y, mse_loss, sparsity_loss = _build(2, "boundaries")
    ref_losses, ref_params = _train(Trainer, model, x, y, mse_loss, sparsity_loss)
    out = str(tmp_path / "rank0.pt")
    mp.spawn(_worker, args=(2, _free_port(), "bins", out, "boundaries"), nprocs=2, join=True)
    got = torch.load(out)
    assert np.allclose(got["losses"], ref_losses, rtol=1e-9), (got["losses"], ref_losses)
    for a, b in zip(got["params"], ref_params):
        assert torch.allclose(a, b, rtol=1e-8, atol=1e-10)


def test_bin_range_partition():
    from flamo_b200.parallel import bin_range

    for M, W in ((48001, 8), (513, 4), (7, 3), (5, 8)):
        cuts = [bin_range(M, r, W) for r in range(W)]
        assert cuts[0][0] == 0 and cuts[-1][1] == M
        assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
        sizes = [b - a for a, b in cuts]
        assert max(sizes) - min(sizes) <= 1


def _build_time_domain(batch):
    """Shell whose OUTPUT LAYER is time-domain (iFFTAntiAlias): under bin sharding it needs the whole spectrum."""
    from flamo_b200 import workloads as W
    from flamo_b200.processor import dsp, system

    torch.manual_seed(7)
    core = W.build(W.fdn(N, delays=[101, 157, 211, 263]), dsp, system, NFFT, 30.0, dtype=torch.float64)
    model = system.Shell(core, dsp.FFT(NFFT, dtype=torch.float64),
                         dsp.iFFTAntiAlias(NFFT, alias_decay_db=30.0, dtype=torch.float64))
    M = NFFT // 2 + 1
    g = torch.Generator().manual_seed(3)
    x = torch.zeros(batch, M, 1, dtype=torch.float64)
    x[:, 0, :] = 1.0
    y = 0.01 * torch.randn(batch, NFFT, 1, generator=g, dtype=torch.float64)  # a time-domain target
    return model, x, y


def _train_td(trainer_cls, model, x, y, **kw):
    tr = trainer_cls(model, max_epochs=1, lr=1e-2, log=False, device="cpu", **kw)
    tr.register_criterion(torch.nn.MSELoss(), 1)
    losses = [tr.train_step((x, y)) for _ in range(STEPS)]
    return losses, [p.detach().clone() for p in model.parameters()]


def _worker_td(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cpu_emulator
    from flamo_b200.parallel import DataParallelTrainer

    cpu_emulator.install()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        model, x, y = _build_time_domain(2)
        losses, params = _train_td(DataParallelTrainer, model, x, y, shard="bins")
        if rank == 0:
            torch.save({"losses": losses, "params": params}, out)
    finally:
        dist.destroy_process_group()


def test_time_domain_output_layer_under_bin_sharding(tmp_path, emulated_backend):
    """SURVEY.md §8e last row: a time-domain output layer (iFFTAntiAlias) needs all bins — under bin sharding the
    spectrum is all-gathered before the inverse FFT, the time-domain criterion is evaluated on every rank and its
    gradient flows back through each rank's own bins: same trajectory as one process."""
    from flamo_b200.optimize.trainer import Trainer

    model, x, y = _build_time_domain(2)
    ref_losses, ref_params = _train_td(Trainer, model, x, y)
    out = str(tmp_path / "rank0.pt")
    mp.spawn(_worker_td, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    assert np.allclose(got["losses"], ref_losses, rtol=1e-9), (got["losses"], ref_losses)
    for a, b in zip(got["params"], ref_params):
        assert torch.allclose(a, b, rtol=1e-8, atol=1e-10)


def test_time_domain_layer_refuses_a_fragment_of_the_spectrum(emulated_backend):
    from flamo_b200 import sweep

    model, x, y = _build_time_domain(1)
    with pytest.raises(RuntimeError, match="needs all"):
        with sweep.bin_shard(10, 200):
            model(x)
