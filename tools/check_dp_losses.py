#!/usr/bin/env python
"""Multi-GPU check (run under torchrun): the captured data-parallel step — loss values exchanged early on a side branch
of the graph and polled by the host, gradients exchanged by the push kernel — must reproduce the eager data-parallel
step (one exchange of gradients + values, read back with a synchronize), loss by loss and parameter by parameter, with
DIFFERENT data on every rank, for both shard modes."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(dev))
import bench  # noqa: E402
from flamo_b200.parallel import DataParallelTrainer  # noqa: E402

ok = True
for shard in ("batch", "bins"):
    runs = {}
    for graph in (False, True):
        model, ds, Trainer, mse_loss, sparsity_loss = bench.build_gpu_model(dev)
        tr = DataParallelTrainer(model, max_epochs=1, lr=1e-3, log=False, device=dev, graph=graph, shard=shard)
        tr.register_criterion(mse_loss(nfft=bench.NFFT, device=dev), 1)
        tr.register_criterion(sparsity_loss(), 0.2, requires_model=True)
        s = 1.0 + (0.1 * rank if shard == "batch" else 0.0)  # bin shards see the same batch
        x, y = (ds.input[:1] * s).to(dev), (ds.target[:1] * (2.0 - s)).to(dev)
        losses = [tr.train_step((x, y)) for _ in range(12)]
        torch.cuda.synchronize()
        tr.check_exchange()
        runs[graph] = (losses, [p.detach().clone() for p in model.parameters()], tr)
    la, lb = runs[False][0], runs[True][0]
    same = np.allclose(la, lb, rtol=2e-5)
    for pa, pb in zip(runs[False][1], runs[True][1]):
        same = same and bool(torch.allclose(pa, pb, rtol=1e-4, atol=1e-5))
    g = list(runs[True][2]._graphs.values())
    early = bool(g) and g[0][6] is not None
    every = [None] * world
    dist.all_gather_object(every, (lb[-1], same, early))
    if rank == 0:
        agree = all(abs(e[0] - every[0][0]) <= 1e-6 * abs(every[0][0]) for e in every)
        print(f"shard={shard} world={world} captured == eager on every rank: {all(e[1] for e in every)}; "
              f"ranks agree on the loss: {agree}; notification slot in use: {all(e[2] for e in every)}; "
              f"last loss {lb[-1]:.6f}")
        ok = ok and agree and all(e[1] for e in every)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
