"""Debug helper (not a test): step through the multi-GPU trainer with per-rank logs."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
log = open(os.path.join(ROOT, "gpurun_out", f"rank{rank}.log"), "w")
def P(*a):
    print(f"[{time.time():.2f}]", *a, file=log, flush=True)
torch.cuda.set_device(local)
dev = f"cuda:{local}"
P("init pg")
dist.init_process_group("nccl", device_id=torch.device(dev))
P("pg ok")
import bench
from flamo_b200.parallel import DataParallelTrainer
for graph in (False, True):
    model, ds, Trainer, mse_loss, sparsity_loss = bench.build_gpu_model(dev)
    tr = DataParallelTrainer(model, max_epochs=1, lr=1e-3, log=False, device=dev, graph=graph)
    tr.register_criterion(mse_loss(nfft=bench.NFFT, device=dev), 1)
    tr.register_criterion(sparsity_loss(), 0.2, requires_model=True)
    x, y = ds.input[:1].to(dev), ds.target[:1].to(dev)
    for i in range(8):
        l = tr.train_step((x, y))
        P("graph", graph, "step", i, "loss", l, "use_graph", tr.use_graph, "n_graphs", len(tr._graphs))
    torch.cuda.synchronize()
    P("barrier")
    dist.barrier()
    P("barrier ok")
dist.destroy_process_group()
P("done")
