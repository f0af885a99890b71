#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/check_p2p_allreduce.py — the one-shot peer-memory all-reduce kernel
(libfsweep fsweep_allreduce_p2p) against NCCL: eager calls and CUDA-graph replays, and its latency."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flamo_b200 import _lib  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm_mem  # noqa: E402

L = _lib.lib()
ok = True
for n in (83, 4227):
    buf = symm_mem.empty(n, dtype=torch.float32, device=dev)
    buf.zero_()
    hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
    epoch = torch.zeros(1, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    dist.barrier()

    def p2p(scale):
        _lib.check(L.fsweep_allreduce_p2p(hdl.buffer_ptrs_dev, hdl.signal_pad_ptrs_dev, hdl.rank, hdl.world_size, n, scale,
                                          epoch.data_ptr(), torch.cuda.current_stream().cuda_stream))

    g = torch.Generator(device=dev).manual_seed(100 + rank)
    for it in range(5):
        x = torch.randn(n, device=dev, generator=g)
        ref = x.clone()
        dist.all_reduce(ref)
        ref /= world
        buf.copy_(x)
        p2p(1.0 / world)
        torch.cuda.synchronize()
        err = float((buf - ref).abs().max())
        ok &= err <= 1e-6 * float(ref.abs().max() + 1)
        gathered = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(gathered, buf.clone())
        ok &= all(torch.equal(gathered[0], t) for t in gathered)  # bit-identical on every rank
    # graph replay
    static = torch.randn(n, device=dev, generator=g)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        buf.copy_(static)
        p2p(1.0)
    torch.cuda.synchronize()
    dist.barrier()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        buf.copy_(static)
        p2p(1.0)
    ref = static.clone()
    dist.all_reduce(ref)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    ok &= float((buf - ref).abs().max()) <= 1e-6 * float(ref.abs().max() + 1)
    # latency: K replays back to back
    dist.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(200):
        graph.replay()
    t1.record()
    torch.cuda.synchronize()
    us_p2p = t0.elapsed_time(t1) * 1e3 / 200
    y = static.clone()
    gn = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        dist.all_reduce(y)
    torch.cuda.synchronize()
    with torch.cuda.graph(gn):
        y.copy_(static)
        dist.all_reduce(y)
    dist.barrier()
    t0.record()
    for _ in range(200):
        gn.replay()
    t1.record()
    torch.cuda.synchronize()
    us_nccl = t0.elapsed_time(t1) * 1e3 / 200
    if rank == 0:
        print(f"n={n} world={world} ok={bool(ok)}  copy+all-reduce per replay: p2p kernel {us_p2p:.1f} us, NCCL {us_nccl:.1f} us", flush=True)
# ---- the push kernel (fsweep_allreduce_push): segments reduced in place, one flag round, double-buffered receive areas
for sizes in ((64, 8, 8, 3), (4096, 64, 64, 3)):
    n = sum(sizes)
    recv = symm_mem.empty(2 * 2 * world * n, dtype=torch.float32, device=dev)  # 8-byte slots {value, epoch}
    recv.zero_()
    hdl = symm_mem.rendezvous(recv, dist.group.WORLD)
    epoch = torch.zeros(2, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    dist.barrier()
    g = torch.Generator(device=dev).manual_seed(500 + rank)
    segs_t = [torch.empty(k, device=dev) for k in sizes]

    def push(scale):
        segs = (_lib.Seg * len(segs_t))(*[_lib.Seg(t.data_ptr(), t.numel()) for t in segs_t])
        _lib.check(L.fsweep_allreduce_push(segs, len(segs_t), hdl.buffer_ptrs_dev, hdl.signal_pad_ptrs_dev, hdl.rank,
                                           hdl.world_size, n, scale, epoch.data_ptr(), torch.cuda.current_stream().cuda_stream))

    for it in range(6):  # consecutive epochs exercise both parities
        src = [torch.randn(k, device=dev, generator=g) for k in sizes]
        ref = torch.cat(src)
        dist.all_reduce(ref)
        ref /= world
        for t, v in zip(segs_t, src):
            t.copy_(v)
        push(1.0 / world)
        torch.cuda.synchronize()
        out = torch.cat(segs_t)
        ok &= float((out - ref).abs().max()) <= 1e-6 * float(ref.abs().max() + 1)
        gathered = [torch.empty_like(out) for _ in range(world)]
        dist.all_gather(gathered, out)
        ok &= all(torch.equal(gathered[0], t) for t in gathered)
    ok &= int(epoch[1].item()) == 0
    static = [torch.randn(k, device=dev, generator=g) for k in sizes]
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for t, v in zip(segs_t, static):
            t.copy_(v)
        push(1.0)
    torch.cuda.synchronize()
    dist.barrier()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for t, v in zip(segs_t, static):
            t.copy_(v)
        push(1.0)
    ref = torch.cat(static)
    dist.all_reduce(ref)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    ok &= float((torch.cat(segs_t) - ref).abs().max()) <= 1e-6 * float(ref.abs().max() + 1)
    dist.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(200):
        graph.replay()
    t1.record()
    torch.cuda.synchronize()
    if rank == 0:
        print(f"push kernel: n={n} ({len(sizes)} segments) world={world} ok={bool(ok)}  copies+all-reduce per replay: "
              f"{t0.elapsed_time(t1) * 1e3 / 200:.1f} us", flush=True)
torch.cuda.synchronize()
dist.barrier()
os._exit(0 if ok else 1)
