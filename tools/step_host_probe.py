#!/usr/bin/env python
"""Where the bench's per-step time of config 2 goes OUTSIDE the kernels: events (behind an L2 flush, like bench.py) around
(A) the bare graph replay, (B) replay + stream synchronize, (C) replay + synchronize + loss read-back, (D) the whole
Trainer.train_step; plus the host-side clock of the same pieces."""
import os
import statistics
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    dev = "cuda:0"
    torch.cuda.set_device(0)
    model, ds, Trainer, mse_loss, sparsity_loss = bench.build_gpu_model(dev)
    tr = Trainer(model, max_epochs=1, lr=1e-3, log=False, device=dev, graph=True)
    tr.register_criterion(mse_loss(nfft=bench.NFFT, device=dev), 1)
    tr.register_criterion(sparsity_loss(), 0.2, requires_model=True)
    x, y = ds.input[:1].to(dev), ds.target[:1].to(dev)
    for _ in range(8):
        tr.train_step((x, y))
    (g,) = tr._graphs.values()
    graph, out_host, slot = g[0], g[3], g[6]
    stream = torch.cuda.current_stream()
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def bump():
        if slot is not None:
            slot["expected"] += 1

    def a():
        graph.replay()
        bump()

    def b():
        graph.replay()
        bump()
        stream.synchronize()

    def c():
        graph.replay()
        if slot is not None:
            return tr._await_losses(slot, dev)
        stream.synchronize()
        return out_host.tolist()

    def d():
        return tr.train_step((x, y))

    rows = []
    for name, fn in (("A replay", a), ("B replay+sync", b), ("C replay+loss read-back", c), ("D train_step", d)):
        ts, hs = [], []
        for i in range(205):
            flush.fill_(i & 0xFF)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            t0 = time.perf_counter()
            fn()
            t1 = time.perf_counter()
            e.record()
            torch.cuda.synchronize()
            if i >= 5:
                ts.append(s.elapsed_time(e) * 1e3)
                hs.append((t1 - t0) * 1e6)
        rows.append((name, statistics.median(ts), statistics.median(hs)))
    # no flush: back-to-back replays, one pair of events around 50
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(50):
        graph.replay()
        bump()
    e.record()
    torch.cuda.synchronize()
    print(f"PDL={os.environ.get('FSWEEP_PDL', '1')}")
    print("| piece | device us (events, median) | host us |\n|---|---:|---:|")
    for r in rows:
        print(f"| {r[0]} | {r[1]:.1f} | {r[2]:.1f} |")
    print(f"| 50 replays back to back, per replay | {s.elapsed_time(e) * 1e3 / 50:.1f} | |")


if __name__ == "__main__":
    main()
