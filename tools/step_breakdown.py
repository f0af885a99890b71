#!/usr/bin/env python
"""Where the captured config-2 training step spends its time: CUDA graphs of growing prefixes of the step
(input FFT | + maps, sweep, criteria | + backward | + Adam), each replayed behind an L2 flush and timed with events.
ncu's per-launch times are cold-cache and serialised; this is the warm, in-graph view.
    python tools/step_breakdown.py [--reps 200]"""
import argparse
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=200)
    args = ap.parse_args()
    dev = "cuda:0"
    torch.cuda.set_device(0)
    model, ds, Trainer, mse_loss, sparsity_loss = bench.build_gpu_model(dev)
    tr = Trainer(model, max_epochs=1, lr=1e-3, log=False, device=dev, graph=False)
    tr.register_criterion(mse_loss(nfft=bench.NFFT, device=dev), 1)
    tr.register_criterion(sparsity_loss(), 0.2, requires_model=True)
    x, y = ds.input[:1].to(dev), ds.target[:1].to(dev)
    for _ in range(4):
        tr.train_step((x, y))
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)
    from flamo_b200 import sweep

    def stage_fft():
        return model.get_inputLayer()(x)

    def stage_fwd():
        return tr._loss_vector(x, y)

    def stage_fwd_bwd():
        for p in model.parameters():
            p.grad = None
        vals = tr._loss_vector(x, y)
        if getattr(tr, "_parts", None) is not None:
            parts, weights = tr._parts
            tr._parts = None
            seeds = [sweep._const_tensor((w,), p.dtype, p.device).reshape(p.shape) for p, w in zip(parts, weights)]
            torch.autograd.backward(parts, grad_tensors=seeds)
        return vals

    def stage_full():
        for p in model.parameters():
            p.grad = None
        return tr._train_core(x, y)

    # a capturable optimizer for the last stage
    tr2 = Trainer(model, max_epochs=1, lr=1e-3, log=False, device=dev, graph=True)
    tr2.criterion, tr2.alpha, tr2.requires_model, tr2.n_loss = tr.criterion, tr.alpha, tr.requires_model, tr.n_loss
    for _ in range(4):
        tr2._eager_train_step(x, y)

    def stage_full2():
        for p in model.parameters():
            p.grad = None
        return tr2._train_core(x, y)

    rows = []
    for name, fn in (("input FFT", stage_fft), ("+ maps, sweep, criteria (forward)", stage_fwd),
                     ("+ backward", stage_fwd_bwd), ("+ Adam (whole step)", stage_full2)):
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            for _ in range(3):
                fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        ts = []
        for i in range(args.reps + 5):
            flush.fill_(i & 0xFF)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            g.replay()
            e.record()
            torch.cuda.synchronize()
            if i >= 5:
                ts.append(s.elapsed_time(e) * 1e3)
        rows.append((name, statistics.median(ts)))
    print("| graph | us (median) | delta |\n|---|---:|---:|")
    prev = 0.0
    for name, t in rows:
        print(f"| {name} | {t:.1f} | {t - prev:+.1f} |")
        prev = t


if __name__ == "__main__":
    main()
