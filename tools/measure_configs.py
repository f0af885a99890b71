#!/usr/bin/env python
"""Device-timed Trainer.train_step of the five BASELINE.json configs (SURVEY.md §8d restatements) on one GPU.

Not the headline bench (bench.py measures config 2, the config the metric is quoted on): this is the parity-case
companion that records how the other configs run on the same engine.  Prints one JSON line per config and a
markdown table.  Usage:  python tools/measure_configs.py [--configs cfg1_biquad,cfg5_fdn64] [--steps 30]
Under torchrun (one rank per GPU) the step runs through flamo_b200.parallel.DataParallelTrainer with --shard bins
(every rank sweeps a contiguous bin range of the same batch: strong scaling) or --shard batch (weak scaling); one
NCCL all-reduce of the flat gradient buffer per step, captured in the step's graph.
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from flamo_b200 import sweep, workloads as W  # noqa: E402
from flamo_b200.optimize.loss import mse_loss, sparsity_loss  # noqa: E402
from flamo_b200.optimize.trainer import Trainer  # noqa: E402
from flamo_b200.processor import dsp, system  # noqa: E402

RANK = int(os.environ.get("RANK", "0"))
WORLD = int(os.environ.get("WORLD_SIZE", "1"))
LOCAL = int(os.environ.get("LOCAL_RANK", "0"))
DEV = f"cuda:{LOCAL}"
SHARD = None  # set by --shard under torchrun: "bins" (strong scaling) | "batch" (weak scaling)


def build(name, scale_bins=1, batch=None):
    desc, nfft, B, seed, n_ch = W.CONFIGS[name]
    nfft //= scale_bins
    B = batch or B
    M = nfft // 2 + 1
    torch.manual_seed(seed)
    core = W.build(desc, dsp, system, nfft, W.ALIAS_DECAY_DB, dtype=torch.float32, device=DEV)
    model = system.Shell(core, dsp.FFT(nfft), dsp.Transform(lambda x: torch.abs(x)))
    n_in, n_out = model.input_channels, model.output_channels
    if name in ("cfg1_biquad", "cfg3_geq16"):
        # impulse in the time domain; target = magnitude response of a second, randomly drawn instance; nn.MSELoss
        x = torch.zeros(B, nfft, n_in, device=DEV)
        x[:, 0, :] = 1
        torch.manual_seed(seed + 1)
        tcore = W.build(desc, dsp, system, nfft, W.ALIAS_DECAY_DB, dtype=torch.float32, device=DEV)
        with torch.no_grad():
            tgt = system.Shell(tcore, dsp.FFT(nfft), dsp.Transform(lambda x: torch.abs(x)))(x).clone()
        crits = [(torch.nn.MSELoss(), 1, False)]
    else:
        # colourless target: bin-domain impulse, flat magnitude, mse_loss (+ sparsity for the FDNs)
        x = torch.zeros(B, M, n_in, device=DEV)
        x[:, 0, :] = 1
        tgt = torch.ones(B, M, 1, device=DEV)
        crits = [(mse_loss(nfft=nfft, device=DEV), 1, False)]
        if name in ("cfg2_fdn8", "cfg5_fdn64"):
            crits.append((sparsity_loss(), 0.2, True))
    if WORLD > 1:
        from flamo_b200.parallel import DataParallelTrainer

        tr = DataParallelTrainer(model, max_epochs=1, lr=1e-3, log=False, device=DEV, shard=SHARD)
    else:
        tr = Trainer(model, max_epochs=1, lr=1e-3, log=False, device=DEV)
    for c, a, rm in crits:
        tr.register_criterion(c, a, requires_model=rm)
    return tr, x, tgt, B, M, n_ch, nfft


def measure(name, steps, scale_bins=1, batch=None):
    tr, x, tgt, B, M, n_ch, nfft = build(name, scale_bins, batch)
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    t0 = time.time()
    for _ in range(6):
        loss = tr.train_step((x, tgt))
    torch.cuda.synchronize()
    t_setup = time.time() - t0
    ts = []
    n0 = sweep.launch_count
    for i in range(steps):
        flush.fill_(i & 0xFF)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        loss = tr.train_step((x, tgt))
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    med = ts[len(ts) // 2]
    if WORLD > 1:  # the step is a collective: report the slowest rank
        import torch.distributed as dist

        v = torch.tensor([med], device=DEV, dtype=torch.float64)
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        med = float(v.item())
        if SHARD == "batch":
            B = B * WORLD  # every rank trains its own batch items
    return {"config": name, "world": WORLD, "shard": SHARD, "nfft": nfft, "bins": M, "batch": B, "channels": n_ch, "ms_per_step": med,
            "bins_ch_per_s": B * M * n_ch / (med * 1e-3), "cuda_graph": bool(tr.use_graph and tr._graphs),
            "sweep_launches_per_step": (sweep.launch_count - n0) / steps, "loss": loss, "setup_s": round(t_setup, 2)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default=",".join(W.CONFIGS))
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--scale-bins", type=int, default=1, help="divide nfft by this (smoke runs)")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--shard", choices=["bins", "batch"], default="bins", help="under torchrun: how ranks split the work")
    args = ap.parse_args()
    global SHARD
    torch.cuda.set_device(LOCAL)
    if WORLD > 1:
        import torch.distributed as dist

        SHARD = args.shard
        dist.init_process_group("nccl", device_id=torch.device(DEV))
    rows = []
    for name in args.configs.split(","):
        try:
            r = measure(name, args.steps, args.scale_bins, args.batch)
        except Exception as ex:  # keep going: one config failing must not hide the others
            r = {"config": name, "error": f"{type(ex).__name__}: {ex}"}
        if RANK == 0:
            print(json.dumps(r), flush=True)
        rows.append(r)
    if WORLD > 1:
        import torch.distributed as dist

        torch.cuda.synchronize()
        dist.barrier()
        if RANK != 0:
            os._exit(0)
    print("\n| config | nfft | bins | batch | N_ch | ms/step | bins*ch/s | captured | sweep launches/step |")
    print("|---|---:|---:|---:|---:|---:|---:|---|---:|")
    for r in rows:
        if "error" in r:
            print(f"| {r['config']} | error: {r['error']} |")
        else:
            print(f"| {r['config']} | {r['nfft']} | {r['bins']} | {r['batch']} | {r['channels']} | {r['ms_per_step']:.3f} | "
                  f"{r['bins_ch_per_s']:.3e} | {r['cuda_graph']} | {r['sweep_launches_per_step']:.1f} |")
    if WORLD > 1:
        sys.stdout.flush()
        os._exit(0)  # a captured graph holding NCCL work blocks ProcessGroupNCCL's teardown (see bench.py)


if __name__ == "__main__":
    main()
