#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's config.

metric   freq-bins*channels/sec of one Trainer.train_step  =  B * M * N_ch / t_step
workload configs[1]: examples/e8_colorless_fdn.py restated with N = 8 delay lines
         (Gain(8,1) -> Recursion(parallelDelay(8), Matrix(8,8, orthogonal)) -> Gain(1,8)),
         nfft = 96000 (M = 48001 bins), batch 1, alias_decay_db = 30, float32 modules,
         Shell(FFT, core, |.|), criteria mse_loss + 0.2 * sparsity_loss, Adam lr 1e-3
         (SURVEY.md §8d config 2).  A "step" is one full train_step: forward, both criteria,
         backward, Adam.

    python bench.py [--gpus N] [--steps K] [--warmup W]            this repo's CUDA engine
    python bench.py --impl reference [...]                         the reference algorithm on the host CPU
    torchrun --nproc-per-node N bench.py --gpus N ...              one rank per GPU (weak scaling: every
                                                                   rank trains its own batch item, gradients
                                                                   are all-reduced each step)

Before anything is timed the bench model's first-step loss and parameter gradients are checked against the CPU
oracle ("correctness gate before any number counts"); the line carries the errors under "parity".
Besides the headline the line carries, under "configs", a device-timed Trainer.train_step of ALL FIVE BASELINE.json
configs with the roofline that bounds each (HBM for the small loops, the MEASURED FP32 FMA rate for the
compute-bound ones), and under --gpus N > 1 the bin-sharded (strong-scaling) step of config 5.

The reference is pure Python and cannot travel to the GPU box, so the reference arm times
oracle/flamo_oracle.py — the bit-exact restatement of the reference's algorithm and cost structure
(zero-padded rffts, materialised (B,M,N,N) loop matrices, torch.linalg.solve, autograd) — on all
host cores, in float32 like the modules of the GPU arm.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NFFT = 96000
N_DELAYS = 8
BATCH = 1
ALIAS_DB = 30.0
SEED = 130709
METRIC = "freq-bins*channels/sec (Trainer.train_step)"
UNIT = "bins*ch/s"
# real flops of one fused backward launch (~3.7 kflop per bin forward + backward for the 8x8 loop:
# LU 8/3 N^3 + build 6 N^2 + forward solve 8 N^2 + adjoint solve 8 N^2 + gradient contractions 14 N^2 + 8 delays)
FLOPS_PER_BIN = 8 / 3 * 512 + (6 + 8 + 8 + 14) * 64 + 8 * 40
FP32_NOMINAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # 148 SMs x 128 FMA lanes x 2 flop x 1.965 GHz = 74.4
WORKLOAD = "cfg2: e8_colorless_fdn 8x8 FDN (Gain->Recursion(parallelDelay,Matrix orthogonal)->Gain), nfft=96000, B=1"
# real flops of one training step of the two compute-bound configs:
#   config 5: per bin one complex 64 x 64 LU (8/3 N^3) + one forward and one adjoint pair of substitutions (2 * 8 N^2)
#             + forming A (2 N^2) + the dW outer product (4 N^2)
#   config 3: 7680 second-order sections per bin, ~85 flops each forward + gradient (63 GFLOP per step)
STEP_FLOPS = {"cfg5_fdn64": 192001 * (8 / 3 * 64 ** 3 + (16 + 2 + 4) * 64 ** 2), "cfg3_geq16": 63e9}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def config_dict(world, cuda_graph, l2, final_loss, device):
    """The SAME keys in both arms (the driver compares the two `config` objects)."""
    M = NFFT // 2 + 1
    return {"workload": WORKLOAD, "global_batch": world * BATCH, "nfft": NFFT, "bins": M, "channels": N_DELAYS,
            "parallelism": f"dp{world}" if world > 1 else "single", "cuda_graph": cuda_graph, "l2": l2,
            "final_loss": final_loss, "device": device}


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU with NVML every `period` s in a thread."""

    def __init__(self, index: int, period: float = 0.02):
        self.period, self.samples, self.reasons, self.max_mhz = period, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        return {
            "sm_mhz": statistics.median(self.samples) if self.samples else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


# ------------------------------------------------------------------------------------- model
def build_gpu_model(device, dtype=torch.float32):
    from flamo_b200 import workloads as W
    from flamo_b200.optimize.dataset import DatasetColorless
    from flamo_b200.optimize.loss import mse_loss, sparsity_loss
    from flamo_b200.optimize.trainer import Trainer
    from flamo_b200.processor import dsp, system

    torch.manual_seed(SEED)
    core = W.build(W.fdn(N_DELAYS), dsp, system, NFFT, ALIAS_DB, dtype=dtype, device=device)
    model = system.Shell(core=core, input_layer=dsp.FFT(NFFT, dtype=dtype),
                         output_layer=dsp.Transform(lambda x: torch.abs(x), dtype=dtype))
    M = NFFT // 2 + 1
    ds = DatasetColorless(input_shape=(1, M, 1), target_shape=(1, M, 1), expand=BATCH, device="cpu", dtype=dtype)
    return model, ds, Trainer, mse_loss, sparsity_loss


def cpu_reference_trainer(dtype=torch.float32, params=None):
    """The reference algorithm (oracle port) for the same workload, float32 like the GPU arm.  `params`: raw parameter
    values to start from (the parity gate hands over the GPU model's), else the seeded draw."""
    from flamo_b200 import workloads as W
    from flamo_b200.processor import dsp, system
    from oracle import flamo_oracle as O

    if params is None:
        torch.manual_seed(SEED)
        core = W.build(W.fdn(N_DELAYS), dsp, system, NFFT, ALIAS_DB, dtype=dtype, device="cpu")
        params = [p.detach().clone().requires_grad_(p.requires_grad) for p in core.parameters()]
    node = O.from_desc(W.fdn(N_DELAYS))
    fb = node.children[1].children[1]
    crit = [(1, lambda est, tgt, ps: O.mse_loss(est, tgt)),
            (0.2, lambda est, tgt, ps: O.sparsity_loss(O.mapped_matrix(fb, ps[2])))]
    tr = O.OracleTrainer(node, params, NFFT, ALIAS_DB, crit, lr=1e-3)
    M = NFFT // 2 + 1
    x = torch.zeros(BATCH, M, 1, dtype=dtype)
    x[:, 0, :] = 1
    y = torch.ones(BATCH, M, 1, dtype=dtype)
    return tr, x, y


def time_cpu_reference(steps: int, warmup: int):
    torch.set_num_threads(os.cpu_count())
    tr, x, y = cpu_reference_trainer()
    loss = None
    for _ in range(warmup):
        tr.train_step(x, y)
    import gc

    ts = []
    gc.collect()
    gc.disable()  # (as in the GPU arm's timed loop)
    try:
        for _ in range(steps):
            t0 = time.perf_counter()
            loss = tr.train_step(x, y)
            ts.append(time.perf_counter() - t0)
    finally:
        gc.enable()
    t = sum(ts) / len(ts)
    M = NFFT // 2 + 1
    return BATCH * M * N_DELAYS / t, t, statistics.median(ts), loss


def reference_arm(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    steps = min(steps, 100)  # ~150 ms per step on 8 cores: keep the run within a few minutes
    value, t, t_med, loss = time_cpu_reference(steps, min(warmup, 10))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": min(warmup, 10), "ms_per_step": t * 1e3, "ms_per_step_median": t_med * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(1, False, "n/a (host CPU)", loss, "host cpu"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                         "sample": f"{steps} full train_step calls of the whole workload (oracle port of the "
                                   "reference algorithm, torch CPU, all host threads)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ parity gate
def parity_gate(device):
    """First-step loss and parameter gradients of the bench model (fresh, seeded) against the CPU oracle evaluated in
    float64 on the same float32 parameter values: loss within 1e-4, every gradient within 1e-3 of its largest entry
    (BASELINE.md §2).  Raises on a mismatch: no number is printed for a wrong kernel."""
    from flamo_b200.optimize.loss import mse_loss, sparsity_loss

    model, ds, _, _, _ = build_gpu_model(device)
    x, y = ds.input[:BATCH].to(device), ds.target[:BATCH].to(device)
    crit_a, crit_b = mse_loss(nfft=NFFT, device=device), sparsity_loss()
    est = model(x)
    loss = crit_a(est, y) + 0.2 * crit_b(est, y, model)
    loss.backward()
    torch.cuda.synchronize()
    params = [p for p in model.parameters()]
    p64 = [p.detach().cpu().double().requires_grad_(p.requires_grad) for p in model.get_core().parameters()]
    tr, xo, yo = cpu_reference_trainer(torch.float64, p64)
    lo, _ = tr.loss(xo, yo)
    go = torch.autograd.grad(lo, [p for p in p64 if p.requires_grad])
    loss, lo = loss.detach(), lo.detach()
    loss_err = abs(float(loss) - float(lo)) / abs(float(lo))
    gerr, k = 0.0, 0
    core_params = [p for p in model.get_core().parameters()]
    for p in core_params:
        if p.requires_grad:
            ref = go[k]
            gerr = max(gerr, float((p.grad.detach().cpu().double() - ref).abs().max() / (ref.abs().max() + 1e-300)))
            k += 1
    out = {"loss": float(loss), "oracle_loss": float(lo), "loss_rel_err": loss_err, "grad_rel_err": gerr,
           "bounds": {"loss": 1e-4, "grad": 1e-3}, "against": "oracle/flamo_oracle.py (CPU, float64, same parameters)"}
    if not (loss_err <= 1e-4 and gerr <= 1e-3):
        raise SystemExit(f"parity gate failed, nothing timed: {json.dumps(out)}")
    del params
    return out


# ------------------------------------------------------------------------------------ all five configs
def build_config_trainer(name, device, world=1, shard=None):
    """Trainer + data for one BASELINE.json config (SURVEY.md §8d restatements, as tools/measure_configs.py)."""
    from flamo_b200 import workloads as W
    from flamo_b200.optimize.loss import mse_loss, sparsity_loss
    from flamo_b200.optimize.trainer import Trainer
    from flamo_b200.processor import dsp, system

    desc, nfft, B, seed, n_ch = W.CONFIGS[name]
    M = nfft // 2 + 1
    torch.manual_seed(seed)
    core = W.build(desc, dsp, system, nfft, W.ALIAS_DECAY_DB, dtype=torch.float32, device=device)
    model = system.Shell(core, dsp.FFT(nfft), dsp.Transform(lambda x: torch.abs(x)))
    n_in = model.input_channels
    if name in ("cfg1_biquad", "cfg3_geq16"):
        x = torch.zeros(B, nfft, n_in, device=device)
        x[:, 0, :] = 1
        torch.manual_seed(seed + 1)
        tcore = W.build(desc, dsp, system, nfft, W.ALIAS_DECAY_DB, dtype=torch.float32, device=device)
        with torch.no_grad():
            tgt = system.Shell(tcore, dsp.FFT(nfft), dsp.Transform(lambda x: torch.abs(x)))(x).clone()
        crits = [(torch.nn.MSELoss(), 1, False)]
    else:
        x = torch.zeros(B, M, n_in, device=device)
        x[:, 0, :] = 1
        tgt = torch.ones(B, M, 1, device=device)
        crits = [(mse_loss(nfft=nfft, device=device), 1, False)]
        if name in ("cfg2_fdn8", "cfg5_fdn64"):
            crits.append((sparsity_loss(), 0.2, True))
    if world > 1:
        from flamo_b200.parallel import DataParallelTrainer

        tr = DataParallelTrainer(model, max_epochs=1, lr=1e-3, log=False, device=device, shard=shard)
    else:
        tr = Trainer(model, max_epochs=1, lr=1e-3, log=False, device=device)
    for c, a, rm in crits:
        tr.register_criterion(c, a, requires_model=rm)
    return tr, x, tgt, B, M, n_ch, n_in, model.output_channels


def time_trainer(tr, data, steps, flush, world=1, device="cuda"):
    """Median and mean device time of `steps` train_step calls (CUDA events on the launching stream, L2 flushed before
    every step, max over ranks)."""
    for _ in range(6):  # lazy state + graph capture
        tr.train_step(data)
    ts = []
    for i in range(steps):
        flush.fill_(i & 0xFF)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        tr.train_step(data)
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    med, mean = statistics.median(ts), statistics.mean(ts)
    if world > 1:
        import torch.distributed as dist

        v = torch.tensor([med, mean], device=device, dtype=torch.float64)
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        med, mean = float(v[0]), float(v[1])
    return med, mean


def measure_fma_peak(device):
    """FP32 FMA rate of this GPU, measured (SURVEY.md §8d "derive + measure"): libfsweep's probe kernel, 16 blocks of
    256 threads per SM, 64 independent FFMAs per thread and round; best of 5 launches timed with CUDA events."""
    from flamo_b200 import _lib

    L = _lib.lib()
    sms = torch.cuda.get_device_properties(device).multi_processor_count
    blocks, iters = sms * 16, 4096
    out = torch.zeros(1, dtype=torch.float32, device=device)
    st = torch.cuda.current_stream(device).cuda_stream
    best = None
    for i in range(7):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        _lib.check(L.fsweep_fma_probe(out.data_ptr(), blocks, iters, st))
        e.record()
        torch.cuda.synchronize()
        t = s.elapsed_time(e) * 1e-3
        if i >= 2:
            best = t if best is None else min(best, t)
    return L.fsweep_fma_probe_flops(blocks, iters) / best / 1e12


def all_configs(device, flush, steps, peaks, fma_tflops, rank):
    """Device-timed Trainer.train_step of the five BASELINE.json configs on ONE GPU, each with the roofline that bounds
    it: HBM (algorithmic bytes of the sweep, x in + |Y| or target in, per step) for the small loops, the measured FP32
    FMA rate for the two compute-bound ones."""
    from flamo_b200 import workloads as W

    out = {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    for name in W.CONFIGS:
        try:
            tr, x, tgt, B, M, n_ch, n_in, n_out = build_config_trainer(name, device)
            med, mean = time_trainer(tr, (x, tgt), steps, flush)
            rec = {"ms_per_step": med, "ms_per_step_mean": mean, "value": B * M * n_ch / (med * 1e-3), "unit": UNIT,
                   "bins": M, "batch": B, "channels": n_ch, "steps": steps, "cuda_graph": bool(tr.use_graph and tr._graphs)}
            if name in STEP_FLOPS:
                ach = STEP_FLOPS[name] / (med * 1e-3) / 1e12
                rec["roofline"] = {"bound": "fp32", "achieved": ach, "peak": fma_tflops, "unit": "TFLOP/s",
                                   "frac": ach / fma_tflops, "peak_source": "measured (fsweep_fma_probe on this GPU)",
                                   "flops_per_step": STEP_FLOPS[name]}
            else:
                byts = B * M * (8 * n_in + 4 * n_out)
                ach = byts / (med * 1e-3) / 1e9
                rec["roofline"] = {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                                   "algorithmic_bytes_per_step": byts,
                                   "note": "launch / latency bound: the whole captured step is timed, not one kernel"}
            del tr
        except Exception as ex:  # one config failing must not hide the others
            rec = {"error": f"{type(ex).__name__}: {ex}"}
        out[name] = rec
        torch.cuda.empty_cache()
    return out


def table_sweep_record(device, flush, peaks, reps=30):
    """The HBM-bound case (the shape of the real examples/e8_active_acoustics.py path): a Series of
    generic FIR filters whose per-bin response tables are streamed from HBM — Filter(100 x 13 x 4) -> parallelFilter(72000
    x 13) -> parallelGain(13) -> Filter(15000 x 4 x 13), nfft = 96000, 4 columns.  Forward and backward sweep launches
    alone (the cuFFT that builds the tables excluded), median of CUDA-event times behind an L2 flush, against the
    measured copy bandwidth.  Algorithmic bytes: tables + x + y forward; tables read + table gradients written + x +
    dL/dy backward."""
    from flamo_b200 import sweep
    from flamo_b200._lib import EPI_NONE
    from flamo_b200.processor import dsp, system

    nfft, alias, n_M, n_L = 96000, 30.0, 4, 13
    M = nfft // 2 + 1
    torch.manual_seed(0)
    kw = dict(nfft=nfft, alias_decay_db=alias, device=device, requires_grad=True)
    core = system.Series(dsp.Filter(size=(100, n_L, n_M), **kw), dsp.parallelFilter(size=(72000, n_L), **kw),
                         dsp.parallelGain(size=(n_L,), **kw), dsp.Filter(size=(15000, n_M, n_L), **kw))
    X = torch.eye(n_M, dtype=torch.complex64, device=device).expand(1, M, n_M, n_M).contiguous()
    with torch.enable_grad():
        prog = sweep.Program(nfft, alias, X.dtype, X.device)
        core._lower(prog, None)
        (tag, payload), = list(prog._segments())
        ops, coefs, n_out = prog.flatten_segment(payload, X.dtype)
    coefs = [c.detach().contiguous() for c in coefs]
    plan = prog.plan_for(ops)
    y = torch.empty((1, M, n_out, n_M), dtype=torch.complex64, device=device)
    gy = torch.ones_like(y)
    grads = [torch.empty_like(c) for c in coefs]
    be = sweep._BACKEND

    def timed(fn):
        """The launch as a one-node CUDA graph (the host side of the ctypes call takes longer than the kernel and must
        not sit inside the timed region), replayed behind an L2 flush."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        ts = []
        for i in range(reps + 5):
            flush.fill_(i & 0xFF)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            g.replay()
            e.record()
            torch.cuda.synchronize()
            if i >= 5:
                ts.append(s.elapsed_time(e) * 1e-3)
        return statistics.median(ts)

    t_f = timed(lambda: be.forward(plan, ops, coefs, X, y, n_M, 0, EPI_NONE))
    t_b = timed(lambda: be.backward(plan, ops, coefs, X, gy, grads, None, n_M, 0, EPI_NONE))
    tab = sum(c.numel() * c.element_size() for c in coefs if c.is_complex())
    io = X.numel() * 8 + y.numel() * 8
    peak = peaks.get("hbm_gbs", 6650.0)
    rec = {"workload": "FIR Series 13x4 -> 13 -> 13 -> 4x13 (generic Filter tables streamed from HBM), nfft=96000, 4 columns",
           "kernel": plan.kernel_family(M, True), "table_bytes": tab}
    for name, t, byts in (("forward", t_f, tab + io), ("backward", t_b, 2 * tab + io)):
        rec[name] = {"us_per_launch": t * 1e6, "algorithmic_bytes": byts,
                     "roofline": {"bound": "hbm", "achieved": byts / t / 1e9, "peak": peak, "unit": "GB/s",
                                  "frac": byts / t / 1e9 / peak}}
    return rec


# ------------------------------------------------------------------------------------ GPU arm
def gpu_arm(args):
    rank, world, local = dist_env()
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device(device))

    from flamo_b200 import sweep

    parity = parity_gate(device) if (rank == 0 and not args.no_parity_gate) else None

    model, ds, Trainer, mse_loss, sparsity_loss = build_gpu_model(device)
    if world > 1:
        from flamo_b200.parallel import DataParallelTrainer as TrainerCls
    else:
        TrainerCls = Trainer
    trainer = TrainerCls(model, max_epochs=1, lr=1e-3, log=False, device=device, graph=not args.no_graph)
    trainer.register_criterion(mse_loss(nfft=NFFT, device=device), 1)
    trainer.register_criterion(sparsity_loss(), 0.2, requires_model=True)

    M = NFFT // 2 + 1
    x_host = ds.input[:BATCH].contiguous().pin_memory()
    y_host = ds.target[:BATCH].contiguous().pin_memory()
    x_dev, y_dev = x_host.to(device), y_host.to(device)

    # L2 flush between timed iterations: write a buffer larger than the 126 MB L2
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=device)

    def barrier():
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    def make_events(steps):
        """torch creates the CUDA event at its first record(): do that here, in front of the barrier, not inside the
        timed loop (cudaEventCreate is a few microseconds of host time per event, and 2 * steps of them skew the ranks)."""
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(steps)] for _ in range(2)]
        for lst in ev:
            for e in lst:
                e.record()
        torch.cuda.synchronize()
        return ev

    def run(steps, data, events=None):
        import gc

        starts, ends = events if events is not None else make_events(steps)
        loss = None
        gc.disable()  # a generation-0 collection in the middle of a 50 us step is a 100+ us outlier (on N ranks: N chances)
        try:
            for i in range(steps):
                flush.fill_(i & 0xFF)
                starts[i].record()
                loss = trainer.train_step(data)
                ends[i].record()
            torch.cuda.synchronize()
        finally:
            gc.enable()
        ts = [s.elapsed_time(e) * 1e-3 for s, e in zip(starts, ends)]
        if os.environ.get("BENCH_DUMP_STEPS"):  # per-step times (us), for looking at outliers
            print(f"[rank {rank}] steps:", " ".join(f"{t * 1e6:.0f}" for t in ts), file=sys.stderr, flush=True)
        return sum(ts), statistics.median(ts), loss

    # lazy state + graph capture happen in the first few calls; they are not part of W
    for _ in range(6):
        trainer.train_step((x_dev, y_dev))
    warmup = max(3, args.warmup)
    run(warmup, (x_dev, y_dev))

    import gc

    sampler = ClockSampler(local)
    ev_dev, ev_e2e = make_events(args.steps), make_events(args.steps)
    gc.collect()  # (in front of the barrier: a collection behind it would skew the ranks' start by milliseconds)
    sampler.start()  # (its thread starts up in front of the barrier too: the first timed steps are not perturbed by it)
    barrier()
    launches0 = sweep.launch_count
    t_dev, med_dev, loss = run(args.steps, (x_dev, y_dev), ev_dev)
    barrier()
    launches = sweep.launch_count - launches0
    # keep the GPU under the same load long enough for a few clock samples; a FIXED number of steps,
    # because every step is a collective when world > 1
    run(max(300, args.steps), (x_dev, y_dev))
    clocks = sampler.stop()

    # end to end: inputs in pinned host memory, H2D inside the timed region, loss read back
    run(warmup, (x_host, y_host))
    gc.collect()
    barrier()
    t_e2e, med_e2e, _ = run(args.steps, (x_host, y_host), ev_e2e)
    barrier()

    def reduce_max(*ts):
        if world == 1:
            return ts
        import torch.distributed as dist

        v = torch.tensor(list(ts), device=device, dtype=torch.float64)
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return tuple(float(a) for a in v)

    t_dev, t_e2e, med_dev, med_e2e = reduce_max(t_dev, t_e2e, med_dev, med_e2e)
    units_per_step = world * BATCH * M * N_DELAYS
    value = units_per_step * args.steps / t_dev
    e2e_value = units_per_step * args.steps / t_e2e

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    fma_tflops = measure_fma_peak(device) if rank == 0 else None
    roof = kernel_roofline(model, x_dev, flush, peaks, fma_tflops) if rank == 0 else None
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, t, t_med, _ = time_cpu_reference(steps=40, warmup=3)
        cpu = {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
               "sample": "40 full train_step calls of the whole workload (oracle port of the reference "
                         f"algorithm, torch CPU float32, all host threads): {t * 1e3:.1f} ms/step"}

    # the multi-GPU step of the bench model holds NCCL / peer-memory state: release its graphs before other trainers run
    extra = {}
    if world == 1 and rank == 0 and not args.no_configs:
        extra["configs"] = all_configs(device, flush, args.config_steps, peaks, fma_tflops, rank)
        try:
            extra["table_sweep"] = table_sweep_record(device, flush, peaks)
        except Exception as ex:
            extra["table_sweep"] = {"error": f"{type(ex).__name__}: {ex}"}
    if world > 1 and not args.no_configs:
        # BASELINE.json configs[4] as north_star describes it: bin ranges sharded over the GPUs of the node, one
        # all-reduce of the gradients per step (strong scaling: the total work is fixed)
        try:
            tr5, x5, t5, B5, M5, n5, _, _ = build_config_trainer("cfg5_fdn64", device, world, "bins")
            med5, mean5 = time_trainer(tr5, (x5, t5), args.config_steps, flush, world, device)
            extra["cfg5_bins_sharded"] = {"ms_per_step": med5, "ms_per_step_mean": mean5, "n_gpus": world,
                                          "value": B5 * M5 * n5 / (med5 * 1e-3), "unit": UNIT, "scaling": "strong",
                                          "bins_per_rank": M5 // world, "batch": B5, "steps": args.config_steps}
            tr5._graphs.clear()
        except Exception as ex:
            extra["cfg5_bins_sharded"] = {"error": f"{type(ex).__name__}: {ex}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": t_dev / args.steps * 1e3, "ms_per_step_median": med_dev * 1e3,
            "value_median": units_per_step / med_dev, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(world, bool(trainer.use_graph), "flushed between timed steps (192 MB fill)", loss,
                                  "B200"),
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": t_e2e / args.steps * 1e3,
                    "ms_per_step_median": med_e2e * 1e3,
                    "h2d_bytes_per_step": x_host.numel() * x_host.element_size() + y_host.numel() * y_host.element_size(),
                    "d2h_bytes_per_step": 4 * (trainer.n_loss + 1)},
            "gpu_launches": launches,
            "clocks": clocks,
            "parity": parity,
            "roofline": roof,
            "cpu_baseline": cpu,
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        # a captured graph that contains NCCL work keeps the communicator busy: ProcessGroupNCCL's teardown
        # (destroy_process_group, or its destructor at interpreter exit) then blocks.  Drop the graphs, sync,
        # meet at a last barrier and leave without running the destructors.
        import torch.distributed as dist

        trainer._graphs.clear()
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


def kernel_roofline(model, x_dev, flush, peaks, fma_tflops, reps=30):
    """Duration of the dominant kernel (the backward sweep) and of the forward sweep, each launched
    alone behind an L2 flush (as a one-node CUDA graph, so that host launch overhead is excluded), CUDA events on the
    launching stream; algorithmic bytes per launch =
    B*M*(8*N_in + 4*N_out) (x read as complex64, |Y| or dL/d|Y| as float32)."""
    from flamo_b200 import sweep
    from flamo_b200._lib import EPI_ABS

    peak, which = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
    # dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from THIS round's `ncu --set full` capture
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            tj = json.load(f)
        traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
    except Exception:
        pass

    core = model.get_core()
    with torch.enable_grad():  # lowered with grad enabled so the ops carry FSWEEP_F_GRAD like a real step
        X = model.get_inputLayer()(x_dev)
        prog = sweep.Program(NFFT, ALIAS_DB, X.dtype, X.device)
        core._lower(prog, None)
        segs = list(prog._segments())
        assert len(segs) == 1 and segs[0][0] == "sweep"
        ops, coefs, n_out = prog.flatten_segment(segs[0][1], X.dtype)
    coefs = [c.detach().contiguous() for c in coefs]
    plan = prog.plan_for(ops)
    x4 = X.detach().reshape(X.shape[0], X.shape[1], X.shape[2], 1).contiguous()
    y = torch.empty((x4.shape[0], x4.shape[1], n_out, 1), dtype=torch.float32, device=X.device)
    M = X.shape[1]
    bytes_per_launch = BATCH * M * (8 * 1 + 4 * 1)
    backend = sweep._BACKEND

    def timed(call):
        """Median duration of `call`'s device work: the launch is captured once into a CUDA graph and the replay is
        timed with events, so the Python / ctypes launch path (tens of microseconds, longer than these kernels) is
        not inside the timed region; L2 is flushed before every replay."""
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            for _ in range(3):
                call()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            call()
        ts = []
        for i in range(reps + 5):
            flush.fill_(i & 0xFF)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            graph.replay()
            e.record()
            torch.cuda.synchronize()
            if i >= 5:
                ts.append(s.elapsed_time(e) * 1e-3)
        return statistics.median(ts)

    t_fwd = timed(lambda: backend.forward(plan, ops, coefs, x4, y, 1, 0, EPI_ABS))

    # The training step's ONE sweep launch: backward kernel with the fused |.| + mse_loss criterion (it recomputes
    # the forward states, forms dL/d|Y| from the target and sums the loss).  No gradient buffers are passed, so the
    # only other launch inside the timed call is the one-warp loss finalize.
    from flamo_b200._lib import CRIT_MSE_CHSUM

    tgt = torch.ones((x4.shape[0], M), dtype=torch.float32, device=X.device)
    loss = torch.empty((), dtype=torch.float32, device=X.device)

    def bwd():
        backend.loss(plan, ops, coefs, x4, tgt, CRIT_MSE_CHSUM, 1.0 / tgt.numel(), loss, [None] * len(coefs), None, 0)

    t_bwd = timed(bwd)
    ach = bytes_per_launch / t_bwd / 1e9
    return {"bound": "hbm", "kernel": plan.kernel_family(M, True) + " (fused |.|+MSE criterion)", "achieved": ach,
            "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "peak_source": which + " (MEASURED_PEAKS.json hbm_gbs)" if which == "measured" else which,
            "traffic": traffic, "traffic_source": traffic_src,
            "us_per_launch": t_bwd * 1e6, "algorithmic_bytes_per_launch": bytes_per_launch,
            "us_per_launch_note": "median over graph replays of the sweep kernel + one-warp loss finalize, launch "
                                  "latency included",
            "forward_kernel": {"kernel": plan.kernel_family(M, False) + " (validation / inference only; not in the training step)",
                               "us_per_launch": t_fwd * 1e6,
                               "achieved": bytes_per_launch / t_fwd / 1e9, "frac": bytes_per_launch / t_fwd / 1e9 / peak},
            "fp32": {"flops_per_launch": FLOPS_PER_BIN * M, "achieved_tflops": FLOPS_PER_BIN * M / t_bwd / 1e12,
                     "measured_peak_tflops": fma_tflops, "nominal_peak_tflops": FP32_NOMINAL_TFLOPS,
                     "frac": FLOPS_PER_BIN * M / t_bwd / 1e12 / (fma_tflops or FP32_NOMINAL_TFLOPS),
                     "peak_source": "measured (fsweep_fma_probe on this GPU)" if fma_tflops else "nominal"},
            "note": "config 2 moves 0.58 MB per launch and does ~3.7 kflop per bin: it is dependent-issue latency "
                    "bound (one bin per thread), not HBM bound (SURVEY.md §8d); the HBM fraction is reported as the "
                    "contract asks, the FP32 fraction (of the measured FMA rate) beside it"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--no-graph", action="store_true", help="run train_step eagerly (profiling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-gate", action="store_true", help="skip the oracle check before timing (profiling)")
    ap.add_argument("--no-configs", action="store_true", help="headline only: skip the per-config sub-records")
    ap.add_argument("--config-steps", type=int, default=30, help="timed steps per config in the sub-records")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
