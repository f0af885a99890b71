#!/usr/bin/env python
"""HBM-bound sweep: a Series of generic FIR filters (TABLE ops) — the shape of the real examples/e8_active_acoustics.py
path: Filter(100 x 13 x 4) -> parallelFilter(72000 x 13) -> parallelGain(13) -> Filter(15000 x 4 x 13),
nfft = 96000, identity input (4 columns).  The per-bin response tables are streamed from HBM by the sweep kernels, so
this is where the bandwidth roofline of BASELINE.json's north_star applies.  Prints the forward / backward kernel
times (CUDA events, L2 flushed, the cuFFT that builds the tables excluded) and achieved GB/s against MEASURED_PEAKS."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flamo_b200 import sweep  # noqa: E402
from flamo_b200._lib import EPI_NONE  # noqa: E402
from flamo_b200.processor import dsp, system  # noqa: E402

DEV = "cuda"


def main():
    nfft, alias = 96000, 30.0
    M = nfft // 2 + 1
    n_M, n_L = 4, 13
    torch.manual_seed(0)
    kw = dict(nfft=nfft, alias_decay_db=alias, device=DEV, requires_grad=True)
    core = system.Series(
        dsp.Filter(size=(100, n_L, n_M), **kw),
        dsp.parallelFilter(size=(72000, n_L), **kw),
        dsp.parallelGain(size=(n_L,), **kw),
        dsp.Filter(size=(15000, n_M, n_L), **kw),
    )
    X = torch.eye(n_M, dtype=torch.complex64, device=DEV).expand(1, M, n_M, n_M).contiguous()
    with torch.enable_grad():
        prog = sweep.Program(nfft, alias, X.dtype, X.device)
        core._lower(prog, None)
        (tag, payload), = list(prog._segments())
        ops, coefs, n_out = prog.flatten_segment(payload, X.dtype)
    coefs = [c.detach().contiguous() for c in coefs]
    plan = prog.plan_for(ops)
    cols = n_M
    y = torch.empty((1, M, n_out, cols), dtype=torch.complex64, device=DEV)
    gy = torch.ones_like(y)
    grads = [torch.empty_like(c) for c in coefs]
    be = sweep._BACKEND
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=DEV)

    def timed(fn, reps=30):
        # the launch as a one-node CUDA graph: the host side of the ctypes call takes longer than the kernel
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
        ts.sort()
        return ts[len(ts) // 2]

    t_f = timed(lambda: be.forward(plan, ops, coefs, X, y, cols, 0, EPI_NONE))
    t_b = timed(lambda: be.backward(plan, ops, coefs, X, gy, grads, None, cols, 0, EPI_NONE))
    tab = sum(c.numel() * c.element_size() for c in coefs if c.is_complex())
    io = X.numel() * 8 + y.numel() * 8
    bytes_f = tab + io
    bytes_b = 2 * tab + io  # tables read, table gradients written, x and dL/dy read
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    out = {"workload": "FIR Series 13x4 -> 13 -> 13 -> 4x13, nfft=96000, 4 columns", "bins": M,
           "table_bytes": tab, "forward": {"us": t_f * 1e6, "bytes": bytes_f, "GBps": bytes_f / t_f / 1e9,
                                           "frac_of_hbm_peak": bytes_f / t_f / 1e9 / peak},
           "backward": {"us": t_b * 1e6, "bytes": bytes_b, "GBps": bytes_b / t_b / 1e9,
                        "frac_of_hbm_peak": bytes_b / t_b / 1e9 / peak},
           "kernel": plan.kernel_family(M, True), "hbm_peak_GBps": peak}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
