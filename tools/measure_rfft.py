#!/usr/bin/env python
"""Input transform of a step: libfsweep's two-launch FFT against cuFFT (torch.fft.rfft), one (1, nfft, 1) float32
signal, CUDA events around 20 back-to-back calls inside a CUDA graph (how a captured step sees them)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flamo_b200 import sweep  # noqa: E402


def graph_time(fn, reps=20, rounds=30):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
    torch.cuda.current_stream().wait_stream(s)
    ts = []
    for _ in range(rounds):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / reps)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    out = {}
    for nfft, B in ((96000, 1), (192000, 1), (384000, 32), (4096, 1)):
        x = torch.randn(B, nfft, 1, device="cuda")
        t_own = graph_time(lambda: sweep.rfft(x, nfft, force=True))
        t_cufft = graph_time(lambda: torch.fft.rfft(x, n=nfft, dim=1))
        out[f"nfft{nfft}_b{B}"] = {"fsweep_rfft_us": round(t_own, 2), "cufft_us": round(t_cufft, 2)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
