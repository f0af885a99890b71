"""Per-item parameter sets vs the reference's per-item loop (examples/e7_biquad_nn.py:149-156): a Shell(FFT -> Biquad ->
|.|) conditioned with B parameter sets, forward + backward to the parameter sets.

  python tools/measure_per_item.py [--nfft 4096] [--batch 64] [--out-ch 4] [--sections 1]

Prints one JSON line: ms per forward+backward for (a) one call with the (B, ...) parameter tensor — one design launch and
ONE sweep launch each way — and (b) the loop of B calls the reference example uses (run through this package's kernels:
B design launches, B sweep launches each way).  CUDA events, warm, median of `--reps`."""
import argparse
import json
import os
import sys
from collections import OrderedDict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nfft", type=int, default=4096)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--out-ch", type=int, default=4)
    ap.add_argument("--sections", type=int, default=1)
    ap.add_argument("--reps", type=int, default=30)
    a = ap.parse_args()
    import per_item_cases as PC
    from flamo_b200 import sweep
    from flamo_b200.processor import dsp, system

    dt = torch.float32
    filt = dsp.Biquad(size=(a.out_ch, 1), n_sections=a.sections, filter_type="highpass", nfft=a.nfft, fs=48000,
                      alias_decay_db=30, device="cuda", dtype=dt)
    model = system.Shell(core=OrderedDict({"biquad": filt}), input_layer=dsp.FFT(a.nfft, dtype=dt),
                         output_layer=dsp.Transform(lambda x: torch.abs(x), dtype=dt))
    P = PC.draw("biquad", (a.batch,) + tuple(filt.param.shape), 1).to(dtype=dt, device="cuda").requires_grad_(True)
    z = torch.zeros(1, a.nfft, 1, dtype=dt, device="cuda")
    z[:, 0] = 1

    def batched():
        Y = model(z, {"biquad": P})
        return torch.autograd.grad(Y.square().sum(), P)[0]

    def looped():
        Y = torch.vstack([model(z, {"biquad": P[i]}) for i in range(a.batch)])
        return torch.autograd.grad(Y.square().sum(), P)[0]

    def timed(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(a.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]

    ga, gl = batched(), looped()
    err = float((ga - gl).abs().max() / gl.abs().max())
    n0 = sweep.launch_count
    batched()
    nb = sweep.launch_count - n0
    looped()
    nl = sweep.launch_count - n0 - nb
    print(json.dumps({"workload": f"Shell(FFT, Biquad 1->{a.out_ch} x {a.sections} sections, abs), nfft {a.nfft}, {a.batch} parameter sets, forward + backward",
                      "per_item_call_ms": round(timed(batched), 4), "per_item_loop_ms": round(timed(looped), 4),
                      "sweep_launches": {"call": nb, "loop": nl}, "grad_rel_diff": err}))


if __name__ == "__main__":
    main()
