"""Runs the UNMODIFIED model class of the reference's NN-in-the-loop example (examples/e7_biquad_nn.py: `nnBiquad`, an MLP
whose output conditions a Biquad through `ext_param`, one Shell call per batch item, :149-156) on the reference itself
or on flamo_b200 installed under the name `flamo` (C ABI emulated on the CPU), and saves output and gradients.  The
script's own main path cannot run on either engine: its `Dataset(..., dtype=...)` call (:178) does not match the class it
defines (:35) — an upstream defect, the same TypeError on both.  With the b200 engine the per-item loop is additionally
replaced by ONE call with the (B, ...) parameter tensor (per-item parameter sets), which must reproduce the loop.

    python tests/run_reference_nn_model.py {reference|b200} <out.npz>
"""
import os
import runpy
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SCRIPT = "/root/reference/examples/e7_biquad_nn.py"


def main():
    engine, out = sys.argv[1:3]
    if engine == "reference":
        for name in ("soundfile", "nnAudio", "nnAudio.features", "pyfar", "matplotlib", "matplotlib.pyplot"):
            if name not in sys.modules:
                try:
                    __import__(name)
                except Exception:
                    sys.modules[name] = types.ModuleType(name)
        sys.path.insert(0, "/root/reference")
    else:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, HERE)
        import flamo_b200

        flamo_b200.install_as_flamo()
        import cpu_emulator

        cpu_emulator.install()
    ns = runpy.run_path(SCRIPT, run_name="nn_example")  # defines the classes, does not run the example
    args = types.SimpleNamespace(nfft=1024, samplerate=48000, dtype=torch.float64, device="cpu")
    in_ch, out_ch, n_sect, n_param, B = 1, 4, 1, 2, 3
    torch.manual_seed(5)
    model = ns["nnBiquad"](n_sect, n_param, in_ch, out_ch, args)
    M = args.nfft // 2 + 1
    target = torch.rand(B, M, out_ch)
    z = torch.zeros(B, args.nfft, in_ch, dtype=args.dtype)
    z[:, 0] = 1
    y = model((target, z))
    loss = ns["dBMSELoss"]()(y, target.to(args.dtype))
    loss.backward()
    res = {"y": y.detach().numpy(), "loss": float(loss)}
    for i, (k, p) in enumerate(model.named_parameters()):
        if p.grad is not None:
            res[f"grad_{i}_{k}"] = p.grad.numpy()
    res["biquad_param"] = [p for k, p in model.named_parameters() if "biquad" in k][0].detach().numpy()
    if engine == "b200":
        # the same forward with ONE Shell call for the whole batch
        x = torch.abs(target).permute(0, 2, 1)
        x = model.final_dense(model.stack(x)).view(-1, n_sect, n_param, out_ch, in_ch)
        x = torch.cat((torch.sigmoid(x[:, :, :1] * 0.25), x[:, :, 1:]), dim=2)
        yb = model.biquad(z[0].unsqueeze(0), {"biquad": x.to(args.dtype)})
        assert yb.shape == y.shape
        res["batched_max_diff"] = float((yb - y).abs().max())
        gb = torch.autograd.grad(ns["dBMSELoss"]()(yb, target.to(args.dtype)), list(model.final_dense.parameters()))
        res["batched_grad_diff"] = max(float((a - p.grad).abs().max()) for a, p in zip(gb, model.final_dense.parameters()))
    np.savez(out, **res)


if __name__ == "__main__":
    main()
