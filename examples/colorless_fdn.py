#!/usr/bin/env python
"""Colourless feedback delay network, the workflow of the reference's examples/e8_colorless_fdn.py, on the B200 engine.

The script is written against the `flamo.*` names on purpose: `flamo_b200.install_as_flamo()` registers this package
under that name, so the body below is what a flamo user already has.  Needs a CUDA device (there is no CPU sweep).

    python examples/colorless_fdn.py --max_epochs 5 --num 64
"""
import argparse
import os
import sys
import tempfile
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flamo_b200  # noqa: E402

flamo_b200.install_as_flamo()

from flamo.optimize.dataset import DatasetColorless, load_dataset  # noqa: E402
from flamo.optimize.loss import mse_loss, sparsity_loss  # noqa: E402
from flamo.optimize.trainer import Trainer  # noqa: E402
from flamo.processor import dsp, system  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nfft", type=int, default=96000)
    ap.add_argument("--samplerate", type=int, default=48000)
    ap.add_argument("--num", type=int, default=64, help="dataset size")
    ap.add_argument("--batch_size", type=int, default=1)
    ap.add_argument("--max_epochs", type=int, default=5)
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--device", default="cuda")
    args = ap.parse_args()
    torch.manual_seed(130709)
    dtype, dev, alias = torch.float32, args.device, 30.0

    delays = torch.tensor([887, 911, 941, 1699, 1951, 2053, 2129, 2287], dtype=dtype)
    N = len(delays)
    kw = dict(nfft=args.nfft, alias_decay_db=alias, device=dev, dtype=dtype)
    input_gain = dsp.Gain(size=(N, 1), requires_grad=True, **kw)
    output_gain = dsp.Gain(size=(1, N), requires_grad=True, **kw)
    delay_lines = dsp.parallelDelay(size=(N,), max_len=int(delays.max()), isint=True, requires_grad=False,
                                    fs=args.samplerate, **kw)
    delay_lines.assign_value(delay_lines.sample2s(delays.to(dev)))
    mixing = dsp.Matrix(size=(N, N), matrix_type="orthogonal", requires_grad=True, **kw)
    loop = system.Recursion(fF=delay_lines, fB=mixing)
    from collections import OrderedDict

    fdn = system.Series(OrderedDict({"input_gain": input_gain, "feedback_loop": loop, "output_gain": output_gain}))
    model = system.Shell(core=fdn, input_layer=dsp.FFT(args.nfft, dtype=dtype),
                         output_layer=dsp.Transform(transform=lambda x: torch.abs(x), dtype=dtype))

    with torch.no_grad():
        ir0 = model.get_time_response(identity=False, fs=args.samplerate).squeeze()
        mag0 = model.get_freq_response(identity=False, fs=args.samplerate).abs().squeeze()

    dataset = DatasetColorless(input_shape=(1, args.nfft // 2 + 1, 1), target_shape=(1, args.nfft // 2 + 1, 1),
                               expand=args.num, device=dev, dtype=dtype)
    train_loader, valid_loader = load_dataset(dataset, batch_size=args.batch_size)
    with tempfile.TemporaryDirectory() as train_dir:
        trainer = Trainer(model, max_epochs=args.max_epochs, lr=args.lr, train_dir=train_dir, device=dev)
        trainer.register_criterion(mse_loss(nfft=args.nfft, device=dev), 1)
        trainer.register_criterion(sparsity_loss(), 0.2, requires_model=True)
        t0 = time.time()
        trainer.train(train_loader, valid_loader)
        torch.cuda.synchronize()
        dt = time.time() - t0
        saved = sorted(os.listdir(os.path.join(train_dir, "checkpoints")))
    with torch.no_grad():
        mag1 = model.get_freq_response(identity=False, fs=args.samplerate).abs().squeeze()
    steps = args.max_epochs * len(train_loader)
    print(f"impulse response: {tuple(ir0.shape)}; spectral flatness (std of |H|): {float(mag0.std()):.4f} -> {float(mag1.std()):.4f}")
    print(f"train loss: {trainer.train_loss[0]:.4f} -> {trainer.train_loss[-1]:.4f}; valid loss: {trainer.valid_loss[-1]:.4f}")
    print(f"{steps} training steps + validation in {dt:.2f} s wall clock; checkpoints: {saved}")
    assert trainer.train_loss[-1] < trainer.train_loss[0], "the loss did not go down"


if __name__ == "__main__":
    main()
