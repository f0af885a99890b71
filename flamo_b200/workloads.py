"""Model descriptions for the BASELINE.json configs and parity cases.

A description is plain data:

    ("Series", [desc, ...], [key, ...] | None)
    ("Recursion", ff_desc, fb_desc)
    ("Parallel", a_desc, b_desc [, sum_output])
    (ClassName, ctor_kwargs [, post])        post: {"delay_samples": [...]} | {"assign": tensor} | {"requires_grad": True}

`build(desc, dsp, system, ...)` instantiates it with any library exposing the flamo class
API — this package's `processor.dsp/system`, or the reference's own modules (used only by
tests/golden/make_golden.py in the build container).  Because both sides are built from the
same description, raw parameters can be copied 1:1 in `nn.Module.parameters()` order.
"""
from __future__ import annotations

from collections import OrderedDict

import torch

FS = 48000


def build(desc, dsp, system, nfft, alias_decay_db, dtype=torch.float32, device=None):
    name = desc[0]
    if name == "Series":
        mods = [build(d, dsp, system, nfft, alias_decay_db, dtype, device) for d in desc[1]]
        keys = desc[2] if len(desc) > 2 and desc[2] else None
        if keys:
            return system.Series(OrderedDict(zip(keys, mods)))
        return system.Series(*mods)
    if name == "Recursion":
        ff = build(desc[1], dsp, system, nfft, alias_decay_db, dtype, device)
        fb = build(desc[2], dsp, system, nfft, alias_decay_db, dtype, device)
        return system.Recursion(fF=ff, fB=fb)
    if name == "Parallel":
        a = build(desc[1], dsp, system, nfft, alias_decay_db, dtype, device)
        b = build(desc[2], dsp, system, nfft, alias_decay_db, dtype, device)
        return system.Parallel(a, b, sum_output=desc[3] if len(desc) > 3 else True)
    kwargs = dict(desc[1])
    mod = getattr(dsp, name)(nfft=nfft, alias_decay_db=alias_decay_db, dtype=dtype, device=device, **kwargs)
    post = desc[2] if len(desc) > 2 else None
    if post:
        if "delay_samples" in post:
            d = torch.tensor(post["delay_samples"], dtype=dtype, device=device)
            mod.assign_value(mod.sample2s(d))
        if "assign" in post:
            mod.assign_value(torch.as_tensor(post["assign"], dtype=dtype, device=device))
        if post.get("requires_grad"):  # modules whose constructor has no requires_grad argument (SOSFilter)
            mod.param.requires_grad_(True)
    return mod


def set_params(module, values):
    """Copy raw parameter values (parameters() order) into a built module tree."""
    ps = list(module.parameters())
    assert len(ps) == len(values), (len(ps), len(values))
    with torch.no_grad():
        for p, v in zip(ps, values):
            p.copy_(torch.as_tensor(v).to(dtype=p.dtype, device=p.device).reshape(p.shape))


# ------------------------------------------------------------------ BASELINE configs


def _primes(lo, hi):
    sieve = [True] * (hi + 1)
    out = []
    for i in range(2, hi + 1):
        if sieve[i]:
            if i >= lo:
                out.append(i)
            for j in range(i * i, hi + 1, i):
                sieve[j] = False
    return out


def fdn_delays(N: int):
    """config 2 uses the example's six delays plus two primes; config 5 uses
    64 distinct primes in [500, 3000] drawn with seed 0."""
    base = [887, 911, 941, 1699, 1951, 2053, 2129, 2287]
    if N <= 8:
        return base[:N]
    g = torch.Generator().manual_seed(0)
    pr = _primes(500, 3000)
    idx = torch.randperm(len(pr), generator=g)[:N].tolist()
    return sorted(pr[i] for i in idx)


def fdn(N: int, delays=None, isint=True):
    """examples/e8_colorless_fdn.py:28-100 with N delay lines."""
    delays = delays or fdn_delays(N)
    return (
        "Series",
        [
            ("Gain", dict(size=(N, 1), requires_grad=True)),
            (
                "Recursion",
                ("parallelDelay", dict(size=(N,), max_len=max(delays), isint=isint, requires_grad=False),
                 {"delay_samples": delays}),
                ("Matrix", dict(size=(N, N), matrix_type="orthogonal", requires_grad=True)),
            ),
            ("Gain", dict(size=(1, N), requires_grad=True)),
        ],
        ["input_gain", "feedback_loop", "output_gain"],
    )


def biquad(out_ch=2, in_ch=1, n_sections=2, filter_type="highpass"):
    """examples/e7_biquad.py:15-60."""
    return ("Biquad", dict(size=(out_ch, in_ch), n_sections=n_sections, filter_type=filter_type,
                           fs=FS, requires_grad=True))


def geq(out_ch=16, in_ch=16, octave_interval=3):
    """examples/e7_geq.py restated per SURVEY §8d config 3."""
    return ("Series", [("GEQ", dict(size=(out_ch, in_ch), octave_interval=octave_interval, fs=FS,
                                    requires_grad=True))])


def active_acoustics(n_M=4, n_L=13):
    """SURVEY §8d config 4 (synthetic restatement of e8_active_acoustics dimensions)."""
    return (
        "Series",
        [
            ("Gain", dict(size=(n_M, 1), requires_grad=True)),
            (
                "Recursion",
                ("Series", [
                    ("SVF", dict(size=(n_L, n_M), n_sections=2, filter_type=None, fs=FS, requires_grad=True)),
                    ("parallelDelay", dict(size=(n_L,), isint=False, fs=FS, requires_grad=True)),
                    ("parallelGain", dict(size=(n_L,), requires_grad=True), {"assign": [0.3] * n_L}),
                ]),
                ("Series", [
                    ("Delay", dict(size=(n_M, n_L), isint=False, fs=FS, requires_grad=False)),
                    ("parallelGain", dict(size=(n_M,), requires_grad=True), {"assign": [0.05] * n_M}),
                ]),
            ),
            ("Gain", dict(size=(1, n_L), requires_grad=True)),
        ],
        ["input_gain", "feedback_loop", "output_gain"],
    )


CONFIGS = {
    # name: (description, nfft, batch, seed, N_ch)
    "cfg1_biquad": (biquad(), 96000, 1, 130709, 2),
    "cfg2_fdn8": (fdn(8), 96000, 1, 130709, 8),
    "cfg3_geq16": (geq(), 192000, 1, 130710, 16),
    "cfg4_active": (active_acoustics(), 96000, 1, 130297, 13),
    "cfg5_fdn64": (fdn(64), 384000, 32, 0, 64),
}
ALIAS_DECAY_DB = 30
