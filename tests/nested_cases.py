"""Module trees that nest a Recursion or a Parallel INSIDE a Recursion path (the reference allows it: its Recursion only
ever applies its paths to signals, system.py:397-425).  The flat sweep program cannot express them in place; they enter
the outer program as streamed response tables (sweep.Program.table_of)."""
import torch

from flamo_b200 import workloads as W

FS = W.FS


def _mat(n, m, scale, seed):
    g = torch.Generator().manual_seed(seed)
    return (scale * torch.randn(n, m, generator=g, dtype=torch.float64)).tolist()


def _matrix(n, m, scale, seed):
    return ("Matrix", dict(size=(n, m), matrix_type="random", requires_grad=True), {"assign": _mat(n, m, scale, seed)})


def _pgain(n, value):
    return ("parallelGain", dict(size=(n,), requires_grad=True), {"assign": [value] * n})


def _pdelay(n, first):
    return ("parallelDelay", dict(size=(n,), max_len=first + 40 * n, isint=True, requires_grad=False),
            {"delay_samples": [first + 37 * i for i in range(n)]})


NESTED = {
    # a loop whose feedforward path contains another loop
    "loop_in_feedforward": ("Series", [
        ("Gain", dict(size=(3, 2), requires_grad=True)),
        ("Recursion",
         ("Series", [_pdelay(3, 11), ("Recursion", _pgain(3, 0.5), _matrix(3, 3, 0.2, 1))]),
         _matrix(3, 3, 0.15, 2)),
        ("Gain", dict(size=(1, 3), requires_grad=True)),
    ]),
    # a loop whose feedback path IS another loop (not wrapped in a Series)
    "loop_as_feedback": ("Series", [
        ("Gain", dict(size=(2, 1), requires_grad=True)),
        ("Recursion", _pdelay(2, 7), ("Recursion", _matrix(2, 2, 0.3, 3), _matrix(2, 2, 0.3, 4))),
        ("Gain", dict(size=(2, 2), requires_grad=True)),
    ]),
    # a Parallel (sum of two branches) in the feedback path
    "parallel_in_feedback": ("Series", [
        ("Gain", dict(size=(3, 1), requires_grad=True)),
        ("Recursion", _pdelay(3, 5),
         ("Parallel", _matrix(3, 3, 0.2, 5), ("Series", [_pgain(3, 0.4), _matrix(3, 3, 0.2, 6)]), True)),
        ("Gain", dict(size=(1, 3), requires_grad=True)),
    ]),
    # a Parallel that concatenates, as a whole feedforward path (rectangular loop: 2 -> 4 -> 2)
    "parallel_cat_as_feedforward": ("Series", [
        ("Gain", dict(size=(2, 1), requires_grad=True)),
        ("Recursion", ("Parallel", _matrix(2, 2, 0.3, 7), _matrix(2, 2, 0.3, 8), False), _matrix(2, 4, 0.2, 9)),
        ("Gain", dict(size=(1, 4), requires_grad=True)),
    ]),
}
