"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

import cases as C
from flamo_b200 import workloads as W
from flamo_b200.processor import dsp, system

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def golden_params(g):
    out, i = [], 0
    while f"param_{i}" in g:
        out.append(g[f"param_{i}"])
        i += 1
    return out


def build_case(name, dtype, device):
    """This package's module tree for a parity case, with the reference's raw parameters."""
    case, g = C.CASES[name], load_golden(name)
    model = W.build(case["desc"], dsp, system, case["nfft"], case["alias"], dtype=dtype, device=device)
    W.set_params(model, golden_params(g))
    return case, g, model


def rel_err(Y, Yref, floor=1e-3):
    """|Y - Yref| / max(|Yref|, floor * max|Yref|)  (BASELINE.md §2 metric)."""
    Y, Yref = np.asarray(Y), np.asarray(Yref)
    den = np.maximum(np.abs(Yref), floor * np.abs(Yref).max())
    return float((np.abs(Y - Yref) / den).max())


def grad_err(g, gref):
    g, gref = np.asarray(g, dtype=np.float64), np.asarray(gref, dtype=np.float64)
    return float(np.abs(g - gref).max() / (np.abs(gref).max() + 1e-300))
