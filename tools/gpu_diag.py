"""Diagnostic (not a test): print forward / gradient errors of every parity case on the GPU."""
import sys
import time
import traceback

import torch

import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases as C  # noqa: E402
from test_gpu_parity import WIDE, run_case  # noqa: E402

only = sys.argv[1:]
for dtype in (torch.float32, torch.float64):
    for name in C.CASES:
        if only and name not in only:
            continue
        if name in WIDE and dtype == torch.float64:
            continue
        t0 = time.time()
        try:
            case, g, Y, ferr, gerrs, missing = run_case(name, dtype)
            worst = max(gerrs.values()) if gerrs else 0.0
            print(f"{str(dtype)[6:]:8s} {name:26s} fwd {ferr:.2e} |mag| {run_case.mag_err:.2e}  grad {worst:.2e} {dict((k, float(f'{v:.1e}')) for k, v in gerrs.items())}"
                  f" missing={missing}  ({time.time() - t0:.1f}s)", flush=True)
        except Exception as e:
            print(f"{str(dtype)[6:]:8s} {name:26s} EXC {type(e).__name__}: {e}", flush=True)
            traceback.print_exc()
            if "CUDA" in str(e) or "cuda" in str(e):
                sys.exit(1)
