"""CPU (build container only): the reference's own example scripts, UNMODIFIED, run on flamo_b200 registered under the
name `flamo` (flamo_b200.install_as_flamo) — and, side by side, on the reference itself with the same seed.  Both runs
train for a few epochs and save their checkpoints through Trainer.save_model; the state dicts must have the same keys
and the same values.  This is the evidence behind "existing examples run unchanged" (BASELINE.json north_star).
The kernels are not involved here (no GPU in the build container: the C ABI is emulated, tests/cpu_emulator.py); the
same module code drives libfsweep.so on a B200 (tests/test_gpu_*.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "examples")), reason="reference checkout not present")

COMMON = ["--nfft", "2048", "--num", "8", "--max_epochs", "3", "--device", "cpu", "--batch_size", "1"]


def run(engine, script, train_dir, extra):
    cmd = [sys.executable, os.path.join(HERE, "run_reference_example.py"), engine, os.path.join(REF, "examples", script),
           str(train_dir)] + COMMON + extra
    env = dict(os.environ, OMP_NUM_THREADS="4", MKL_NUM_THREADS="4")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, f"{engine} run of {script} failed:\n{r.stdout[-2000:]}\n{r.stderr[-4000:]}"
    return r.stdout


def last_checkpoint(train_dir):
    d = os.path.join(train_dir, "checkpoints")
    files = sorted(os.listdir(d), key=lambda f: int(f.split("_e")[1].split(".")[0]))
    assert files, "the Trainer wrote no checkpoint"
    return torch.load(os.path.join(d, files[-1]), map_location="cpu"), len(files)


@pytest.mark.parametrize("script,extra,tol", [
    ("e8_colorless_fdn.py", ["--dtype", "float64"], 1e-7),   # Shell(FFT, FDN, |.|), mse_loss + sparsity_loss, get_time_response
    ("e7_biquad.py", ["--dtype", "float64"], 1e-7),          # parallelBiquad, nn.MSELoss, get_freq_response
])
def test_unmodified_reference_example_runs_and_matches(script, extra, tol, tmp_path):
    ours_dir, ref_dir = tmp_path / "b200", tmp_path / "reference"
    out_ours = run("b200", script, ours_dir, extra)
    out_ref = run("reference", script, ref_dir, extra)
    sd_o, n_o = last_checkpoint(ours_dir)
    sd_r, n_r = last_checkpoint(ref_dir)
    assert n_o == n_r  # same number of epochs ran (early stopping behaves alike)
    assert list(sd_o.keys()) == list(sd_r.keys())
    for k in sd_r:
        a, b = sd_o[k].double().numpy(), sd_r[k].double().numpy()
        assert a.shape == b.shape, k
        assert np.abs(a - b).max() <= tol * (1.0 + np.abs(b).max()), (k, float(np.abs(a - b).max()))
    # the per-epoch losses both Trainers print agree as well
    def losses(out):
        return [float(l.split("train_loss:")[1].split()[0]) for l in out.splitlines() if "train_loss:" in l]
    lo, lr = losses(out_ours), losses(out_ref)
    assert len(lo) == len(lr) and len(lo) >= 1
    assert np.allclose(lo, lr, rtol=1e-3, atol=1e-4)
    for f in sorted(os.listdir(ref_dir)):
        if f not in ("checkpoints",):
            assert os.path.exists(os.path.join(ours_dir, f)), f"the reference run wrote {f}, ours did not"


def test_unmodified_nn_in_the_loop_model_matches_and_batches(tmp_path):
    """examples/e7_biquad_nn.py's `nnBiquad` (an MLP conditioning a Biquad through ext_param, one Shell call per batch
    item) — the class as the reference ships it — on both engines: same output, loss and gradients; and on this engine
    the per-item loop replaced by ONE call with the batched parameter tensor gives the same numbers."""
    res = {}
    for engine in ("b200", "reference"):
        out = tmp_path / f"{engine}.npz"
        r = subprocess.run([sys.executable, os.path.join(HERE, "run_reference_nn_model.py"), engine, str(out)],
                           capture_output=True, text=True, timeout=900, env=dict(os.environ, OMP_NUM_THREADS="4"))
        assert r.returncode == 0, f"{engine}:\n{r.stdout[-2000:]}\n{r.stderr[-4000:]}"
        res[engine] = dict(np.load(out))
    o, r = res["b200"], res["reference"]
    assert abs(float(o["loss"]) - float(r["loss"])) <= 1e-9 * abs(float(r["loss"]))
    assert np.abs(o["y"] - r["y"]).max() <= 1e-9 * np.abs(r["y"]).max()
    assert np.array_equal(o["biquad_param"], r["biquad_param"])  # both leave the LAST item's parameters in the module
    grads = [k for k in r if k.startswith("grad_")]
    assert grads and sorted(grads) == sorted(k for k in o if k.startswith("grad_"))
    for k in grads:
        assert np.abs(o[k] - r[k]).max() <= 1e-6 * (np.abs(r[k]).max() + 1e-12), k
    assert float(o["batched_max_diff"]) <= 1e-12 and float(o["batched_grad_diff"]) <= 1e-9
