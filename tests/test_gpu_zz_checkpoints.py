"""GPU: the 28 checkpoints the reference ships (notebooks/output/ex_fdn, ex_biquad — SURVEY.md §8c's fixtures) loaded
into this package's models with load_state_dict; forward |.|, Shell.get_freq_response and Shell.get_time_response
(reference system.py:1012-1153) against what the unmodified reference returned for the same checkpoint
(tests/golden/reference_checkpoint_responses.npz, written by tests/golden/make_golden.py).

float64 kernels: 1e-9.  float32 kernels: 1e-4 on the magnitude (floored relative metric); the
de-aliased responses are held to 1e-3 of their peak because the rising envelope of get_freq_response multiplies the
float32 rounding of the late impulse-response samples by up to 10**(30/20)."""
import pytest
import torch

from helpers import CHECKPOINTS, check_checkpoint

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag,epoch", CHECKPOINTS)
def test_checkpoint_float64(tag, epoch):
    check_checkpoint(tag, epoch, torch.float64, "cuda", tol_mag=1e-9, tol_resp=1e-9)


@pytest.mark.parametrize("tag,epoch", CHECKPOINTS)
def test_checkpoint_float32(tag, epoch):
    check_checkpoint(tag, epoch, torch.float32, "cuda", tol_mag=1e-4, tol_resp=1e-3)
