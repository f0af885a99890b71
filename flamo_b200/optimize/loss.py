"""The two criteria on the sweep path (reference flamo/optimize/loss.py:12-103).  Both are a
handful of PyTorch ops on (B, M, N_out) magnitudes / an (N, N) matrix and stay in PyTorch."""
import numpy as np
import torch
import torch.nn as nn


class sparsity_loss(nn.Module):
    """-(sum|A| - N sqrt N) / (N (sqrt N - 1)) of the mapped feedback matrix (loss.py:36-63)."""

    uses_prediction = False  # depends on the model only: the Trainer may skip materialising y_pred for it

    def forward(self, y_pred, y_target, model):
        core = model.get_core()
        mm = None
        for path in (lambda: core.feedback_loop.feedback,
                     lambda: core.feedback_loop.feedback.mixing_matrix,
                     lambda: core.branchA.feedback_loop.feedback.mixing_matrix):
            try:
                cand = path()
                A = cand.map(cand.param)
                mm = cand
                break
            except Exception:
                continue
        if mm is None:
            raise AttributeError("sparsity_loss: could not locate the feedback mixing matrix in the model")
        from ..processor.dsp import HouseholderMatrix

        if isinstance(mm, HouseholderMatrix):  # the parameter is the unit vector u: A = I - 2 u u^T (loss.py:53-55)
            A = mm._matrix(mm.param)
        N = A.shape[-1]
        from .. import sweep

        hit = getattr(mm, "_expm_cache", None)
        if hit is not None and len(hit) > 2 and hit[1] is A and hit[2] is not None and A.dim() == 2 and N >= 2:
            return hit[2]  # evaluated by the orthogonal map's own kernel (sweep.OrthogonalMap)
        if sweep.SparsityFunction.supported(A):
            return sweep.SparsityFunction.apply(A)
        if A.dim() == 3:
            return torch.mean((torch.sum(torch.abs(A), dim=(-2, -1)) - N * np.sqrt(N)) / (N * (1 - np.sqrt(N))))
        return -(torch.sum(torch.abs(A)) - N * np.sqrt(N)) / (N * (np.sqrt(N) - 1))


class mse_loss(nn.Module):
    """MSE between the channel-summed prediction and the target (loss.py:66-103)."""

    def __init__(self, nfft: int = None, device: str = "cpu"):
        super().__init__()
        self.nfft, self.device = nfft, device
        self.mse_loss = nn.MSELoss()
        self.name = "MSE"

    def forward(self, y_pred, y_true):
        return self.mse_loss(torch.sum(y_pred, dim=-1), y_true.squeeze(-1))
