"""Small helpers mirroring flamo.utils (reference: flamo/utils.py:7-22)."""
import torch


def get_device():
    return torch.device("cuda" if torch.cuda.is_available() else "cpu")


def to_complex(x: torch.Tensor) -> torch.Tensor:
    """Real tensor -> complex tensor with zero imaginary part (flamo/utils.py:12-22)."""
    return torch.complex(x, torch.zeros_like(x))
