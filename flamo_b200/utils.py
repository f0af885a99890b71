"""Small helpers mirroring flamo.utils (reference: flamo/utils.py:7-22)."""
import torch


def get_device():
    return torch.device("cuda" if torch.cuda.is_available() else "cpu")


def to_complex(x: torch.Tensor) -> torch.Tensor:
    """Real tensor -> complex tensor with zero imaginary part (flamo/utils.py:12-22)."""
    return torch.complex(x, torch.zeros_like(x))


def save_audio(filepath, x, fs=48000, subtype="PCM_24"):
    """flamo.utils.save_audio (flamo/utils.py:25-30): write `x` (samples[, channels]) as a WAV file.  Uses soundfile
    when it is installed (as the reference does), else the standard library's `wave` (24- or 16-bit PCM).  Not on the
    sweep path — kept so that the reference's example scripts run unchanged."""
    import os

    folder = os.path.dirname(filepath)
    if folder and not os.path.exists(folder):
        os.makedirs(folder)
    data = x.detach().cpu().numpy() if torch.is_tensor(x) else x
    try:
        import soundfile as sf

        sf.write(filepath, data, fs, subtype=subtype)
        return
    except ImportError:
        pass
    import wave

    import numpy as np

    data = np.asarray(data, dtype=np.float64)
    if data.ndim == 1:
        data = data[:, None]
    width = 3 if subtype == "PCM_24" else 2
    full = float(2 ** (8 * width - 1) - 1)
    q = np.clip(np.round(data * full), -full - 1, full).astype("<i4")
    raw = q.reshape(-1, 1).view(np.uint8).reshape(-1, 4)[:, :width].tobytes()
    with wave.open(filepath, "wb") as w:
        w.setnchannels(data.shape[1])
        w.setsampwidth(width)
        w.setframerate(int(fs))
        w.writeframes(raw)
