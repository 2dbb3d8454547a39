"""Log-mel spectrogram of generated audio on the device (reference utils/vocoder/dataset.py:43-69, called from
utils/vocoder/inference.py:54 and the training loop) - the validation metric beside the 1e-3 sample tolerance.

The STFT, magnitude, mel projection and log run in one sm_100a kernel (``pb_mel_spectrogram``).  The mel basis is
``librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)`` with its defaults (Slaney scale, Slaney area normalisation), restated here
from the published algorithm because ``librosa`` is not installed (parity-unpinned for that matrix)."""
import ctypes as C

import numpy as np
import torch

from .. import _lib


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr, n_fft, n_mels, fmin=0.0, fmax=None):
    """``librosa.filters.mel(sr=, n_fft=, n_mels=, fmin=, fmax=)`` -> float32 [n_mels, 1 + n_fft // 2]: triangular filters on the
    Slaney mel scale (linear below 1 kHz, logarithmic above), each scaled to unit area (``norm='slaney'``)."""
    fmax = float(sr) / 2 if fmax is None else float(fmax)
    fftfreqs = np.linspace(0.0, float(sr) / 2, 1 + n_fft // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    weights = np.maximum(0.0, np.minimum(lower, upper))
    weights *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return weights.astype(np.float32)


_basis = {}


def mel_spectrogram(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, center=False):
    """Same signature and result layout as the reference: y f32 [B, N] on a CUDA device -> f32 [B, num_mels, frames]."""
    if center:
        raise NotImplementedError("center=True is not implemented (the reference always calls it with center=False)")
    if not y.is_cuda:
        raise _lib.ParrotB200Error("mel_spectrogram needs a CUDA tensor (no CPU path)")
    y = y.to(torch.float32).contiguous()
    if y.dim() != 2:
        raise ValueError("mel_spectrogram: y must be [B, N]")
    key = (int(sampling_rate), int(n_fft), int(num_mels), float(fmin), None if fmax is None else float(fmax), str(y.device))
    if key not in _basis:
        _basis[key] = torch.from_numpy(mel_filterbank(sampling_rate, n_fft, num_mels, fmin, fmax)).to(y.device)
    B, N = y.shape
    lib = _lib.load()
    frames = C.c_int32(0)
    with torch.cuda.device(y.device):
        st = _lib.stream_ptr(y.device)
        _lib.check(lib.pb_mel_spectrogram(None, B, N, None, n_fft, hop_size, win_size, num_mels, None, C.byref(frames), st))
        out = torch.empty(B, num_mels, frames.value, device=y.device, dtype=torch.float32)
        _lib.check(lib.pb_mel_spectrogram(_lib.ptr(y), B, N, _lib.ptr(_basis[key]), n_fft, hop_size, win_size, num_mels, _lib.ptr(out),
                                          C.byref(frames), st))
    return out
