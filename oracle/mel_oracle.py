"""ORACLE (test infrastructure - never imported by the product): CPU restatement of the reference's log-mel spectrogram.

  mel_spectrogram()          utils/vocoder/dataset.py:43-69 (+ spectral_normalize_torch :85-95), on torch.stft
  slaney_mel_filterbank()    librosa.filters.mel with its defaults (htk=False, norm='slaney'), the third-party function the reference
                             calls at dataset.py:51.  librosa (requirements.txt:8, unpinned) is neither vendored nor installed here:
                             restated from its published algorithm, loop form - PARITY UNPINNED for this matrix.
Pinned: tests/test_oracle_vs_reference.py runs the reference's own mel_spectrogram (oracle/ref_shim.load_reference_mel_spectrogram,
which supplies this filterbank for the missing librosa call) against mel_spectrogram() below; tests/golden/mel_*.npz holds its output.
"""
import math

import numpy as np
import torch


def _hz_to_mel(f):
    f_sp = 200.0 / 3
    if f >= 1000.0:
        return 1000.0 / f_sp + math.log(f / 1000.0) / (math.log(6.4) / 27.0)
    return f / f_sp


def _mel_to_hz(m):
    f_sp = 200.0 / 3
    min_log_mel = 1000.0 / f_sp
    if m >= min_log_mel:
        return 1000.0 * math.exp((math.log(6.4) / 27.0) * (m - min_log_mel))
    return f_sp * m


def slaney_mel_filterbank(sr, n_fft, n_mels, fmin=0.0, fmax=None):
    fmax = sr / 2.0 if fmax is None else float(fmax)
    n_bins = 1 + n_fft // 2
    fft_f = [sr / 2.0 * k / (n_bins - 1) for k in range(n_bins)]
    lo, hi = _hz_to_mel(float(fmin)), _hz_to_mel(fmax)
    mel_f = [_mel_to_hz(lo + (hi - lo) * i / (n_mels + 1)) for i in range(n_mels + 2)]
    w = np.zeros((n_mels, n_bins), dtype=np.float64)
    for i in range(n_mels):
        for k in range(n_bins):
            lower = (fft_f[k] - mel_f[i]) / (mel_f[i + 1] - mel_f[i])
            upper = (mel_f[i + 2] - fft_f[k]) / (mel_f[i + 2] - mel_f[i + 1])
            w[i, k] = max(0.0, min(lower, upper)) * 2.0 / (mel_f[i + 2] - mel_f[i])
    return w.astype(np.float32)


def mel_spectrogram(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, center=False):
    """y f32 [B, N] -> f32 [B, num_mels, frames]   (dataset.py:43-69)."""
    mel = torch.from_numpy(slaney_mel_filterbank(sampling_rate, n_fft, num_mels, fmin, fmax))
    window = torch.hann_window(win_size)
    p = int((n_fft - hop_size) / 2)
    y = torch.nn.functional.pad(y.unsqueeze(1), (p, p), mode="reflect").squeeze(1)
    spec = torch.stft(y, n_fft, hop_length=hop_size, win_length=win_size, window=window, center=center, pad_mode="reflect",
                      normalized=False, onesided=True, return_complex=True)
    spec = torch.sqrt(spec.real.pow(2) + spec.imag.pow(2) + 1e-9)
    return torch.log(torch.clamp(torch.matmul(mel, spec), min=1e-5))
