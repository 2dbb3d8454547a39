"""The reference's vocoder inference helpers that sit on the hot path (reference utils/vocoder/inference.py:65-74,
104-109,136-137), on the sm_100a backend."""
import time

import torch

from .. import ops

MAX_WAV_VALUE = 32768.0  # reference utils/vocoder/dataset.py:22


def load_checkpoint(filepath, device="cpu"):
    """torch.load of a ``g_XXXXXXXX`` file -> dict with key 'generator' (reference utils/vocoder/inference.py:48-53)."""
    return torch.load(filepath, map_location=device, weights_only=True)


def generate(h, generator, code):
    """(int16 numpy audio, rtf) for ``generator(**code)`` - reference utils/vocoder/inference.py:65-74.

    Same outputs as the reference; the x32768 and the int16 cast (truncation toward zero, two's-complement wrap)
    run on the device so only 2 bytes per sample cross PCIe, and the wall clock brackets a synchronised region
    (the reference's timing omits ``cuda.synchronize`` and under-reports GPU time).
    """
    start = time.time()
    y_g_hat = generator(**code)
    if type(y_g_hat) is tuple:
        y_g_hat = y_g_hat[0]
    audio = ops.wav_to_pcm16(y_g_hat.squeeze()).cpu().numpy()
    rtf = (time.time() - start) / (y_g_hat.shape[-1] / h.sampling_rate)
    return audio, rtf


def normalize(audio, device="cuda"):
    """Drop-in for ``librosa.util.normalize(audio.astype(np.float32))`` as the reference calls it on generated audio
    (utils/vocoder/inference.py:169, demo.ipynb cell 13): int16 PCM -> float32 divided by the utterance's peak.
    numpy in -> numpy out (the reduction and the division run on the device); a device tensor [n] / [B, n] stays there."""
    if torch.is_tensor(audio):
        return ops.peak_normalize(audio)
    import numpy as np
    a = np.ascontiguousarray(audio)
    if a.dtype != np.int16:
        raise TypeError("normalize expects the int16 audio that generate() returns")
    return ops.peak_normalize(torch.from_numpy(a).to(device)).cpu().numpy()


def stream_chunks(n_units, chunk_units=1024, halo_units=32):
    """[(lo, hi, a, b)]: chunk i synthesises units [a, b) from the window [lo, hi) = [a - halo, b + halo) clipped to the
    utterance.  The generator is purely convolutional with a receptive-field radius of ~23 units,
    so with halo >= 24 every kept sample equals the unchunked result."""
    out = []
    for a in range(0, n_units, chunk_units):
        b = min(a + chunk_units, n_units)
        out.append((max(a - halo_units, 0), min(b + halo_units, n_units), a, b))
    return out


def stream_generate(generator, code, spkr=None, chunk_units=1024, halo_units=32, pcm16=False):
    """Long-form synthesis in chunks (BASELINE configs[4]): yields (first_sample, audio) per chunk, audio f32 [B, n] on the
    device (or int16 with ``pcm16=True``), n = hop * (b - a).  ``code`` i64 [B, T] on the device.  Each chunk is one
    ``generator(code=window, spkr=spkr)`` call on the window of units it needs; nothing is recomputed beyond the two halos."""
    T = code.shape[1]
    hop = generator.hop
    for lo, hi, a, b in stream_chunks(T, chunk_units, halo_units):
        kw = {"code": code[:, lo:hi].contiguous()}
        if spkr is not None:
            kw["spkr"] = spkr
        wav = generator(**kw)[:, 0, (a - lo) * hop:(b - lo) * hop]
        yield a * hop, (ops.wav_to_pcm16(wav.contiguous()) if pcm16 else wav)
