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
