"""script-mode ``from dataset import MAX_WAV_VALUE`` (reference utils/vocoder/inference.py:24-25); the dataset classes are training-side"""
from parrot_tts_b200.vocoder import MAX_WAV_VALUE  # noqa: F401
from parrot_tts_b200.vocoder.mel import mel_spectrogram  # noqa: F401  (utils/vocoder/inference.py:24,54)
