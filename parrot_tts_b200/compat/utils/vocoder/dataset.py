"""reference utils/vocoder/dataset.py:22"""
from parrot_tts_b200.vocoder import MAX_WAV_VALUE  # noqa: F401
