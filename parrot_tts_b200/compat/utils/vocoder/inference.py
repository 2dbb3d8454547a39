"""reference utils/vocoder/inference.py:48-53,65-74"""
from parrot_tts_b200.vocoder.inference import MAX_WAV_VALUE, generate, load_checkpoint  # noqa: F401
