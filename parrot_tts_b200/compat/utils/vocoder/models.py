"""reference utils/vocoder/models.py"""
from parrot_tts_b200.vocoder.models import CodeGenerator, Generator, ResBlock1, ResBlock2  # noqa: F401
