"""script-mode ``from models import CodeGenerator`` (reference utils/vocoder/inference.py:27)"""
from parrot_tts_b200.vocoder.models import CodeGenerator, Generator, ResBlock1, ResBlock2  # noqa: F401
