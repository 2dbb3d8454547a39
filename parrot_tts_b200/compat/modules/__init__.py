"""reference modules/__init__.py (the dataset classes are training-side and not mirrored)"""
from parrot_tts_b200.tte import Parrot  # noqa: F401
