"""reference inference.py:10-23"""
from parrot_tts_b200.tte import LitParrot, Parrot  # noqa: F401
